// km_solver_tpe.cuh -- constraint solver of the thread-per-env mapping (G == 1).
//
// Same algorithm as the lane-group solver of km_sim.cuh (mj_solNewton on the primal problem, SURVEY.md A5: warm-start
// choice, Newton direction from H = M + J^T D_active J, exact 1-D Newton line search with bracketing, the same stopping
// tests), restated for one thread per env:
//   * linear algebra by structure: while no finger pad touches the cube ("uncoupled", the case this file handles) H is
//     block diagonal -- one dense block per kinematic chain (known at compile time from the scene header) plus the 6x6
//     cube block -- and every block is factorised fully unrolled in registers;
//   * constraint rows by kind instead of a generic row list: friction-loss rows (fixed dofs), joint-limit rows (one
//     per joint, possibly inactive), and the four base rows (normal, two tangents, torsion) of each table-corner
//     contact from which its six pyramid rows are formed on the fly; those contacts only have cube columns;
//   * control flow that keeps the 32 envs of a warp together: the line search is a state machine with ONE evaluation
//     site per iteration, so envs that need a different number of evaluations or Newton iterations idle in place
//     instead of serialising whole code paths.
// While a pad touches the cube ("coupled", well under 1 % of env-steps) the pad contacts add arm columns to their base
// rows; the touched chain block is then eliminated into the cube block (Schur complement), so the factorisations stay
// block-sized and in registers.
#pragma once

// KM_TPE_K_ROLL = 1 keeps the loops over the three pyramid edge pairs of a contact rolled (smaller hot loop bodies)
#ifndef KM_TPE_K_ROLL
#define KM_TPE_K_ROLL 0
#endif
#if KM_TPE_K_ROLL
#define KM_K_LOOP _Pragma("unroll 1")
#else
#define KM_K_LOOP
#endif

namespace km {

KM_HD constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }
// dof of friction-loss row r (a constexpr function: static constexpr arrays of the scene cannot be odr-used in device code)
template <class S> constexpr int fric_dof_of(int r) { return S::fric_dof[r]; }

// In-register Cholesky of an SPD matrix given by its packed lower triangle: L <- factor (off-diagonal entries scaled),
// dinv <- reciprocal pivots; apply solves one right-hand side with it.
template <typename T, int n> KM_HD void chol_factor_packed(T* L, T* dinv) {
  typedef Num<T> N;
  sfor<0, n>([&](auto J) {
    constexpr int j = decltype(J)::value;
    T d = L[tri(j, j)];
    sfor<0, j>([&](auto K) { constexpr int k = decltype(K)::value; d -= L[tri(j, k)] * L[tri(j, k)]; });
    const T inv = N::rsqrt(tmax(d, N::minval()));
    dinv[j] = inv;
    sfor<j + 1, n>([&](auto I) {
      constexpr int i = decltype(I)::value;
      T s = L[tri(i, j)];
      sfor<0, j>([&](auto K) { constexpr int k = decltype(K)::value; s -= L[tri(i, k)] * L[tri(j, k)]; });
      L[tri(i, j)] = s * inv;
    });
  });
}
template <typename T, int n> KM_HD void chol_apply_packed(const T* L, const T* dinv, T* x) {
  sfor<0, n>([&](auto J) {
    constexpr int j = decltype(J)::value;
    T s = x[j];
    sfor<0, j>([&](auto K) { constexpr int k = decltype(K)::value; s -= L[tri(j, k)] * x[k]; });
    x[j] = s * dinv[j];
  });
  sfor_rev<n>([&](auto J) {
    constexpr int j = decltype(J)::value;
    T s = x[j];
    sfor<j + 1, n>([&](auto K) { constexpr int k = decltype(K)::value; s -= L[tri(k, j)] * x[k]; });
    x[j] = s * dinv[j];
  });
}
template <typename T, int n> KM_HD void chol_solve_packed(T* L, T* x) {
  T dinv[n];
  chol_factor_packed<T, n>(L, dinv);
  chol_apply_packed<T, n>(L, dinv, x);
}

// dst[B0..B1) = (M + diag(hd))^{-1} rhs on the chain block [B0, B1), then the blocks after it
template <class S, typename T, int B0, class E> KM_HD void tpe_arm_solve(const E& e, const T* hd, const T* rhs, T* dst) {
  if constexpr (B0 < S::NVA) {
    constexpr int B1 = blk_end<S>(B0), n = B1 - B0;
    T L[n * (n + 1) / 2], x[n];
    sfor<0, n>([&](auto I) {
      constexpr int i = decltype(I)::value;
      sfor<0, i + 1>([&](auto J) { constexpr int j = decltype(J)::value; L[tri(i, j)] = e.M[B0 + i][B0 + j]; });
      if (hd) L[tri(i, i)] += hd[B0 + i];
      x[i] = rhs[B0 + i];
    });
    chol_solve_packed<T, n>(L, x);
    sfor<0, n>([&](auto I) { dst[B0 + decltype(I)::value] = x[decltype(I)::value]; });
    tpe_arm_solve<S, T, B1>(e, hd, rhs, dst);
  }
}

// out = M x (chain blocks dense, cube block diagonal)
template <class S, typename T, int B0, class E> KM_HD void tpe_arm_mulM(const E& e, const T* x, T* out) {
  if constexpr (B0 < S::NVA) {
    constexpr int B1 = blk_end<S>(B0), n = B1 - B0;
    T xv[n];
    sfor<0, n>([&](auto J) { xv[decltype(J)::value] = x[B0 + decltype(J)::value]; });
    sfor<0, n>([&](auto I) {
      constexpr int i = decltype(I)::value;
      T s = 0;
      // lower triangle only: M is symmetric by construction (com_crb stores one value in both halves), so the result is
      // identical and the upper half drops out of the solver's hot set in local memory
      sfor<0, n>([&](auto J) { constexpr int j = decltype(J)::value; s += e.M[B0 + (i > j ? i : j)][B0 + (i > j ? j : i)] * xv[j]; });
      out[B0 + i] = s;
    });
    tpe_arm_mulM<S, T, B1>(e, x, out);
  }
}
template <class S, typename T, class E> KM_HD void tpe_mulM(const E& e, const Model<S, T>& m, const T* x, T* out) {
  tpe_arm_mulM<S, T, 0>(e, x, out);
  for (int k = 0; k < 3; k++) { out[S::NVA + k] = m.cube_mass * x[S::NVA + k]; out[S::NVA + 3 + k] = m.cube_inertia[k] * x[S::NVA + 3 + k]; }
}

// M^{-1} rhs (mj_fwdAcceleration / qacc_smooth) in the thread-per-env mapping
template <class S, typename T, class E> KM_HD void tpe_solveM(const E& e, const Model<S, T>& m, const T* rhs, T* dst) {
  tpe_arm_solve<S, T, 0>(e, (const T*)0, rhs, dst);
  for (int k = 0; k < 3; k++) { dst[S::NVA + k] = rhs[S::NVA + k] / m.cube_mass; dst[S::NVA + 3 + k] = rhs[S::NVA + 3 + k] / m.cube_inertia[k]; }
}

// cost / force of one non-friction row (limit, pyramid edge) at residual x
template <typename T> KM_HD T uni_cost(T x, T Dr, T* force) {
  if (x < T(0)) { *force = -Dr * x; return T(0.5) * Dr * x * x; }
  *force = 0;
  return 0;
}
template <typename T> KM_HD T fric_cost(T x, T Dr, T rf, T floss, T* force) {
  if (x <= -rf) { *force = floss; return -floss * (T(0.5) * rf + x); }
  if (x >= rf) { *force = -floss; return -floss * (T(0.5) * rf - x); }
  *force = -Dr * x;
  return T(0.5) * Dr * x * x;
}

template <class S, typename T, class E> struct TpeSolver {
  typedef Dim<S> D;
  typedef Num<T> N;
  static constexpr int NVA = D::NVA, NV = D::NV, NF = D::NFRIC;
  E& e;
  const Model<S, T>& m;
  int nc, base;
  // joints whose limit row is active / whose active row has J = -e, gathered once per solve: every routine below walks
  // the joints, and testing a register bit replaces the dof_lim[j] (and efc_desc) loads from the env record -- most
  // joints are away from their limits
  unsigned limmask, negmask;
  int evals;   // line-search evaluations of this solve (added to the env's diagnostic counter once, at the end)
  KM_HD TpeSolver(E& e_, const Model<S, T>& m_) : e(e_), m(m_), nc(e_.ncon), base(D::NFRIC + e_.nlim), limmask(0), negmask(0), evals(0) {
    for (int j = 0; j < NVA; j++) {
      const int r = e.dof_lim[j];
      if (r >= 0) {
        limmask |= 1u << j;
        if (efc_neg(e.efc_desc[r])) negmask |= 1u << j;
      }
    }
  }
  KM_HD bool lim_on(int j) const { return (limmask >> j) & 1u; }
  // friction coefficients of contact c: the model's (shared memory) -- make_constraint copies exactly these into
  // e.con_mu, and reading them there would be three more local-memory words per contact in every routine
  KM_HD const T* mu_of(int c) const { const int sl = e.con_slot[c]; return sl < D::NPAD ? m.pad_mu[sl] : m.tab_mu; }

  KM_HD void mulM(const T* x, T* out) const { tpe_mulM<S, T>(e, m, x, out); }
  KM_HD T lim_sign(int j) const { return ((negmask >> j) & 1u) ? T(-1) : T(1); }

  // J x without the reference acceleration: friction rows, limit rows (signed), contact base rows
  KM_HD void rows(const T* x, T* xf, T* xl, T (*xb)[4]) const {
    sfor<0, NF>([&](auto R) { constexpr int r = decltype(R)::value; constexpr int d = fric_dof_of<S>(r); xf[r] = x[d]; });
    for (int j = 0; j < NVA; j++) xl[j] = lim_on(j) ? lim_sign(j) * x[j] : T(0);
    for (int c = 0; c < nc; c++)
      for (int b = 0; b < 4; b++) xb[c][b] = brow(c, b, x);
  }
  // base row b of contact c times a dof-space vector: cube columns always, arm columns for finger-pad contacts
  // A table-corner contact has the fixed frame of the plane normal (0,0,1): normal (0,0,1), tangents (0,1,0) and
  // (-1,0,0) (makeframe), so the cube-translation columns of its four base rows are the exact constants 0 / +-1
  // (make_constraint forms them as dot products with unit vectors).  Using the constants instead of loading those
  // twelve words gives bit-identical sums and takes them out of the solver's hot set in local memory.
  KM_HD T brow(int c, int b, const T* x) const {
    T s = 0;
    const int sl = e.con_slot[c];
    if (sl >= D::NPAD) {
      s = b == 0 ? x[NVA + 2] : (b == 1 ? x[NVA + 1] : (b == 2 ? -x[NVA] : T(0)));
      for (int k = 3; k < 6; k++) s += e.Jq[c][b][k] * x[NVA + k];
      return s;
    }
    for (int k = 0; k < 6; k++) s += e.Jq[c][b][k] * x[NVA + k];
    if (sl < D::NPAD) {
      const unsigned sup = e.con_sup[c];
      for (int j = 0; j < NVA; j++) if ((sup >> j) & 1u) s += e.Ja[sl][b][j] * x[j];
    }
    return s;
  }

  // total cost (constraint + Gauss) at acceleration a (warm-start choice); uses t.Mv as scratch
  KM_HD T cost_at(const T* a) {
    auto& t = e.t;
    mulM(a, t.Mv);
    T c = 0, f;
    sfor<0, NF>([&](auto R) {
      constexpr int r = decltype(R)::value;
      constexpr int d = fric_dof_of<S>(r);
      c += fric_cost(a[d] - e.efc_aref[r], m.fr_D[r], m.fr_Rf[r], m.fr_loss[r], &f);
    });
    for (int j = 0; j < NVA; j++) {
      if (lim_on(j)) { const int r = e.dof_lim[j]; c += uni_cost(lim_sign(j) * a[j] - e.efc_aref[r], e.efc_D[r], &f); }
    }
    for (int ci = 0; ci < nc; ci++) {
      T pb[4];
      for (int b = 0; b < 4; b++) pb[b] = brow(ci, b, a);
      const T Dc = e.con_D[ci];
      const T* mu3 = mu_of(ci);
      KM_K_LOOP
      for (int k = 0; k < 3; k++) {
        const T tk = mu3[k] * pb[1 + k];
        c += uni_cost(pb[0] + tk - e.efc_aref[base + 6 * ci + 2 * k], Dc, &f);
        c += uni_cost(pb[0] - tk - e.efc_aref[base + 6 * ci + 2 * k + 1], Dc, &f);
      }
    }
    T gs = 0;
    for (int i = 0; i < NV; i++) gs += (t.Mv[i] - e.qfrc_smooth[i]) * (a[i] - e.qacc_smooth[i]);
    return c + T(0.5) * gs;
  }

  // forces, cost and gradient at the current jar / Ma / qacc
  KM_HD T update() {
    auto& t = e.t;
    T c = 0, qfc[NV];
    for (int i = 0; i < NV; i++) qfc[i] = 0;
    sfor<0, NF>([&](auto R) {
      constexpr int r = decltype(R)::value;
      T f;
      c += fric_cost(t.jar_f[r], m.fr_D[r], m.fr_Rf[r], m.fr_loss[r], &f);
      constexpr int d = fric_dof_of<S>(r);
      qfc[d] += f;
    });
    sfor<0, NVA>([&](auto Jj) {
      constexpr int j = decltype(Jj)::value;
      if (lim_on(j)) {
        T f;
        c += uni_cost(t.jar_l[j], e.efc_D[e.dof_lim[j]], &f);
        qfc[j] += lim_sign(j) * f;
      }
    });
    for (int ci = 0; ci < nc; ci++) {
      const T Dc = e.con_D[ci], p0 = t.jarb[ci][0];
      const T* mu3 = mu_of(ci);
      T fb[4] = {0, 0, 0, 0};
      KM_K_LOOP
      for (int k = 0; k < 3; k++) {
        const T mu = mu3[k], tk = mu * t.jarb[ci][1 + k];
        T fp, fn;
        c += uni_cost(p0 + tk - e.efc_aref[base + 6 * ci + 2 * k], Dc, &fp);
        c += uni_cost(p0 - tk - e.efc_aref[base + 6 * ci + 2 * k + 1], Dc, &fn);
        fb[0] += fp + fn;
        fb[1 + k] = mu * (fp - fn);
      }
      const int sl = e.con_slot[ci];
      if (sl >= D::NPAD) {   // table corner: constant translation columns (see brow)
        qfc[NVA + 0] += -fb[2];
        qfc[NVA + 1] += fb[1];
        qfc[NVA + 2] += fb[0];
        sfor<3, 6>([&](auto K) {
          constexpr int k = decltype(K)::value;
          qfc[NVA + k] += e.Jq[ci][0][k] * fb[0] + e.Jq[ci][1][k] * fb[1] + e.Jq[ci][2][k] * fb[2] + e.Jq[ci][3][k] * fb[3];
        });
      } else {
        sfor<0, 6>([&](auto K) {
          constexpr int k = decltype(K)::value;
          qfc[NVA + k] += e.Jq[ci][0][k] * fb[0] + e.Jq[ci][1][k] * fb[1] + e.Jq[ci][2][k] * fb[2] + e.Jq[ci][3][k] * fb[3];
        });
      }
      if (sl < D::NPAD) {
        const unsigned sup = e.con_sup[ci];
        sfor<0, NVA>([&](auto Jj) {
          constexpr int j = decltype(Jj)::value;
          if ((sup >> j) & 1u) qfc[j] += e.Ja[sl][0][j] * fb[0] + e.Ja[sl][1][j] * fb[1] + e.Ja[sl][2][j] * fb[2] + e.Ja[sl][3][j] * fb[3];
        });
      }
    }
    T gs = 0;
    sfor<0, NV>([&](auto I) {
      constexpr int i = decltype(I)::value;
      const T r = t.Ma[i] - e.qfrc_smooth[i];
      gs += r * (e.qacc[i] - e.qacc_smooth[i]);
      t.grad[i] = r - qfc[i];
    });
    return c + T(0.5) * gs;
  }

  // search = -H^{-1} grad with H = M + J^T diag(D of the quadratic rows) J, block by block
  KM_HD void direction() {
    auto& t = e.t;
    T hd[NV];
    for (int i = 0; i < NV; i++) hd[i] = 0;
    sfor<0, NF>([&](auto R) {
      constexpr int r = decltype(R)::value;
      const T x = t.jar_f[r];
      constexpr int d = fric_dof_of<S>(r);
      if (x > -m.fr_Rf[r] && x < m.fr_Rf[r]) hd[d] += m.fr_D[r];
    });
    sfor<0, NVA>([&](auto Jj) {
      constexpr int j = decltype(Jj)::value;
      if (lim_on(j) && t.jar_l[j] < T(0)) hd[j] += e.efc_D[e.dof_lim[j]];
    });
    // cube block first: C = diag + sum over ALL contacts of Jq^T W Jq, right-hand side g_c
    T Hc[21], xc[6];
    for (int i = 0; i < 21; i++) Hc[i] = 0;
    sfor<0, 6>([&](auto K) {
      constexpr int k = decltype(K)::value;
      Hc[tri(k, k)] = (k < 3 ? m.cube_mass : m.cube_inertia[k < 3 ? 0 : k - 3]) + hd[NVA + k];
      xc[k] = t.grad[NVA + k];
    });
    for (int ci = 0; ci < nc; ci++) {
      T w00, w0[3], wk[3];
      contact_weights(ci, &w00, w0, wk);
      T Jr[4][6], Y[4][6];
      const bool corner = e.con_slot[ci] >= D::NPAD;   // table corner: constant translation columns (see brow)
      sfor<0, 6>([&](auto K) {
        constexpr int k = decltype(K)::value;
        if (k < 3 && corner) {
          Jr[0][k] = k == 2 ? T(1) : T(0); Jr[1][k] = k == 1 ? T(1) : T(0); Jr[2][k] = k == 0 ? T(-1) : T(0); Jr[3][k] = T(0);
        } else
          for (int b = 0; b < 4; b++) Jr[b][k] = e.Jq[ci][b][k];
        Y[0][k] = w00 * Jr[0][k] + w0[0] * Jr[1][k] + w0[1] * Jr[2][k] + w0[2] * Jr[3][k];
        for (int b = 1; b < 4; b++) Y[b][k] = w0[b - 1] * Jr[0][k] + wk[b - 1] * Jr[b][k];
      });
      sfor<0, 6>([&](auto I) {
        constexpr int i = decltype(I)::value;
        sfor<0, i + 1>([&](auto Jj) {
          constexpr int j = decltype(Jj)::value;
          Hc[tri(i, j)] += Jr[0][i] * Y[0][j] + Jr[1][i] * Y[1][j] + Jr[2][i] * Y[2][j] + Jr[3][i] * Y[3][j];
        });
      });
    }
    // chain blocks; a block touched by a pad contact is eliminated into the cube block (Schur complement)
    unsigned schur = 0;
    chain_blocks<0>(hd, Hc, xc, &schur);
    chol_solve_packed<T, 6>(Hc, xc);
    sfor<0, 6>([&](auto K) { t.search[NVA + decltype(K)::value] = xc[decltype(K)::value]; });
    if (schur)
      for (int i = 0; i < NVA; i++)
        if ((schur >> i) & 1u) {
          T sacc = t.search[i];
          for (int k = 0; k < 6; k++) sacc -= t.Z[i][k] * xc[k];
          t.search[i] = sacc;
        }
    for (int i = 0; i < NV; i++) t.search[i] = -t.search[i];
  }

  // Chain block [B0, B1) of H:  A = M + diag(hd) + sum_pads Ja^T W Ja,  B = sum_pads Ja^T W Jq  (n x 6).
  // y = A^{-1} g_a goes to t.search; with pads also Z = A^{-1} B goes to t.Z, the cube block becomes C - B^T Z and its
  // right-hand side g_c - B^T y; the caller finishes x_a = y - Z x_c for the blocks flagged in *schur.
  template <int B0> KM_HD void chain_blocks(const T* hd, T* Hc, T* xc, unsigned* schur) {
    if constexpr (B0 < NVA) {
      auto& t = e.t;
      constexpr int B1 = blk_end<S>(B0), n = B1 - B0;
      T L[n * (n + 1) / 2], dinv[n], x[n];
      sfor<0, n>([&](auto I) {
        constexpr int i = decltype(I)::value;
        sfor<0, i + 1>([&](auto J) { constexpr int j = decltype(J)::value; L[tri(i, j)] = e.M[B0 + i][B0 + j]; });
        L[tri(i, i)] += hd[B0 + i];
        x[i] = t.grad[B0 + i];
      });
      bool pads = false;
      if (e.coupled)
        for (int ci = 0; ci < nc; ci++) {
          const int sl = e.con_slot[ci];
          if (sl >= D::NPAD) break;                         // pad contacts come first
          const int pl = m.pad_link[sl];
          if (pl < B0 || pl >= B1) continue;
          if (!pads) {
            pads = true;
            for (int i = 0; i < n; i++) for (int k = 0; k < 6; k++) t.Z[B0 + i][k] = 0;
          }
          T w00, w0[3], wk[3];
          contact_weights(ci, &w00, w0, wk);
          const unsigned sup = e.con_sup[ci];
          T Ja[4][n], Ya[4][n];
          sfor<0, n>([&](auto I) {
            constexpr int i = decltype(I)::value;
            const bool on = (sup >> (B0 + i)) & 1u;
            for (int b = 0; b < 4; b++) Ja[b][i] = on ? e.Ja[sl][b][B0 + i] : T(0);
            Ya[0][i] = w00 * Ja[0][i] + w0[0] * Ja[1][i] + w0[1] * Ja[2][i] + w0[2] * Ja[3][i];
            for (int b = 1; b < 4; b++) Ya[b][i] = w0[b - 1] * Ja[0][i] + wk[b - 1] * Ja[b][i];
          });
          sfor<0, n>([&](auto I) {
            constexpr int i = decltype(I)::value;
            sfor<0, i + 1>([&](auto J) {
              constexpr int j = decltype(J)::value;
              L[tri(i, j)] += Ja[0][i] * Ya[0][j] + Ja[1][i] * Ya[1][j] + Ja[2][i] * Ya[2][j] + Ja[3][i] * Ya[3][j];
            });
          });
          for (int k = 0; k < 6; k++) {
            const T q0 = e.Jq[ci][0][k], q1 = e.Jq[ci][1][k], q2 = e.Jq[ci][2][k], q3 = e.Jq[ci][3][k];
            sfor<0, n>([&](auto I) {
              constexpr int i = decltype(I)::value;
              t.Z[B0 + i][k] += Ya[0][i] * q0 + Ya[1][i] * q1 + Ya[2][i] * q2 + Ya[3][i] * q3;
            });
          }
        }
      chol_factor_packed<T, n>(L, dinv);
      T bcol[n];
      if (pads) {   // right-hand side of the cube block before y overwrites x:  g_c - B^T y  needs y, so solve first
        chol_apply_packed<T, n>(L, dinv, x);
        T by[6];
        for (int k = 0; k < 6; k++) {
          T sacc = 0;
          sfor<0, n>([&](auto I) { constexpr int i = decltype(I)::value; bcol[i] = t.Z[B0 + i][k]; sacc += bcol[i] * x[i]; });
          by[k] = sacc;
          sfor<0, n>([&](auto I) { constexpr int i = decltype(I)::value; t.Bm[k][i] = bcol[i]; });   // keep B for the Schur update
          chol_apply_packed<T, n>(L, dinv, bcol);
          sfor<0, n>([&](auto I) { constexpr int i = decltype(I)::value; t.Z[B0 + i][k] = bcol[i]; });
        }
        sfor<0, 6>([&](auto K) { xc[decltype(K)::value] -= by[decltype(K)::value]; });
        // C -= B^T Z (symmetric 6 x 6)
        sfor<0, 6>([&](auto K) {
          constexpr int k = decltype(K)::value;
          sfor<0, k + 1>([&](auto Ll) {
            constexpr int l = decltype(Ll)::value;
            T sacc = 0;
            for (int i = 0; i < n; i++) sacc += t.Bm[k][i] * t.Z[B0 + i][l];
            Hc[tri(k, l)] -= sacc;
          });
        });
        *schur |= ((1u << n) - 1u) << B0;
      } else chol_apply_packed<T, n>(L, dinv, x);
      sfor<0, n>([&](auto I) { t.search[B0 + decltype(I)::value] = x[decltype(I)::value]; });
      chain_blocks<B1>(hd, Hc, xc, schur);
    }
  }

  // arrow-head weights of the four base rows of contact ci from the states of its six pyramid rows
  KM_HD void contact_weights(int ci, T* w00, T* w0, T* wk) const {
    auto& t = e.t;
    const T Dc = e.con_D[ci], p0 = t.jarb[ci][0];
    const T* mu3 = mu_of(ci);
    T n = 0;
    KM_K_LOOP
    for (int k = 0; k < 3; k++) {
      const T mu = mu3[k], tk = mu * t.jarb[ci][1 + k];
      const T p = p0 + tk - e.efc_aref[base + 6 * ci + 2 * k] < T(0) ? T(1) : T(0);
      const T q = p0 - tk - e.efc_aref[base + 6 * ci + 2 * k + 1] < T(0) ? T(1) : T(0);
      n += p + q;
      w0[k] = Dc * mu * (p - q);
      wk[k] = Dc * mu * mu * (p + q);
    }
    *w00 = Dc * n;
  }

  // first and second derivative of the 1-D cost along the search direction at step alpha
  KM_HD void ls_eval(T qg1, T qg2, T alpha, T* d1, T* d2) {
    auto& t = e.t;
    T q1 = 0, q2 = 0;
    // Rolled on purpose (KM_TPE_LS_UNROLL = 1 restores the unrolled form): this function is two thirds of the kernel's
    // dynamic instructions, and unrolled over the 8 friction and 10 limit rows its loop body is 36 KB of SASS, more than
    // the 32 KB instruction cache the warps of an SM share.  Same rows in the same order: identical results.
#ifndef KM_TPE_LS_UNROLL
#define KM_TPE_LS_UNROLL 0
#endif
#if KM_TPE_LS_UNROLL
#pragma unroll
#else
#pragma unroll 1
#endif
    for (int r = 0; r < NF; r++) {
      const T jv = t.jv_f[r], jar = t.jar_f[r], x = jar + alpha * jv, rf = m.fr_Rf[r];
      if (x <= -rf) q1 += -m.fr_loss[r] * jv;
      else if (x >= rf) q1 += m.fr_loss[r] * jv;
      else { q1 += m.fr_D[r] * jar * jv; q2 += T(0.5) * m.fr_D[r] * jv * jv; }
    }
#if KM_TPE_LS_UNROLL
#pragma unroll
#else
#pragma unroll 1
#endif
    for (int j = 0; j < NVA; j++) {
      if (lim_on(j)) {
        const T jv = t.jv_l[j], jar = t.jar_l[j];
        if (jar + alpha * jv < T(0)) { const T Dr = e.efc_D[e.dof_lim[j]]; q1 += Dr * jar * jv; q2 += T(0.5) * Dr * jv * jv; }
      }
    }
    for (int ci = 0; ci < nc; ci++) {
      const T Dc = e.con_D[ci], a0 = t.jarb[ci][0], v0 = t.jvb[ci][0];
      const T* mu3 = mu_of(ci);
      KM_K_LOOP
      for (int k = 0; k < 3; k++) {
        const T mu = mu3[k], ak = mu * t.jarb[ci][1 + k], vk = mu * t.jvb[ci][1 + k];
        const T jarp = a0 + ak - e.efc_aref[base + 6 * ci + 2 * k], jvp = v0 + vk;
        const T jarn = a0 - ak - e.efc_aref[base + 6 * ci + 2 * k + 1], jvn = v0 - vk;
        if (jarp + alpha * jvp < T(0)) { q1 += Dc * jarp * jvp; q2 += T(0.5) * Dc * jvp * jvp; }
        if (jarn + alpha * jvn < T(0)) { q1 += Dc * jarn * jvn; q2 += T(0.5) * Dc * jvn * jvn; }
      }
    }
    q1 += qg1;
    q2 += qg2;
    *d1 = T(2) * alpha * q2 + q1;
    *d2 = T(2) * q2;
    evals++;
  }

  // exact line search; returns the step (0: no progress possible).  One ls_eval site shared by all phases.
  KM_HD T linesearch(T scale) {
    auto& t = e.t;
    T sn = 0;
    for (int i = 0; i < NV; i++) sn += t.search[i] * t.search[i];
    const T snorm = N::sqrt(sn);
    if (snorm < N::minval()) return 0;
    mulM(t.search, t.Mv);
    rows(t.search, t.jv_f, t.jv_l, t.jvb);
    T qg1 = 0, qg2 = 0;
    for (int i = 0; i < NV; i++) {
      qg1 += t.search[i] * (t.Ma[i] - e.qfrc_smooth[i]);
      qg2 += T(0.5) * t.search[i] * t.Mv[i];
    }
    T d1, d2;
    ls_eval(qg1, qg2, T(0), &d1, &d2);
    const T gtol = tmax(m.tol * m.ls_tol * snorm / scale, T(64) * N::eps() * N::abs(d1));
    if (N::abs(d1) < gtol || d1 > T(0)) return 0;
    T lo = 0, lo_d1 = d1, lo_d2 = d2, hi = 0, hi_d1 = 0, hi_d2 = 0, result = 0;
    int phase = 1, it = 0;   // 1: Newton steps to the right until the slope changes sign; 2: safeguarded Newton in the bracket
    while (phase != 0) {
      T a = 0;
      if (it >= m.ls_iterations) {
        result = phase == 1 ? lo : (N::abs(lo_d1) < N::abs(hi_d1) ? lo : hi);
        phase = 0;
      } else if (phase == 1) a = lo - lo_d1 / lo_d2;
      else {
        a = N::abs(lo_d1) < N::abs(hi_d1) ? lo - lo_d1 / lo_d2 : hi - hi_d1 / hi_d2;
        if (!(a > lo && a < hi)) a = T(0.5) * (lo + hi);
        if (a == lo || a == hi) { result = N::abs(lo_d1) < N::abs(hi_d1) ? lo : hi; phase = 0; }
      }
      if (phase != 0) {
        ls_eval(qg1, qg2, a, &d1, &d2);
        if (N::abs(d1) < gtol) { result = a; phase = 0; }
        else if (d1 > T(0)) {
          hi = a; hi_d1 = d1; hi_d2 = d2;
          if (phase == 1) phase = 2; else it++;     // the bracketing evaluation does not advance the count
        } else { lo = a; lo_d1 = d1; lo_d2 = d2; it++; }
      }
    }
    return result;
  }

  KM_HD void solve() {
    auto& t = e.t;
#ifndef KM_TPE_DEBUG
    e.solver_niter = 0;
#endif
    const T cw = cost_at(e.warm), cs = cost_at(e.qacc_smooth);
    for (int i = 0; i < NV; i++) e.qacc[i] = cw > cs ? e.qacc_smooth[i] : e.warm[i];
    mulM(e.qacc, t.Ma);
    rows(e.qacc, t.jar_f, t.jar_l, t.jarb);
    sfor<0, NF>([&](auto R) { t.jar_f[decltype(R)::value] -= e.efc_aref[decltype(R)::value]; });
    for (int j = 0; j < NVA; j++) if (lim_on(j)) t.jar_l[j] -= e.efc_aref[e.dof_lim[j]];
    const T scale = T(1) / (m.meaninertia * T(NV));
    T cost = update();
    int niter = 0;
    bool done = niter >= m.iterations;
    while (!done) {
      direction();
      const T alpha = linesearch(scale);
      if (alpha == T(0)) done = true;
      else {
        for (int i = 0; i < NV; i++) { e.qacc[i] += alpha * t.search[i]; t.Ma[i] += alpha * t.Mv[i]; }
        sfor<0, NF>([&](auto R) { t.jar_f[decltype(R)::value] += alpha * t.jv_f[decltype(R)::value]; });
        for (int j = 0; j < NVA; j++) t.jar_l[j] += alpha * t.jv_l[j];
        for (int c = 0; c < nc; c++) for (int b = 0; b < 4; b++) t.jarb[c][b] += alpha * t.jvb[c][b];
        const T oldcost = cost;
        cost = update();
        T gn = 0;
        for (int i = 0; i < NV; i++) gn += t.grad[i] * t.grad[i];
        niter++;
        done = scale * (oldcost - cost) < m.tol || scale * N::sqrt(gn) < m.tol || niter >= m.iterations;
      }
    }
    for (int i = 0; i < NV; i++) e.warm[i] = e.qacc[i];
    e.ls_evals += evals;
#ifdef KM_TPE_DEBUG
    e.solver_niter += niter + (e.coupled ? 65536 : 0);   // debug build: totals over the env step
#else
    e.solver_niter = niter;
#endif
  }
};

// mj_fwdConstraint in the thread-per-env mapping
template <class S, typename T, class E> KM_FN void fwd_constraint_tpe(E& e, const Model<S, T>& m) {
  TpeSolver<S, T, E> s(e, m);
  s.solve();
}

}  // namespace km
