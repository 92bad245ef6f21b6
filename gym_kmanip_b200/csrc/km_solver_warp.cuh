// km_solver_warp.cuh -- acceleration stage + constraint solver of the warp-per-env mapping (G == 32), register resident.
//
// Same algorithm and iteration path as fwd_constraint of km_sim.cuh (mj_fwdAcceleration + mj_solNewton on the primal
// problem, SURVEY.md A4/A5: warm-start choice, Newton direction from H = M + J^T D_active J, exact 1-D Newton line search
// with bracketing, the same stopping tests), restated for one warp per env with the solver state in registers:
//   * lane i < NV owns dof i: qacc, M qacc, gradient, search direction, its friction-loss row and its joint-limit row
//     (every dof has at most one of each) -- the forces of those rows never leave the lane;
//   * sixteen lanes own the base rows (normal, two tangents, torsion) of the up to four table-corner contacts, one row
//     each, with the row's cube columns in registers; an edge lane k = 1..3 carries the two pyramid rows n +- mu_k t_k.
//     For the solo-arm scene (NV = 16) these are lanes 16..31; the 26-dof scenes put them on lanes 0..15 as extra slots;
//   * H is block diagonal while no finger pad touches the cube (one block per kinematic chain + the 6 x 6 cube block).
//     All blocks are factorised at once, one lane per row with block-local columns in registers and pivots moved by
//     shuffles (10 steps for the largest chain instead of NV).  The chain blocks only change when a friction-loss or
//     limit row of the arm changes state, so their factor is cached across Newton iterations; the cube block (the
//     contacts) is rebuilt and refactorised by its six lanes in every iteration;
//   * one line-search evaluation is a handful of selects per lane and one two-value butterfly.
// While a finger pad touches the cube (well under 1 % of env-steps) the generic solver of km_sim.cuh runs instead.
#pragma once

namespace km {

#if defined(__CUDACC__) || defined(KM_WARP_EMU)


// CPL: the instantiation for envs whose cube touches a finger pad (pad contacts carry arm columns, H is dense).  It is a
// separate instantiation so that the common one carries none of its state or code (28 warps per SM: 72 registers).
template <class S, typename T, class E, bool CPL> struct WarpSolver {
  typedef Dim<S> D;
  typedef Num<T> N;
  static constexpr int NV = D::NV, NVA = D::NVA, BS = max_block<S>();
  static constexpr int CL0 = NV <= 16 ? 16 : 0;       // first lane of the contact role
  static constexpr bool SH = CL0 == 0;                // contact rows share lanes with dofs (extra slots)
  // contact sets: a contact lane carries one base row of contact cc (set 0) and, in the coupled instantiation, of contact
  // cc + 4 (set 1: finger pads + table corners can add up to eight contacts); EA + 2 s = first slot of set s's edge pair
  static constexpr int NCS = CPL ? 2 : 1, EA = SH ? 2 : 0, NS = EA + 2 * NCS;
  static_assert(NV <= 32 && D::NFRIC <= NV, "one lane per dof");

  E& e;
  const Model<S, T>& m;
  const Grp<32>& g;
  // roles
  int lane, dofi, b0, bn, li, cl, cc, cb, ncon, cslot[NCS];
  unsigned csup[NCS];
  bool isdof, isarm, iscube, iscon[NCS], isedge[NCS], has_f, has_l;
  // row constants: slot 0 = friction-loss row of the dof (or a pyramid edge on the solo-arm contact lanes), slot 1 = limit
  T cdiag, rf0, fl0, sg, mu[NCS], Dr[NS];
  // solver state
  T qacc, Ma, grad, search, Mv, qs, as, jar[NS], jv[NS], dinv, hd_cached;
  int evals;

  KM_DI WarpSolver(E& e_, const Model<S, T>& m_, const Grp<32>& g_) : e(e_), m(m_), g(g_) {}

  KM_DI T* xs() const { return e.c.search; }   // broadcast buffer of a dof-space vector
  KM_DI T* fbs() const { return CPL ? e.c.efc_force : e.c.Mv; }   // base-row forces of the contacts (16 per contact set)

  // M x for this lane's dof (x of every dof is in xs())
  KM_DI T mulM(T xi) const {
    T s = 0;
#pragma unroll
    for (int j = 0; j < NVA; j++) s += e.M[isarm ? dofi : 0][j] * xs()[j];
    return isarm ? s : (iscube ? cdiag * xi : T(0));
  }
  // J x for this lane's rows (cube part of x in xs())
  KM_DI void jrows(T xi, T* out) const {
    out[0] = has_f ? xi : T(0);
    out[1] = has_l ? sg * xi : T(0);
    sfor<0, NCS>([&](auto Cs) {
      constexpr int cs = decltype(Cs)::value;
      T pb = 0;
      const T* jr = e.Jq[iscon[cs] ? cc + 4 * cs : 0][iscon[cs] ? cb : 0];
#pragma unroll
      for (int k = 0; k < 6; k++) pb += jr[k] * xs()[NVA + k];
      pb = iscon[cs] ? pb : T(0);
      if constexpr (CPL) {             // finger-pad contacts: arm columns of the base row (its dof support only)
        if (iscon[cs] && cslot[cs] < D::NPAD)
          for (int j = 0; j < NVA; j++) pb += ((csup[cs] >> j) & 1u) ? e.Ja[cslot[cs]][cb][j] * xs()[j] : T(0);
      }
      const T p0 = __shfl_sync(0xffffffffu, pb, lane & ~3);
      constexpr int ea = EA + 2 * cs;
      if (SH || cs > 0) { out[ea] = isedge[cs] ? p0 + mu[cs] * pb : T(0); out[ea + 1] = isedge[cs] ? p0 - mu[cs] * pb : T(0); }
      else if (isedge[cs]) { out[0] = p0 + mu[cs] * pb; out[1] = p0 - mu[cs] * pb; }
    });
  }
  // cost and force of slot s at residual x
  template <int s> KM_DI T rowcost(T x, T* f) const {
    if (s == 0 && has_f) {
      if (x <= -rf0) { *f = fl0; return -fl0 * (T(0.5) * rf0 + x); }
      if (x >= rf0) { *f = -fl0; return -fl0 * (T(0.5) * rf0 - x); }
      *f = -Dr[0] * x;
      return T(0.5) * Dr[0] * x * x;
    }
    if (x < T(0)) { *f = -Dr[s] * x; return T(0.5) * Dr[s] * x * x; }
    *f = 0;
    return 0;
  }
  KM_DI void sum2(T& a, T& b) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
  }

  // forces, cost and gradient at the current jar / Ma / qacc; *gn receives |grad|^2
  KM_DI T update(T* gn) {
    T c = 0, f[NS];
    sfor<0, NS>([&](auto Sx) { constexpr int s = decltype(Sx)::value; c += rowcost<s>(jar[s], &f[s]); });
    T qfc = (SH || isdof) ? f[0] + sg * f[1] : T(0);
    sfor<0, NCS>([&](auto Cs) {
      constexpr int cs = decltype(Cs)::value;
      const T fp = isedge[cs] ? f[EA + 2 * cs] : T(0), fn = isedge[cs] ? f[EA + 2 * cs + 1] : T(0);
      T sn = fp + fn;
      sn += __shfl_xor_sync(0xffffffffu, sn, 1);
      sn += __shfl_xor_sync(0xffffffffu, sn, 2);
      if (cl >= 0 && cl < 16) fbs()[16 * cs + cl] = iscon[cs] ? (cb == 0 ? sn : mu[cs] * (fp - fn)) : T(0);
    });
    g.sync();
    if (iscube)
      for (int c2 = 0; c2 < ncon; c2++)
#pragma unroll
        for (int b = 0; b < 4; b++) qfc += e.Jq[c2][b][li] * fbs()[4 * c2 + b];
    if constexpr (CPL) if (isarm)
      for (int c2 = 0; c2 < ncon; c2++) {
        const int sl = e.con_slot[c2];
        if (sl < D::NPAD && ((e.con_sup[c2] >> lane) & 1u))
#pragma unroll
          for (int b = 0; b < 4; b++) qfc += e.Ja[sl][b][lane] * fbs()[4 * c2 + b];
      }
    const T r = Ma - qs;
    grad = isdof ? r - qfc : T(0);
    c += T(0.5) * r * (qacc - as);
    T g2 = grad * grad;
    sum2(c, g2);
    *gn = g2;
    return c;
  }

  // Branch-free on purpose: ptxas wraps every shuffle that may follow a divergent branch in a WARPSYNC.COLLECTIVE
  // sequence, so loops that contain shuffles are controlled by votes (warp-uniform by construction) and lanes without
  // a row carry dummy values instead of branching around the collective code.
  static constexpr unsigned FULL = 0xffffffffu;
  // A zero the compiler cannot see through.  Shuffle source lanes (b0 + j) and similar per-lane constants are loop
  // invariant, so ptxas hoists all of them out of the Newton loop into registers it does not have (the kernel runs 28
  // warps per SM: 72 registers) and then spills other state to local memory, which thrashes the small L1 left beside the
  // shared-memory carve-out.  Adding this zero inside the loop makes them cheap to recompute instead.
  KM_DI int opaque_zero() const {
#if defined(__CUDA_ARCH__)
    int z;
    asm volatile("mov.u32 %0, 0;" : "=r"(z));
    return z;
#else
    return 0;
#endif
  }
  // Cholesky of every diagonal block at once: one lane per row, block-local columns in row[], pivots by shuffles
  // (columns past the end of a lane's block receive garbage from foreign lanes; nothing reads them)
  KM_DI void factor_blocks(T* row) {
    const int s0 = b0 + opaque_zero();
    sfor<0, BS>([&](auto J) {
      constexpr int j = decltype(J)::value;
      const T ajj = __shfl_sync(FULL, row[j], s0 + j);
      const T inv = N::rsqrt(tmax(ajj, N::minval()));
      const T lij = row[j] * inv;
      row[j] = lij;
      dinv = li == j ? inv : dinv;
      sfor<j + 1, BS>([&](auto K) {
        constexpr int k = decltype(K)::value;
        row[k] -= lij * __shfl_sync(FULL, lij, s0 + k);
      });
    });
  }
  // Cholesky of the 6 x 6 cube block alone, rows in w[] on the cube lanes (the other lanes carry dummies)
  KM_DI void factor_cube(T* w, T* dw) {
    sfor<0, 6>([&](auto J) {
      constexpr int j = decltype(J)::value;
      const T ajj = __shfl_sync(FULL, w[j], NVA + j);
      const T inv = N::rsqrt(tmax(ajj, N::minval()));
      const T lij = w[j] * inv;
      w[j] = lij;
      *dw = li == j ? inv : *dw;
      sfor<j + 1, 6>([&](auto K) {
        constexpr int k = decltype(K)::value;
        w[k] -= lij * __shfl_sync(FULL, lij, NVA + k);
      });
    });
  }
  // x = H^{-1} rhs with the factor in row[] / dinv (the rows are also in e.c.H for the transposed access)
  KM_DI T solve(T rhs) const {
    const int s0 = b0 + opaque_zero();
    const T* hrow = e.c.H[dofi];                 // row i of the factor: L[i][j], j < li
    const T* hcol = &e.c.H[b0][li];              // column li of the factor: L[b0 + j][li], j > li (row stride HS)
    T acc = rhs, y = 0;
    sfor<0, BS>([&](auto J) {
      constexpr int j = decltype(J)::value;
      const T yj = __shfl_sync(FULL, acc * dinv, s0 + j);
      y = li == j ? yj : y;
      const T l = hrow[j];
      acc = li > j ? acc - l * yj : acc;         // (select, not a product with zero: yj of a foreign lane may be anything)
    });
    T acc2 = y, x = 0;
    sfor_rev<BS>([&](auto J) {
      constexpr int j = decltype(J)::value;
      const T xj = __shfl_sync(FULL, acc2 * dinv, s0 + j);
      x = li == j ? xj : x;
      const T l = hcol[(j < bn ? j : 0) * D::HS];
      acc2 = (j > li && j < bn) ? acc2 - l * xj : acc2;
    });
    return x;
  }
  // rows of the factor (N block-local columns) to e.c.H (lower triangle) for the lanes of `which`
  template <int N> KM_DI void store_rows(const T* row, bool which) {
    sfor<0, N>([&](auto J) { constexpr int j = decltype(J)::value; if (which && j <= li) e.c.H[dofi][j] = row[j]; });
    g.sync();
  }

  // Coupled case (a finger pad touches the cube, well under 1 % of the env-steps -- but the kernel ends with its slowest
  // env, and that is a coupled one): dense H = M + diag + sum_c J_c^T W_c J_c over all NV dofs.  One row per lane: for
  // every contact the lane forms u = W_c J_c[:, i] from its own column of the base rows (W_c is the 4 x 4 arrow-head of
  // the active pyramid edges) and adds u . J_c[:, j] for the columns j of the contact's support (broadcast loads);
  // table-corner contacts only touch the cube block.  The factor stays in e.c.H / dinv and, like the block factors of
  // the common case, is reused while the active set (hd, pm, nm of both contact sets) is unchanged.
  KM_DI void dense_factor(T hd, unsigned pm0, unsigned nm0, unsigned pm1, unsigned nm1) {
    T row[NV];
    sfor<0, NV>([&](auto J) {
      constexpr int j = decltype(J)::value;
      T v = 0;
      if constexpr (j < NVA) v = isarm ? e.M[isarm ? dofi : 0][j] : T(0);
      row[j] = j == dofi ? v + hd + cdiag : v;
    });
    for (int c = 0; c < ncon; c++) {   // warp-uniform
      const int slot = e.con_slot[c];
      const unsigned sup = e.con_sup[c];
      const bool pad = slot < D::NPAD;
      const bool in = isdof && ((sup >> dofi) & 1u);
      const int sa = pad ? slot : 0;
      T own[4];
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const T v = iscube ? e.Jq[c][b][iscube ? li : 0] : e.Ja[sa][b][isarm ? dofi : 0];
        own[b] = in ? v : T(0);
      }
      const unsigned pmc = c < 4 ? pm0 : pm1, nmc = c < 4 ? nm0 : nm1;   // contact set of c: lanes CL0 + 4 (c mod 4) + k
      const int sh = CL0 + 4 * (c & 3);
      const T Dc = e.con_D[c];
      T cnt = 0, u[4] = {0, 0, 0, 0};
#pragma unroll
      for (int k = 1; k < 4; k++) {
        const T p = (pmc >> (sh + k)) & 1u ? T(1) : T(0), q = (nmc >> (sh + k)) & 1u ? T(1) : T(0);
        const T muk = e.con_mu[c][k - 1];
        const T w0 = Dc * muk * (p - q), wk = Dc * muk * muk * (p + q);
        cnt += p + q;
        u[0] += w0 * own[k];
        u[k] = w0 * own[0] + wk * own[k];
      }
      u[0] += Dc * cnt * own[0];
      if (pad) {
        sfor<0, NVA>([&](auto J) {
          constexpr int j = decltype(J)::value;
          if ((sup >> j) & 1u) row[j] += u[0] * e.Ja[sa][0][j] + u[1] * e.Ja[sa][1][j] + u[2] * e.Ja[sa][2][j] + u[3] * e.Ja[sa][3][j];
        });
      }
      sfor<NVA, NV>([&](auto J) {
        constexpr int j = decltype(J)::value, k = j - NVA;
        row[j] += u[0] * e.Jq[c][0][k] + u[1] * e.Jq[c][1][k] + u[2] * e.Jq[c][2][k] + u[3] * e.Jq[c][3][k];
      });
    }
    T di = 1;
    sfor<0, NV>([&](auto J) {
      constexpr int j = decltype(J)::value;
      const T ajj = __shfl_sync(FULL, row[j], j);
      const T inv = N::rsqrt(tmax(ajj, N::minval()));
      const T lij = row[j] * inv;
      row[j] = lij;
      di = dofi == j ? inv : di;
      sfor<j + 1, NV>([&](auto K) {
        constexpr int k = decltype(K)::value;
        row[k] -= lij * __shfl_sync(FULL, lij, k);
      });
    });
    dinv = di;
    sfor<0, NV>([&](auto J) { constexpr int j = decltype(J)::value; if (isdof && j <= dofi) e.c.H[dofi][j] = row[j]; });
    g.sync();
  }
  // x = H^{-1} rhs with the dense factor in e.c.H (row i: L[i][j], j < i; column i: L[j][i], j > i) and dinv
  KM_DI T dense_solve(T rhs) const {
    const T* hrow = e.c.H[dofi];
    T acc = rhs, y = 0;
    sfor<0, NV>([&](auto J) {
      constexpr int j = decltype(J)::value;
      const T yj = __shfl_sync(FULL, acc * dinv, j);
      y = dofi == j ? yj : y;
      const T l = hrow[j];
      acc = dofi > j ? acc - l * yj : acc;
    });
    T acc2 = y, x = 0;
    sfor_rev<NV>([&](auto J) {
      constexpr int j = decltype(J)::value;
      const T xj = __shfl_sync(FULL, acc2 * dinv, j);
      x = dofi == j ? xj : x;
      const T l = e.c.H[j][dofi];
      acc2 = j > dofi ? acc2 - l * xj : acc2;
    });
    return x;
  }

  // Newton direction: search = -H^{-1} grad
  KM_DI void direction() {
    const bool q0 = has_f ? (jar[0] > -rf0 && jar[0] < rf0) : (jar[0] < T(0));
    const bool q1 = jar[1] < T(0);
    const T hd = ((has_f && q0) ? Dr[0] : T(0)) + ((has_l && q1) ? Dr[1] : T(0));
    const unsigned pm = __ballot_sync(FULL, isedge[0] && (SH ? jar[EA] < T(0) : q0));
    const unsigned nm = __ballot_sync(FULL, isedge[0] && (SH ? jar[EA + 1] < T(0) : q1));
    if constexpr (CPL) {
      const unsigned pm1 = __ballot_sync(FULL, isedge[1] && jar[EA + 2] < T(0)), nm1 = __ballot_sync(FULL, isedge[1] && jar[EA + 3] < T(0));
      int* key = e.c.efc_state;
      const bool rebuild = __any_sync(FULL, (isdof && hd != hd_cached) || (unsigned)key[0] != pm || (unsigned)key[1] != nm ||
                                                (unsigned)key[2] != pm1 || (unsigned)key[3] != nm1);
      if (rebuild) {
        g.sync();
        if (lane == 0) { key[0] = (int)pm; key[1] = (int)nm; key[2] = (int)pm1; key[3] = (int)nm1; }
        hd_cached = hd;
        dense_factor(hd, pm, nm, pm1, nm1);
      }
      { const T x = dense_solve(grad); search = isdof ? -x : T(0); }   // every lane takes part in the shuffles
      return;
    }
    const bool refactor = __any_sync(FULL, isarm && hd != hd_cached);
    // The cube block depends on the states of the pyramid rows (pm, nm) and of the cube's friction-loss rows only: when
    // none of them changed since the previous iteration (the usual case once the active set has settled) its factor in
    // e.c.H / dinv is still valid and the rebuild is skipped
    int* key = e.c.efc_state;
    const bool rebuild = __any_sync(FULL, (iscube && hd != hd_cached) || (unsigned)key[0] != pm || (unsigned)key[1] != nm);
    if (refactor || rebuild) g.sync();
#ifdef KM_PHASE_CLOCKS
    if (lane == 0) { e.clk[14] += rebuild ? 1u : 0u; e.clk[15] += refactor ? 1u : 0u; }   // profiling build: event counts
#endif
    if (rebuild) {
    if (lane == 0) { key[0] = (int)pm; key[1] = (int)nm; }
    // cube block: diag + sum over contacts of Jq^T W Jq, one lower-triangle entry (ei, ej) per lane
    const int t = (lane < 21 ? lane : 20) + opaque_zero();
    const int ei = (t >= 1) + (t >= 3) + (t >= 6) + (t >= 10) + (t >= 15), ej = t - ei * (ei + 1) / 2;
    const T hdi = __shfl_sync(FULL, hd, NVA + ei);
    if (lane < 21) {
      T h = ei == ej ? (ei < 3 ? m.cube_mass : m.cube_inertia[ei < 3 ? 0 : ei - 3]) + hdi : T(0);
      for (int c = 0; c < ncon; c++) {
        const T Dc = e.con_D[c];
        const T ni = e.Jq[c][0][ei], nj = e.Jq[c][0][ej];
        T cnt = 0, acc = 0;
#pragma unroll
        for (int k = 1; k < 4; k++) {
          const T p = (pm >> (CL0 + 4 * c + k)) & 1u ? T(1) : T(0), q = (nm >> (CL0 + 4 * c + k)) & 1u ? T(1) : T(0);
          const T muk = e.con_mu[c][k - 1];
          const T ti = e.Jq[c][k][ei], tj = e.Jq[c][k][ej];
          cnt += p + q;
          acc += Dc * muk * (p - q) * (ni * tj + ti * nj) + Dc * muk * muk * (p + q) * ti * tj;
        }
        h += Dc * cnt * ni * nj + acc;
      }
      e.c.H[NVA + ei][ej] = h;
      e.c.H[NVA + ej][ei] = h;
    }
    g.sync();
    }
    if (refactor) {   // warp-uniform (a vote): a friction-loss or limit row of the arm changed state
      T row[BS], dsave = dinv;
      sfor<0, BS>([&](auto J) {
        constexpr int j = decltype(J)::value;
        const T mv = e.M[isarm ? dofi : 0][isarm ? b0 + (j < bn ? j : 0) : 0];
        row[j] = isarm ? (j < bn ? mv + (j == li ? hd : T(0)) : T(0)) : (j == li ? T(1) : T(0));
      });
      hd_cached = isarm ? hd : hd_cached;
      factor_blocks(row);
      dinv = isarm ? dinv : dsave;
      store_rows<BS>(row, isarm);
    }
    if (rebuild) {
      T w[6], dw = 1;
      sfor<0, 6>([&](auto J) { constexpr int j = decltype(J)::value; w[j] = e.c.H[iscube ? dofi : NVA][j]; });
      g.sync();
      hd_cached = iscube ? hd : hd_cached;
      factor_cube(w, &dw);
      dinv = iscube ? dw : dinv;
      store_rows<6>(w, iscube);
    }
    { const T x = solve(grad); search = isdof ? -x : T(0); }   // every lane takes part in the shuffles
  }

  // exact line search along `search` (mj_solNewton's, as sol_linesearch of km_sim.cuh); returns the step, 0 = no progress
  KM_DI T linesearch(T scale) {
    if (isdof) xs()[lane] = search;
    g.sync();
    Mv = mulM(search);
    jrows(search, jv);
    g.sync();
    T sn = search * search, a1 = search * (Ma - qs), a2 = T(0.5) * search * Mv;
    sum2(sn, a1);
    a2 = g.sum(a2);
    const T snorm = N::sqrt(sn);
    if (__all_sync(FULL, snorm < N::minval())) return 0;
    const T qg1 = a1, qg2 = a2;
    // per-row quadratic coefficients of the 1-D cost while the row is in its quadratic zone
    T ra[NS], rb[NS];
    sfor<0, NS>([&](auto Sx) { constexpr int s = decltype(Sx)::value; ra[s] = Dr[s] * jar[s] * jv[s]; rb[s] = T(0.5) * Dr[s] * jv[s] * jv[s]; });
    auto eval = [&](T alpha, T* d1, T* d2) {
      T q1 = 0, q2 = 0;
      sfor<0, NS>([&](auto Sx) {
        constexpr int s = decltype(Sx)::value;
        const T x = jar[s] + alpha * jv[s];
        if (s == 0 && has_f) {
          if (x <= -rf0) q1 += -fl0 * jv[0];
          else if (x >= rf0) q1 += fl0 * jv[0];
          else { q1 += ra[0]; q2 += rb[0]; }
        } else if (x < T(0)) { q1 += ra[s]; q2 += rb[s]; }
      });
      sum2(q1, q2);
      q1 += qg1; q2 += qg2;
      *d1 = T(2) * alpha * q2 + q1;
      *d2 = T(2) * q2;
      evals++;
    };
    // one evaluation site shared by all phases (the loop body stays small): phase 0 = slope at alpha 0, 1 = Newton steps
    // to the right until the slope changes sign, 2 = safeguarded Newton inside the bracket
    T d1 = 0, d2 = 0, gtol = 0, lo = 0, lo_d1 = 0, lo_d2 = 1, hi = 0, hi_d1 = 0, hi_d2 = 1, result = 0;
    int phase = 0, it = 0;
    bool run = true;
#pragma unroll 1
    while (__any_sync(FULL, run)) {
      T a = 0;
      if (phase == 1) {
        if (it >= m.ls_iterations) { result = lo; run = false; }
        else a = lo - lo_d1 / lo_d2;
      } else if (phase == 2) {
        const bool lo_closer = N::abs(lo_d1) < N::abs(hi_d1);
        if (it >= m.ls_iterations) { result = lo_closer ? lo : hi; run = false; }
        else {
          a = lo_closer ? lo - lo_d1 / lo_d2 : hi - hi_d1 / hi_d2;
          if (!(a > lo && a < hi)) a = T(0.5) * (lo + hi);
          if (a == lo || a == hi) { result = lo_closer ? lo : hi; run = false; }
        }
      }
      if (__any_sync(FULL, run)) {
        eval(a, &d1, &d2);
        if (phase == 0) {
          gtol = tmax(m.tol * m.ls_tol * snorm / scale, T(64) * N::eps() * N::abs(d1));
          if (N::abs(d1) < gtol || d1 > T(0)) { result = 0; run = false; }
          else { lo = 0; lo_d1 = d1; lo_d2 = d2; phase = 1; }
        } else if (N::abs(d1) < gtol) { result = a; run = false; }
        else if (d1 > T(0)) {
          hi = a; hi_d1 = d1; hi_d2 = d2;
          if (phase == 1) phase = 2; else it++;     // the bracketing evaluation does not advance the count
        } else { lo = a; lo_d1 = d1; lo_d2 = d2; it++; }
      }
    }
    return result;
  }

  // mj_fwdAcceleration (qacc_smooth = M^{-1} qfrc_smooth) + mj_fwdConstraint
  KM_DI void run() {
    lane = g.lane;
    isdof = lane < NV; isarm = lane < NVA; iscube = isdof && !isarm;
    dofi = isdof ? lane : NV - 1;
    b0 = m.blk0[dofi]; bn = m.blkn[dofi]; li = dofi - b0;
    ncon = e.ncon;
    cl = lane - CL0; cc = (cl >> 2) & 3; cb = cl & 3;
    sfor<0, NCS>([&](auto Cs) {
      constexpr int cs = decltype(Cs)::value;
      iscon[cs] = cl >= 0 && cl < 16 && cc + 4 * cs < ncon;
      cslot[cs] = (CPL && iscon[cs]) ? e.con_slot[cc + 4 * cs] : D::NPAD;
      csup[cs] = (CPL && iscon[cs]) ? e.con_sup[cc + 4 * cs] : 0u;
      isedge[cs] = iscon[cs] && cb > 0;
    });
    const int base = D::NFRIC + e.nlim;
    cdiag = iscube ? (li < 3 ? m.cube_mass : m.cube_inertia[li < 3 ? 0 : li - 3]) : T(0);
    // rows of this lane
    const int fr = isdof ? m.dof_fric[dofi] : -1, lr = isarm ? e.dof_lim[dofi] : -1;
    has_f = fr >= 0; has_l = lr >= 0;
    T ar[NS];
    sfor<0, NS>([&](auto Sx) { constexpr int s = decltype(Sx)::value; Dr[s] = 0; ar[s] = 0; jar[s] = 0; jv[s] = 0; });
    rf0 = has_f ? m.fr_Rf[fr] : T(0); fl0 = has_f ? m.fr_loss[fr] : T(0);
    if (has_f) { Dr[0] = m.fr_D[fr]; ar[0] = e.efc_aref[fr]; }
    sg = T(1);
    if (has_l) { Dr[1] = e.efc_D[lr]; ar[1] = e.efc_aref[lr]; sg = efc_neg(e.efc_desc[lr]) ? T(-1) : T(1); }
    sfor<0, NCS>([&](auto Cs) {
      constexpr int cs = decltype(Cs)::value;
      mu[cs] = 0;
      if (isedge[cs]) {
        const int c = cc + 4 * cs, rp = base + 6 * c + 2 * (cb - 1);
        mu[cs] = e.con_mu[c][cb - 1];
        Dr[EA + 2 * cs] = e.con_D[c]; Dr[EA + 2 * cs + 1] = e.con_D[c];
        ar[EA + 2 * cs] = e.efc_aref[rp]; ar[EA + 2 * cs + 1] = e.efc_aref[rp + 1];
      }
    });
    evals = 0;
    qs = isdof ? e.qfrc_smooth[dofi] : T(0);
    // ---- smooth acceleration: factor M (every block), qacc_smooth = M^{-1} qfrc_smooth
    dinv = 1; hd_cached = 0;
    if (lane == 0) { e.c.efc_state[0] = -1; e.c.efc_state[1] = -1; }   // no cube-block factor yet (pm = ~0 cannot occur)
    {
      T row[BS];
      sfor<0, BS>([&](auto J) {
        constexpr int j = decltype(J)::value;
        const T mv = e.M[isarm ? dofi : 0][isarm ? b0 + (j < bn ? j : 0) : 0];
        row[j] = isarm ? (j < bn ? mv : T(0)) : (j == li ? (iscube ? cdiag : T(1)) : T(0));
      });
      factor_blocks(row);
      store_rows<BS>(row, isdof);
    }
    { const T x = solve(qs); as = isdof ? x : T(0); }   // every lane takes part in the shuffles
    if (isdof) e.qacc_smooth[lane] = as;
    // ---- warm start: the previous qacc if it is cheaper than the unconstrained acceleration (rolled: one code site)
    const T wi = isdof ? e.warm[dofi] : T(0);
    T cw = 0;
    qacc = wi;
#pragma unroll 1
    for (int cand = 0; cand < 2; cand++) {
      const T ai = cand ? as : wi;
      if (isdof) xs()[lane] = ai;
      g.sync();
      const T mv = mulM(ai);
      T r[NS], f, c = 0;
      jrows(ai, r);
      sfor<0, NS>([&](auto Sx) { constexpr int s = decltype(Sx)::value; r[s] -= ar[s]; c += rowcost<s>(r[s], &f); });
      c += T(0.5) * (mv - qs) * (ai - as);
      g.sync();
      c = g.sum(c);
      if (cand == 0 || cw > c) {   // candidate 1 (qacc_smooth) replaces the warm start only if strictly cheaper
        qacc = ai; Ma = mv;
        sfor<0, NS>([&](auto Sx) { constexpr int s = decltype(Sx)::value; jar[s] = r[s]; });
      }
      if (cand == 0) cw = c;
    }
    const T scale = T(1) / (m.meaninertia * T(NV));
    KM_CLK(CLK_SOL_SETUP);
    // Newton iterations; the loop is rotated so that update / direction / linesearch each have one code site
    T gn, cost = 0;
    int niter = 0;
#pragma unroll 1
    while (true) {
      const T newcost = update(&gn);
      bool done = niter >= m.iterations;
      if (niter > 0) done = done || scale * (cost - newcost) < m.tol || scale * N::sqrt(gn) < m.tol;
      cost = newcost;
      KM_CLK(CLK_SOL_UPD);
      if (__any_sync(FULL, done)) break;
      direction();
      KM_CLK(CLK_SOL_DIR);
      const T alpha = linesearch(scale);
      KM_CLK(CLK_SOL_LS);
      if (__all_sync(FULL, alpha == T(0))) break;
      qacc += alpha * search; Ma += alpha * Mv;
      sfor<0, NS>([&](auto Sx) { constexpr int s = decltype(Sx)::value; jar[s] += alpha * jv[s]; });
      niter++;
#ifdef KM_PHASE_CLOCKS
      if (lane == 0) e.clk[9] += 1u;   // CLK_SOL_VOTE slot (unused by this solver): Newton iterations of the env step
#endif
    }
    if (isdof) { e.qacc[lane] = qacc; e.warm[lane] = qacc; }
    if (lane == 0) { e.solver_niter = niter; e.ls_evals += evals; }
    g.sync();
  }
};

template <class S, typename T, bool CPL, class E> KM_DN void fwd_acc_constraint_w(E& e, const Model<S, T>& m, const Grp<32>& g) {
  WarpSolver<S, T, E, CPL> s(e, m, g);
  s.run();
}

#endif  // __CUDACC__ || KM_WARP_EMU

}  // namespace km
