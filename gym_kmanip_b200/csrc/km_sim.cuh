// km_sim.cuh -- the env-step hot path as lane-group cooperative code (see km_common.cuh for the model).
//
// What each routine replaces in the reference (gym-kmanip) is cited at the routine; third-party MuJoCo /
// dm_control semantics follow SURVEY.md Appendix A.  Per env-step (dm_control legacy_step ordering, A1):
//     load state -> step1 -> before_step (action decode + IK) -> step2 -> 9 x (step1, step2) -> step1(light)
//     -> reward, observation, truncation / autoreset -> store state
// where step1 = position + velocity stage, step2 = actuation + smooth acceleration + Newton constraint
// solve + semi-implicit Euler.
#pragma once
#include "km_model.cuh"

namespace km {

// E is the env working-set type (Env<S, T, TPE>); it is always deduced from the argument
// A/B switches of the one-link-per-lane formulations of the smooth phases (warp-per-env kernels)
#ifndef KM_PAR_KIN
#define KM_PAR_KIN 1
#endif
#ifndef KM_PAR_CRB
#define KM_PAR_CRB 1
#endif
#ifndef KM_PAR_VEL
#define KM_PAR_VEL 1
#endif
#define KM_TPL template <class S, typename T, int G, class E>
#define KM_ARGS E& e, const Model<S, T>& m, const Grp<G>& g

// =========================================================================================== position stage
// mj_kinematics for the articulated links (level by level) and the cube.
KM_TPL KM_FN void kinematics(KM_ARGS) {
  typedef Dim<S> D;
  typedef Num<T> N;
  if (g.lane == G - 1) {   // cube: normalise the quaternion in place as mj_kinematics does
    qnormalize(e.qpos + D::NVA + 3);
    q2mat(e.cmat, e.qpos + D::NVA + 3);
  }
#if KM_WARP_CODE && KM_PAR_KIN
  if constexpr (G >= D::NVA && G > 1) {
    // One link per lane.  Every link first forms its transform relative to its parent (joint included), then the
    // transforms are composed up the tree by pointer jumping: after round r a link's transform is relative to its
    // 2^r-th ancestor (or the world), so four rounds cover chains of up to 16 links -- ~300 warp-instructions with all
    // links busy instead of one lane walking the ten levels (~1 400).  Compared with walking the chain link by link the
    // products associate differently (and the quaternion is normalised once, at the end): differences of a few ulp.
    static_assert(D::MAXLEVEL <= 16, "four pointer-jumping rounds");
    const int l = g.lane < D::NVA ? g.lane : D::NVA - 1;   // surplus lanes shadow the last link (they take part in the shuffles)
    const T th = e.qpos[l];
    T q[4] = {m.lquat[l][0], m.lquat[l][1], m.lquat[l][2], m.lquat[l][3]}, p[3] = {m.lpos[l][0], m.lpos[l][1], m.lpos[l][2]};
    if (m.jtype[l] == JT_HINGE) {   // rotate about local z: q <- q * (c, 0, 0, s)
      T s, c;
      N::sincos(th * T(0.5), &s, &c);
      const T t0 = q[0] * c - q[3] * s, t1 = q[1] * c + q[2] * s, t2 = q[2] * c - q[1] * s, t3 = q[3] * c + q[0] * s;
      q[0] = t0; q[1] = t1; q[2] = t2; q[3] = t3;
    } else {                        // slide along local z
      const T z[3] = {0, 0, th};
      T d[3];
      qrot(d, q, z);
      p[0] += d[0]; p[1] += d[1]; p[2] += d[2];
    }
    int anc = m.parent[l];
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const int src = anc >= 0 ? anc : l;
      T qa[4], pa[3];
      for (int i = 0; i < 4; i++) qa[i] = g.shfl(q[i], src);
      for (int i = 0; i < 3; i++) pa[i] = g.shfl(p[i], src);
      const int anc2 = g.shfl(anc, src);
      if (anc >= 0) {
        T t[3], qq[4];
        qrot(t, qa, p);
        p[0] = pa[0] + t[0]; p[1] = pa[1] + t[1]; p[2] = pa[2] + t[2];
        qmul(qq, qa, q);
        q[0] = qq[0]; q[1] = qq[1]; q[2] = qq[2]; q[3] = qq[3];
        anc = anc2;
      }
    }
    qnormalize(q);
    T mat[9];
    q2mat(mat, q);
    if (g.lane < D::NVA) {
      for (int i = 0; i < 4; i++) e.xquat[l][i] = q[i];
      for (int i = 0; i < 9; i++) e.xmat[l][i] = mat[i];
      for (int i = 0; i < 3; i++) e.xpos[l][i] = p[i];
    }
    g.sync();
    return;
  }
#endif
  for (int lv = 0; lv < m.nlevel; lv++) {
    for (int k = m.level_adr[lv] + g.lane; k < m.level_adr[lv + 1]; k += G) {
      const int l = m.level_link[k], p = m.parent[l];
      T q[4], pos[3], mat[9];
      if (p < 0) {
        for (int i = 0; i < 4; i++) q[i] = m.lquat[l][i];
        for (int i = 0; i < 3; i++) pos[i] = m.lpos[l][i];
      } else {
        qmul(q, e.xquat[p], m.lquat[l]);
        mulv3(pos, e.xmat[p], m.lpos[l]);
        for (int i = 0; i < 3; i++) pos[i] += e.xpos[p][i];
      }
      const T th = e.qpos[l];
      if (m.jtype[l] == JT_HINGE) {   // rotate about local z: q <- q * (c, 0, 0, s)
        T s, c;
        N::sincos(th * T(0.5), &s, &c);
        T t0 = q[0] * c - q[3] * s, t1 = q[1] * c + q[2] * s, t2 = q[2] * c - q[1] * s, t3 = q[3] * c + q[0] * s;
        q[0] = t0; q[1] = t1; q[2] = t2; q[3] = t3;
      }
      qnormalize(q);
      q2mat(mat, q);
      if (m.jtype[l] == JT_SLIDE) { pos[0] += mat[2] * th; pos[1] += mat[5] * th; pos[2] += mat[8] * th; }
      for (int i = 0; i < 4; i++) e.xquat[l][i] = q[i];
      for (int i = 0; i < 9; i++) e.xmat[l][i] = mat[i];
      for (int i = 0; i < 3; i++) e.xpos[l][i] = pos[i];
    }
    g.sync();
  }
}

#if KM_WARP_CODE
// One link per lane (G >= NVA).  v[0..N) <- sum of v over the links of the own link's SUBTREE: links are numbered depth
// first, so a subtree is the index range [l, sub_end[l]) inside the link's kinematic chain block; a suffix scan over the
// block (four rounds) minus the suffix that starts behind the subtree (only the later siblings' subtrees: comparable
// magnitudes, no cancellation against the rest of the arm).
template <class S, typename T, int G, int N> KM_HD void subtree_sum(T* v, int l, const Model<S, T>& m, const Grp<G>& g) {
  const int be = m.blk0[l] + m.blkn[l], se = m.sub_end[l];
#pragma unroll
  for (int d = 1; d < max_block<S>(); d *= 2) {
    const bool ok = l + d < be && g.lane < S::NVA;
    const int src = ok ? g.lane + d : g.lane;
#pragma unroll
    for (int k = 0; k < N; k++) { const T t = g.shfl(v[k], src); v[k] = ok ? v[k] + t : v[k]; }
  }
  const bool cut = se < be && g.lane < S::NVA;
  const int src = cut ? se : g.lane;
#pragma unroll
  for (int k = 0; k < N; k++) { const T t = g.shfl(v[k], src); v[k] = cut ? v[k] - t : v[k]; }
}
// v[0..N) <- sum of v over the own link and all its ancestors (pointer jumping up the tree, four rounds)
template <class S, typename T, int G, int N> KM_HD void path_sum(T* v, int l, const Model<S, T>& m, const Grp<G>& g) {
  int anc = m.parent[l];
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const int src = anc >= 0 ? anc : g.lane;
    const int anc2 = g.shfl(anc, src);
#pragma unroll
    for (int k = 0; k < N; k++) { const T t = g.shfl(v[k], src); v[k] = anc >= 0 ? v[k] + t : v[k]; }
    anc = anc >= 0 ? anc2 : anc;
  }
}
#endif

// mj_comPos: robot-tree centre of mass, cinert and cdof about it; mj_crb: mass matrix.
KM_TPL KM_FN void com_crb(KM_ARGS) {
  typedef Dim<S> D;
#if KM_WARP_CODE && KM_PAR_CRB
  if constexpr (G >= D::NVA && G > 1) {
    // one link per lane: centre of mass by a group sum, composite inertias by subtree sums in registers, then the mass
    // matrix entries (link, ancestor) dealt over all lanes
    const bool on = g.lane < D::NVA;
    const int l = on ? g.lane : D::NVA - 1;
    T xi[3];
    mulv3(xi, e.xmat[l], m.ipos[l]);
    for (int k = 0; k < 3; k++) xi[k] += e.xpos[l][k];
    T com[3];
    for (int k = 0; k < 3; k++) com[k] = g.sum(on ? m.mass[l] * xi[k] : T(0)) * m.total_mass_inv;
    if (g.lane < 3) e.com[g.lane] = com[g.lane];
    T off[3] = {xi[0] - com[0], xi[1] - com[1], xi[2] - com[2]}, ci[10], cd[6];
    inert_com(ci, m.inertia[l], e.xmat[l], off, m.mass[l]);
    const T ax[3] = {e.xmat[l][2], e.xmat[l][5], e.xmat[l][8]};
    if (m.jtype[l] == JT_SLIDE) { cd[0] = 0; cd[1] = 0; cd[2] = 0; cd[3] = ax[0]; cd[4] = ax[1]; cd[5] = ax[2]; }
    else {
      const T o[3] = {com[0] - e.xpos[l][0], com[1] - e.xpos[l][1], com[2] - e.xpos[l][2]};
      T c[3];
      cross3(c, ax, o);
      cd[0] = ax[0]; cd[1] = ax[1]; cd[2] = ax[2]; cd[3] = c[0]; cd[4] = c[1]; cd[5] = c[2];
    }
    if (on) {
      for (int i = 0; i < 10; i++) e.a.cinert[l][i] = ci[i];
      for (int i = 0; i < 6; i++) e.a.cdof[l][i] = cd[i];
    }
    g.sync();
    // composite inertia of every link's subtree: (link, component) items dealt over all lanes, summed link by link in
    // index order (no scan here: a suffix-minus-suffix form cancels in float32 and the mass matrix of the light distal
    // links is already a difference of large parallel-axis terms)
    T* crb = &e.a.cvel[0][0];                        // cvel / cdof_dot are free until the velocity stage: 12 NVA words
    KM_FOR(w, D::NVA * 10) {
      const int i = w / 10, k = w - i * 10;
      T sacc = e.a.cinert[i][k];
      for (int c = i + 1; c < m.sub_end[i]; c++) sacc += e.a.cinert[c][k];
      crb[w] = sacc;
    }
    g.sync();
    for (int k = 0; k < 10; k++) ci[k] = crb[l * 10 + k];
    T buf[6];
    mul_inert_vec(buf, ci, cd);
    if (on) for (int i = 0; i < 6; i++) e.a.cfrc[l][i] = buf[i];   // (cfrc is free until the velocity stage)
    g.sync();
    KM_FOR(w, D::NVA * (D::NVA + 1) / 2) {
      const int ij = m.pair_ij[w], i = ij >> 8, j = ij & 255;    // the first NVA (NVA + 1) / 2 pairs are the links' lower triangle
      if ((m.ancmask[i] >> j) & 1u) {
        T s = 0;
        for (int k = 0; k < 6; k++) s += e.a.cdof[j][k] * e.a.cfrc[i][k];
        e.M[i][j] = s;
        e.M[j][i] = s;
      }
    }
    g.sync();
    return;
  }
#endif
  KM_FOR(k, 3) {
    T s = 0;
    for (int l = 0; l < D::NVA; l++)
      s += m.mass[l] * (e.xpos[l][k] + e.xmat[l][3 * k] * m.ipos[l][0] + e.xmat[l][3 * k + 1] * m.ipos[l][1] + e.xmat[l][3 * k + 2] * m.ipos[l][2]);
    e.com[k] = s * m.total_mass_inv;
  }
  g.sync();
  KM_FOR(l, D::NVA) {
    T off[3], ci[10];
    mulv3(off, e.xmat[l], m.ipos[l]);
    for (int k = 0; k < 3; k++) off[k] += e.xpos[l][k] - e.com[k];
    inert_com(ci, m.inertia[l], e.xmat[l], off, m.mass[l]);
    for (int i = 0; i < 10; i++) e.a.cinert[l][i] = ci[i];
    const T ax[3] = {e.xmat[l][2], e.xmat[l][5], e.xmat[l][8]};
    if (m.jtype[l] == JT_SLIDE) {
      e.a.cdof[l][0] = 0; e.a.cdof[l][1] = 0; e.a.cdof[l][2] = 0;
      e.a.cdof[l][3] = ax[0]; e.a.cdof[l][4] = ax[1]; e.a.cdof[l][5] = ax[2];
    } else {
      T o[3] = {e.com[0] - e.xpos[l][0], e.com[1] - e.xpos[l][1], e.com[2] - e.xpos[l][2]}, c[3];
      cross3(c, ax, o);
      e.a.cdof[l][0] = ax[0]; e.a.cdof[l][1] = ax[1]; e.a.cdof[l][2] = ax[2];
      e.a.cdof[l][3] = c[0]; e.a.cdof[l][4] = c[1]; e.a.cdof[l][5] = c[2];
    }
  }
  g.sync();
  // composite inertia of link i's subtree times its own motion axis, projected on the ancestors' axes
  KM_FOR(i, D::NVA) {
    T crb[10];
    for (int k = 0; k < 10; k++) crb[k] = e.a.cinert[i][k];
    for (int c = i + 1; c < m.sub_end[i]; c++)
      for (int k = 0; k < 10; k++) crb[k] += e.a.cinert[c][k];
    T buf[6];
    mul_inert_vec(buf, crb, e.a.cdof[i]);
    for (int j = i; j >= 0; j = m.parent[j]) {
      T s = 0;
      for (int k = 0; k < 6; k++) s += e.a.cdof[j][k] * buf[k];
      e.M[i][j] = s;
      if constexpr (!E::TPE) e.M[j][i] = s;   // the thread-per-env solver reads the lower triangle only (km_solver_tpe.cuh)
    }
  }
  g.sync();
}

// Dense Cholesky A = L L^T of the leading n x n block, one lane per row, in place (strict lower triangle
// holds L, diag[] holds the pivots).  A has row stride `ld`.
template <class S, typename T, int G> KM_FN void chol_factor(T* A, T* diag, int n, int ld, const Grp<G>& g) {
  for (int j = 0; j < n; j++) {
    for (int i = j + g.lane; i < n; i += G) {
      T s = A[i * ld + j];
      for (int k = 0; k < j; k++) s -= A[i * ld + k] * A[j * ld + k];
      A[i * ld + j] = s;
    }
    g.sync();
    const T d = Num<T>::sqrt(tmax(A[j * ld + j], Num<T>::minval())), inv = T(1) / d;   // A[j][j] keeps the unscaled pivot
    for (int i = j + g.lane; i < n; i += G) {
      if (i == j) diag[j] = d;
      else A[i * ld + j] *= inv;
    }
    g.sync();
  }
}
// x <- A^{-1} x for a factor produced by chol_factor
template <class S, typename T, int G> KM_FN void chol_solve(const T* A, const T* diag, T* x, int n, int ld, const Grp<G>& g) {
  for (int j = 0; j < n; j++) {
    const T y = x[j] / diag[j];
    g.sync();
    for (int i = j + g.lane; i < n; i += G) {
      if (i == j) x[j] = y;
      else x[i] -= A[i * ld + j] * y;
    }
    g.sync();
  }
  for (int j = n - 1; j >= 0; j--) {
    const T y = x[j] / diag[j];
    g.sync();
    for (int i = g.lane; i <= j; i += G) {
      if (i == j) x[j] = y;
      else x[i] -= A[j * ld + i] * y;
    }
    g.sync();
  }
}

// contact Jacobian base row b of contact c at dof col (col must lie in the contact's support)
template <class S, typename T, class E> KM_HD T jc(const E& e, int c, int b, int col) {
  typedef Dim<S> D;
  return col >= D::NVA ? e.Jq[c][b][col - D::NVA] : e.Ja[e.con_slot[c] < D::NPAD ? e.con_slot[c] : 0][b][col];
}

// dst = A^{-1} rhs for the SPD matrix whose lower triangle has been assembled in e.c.H.
// Device: lane i keeps row i of the Cholesky factor in registers; pivots and multipliers travel by warp
// shuffles, so the factorisation and both triangular solves need no shared-memory round trips and no barriers
// except one transpose through e.c.H.  DENSE = false skips the blocks that are structurally zero at compile time
// (other kinematic chains; the cube while no finger pad touches it).  The body is branch-free on purpose: ptxas
// wraps every shuffle that follows a potentially divergent branch in a WARPSYNC.COLLECTIVE sequence.
// Host build (tests only): plain dense Cholesky.
template <class S, typename T, int G, bool DENSE, class E> KM_FN void solve_spd(KM_ARGS, const T* rhs, T* dst) {
  typedef Dim<S> D;
  typedef Num<T> N;
  constexpr int NV = D::NV;
#if KM_WARP_CODE
  if constexpr (G > 1) {
  static_assert(G >= NV, "one lane per dof row");
  g.sync();
  const int i = g.lane, ic = i < NV ? i : NV - 1;   // surplus lanes shadow the last row
  T row[NV];
  sfor<0, NV>([&](auto J) { row[decltype(J)::value] = e.c.H[ic][decltype(J)::value]; });
  T dinv = 1;
  sfor<0, NV>([&](auto J) {
    constexpr int j = decltype(J)::value;
    const T ajj = g.shfl(row[j], j);
    const T inv = N::rsqrt(tmax(ajj, N::minval()));
    const T lij = row[j] * inv;
    row[j] = lij;
    dinv = ic == j ? inv : dinv;
    constexpr int KE = DENSE ? NV : blk_end<S>(j);
    sfor<j + 1, KE>([&](auto K) {
      constexpr int k = decltype(K)::value;
      row[k] -= lij * g.shfl(lij, k);
    });
  });
  // forward substitution, row-oriented
  T acc = rhs[ic], y = 0;
  sfor<0, NV>([&](auto J) {
    constexpr int j = decltype(J)::value;
    const T yj = g.shfl(acc * dinv, j);
    y = ic == j ? yj : y;
    acc -= (ic > j ? row[j] : T(0)) * yj;
  });
  // transpose the factor through shared memory: lane i then holds column i (entries below the diagonal)
  sfor<0, NV>([&](auto J) { e.c.H[ic][decltype(J)::value] = row[decltype(J)::value]; });
  g.sync();
  T col[NV];
  sfor<0, NV>([&](auto K) {
    constexpr int k = decltype(K)::value;
    const T v = e.c.H[k][ic];
    col[k] = k > ic ? v : T(0);
  });
  T acc2 = y, x = 0;
  sfor_rev<NV>([&](auto J) {
    constexpr int j = decltype(J)::value;
    const T xj = g.shfl(acc2 * dinv, j);
    x = ic == j ? xj : x;
    acc2 -= col[j] * xj;
  });
  dst[ic] = x;
  g.sync();
  return;
  }
#endif
  if constexpr (G == 1) {
    if (dst != rhs)
      for (int i = 0; i < NV; i++) dst[i] = rhs[i];
    chol_factor<S, T, G>(&e.c.H[0][0], e.c.Hd, NV, D::HS, g);
    chol_solve<S, T, G>(&e.c.H[0][0], e.c.Hd, dst, NV, D::HS, g);
  }
}

// one lower-triangle entry (i >= j) of J^T diag(D_active) J restricted to the contacts
KM_TPL KM_HD T hess_contacts(const E& e, int i, int j) {
  T h = 0;
  for (int c = 0; c < e.ncon; c++) {
    const unsigned sup = e.con_sup[c];
    if (((sup >> i) & 1u) && ((sup >> j) & 1u)) {
      const T ni = jc<S, T>(e, c, 0, i), nj = jc<S, T>(e, c, 0, j);
      T acc = e.cb[c][0] * ni * nj;
      for (int k = 1; k < 4; k++) {
        const T ki = jc<S, T>(e, c, k, i), kj = jc<S, T>(e, c, k, j);
        acc += e.cb[c][k] * (ni * kj + ki * nj) + e.con_W[c][k - 1] * ki * kj;
      }
      h += acc;
    }
  }
  return h;
}

// lower triangle of e.c.H = M + diag(hdiag) [+ contact terms when with_contacts]
KM_TPL KM_HD void assemble_h(KM_ARGS, bool with_contacts) {
  typedef Dim<S> D;
  KM_FOR(w, D::NV * (D::NV + 1) / 2) {
    const int ij = m.pair_ij[w], i = ij >> 8, j = ij & 255;
    T h;
    if (i < D::NVA) h = e.M[i][j];
    else h = i == j ? (i - D::NVA < 3 ? m.cube_mass : m.cube_inertia[i - D::NVA - 3]) : T(0);
    if (i == j) h += e.c.hdiag[i];
    // contact terms: the cube block always (every contact has cube columns: direct loads, no support test);
    // arm rows only while a finger pad touches the cube
    if (with_contacts && j >= D::NVA) {
      const int ki = i - D::NVA, kj = j - D::NVA;
      for (int c = 0; c < e.ncon; c++) {
        const T ni = e.Jq[c][0][ki], nj = e.Jq[c][0][kj];
        T acc = e.cb[c][0] * ni * nj;
        for (int k = 1; k < 4; k++) {
          const T ti = e.Jq[c][k][ki], tj = e.Jq[c][k][kj];
          acc += e.cb[c][k] * (ni * tj + ti * nj) + e.con_W[c][k - 1] * ti * tj;
        }
        h += acc;
      }
    } else if (with_contacts && e.coupled) h += hess_contacts<S, T, G>(e, i, j);
    e.c.H[i][j] = h;
  }
  g.sync();
}

// mj_collision on the primitive pairs of the completed model: finger-pad spheres vs the cube box, then the
// table plane vs the cube's corners (at most four, in corner order).  Contacts are compacted in pair order.
KM_TPL KM_FN void collision(KM_ARGS) {
  typedef Dim<S> D;
  typedef Num<T> N;
  const T* cpos = e.qpos + D::NVA;
  KM_FOR(s, D::NSLOT) {
    int on = 0;
    T dist = 0, pos[3] = {0, 0, 0}, nrm[3] = {0, 0, 1};
    if (s < D::NPAD) {   // mjc_SphereBox, sphere = geom1
      const int l = m.pad_link[s];
      T c1[3], tmp[3], center[3], cl[3], n[3], pb[3];
      mulv3(c1, e.xmat[l], m.pad_pos[s]);
      for (int i = 0; i < 3; i++) tmp[i] = (c1[i] + e.xpos[l][i] - cpos[i]) - e.cube_lo[i];
      mulTv3(center, e.cmat, tmp);
      for (int i = 0; i < 3; i++) { cl[i] = tclip(center[i], -m.cube_size[i], m.cube_size[i]); n[i] = cl[i] - center[i]; }
      const T d = N::sqrt(dot3(n, n)), radius = m.pad_rad[s];
      if (!(d - radius > T(0))) {
        on = 1;
        if (d <= N::minval()) {   // centre inside the box: push out through the nearest face
          T closest = T(2) * (m.cube_size[0] + m.cube_size[1] + m.cube_size[2]);
          int k = 0;
          for (int i = 0; i < 6; i++) {
            T face = (i % 2 ? T(1) : T(-1)) * m.cube_size[i / 2];
            T t = N::abs(face - center[i / 2]);
            if (t < closest) { closest = t; k = i; }
          }
          T fo[3] = {0, 0, 0};
          fo[k / 2] = k % 2 ? T(1) : T(-1);
          for (int i = 0; i < 3; i++) { n[i] = -fo[i]; pb[i] = center[i] + fo[i] * (closest - radius) * T(0.5); }
          dist = -closest - radius;
        } else {
          for (int i = 0; i < 3; i++) n[i] /= d;
          dist = d - radius;
          for (int i = 0; i < 3; i++) pb[i] = center[i] + n[i] * (radius + T(0.5) * dist);
        }
        mulv3(pos, e.cmat, pb);
        for (int i = 0; i < 3; i++) pos[i] += cpos[i];
        mulv3(nrm, e.cmat, n);
      }
    } else {             // mjc_PlaneBox: corner i of the cube against z = tab_z, normal (0,0,1)
      const int i = s - D::NPAD;
      const T pd = (cpos[2] - m.tab_z) + e.cube_lo[2];   // the difference of the high parts is exact
      T vec[3] = {(i & 1 ? T(1) : T(-1)) * m.cube_size[0], (i & 2 ? T(1) : T(-1)) * m.cube_size[1],
                  (i & 4 ? T(1) : T(-1)) * m.cube_size[2]}, corner[3];
      mulv3(corner, e.cmat, vec);
      const T ld = corner[2];
      if (!(pd + ld > T(0) || ld > T(0))) {
        on = 1;
        dist = pd + ld;
        pos[0] = corner[0] + cpos[0]; pos[1] = corner[1] + cpos[1]; pos[2] = corner[2] + cpos[2] - dist * T(0.5);
      }
    }
    e.sl_on[s] = on;
    if (on) {
      e.sl_dist[s] = dist;
      T f[9] = {nrm[0], nrm[1], nrm[2], 0, 0, 0, 0, 0, 0};
      makeframe(f);
      for (int i = 0; i < 3; i++) e.sl_pos[s][i] = pos[i];
      for (int i = 0; i < 9; i++) e.sl_frame[s][i] = f[i];
    }
  }
  g.sync();
  if (g.lane == 0) {
    int n = 0, corners = 0;
    for (int s = 0; s < D::NSLOT; s++) {
      if (!e.sl_on[s]) continue;
      if (s >= D::NPAD && ++corners > 4) break;
      e.con_slot[n++] = s;
    }
    e.ncon = n;
    e.coupled = n > 0 && e.con_slot[0] < D::NPAD;
  }
  g.sync();
}

// solimp -> impedance at penetration `pos` (margin 0)  (SURVEY.md A5).  solimp carries two extra entries,
// 1 - dmin and 1 - dmax rounded from double, so that 1 - imp keeps full relative precision in fp32 when the
// impedance is close to one (the cube's solimp gives imp up to 0.9999).  *omi receives 1 - imp.
template <typename T> KM_HD T impedance(const T* solimp, T pos, T* omi) {
  typedef Num<T> N;
  const T dmin = solimp[0], dmax = solimp[1], width = solimp[2], mid = solimp[3], power = solimp[4];
  const T omin = solimp[5], omax = solimp[6];
  if (dmin == dmax || width <= N::minval()) { *omi = T(0.5) * (omin + omax); return T(0.5) * (dmin + dmax); }
  T x = pos / width;
  if (x < T(0)) x = -x;
  if (x >= T(1)) { *omi = omax; return dmax; }
  if (x == T(0)) { *omi = omin; return dmin; }
  T y;
  if (power == T(1)) y = x;
  else y = x <= mid ? x * x / mid : T(1) - (T(1) - x) * (T(1) - x) / (T(1) - mid);   // power == 2 (km_fill.h checks)
  *omi = omin - y * (dmax - dmin);
  return dmin + y * (dmax - dmin);
}

// mj_makeConstraint: friction-loss rows (constant, set up once), active joint limits, pyramidal contact rows,
// and the contact Jacobian base rows.
KM_TPL KM_FN void make_constraint(KM_ARGS) {
  typedef Dim<S> D;
  typedef Num<T> N;
  const T* cpos = e.qpos + D::NVA;
#if KM_WARP_CODE
  constexpr bool kLaneLimits = G >= D::NVA && G > 1;
#else
  constexpr bool kLaneLimits = false;
#endif
  if constexpr (kLaneLimits) {
    // one joint per lane (a joint violates at most one side of its range, km_fill.h checks lo < hi); the active rows are
    // numbered in joint order by a ballot, as the serial loop below numbers them
    const int j = g.lane < D::NVA ? g.lane : D::NVA - 1;
    const T q = e.qpos[j];
    const T d0 = q - m.range[j][0], d1 = m.range[j][1] - q;
    const bool on = g.lane < D::NVA && (d0 < T(0) || d1 < T(0));
    const int side = d0 < T(0) ? 0 : 1;
    const T dist = side == 0 ? d0 : d1;
    const unsigned act = g.ballot(on);
    const int r = D::NFRIC + popc(act & ((1u << g.lane) - 1u));
    if (g.lane < D::NVA) e.dof_lim[j] = on ? r : -1;
    if (on) {
      T omi;
      const T imp = impedance(m.lim_solimp[j], dist, &omi);
      const T R = tmax(N::minval(), omi * m.lim_invw[j] / imp);
      const T tc = tmax(m.lim_solref[j][0], T(2) * m.h), dr = m.lim_solref[j][1], dmax = m.lim_solimp[j][1];
      e.efc_desc[r] = efc_pack(EFC_LIMIT, j, 0, side);
      e.efc_D[r] = T(1) / R;
      e.lim_B[r - D::NFRIC] = T(2) / (dmax * tc);
      e.lim_Kip[r - D::NFRIC] = (T(1) / (dmax * dmax * tc * tc * dr * dr)) * imp * dist;
    }
    if (g.lane == 0) { e.nlim = popc(act); e.nefc = D::NFRIC + popc(act) + 6 * e.ncon; }
  } else if (g.lane == 0) {
    int r = D::NFRIC;
    for (int j = 0; j < D::NVA; j++) {
      const T q = e.qpos[j];
      e.dof_lim[j] = -1;
      for (int side = 0; side < 2; side++) {
        const T dist = side == 0 ? q - m.range[j][0] : m.range[j][1] - q;
        if (dist < T(0)) {
          T omi;
          const T imp = impedance(m.lim_solimp[j], dist, &omi);
          const T R = tmax(N::minval(), omi * m.lim_invw[j] / imp);
          const T tc = tmax(m.lim_solref[j][0], T(2) * m.h), dr = m.lim_solref[j][1], dmax = m.lim_solimp[j][1];
          e.efc_desc[r] = efc_pack(EFC_LIMIT, j, 0, side);
          e.dof_lim[j] = r;
          e.efc_D[r] = T(1) / R;
          e.lim_B[r - D::NFRIC] = T(2) / (dmax * tc);
          e.lim_Kip[r - D::NFRIC] = (T(1) / (dmax * dmax * tc * tc * dr * dr)) * imp * dist;
          r++;
        }
      }
    }
    e.nlim = r - D::NFRIC;
    e.nefc = r + 6 * e.ncon;
  }
  g.sync();
  const int base = D::NFRIC + e.nlim;
  KM_FOR(c, e.ncon) {
    const int s = e.con_slot[c];
    const T *solref, *solimp, *mu3;
    T tran;
    if (s < D::NPAD) { solref = m.pad_solref[s]; solimp = m.pad_solimp[s]; mu3 = m.pad_mu[s]; tran = m.pad_tran[s]; }
    else { solref = m.tab_solref; solimp = m.tab_solimp; mu3 = m.tab_mu; tran = m.tab_tran; }
    const T dist = e.sl_dist[s];
    T omi;
    const T imp = impedance(solimp, dist, &omi);
    const T diag1 = tran + mu3[0] * mu3[0] * tran;
    const T R1 = tmax(N::minval(), omi * diag1 / imp);
    const T mu = mu3[0] / N::sqrt(m.impratio);
    const T R = T(2) * mu * mu * R1, Dc = T(1) / R;
    const T tc = tmax(solref[0], T(2) * m.h), dr = solref[1], dmax = solimp[1];
    const T B = T(2) / (dmax * tc), Kip = (T(1) / (dmax * dmax * tc * tc * dr * dr)) * imp * dist;
    e.con_D[c] = Dc; e.con_B[c] = B; e.con_Kip[c] = Kip;
    for (int k = 0; k < 3; k++) e.con_mu[c][k] = mu3[k];
    e.con_sup[c] = (s < D::NPAD ? m.ancmask[m.pad_link[s]] : 0u) | (63u << D::NVA);
    for (int k = 0; k < 6; k++) {
      const int r = base + 6 * c + k;
      e.efc_desc[r] = efc_pack(EFC_CONTACT, c, 1 + k / 2, k & 1);
      e.efc_D[r] = Dc;
    }
  }
  // contact Jacobian base rows, J = J(cube) - J(pad link), one (contact, dof) per work item
  KM_FOR(w, e.ncon * D::NV) {
    const int c = w / D::NV, col = w - c * D::NV, s = e.con_slot[c];
    const T* f = e.sl_frame[s];
    const T* p = e.sl_pos[s];
    T jp[3] = {0, 0, 0}, jr[3] = {0, 0, 0};
    bool on = true;
    if (col >= D::NVA) {
      const int k = col - D::NVA;
      if (k < 3) jp[k] = 1;
      else {
        const T a[3] = {e.cmat[k - 3], e.cmat[3 + k - 3], e.cmat[6 + k - 3]};
        const T o[3] = {p[0] - cpos[0], p[1] - cpos[1], p[2] - cpos[2]};
        cross3(jp, a, o);
        jr[0] = a[0]; jr[1] = a[1]; jr[2] = a[2];
      }
    } else if (s < D::NPAD && ((m.ancmask[m.pad_link[s]] >> col) & 1u)) {
      const T a[3] = {e.xmat[col][2], e.xmat[col][5], e.xmat[col][8]};
      if (m.jtype[col] == JT_SLIDE) { jp[0] = -a[0]; jp[1] = -a[1]; jp[2] = -a[2]; }
      else {
        const T o[3] = {p[0] - e.xpos[col][0], p[1] - e.xpos[col][1], p[2] - e.xpos[col][2]};
        T t[3];
        cross3(t, a, o);
        jp[0] = -t[0]; jp[1] = -t[1]; jp[2] = -t[2];
        jr[0] = -a[0]; jr[1] = -a[1]; jr[2] = -a[2];
      }
    } else on = false;
    if (on) {
      T* dst = col >= D::NVA ? &e.Jq[c][0][col - D::NVA] : &e.Ja[s][0][col];
      const int st = col >= D::NVA ? 6 : D::NVA;   // stride between base rows
      dst[0] = dot3(f, jp);
      dst[st] = dot3(f + 3, jp);
      dst[2 * st] = dot3(f + 6, jp);
      dst[3 * st] = dot3(f, jr);
    }
  }
  g.sync();
}

// out[r] = J_r . x for every constraint row (x is a dof-space vector in shared memory)
KM_TPL KM_FN void mul_J(KM_ARGS, const T* x, T* out) {
  typedef Dim<S> D;
  KM_FOR(w, e.ncon * 4) {
    const int c = w >> 2, b = w & 3;
    T s = 0;
    for (int k = 0; k < 6; k++) s += e.Jq[c][b][k] * x[D::NVA + k];
    const int sl = e.con_slot[c];
    if (sl < D::NPAD)
      for (int j = m.pad_link[sl]; j >= 0; j = m.parent[j]) s += e.Ja[sl][b][j] * x[j];
    e.cb[c][b] = s;
  }
  g.sync();
  KM_FOR(r, e.nefc) {
    const int d = e.efc_desc[r], id = efc_id(d);
    T v;
    if (efc_type(d) == EFC_CONTACT) {
      const T t = e.con_mu[id][efc_k(d) - 1] * e.cb[id][efc_k(d)];
      v = e.cb[id][0] + (efc_neg(d) ? -t : t);
    } else v = efc_neg(d) ? -x[id] : x[id];
    out[r] = v;
  }
  g.sync();
}

// e.c.qfc = J^T efc_force
KM_TPL KM_FN void mul_JT_force(KM_ARGS) {
  typedef Dim<S> D;
  const int base = D::NFRIC + e.nlim;
  KM_FOR(c, e.ncon) {
    const T* f = e.c.efc_force + base + 6 * c;
    e.cb[c][0] = f[0] + f[1] + f[2] + f[3] + f[4] + f[5];
    e.cb[c][1] = e.con_mu[c][0] * (f[0] - f[1]);
    e.cb[c][2] = e.con_mu[c][1] * (f[2] - f[3]);
    e.cb[c][3] = e.con_mu[c][2] * (f[4] - f[5]);
  }
  g.sync();
  KM_FOR(i, D::NV) {
    T s = 0;
    const int fr = m.dof_fric[i], lr = i < D::NVA ? e.dof_lim[i] : -1;
    if (fr >= 0) s += e.c.efc_force[fr];
    if (lr >= 0) s += efc_neg(e.efc_desc[lr]) ? -e.c.efc_force[lr] : e.c.efc_force[lr];
    if (i >= D::NVA) {   // cube columns: every contact, direct loads
      const int k = i - D::NVA;
      for (int c = 0; c < e.ncon; c++)
        s += e.Jq[c][0][k] * e.cb[c][0] + e.Jq[c][1][k] * e.cb[c][1] + e.Jq[c][2][k] * e.cb[c][2] + e.Jq[c][3][k] * e.cb[c][3];
    } else if (e.coupled) {
      for (int c = 0; c < e.ncon; c++)
        if ((e.con_sup[c] >> i) & 1u)
          s += jc<S, T>(e, c, 0, i) * e.cb[c][0] + jc<S, T>(e, c, 1, i) * e.cb[c][1] + jc<S, T>(e, c, 2, i) * e.cb[c][2] +
               jc<S, T>(e, c, 3, i) * e.cb[c][3];
    }
    e.c.qfc[i] = s;
  }
  g.sync();
}

// out = M x  (articulated block: ancestors and subtree of each dof; cube block: constant diagonal)
KM_TPL KM_FN void mul_M(KM_ARGS, const T* x, T* out) {
  typedef Dim<S> D;
  KM_FOR(i, D::NV) {
    T s;
    if (i >= D::NVA) {
      const int k = i - D::NVA;
      s = (k < 3 ? m.cube_mass : m.cube_inertia[k - 3]) * x[i];
    } else {
      // dense row: entries between unrelated links are structural zeros (init_env), and independent loads beat
      // walking the parent chain (a dependent load per ancestor)
      s = 0;
#pragma unroll
      for (int j = 0; j < D::NVA; j++) s += e.M[i][j] * x[j];
    }
    out[i] = s;
  }
  g.sync();
}

KM_TPL KM_HD void fwd_position(KM_ARGS) {
  typedef Dim<S> D;
  g.cta_sync();
  KM_CLK(CLK_BARRIER);
  kinematics<S, T, G>(e, m, g);
  KM_CLK(CLK_KIN);
  g.cta_sync();
  KM_CLK(CLK_BARRIER);
  com_crb<S, T, G>(e, m, g);
  KM_CLK(CLK_CRB);
  g.cta_sync();
  KM_CLK(CLK_BARRIER);
  collision<S, T, G>(e, m, g);
  make_constraint<S, T, G>(e, m, g);
  KM_FOR(i, D::NU) e.actlen[i] = e.qpos[i];   // mj_transmission: joint transmissions, gear 1
  g.sync();
  KM_CLK(CLK_COLL);
}

// =========================================================================================== velocity stage
// mj_comVel + mj_rne(flg_acc = 0) + mj_referenceConstraint.
KM_TPL KM_FN void fwd_velocity(KM_ARGS) {
  typedef Dim<S> D;
#if KM_WARP_CODE && KM_PAR_VEL
  constexpr bool kLanePerLink = G >= D::NVA && G > 1;
#else
  constexpr bool kLanePerLink = false;
#endif
  if constexpr (kLanePerLink) {
#if KM_WARP_CODE
    // one link per lane, everything in registers: link velocities and the velocity-product accelerations by sums along the
    // path to the root, the bias forces by subtree sums of the links' spatial forces
    const bool on = g.lane < D::NVA;
    const int l = on ? g.lane : D::NVA - 1, p = m.parent[l];
    const T qv = e.qvel[l];
    T cd[6], ci[10], v[6], y[6], a[6], f[6], t1[6], t2[6];
    for (int k = 0; k < 6; k++) cd[k] = e.a.cdof[l][k];
    for (int k = 0; k < 10; k++) ci[k] = e.a.cinert[l][k];
    for (int k = 0; k < 6; k++) v[k] = cd[k] * qv;
    path_sum<S, T, G, 6>(v, l, m, g);                       // cvel of the link
    T vp[6], dd[6] = {0, 0, 0, 0, 0, 0};
    for (int k = 0; k < 6; k++) vp[k] = g.shfl(v[k], p >= 0 ? p : g.lane);
    if (p >= 0) cross_motion(dd, vp, cd);                   // cdof_dot
    for (int k = 0; k < 6; k++) y[k] = dd[k] * qv;
    path_sum<S, T, G, 6>(y, l, m, g);
    for (int k = 0; k < 3; k++) { a[k] = y[k]; a[3 + k] = y[3 + k] - m.grav[k]; }
    mul_inert_vec(f, ci, a);
    mul_inert_vec(t1, ci, v);
    cross_force(t2, v, t1);
    for (int k = 0; k < 6; k++) f[k] = on ? f[k] + t2[k] : T(0);
    subtree_sum<S, T, G, 6>(f, l, m, g);
    T s = 0;
    for (int k = 0; k < 6; k++) s += cd[k] * f[k];
    if (on) e.bias[l] = s;
    if (g.lane >= D::NVA && g.lane < D::NV) {
      // free cube with its COM at the body origin: bias = [-m g ; w x (I w)] (w in the body frame)
      const int k = g.lane - D::NVA;
      T sc;
      if (k < 3) sc = -m.cube_mass * m.grav[k];
      else {
        const T* w = e.qvel + D::NVA + 3;
        const T Iw[3] = {m.cube_inertia[0] * w[0], m.cube_inertia[1] * w[1], m.cube_inertia[2] * w[2]};
        T c[3];
        cross3(c, w, Iw);
        sc = c[k - 3];
      }
      e.bias[g.lane] = sc;
    }
#endif
  } else {
  KM_FOR(l, D::NVA) {
    T v[6] = {0, 0, 0, 0, 0, 0};
    for (int j = l; j >= 0; j = m.parent[j]) {
      const T qv = e.qvel[j];
      for (int k = 0; k < 6; k++) v[k] += e.a.cdof[j][k] * qv;
    }
    for (int k = 0; k < 6; k++) e.a.cvel[l][k] = v[k];
  }
  g.sync();
  KM_FOR(l, D::NVA) {
    const int p = m.parent[l];
    T r[6] = {0, 0, 0, 0, 0, 0};
    if (p >= 0) cross_motion(r, e.a.cvel[p], e.a.cdof[l]);
    for (int k = 0; k < 6; k++) e.a.cdof_dot[l][k] = r[k];
  }
  g.sync();
  KM_FOR(l, D::NVA) {
    T a[6] = {0, 0, 0, -m.grav[0], -m.grav[1], -m.grav[2]};
    for (int j = l; j >= 0; j = m.parent[j]) {
      const T qv = e.qvel[j];
      for (int k = 0; k < 6; k++) a[k] += e.a.cdof_dot[j][k] * qv;
    }
    T f[6], t1[6], t2[6];
    mul_inert_vec(f, e.a.cinert[l], a);
    mul_inert_vec(t1, e.a.cinert[l], e.a.cvel[l]);
    cross_force(t2, e.a.cvel[l], t1);
    for (int k = 0; k < 6; k++) e.a.cfrc[l][k] = f[k] + t2[k];
  }
  g.sync();
  KM_FOR(i, D::NV) {
    T s = 0;
    if (i < D::NVA) {
      T f[6] = {0, 0, 0, 0, 0, 0};
      for (int c = i; c < m.sub_end[i]; c++)
        for (int k = 0; k < 6; k++) f[k] += e.a.cfrc[c][k];
      for (int k = 0; k < 6; k++) s += e.a.cdof[i][k] * f[k];
    } else {
      // free cube with its COM at the body origin: bias = [-m g ; w x (I w)] (w in the body frame)
      const int k = i - D::NVA;
      if (k < 3) s = -m.cube_mass * m.grav[k];
      else {
        const T* w = e.qvel + D::NVA + 3;
        const T Iw[3] = {m.cube_inertia[0] * w[0], m.cube_inertia[1] * w[1], m.cube_inertia[2] * w[2]};
        T c[3];
        cross3(c, w, Iw);
        s = c[k - 3];
      }
    }
    e.bias[i] = s;
  }
  }
  // reference acceleration of every row: aref = -B (J qvel) - K imp pos
  mul_J<S, T, G>(e, m, g, e.qvel, e.efc_jv);
  const int base = D::NFRIC + e.nlim;
  KM_FOR(r, e.nefc) {
    T B, Kip;
    if (r < D::NFRIC) { B = m.fr_B[r]; Kip = 0; }
    else if (r < base) { B = e.lim_B[r - D::NFRIC]; Kip = e.lim_Kip[r - D::NFRIC]; }
    else { const int c = (r - base) / 6; B = e.con_B[c]; Kip = e.con_Kip[c]; }
    e.efc_aref[r] = -B * e.efc_jv[r] - Kip;
  }
  g.sync();
}

}  // namespace km
#include "km_solver_tpe.cuh"
namespace km {

// =========================================================================================== acceleration stage
// mj_fwdActuation (<position kp> servos): qfrc_smooth = actuator force - bias
KM_TPL KM_FN void fwd_actuation(KM_ARGS) {
  typedef Dim<S> D;
  KM_FOR(i, D::NV) {
    T f = 0;
    if (i < D::NVA) {
      const T c = tclip(e.ctrl[i], m.ctrl_lo[i], m.ctrl_hi[i]);
      f = tclip(m.kp[i] * c - m.kp[i] * e.actlen[i], m.frc_lo[i], m.frc_hi[i]);
    }
    e.qfrc_smooth[i] = f - e.bias[i];
  }
  g.sync();
}
// mj_fwdAcceleration: qacc_smooth = M^{-1} qfrc_smooth
KM_TPL KM_FN void fwd_acceleration(KM_ARGS) {
  typedef Dim<S> D;
  if constexpr (G == 1) tpe_solveM<S, T>(e, m, e.qfrc_smooth, e.qacc_smooth);
  else {
    KM_FOR(i, D::NV) e.c.hdiag[i] = 0;
    g.sync();
    assemble_h<S, T, G>(e, m, g, false);
    solve_spd<S, T, G, false>(e, m, g, e.qfrc_smooth, e.qacc_smooth);
  }
}

// ---- Newton solver (mj_solNewton on the primal problem)
// constraint cost pieces of one row at residual x: returns the cost, sets state and force
template <typename T> KM_HD T row_cost(bool friction, T x, T Dr, T rf, T floss, int* state, T* force) {
  if (friction) {
    if (x <= -rf) { *state = ST_LINEARNEG; *force = floss; return -floss * (T(0.5) * rf + x); }
    if (x >= rf) { *state = ST_LINEARPOS; *force = -floss; return -floss * (T(0.5) * rf - x); }
    *state = ST_QUADRATIC; *force = -Dr * x; return T(0.5) * Dr * x * x;
  }
  if (x < T(0)) { *state = ST_QUADRATIC; *force = -Dr * x; return T(0.5) * Dr * x * x; }
  *state = ST_SATISFIED; *force = 0; return 0;
}

// total cost (constraint + Gauss) at acceleration `a`; uses efc_jv and Mv as scratch
KM_TPL KM_FN T total_cost(KM_ARGS, const T* a) {
  typedef Dim<S> D;
  mul_J<S, T, G>(e, m, g, a, e.efc_jv);
  mul_M<S, T, G>(e, m, g, a, e.c.Mv);
  T c = 0;
  KM_FOR(r, e.nefc) {
    int st; T f;
    const bool fr = r < D::NFRIC;
    c += row_cost(fr, e.efc_jv[r] - e.efc_aref[r], e.efc_D[r], fr ? m.fr_Rf[r] : T(0), fr ? m.fr_loss[r] : T(0), &st, &f);
  }
  T gs = 0;
  KM_FOR(i, D::NV) gs += (e.c.Mv[i] - e.qfrc_smooth[i]) * (a[i] - e.qacc_smooth[i]);
  c = g.sum(c + T(0.5) * gs);
  g.sync();
  return c;
}

// states, forces, cost and gradient at the current jar / Ma / qacc; returns the total cost, *gauss the Gauss term
KM_TPL KM_FN T sol_update(KM_ARGS, T* gauss) {
  typedef Dim<S> D;
  T c = 0;
  KM_FOR(r, e.nefc) {
    int st; T f;
    const bool fr = r < D::NFRIC;
    c += row_cost(fr, e.c.efc_jar[r], e.efc_D[r], fr ? m.fr_Rf[r] : T(0), fr ? m.fr_loss[r] : T(0), &st, &f);
    e.c.efc_state[r] = st; e.c.efc_force[r] = f;
  }
  g.sync();
  mul_JT_force<S, T, G>(e, m, g);
  T gs = 0;
  KM_FOR(i, D::NV) {
    gs += (e.c.Ma[i] - e.qfrc_smooth[i]) * (e.qacc[i] - e.qacc_smooth[i]);
    e.c.grad[i] = e.c.Ma[i] - e.qfrc_smooth[i] - e.c.qfc[i];
  }
  c = g.sum(c);
  gs = T(0.5) * g.sum(gs);
  g.sync();
  *gauss = gs;
  return c + gs;
}

// H = M + J^T diag(D_active) J, Cholesky, Mgrad = H^{-1} grad
KM_TPL KM_FN void sol_hessian_dir(KM_ARGS) {
  typedef Dim<S> D;
  const int base = D::NFRIC + e.nlim;
  // per-contact weights of the base rows: W = sum_active D w w^T, w = e0 +- mu_k e_k  (arrow-head 4x4)
  KM_FOR(c, e.ncon) {
    const int* st = e.c.efc_state + base + 6 * c;
    const T Dc = e.con_D[c];
    T n = 0;
    for (int k = 0; k < 3; k++) {
      const T p = st[2 * k] == ST_QUADRATIC ? T(1) : T(0), q = st[2 * k + 1] == ST_QUADRATIC ? T(1) : T(0);
      const T mu = e.con_mu[c][k];
      n += p + q;
      e.cb[c][1 + k] = Dc * mu * (p - q);        // W[0][k]
      e.con_W[c][k] = Dc * mu * mu * (p + q);    // W[k][k]
    }
    e.cb[c][0] = Dc * n;                         // W[0][0]
  }
  // diagonal contributions of the friction-loss and limit rows (each touches one dof)
  KM_FOR(i, D::NV) {
    T d = 0;
    const int fr = m.dof_fric[i], lr = i < D::NVA ? e.dof_lim[i] : -1;
    if (fr >= 0 && e.c.efc_state[fr] == ST_QUADRATIC) d += e.efc_D[fr];
    if (lr >= 0 && e.c.efc_state[lr] == ST_QUADRATIC) d += e.efc_D[lr];
    e.c.hdiag[i] = d;
  }
  g.sync();
  assemble_h<S, T, G>(e, m, g, true);
  if (e.coupled) solve_spd<S, T, G, true>(e, m, g, e.c.grad, e.c.Mgrad);
  else solve_spd<S, T, G, false>(e, m, g, e.c.grad, e.c.Mgrad);
}

// derivatives of the 1-D cost along the search direction at step alpha
KM_TPL KM_FN void ls_eval(KM_ARGS, T qg1, T qg2, T alpha, T* d1, T* d2) {
  T q1 = 0, q2 = 0;
  KM_FOR(r, e.nefc) {
    const T jv = e.efc_jv[r], jar = e.c.efc_jar[r], x = jar + alpha * jv, Dr = e.efc_D[r];
    if (r < Dim<S>::NFRIC) {
      const T f = m.fr_loss[r], rf = m.fr_Rf[r];
      if (x <= -rf) { q1 += -f * jv; continue; }
      if (x >= rf) { q1 += f * jv; continue; }
    } else if (x >= T(0)) continue;
    q1 += Dr * jar * jv;
    q2 += T(0.5) * Dr * jv * jv;
  }
  q1 = g.sum(q1) + qg1;
  q2 = g.sum(q2) + qg2;
  *d1 = T(2) * alpha * q2 + q1;
  *d2 = T(2) * q2;
  if (g.lane == 0) e.ls_evals++;
}

KM_TPL KM_FN T sol_linesearch(KM_ARGS, T scale) {
  typedef Dim<S> D;
  typedef Num<T> N;
  T sn = 0;
  KM_FOR(i, D::NV) sn += e.c.search[i] * e.c.search[i];
  const T snorm = N::sqrt(g.sum(sn));
  if (snorm < N::minval()) return 0;
  mul_M<S, T, G>(e, m, g, e.c.search, e.c.Mv);
  mul_J<S, T, G>(e, m, g, e.c.search, e.efc_jv);
  T a1 = 0, a2 = 0;
  KM_FOR(i, D::NV) {
    a1 += e.c.search[i] * (e.c.Ma[i] - e.qfrc_smooth[i]);
    a2 += T(0.5) * e.c.search[i] * e.c.Mv[i];
  }
  const T qg1 = g.sum(a1), qg2 = g.sum(a2);
  T d1, d2;
  ls_eval<S, T, G>(e, m, g, qg1, qg2, T(0), &d1, &d2);
  const T gtol = tmax(m.tol * m.ls_tol * snorm / scale, T(64) * N::eps() * N::abs(d1));
  if (N::abs(d1) < gtol || d1 > T(0)) return 0;
  T lo = 0, lo_d1 = d1, lo_d2 = d2, hi = 0, hi_d1 = 0, hi_d2 = 0;
  int bracket = 0, it = 0;
  for (; it < m.ls_iterations; it++) {   // Newton steps to the right until the slope changes sign
    const T a = lo - lo_d1 / lo_d2;
    ls_eval<S, T, G>(e, m, g, qg1, qg2, a, &d1, &d2);
    if (N::abs(d1) < gtol) return a;
    if (d1 > T(0)) { hi = a; hi_d1 = d1; hi_d2 = d2; bracket = 1; break; }
    lo = a; lo_d1 = d1; lo_d2 = d2;
  }
  if (!bracket) return lo;
  for (; it < m.ls_iterations; it++) {   // safeguarded Newton inside the bracket
    T a = N::abs(lo_d1) < N::abs(hi_d1) ? lo - lo_d1 / lo_d2 : hi - hi_d1 / hi_d2;
    if (!(a > lo && a < hi)) a = T(0.5) * (lo + hi);
    if (a == lo || a == hi) break;
    ls_eval<S, T, G>(e, m, g, qg1, qg2, a, &d1, &d2);
    if (N::abs(d1) < gtol) return a;
    if (d1 > T(0)) { hi = a; hi_d1 = d1; hi_d2 = d2; } else { lo = a; lo_d1 = d1; lo_d2 = d2; }
  }
  return N::abs(lo_d1) < N::abs(hi_d1) ? lo : hi;
}

KM_TPL KM_FN void fwd_constraint(KM_ARGS) {
  typedef Dim<S> D;
  typedef Num<T> N;
  if (g.lane == 0) { e.solver_niter = 0; }
  // warm start: the previous qacc if it is cheaper than the unconstrained acceleration
  const T cw = total_cost<S, T, G>(e, m, g, e.warm), cs = total_cost<S, T, G>(e, m, g, e.qacc_smooth);
  KM_FOR(i, D::NV) e.qacc[i] = cw > cs ? e.qacc_smooth[i] : e.warm[i];
  g.sync();
  mul_M<S, T, G>(e, m, g, e.qacc, e.c.Ma);
  mul_J<S, T, G>(e, m, g, e.qacc, e.c.efc_jar);
  KM_FOR(r, e.nefc) e.c.efc_jar[r] -= e.efc_aref[r];
  g.sync();
  const T scale = T(1) / (m.meaninertia * T(D::NV));
  T gauss;
  T cost = sol_update<S, T, G>(e, m, g, &gauss);
  KM_CLK(CLK_SOL_SETUP);
  int niter = 0;
  bool done = niter >= m.iterations;
  // Newton iterations, marched CTA-wide: envs that have converged idle at the vote until the slowest is done
  while (true) {
    if (!done) {
      sol_hessian_dir<S, T, G>(e, m, g);
      KM_FOR(i, D::NV) e.c.search[i] = -e.c.Mgrad[i];
      g.sync();
      KM_CLK(CLK_SOL_DIR);
      const T alpha = sol_linesearch<S, T, G>(e, m, g, scale);
      KM_CLK(CLK_SOL_LS);
      if (alpha == T(0)) done = true;
      else {
        KM_FOR(i, D::NV) { e.qacc[i] += alpha * e.c.search[i]; e.c.Ma[i] += alpha * e.c.Mv[i]; }
        KM_FOR(r, e.nefc) e.c.efc_jar[r] += alpha * e.efc_jv[r];
        g.sync();
        const T oldcost = cost;
        cost = sol_update<S, T, G>(e, m, g, &gauss);
        T gn = 0;
        KM_FOR(i, D::NV) gn += e.c.grad[i] * e.c.grad[i];
        gn = g.sum(gn);
        niter++;
        done = scale * (oldcost - cost) < m.tol || scale * N::sqrt(gn) < m.tol || niter >= m.iterations;
        KM_CLK(CLK_SOL_UPD);
      }
    }
    const bool more = g.cta_any(!done);
    KM_CLK(CLK_SOL_VOTE);
    if (!more) break;
  }
  g.converge();
  KM_FOR(i, D::NV) e.warm[i] = e.qacc[i];
  if (g.lane == 0) e.solver_niter = niter;
  g.sync();
}

}  // namespace km
#include "km_solver_warp.cuh"
namespace km {

// mj_Euler (no joint damping anywhere): semi-implicit, free-joint quaternion integrated on the manifold (A6)
KM_TPL KM_FN void euler(KM_ARGS) {
  typedef Dim<S> D;
  KM_FOR(i, D::NV) {
    const T v = e.qvel[i] + m.h * e.qacc[i];
    e.qvel[i] = v;
    if (sizeof(T) == 4 && i >= D::NVA && i < D::NVA + 3) {
      // cube translation, float-float: (hi, lo) += h v by an error-free two-sum, then renormalise
      const T hi = e.qpos[i], p = m.h * v, s = hi + p, bb = s - hi;
      const T lo = e.cube_lo[i - D::NVA] + ((hi - (s - bb)) + (p - bb));
      const T hi2 = s + lo;
      e.qpos[i] = hi2;
      e.cube_lo[i - D::NVA] = lo - (hi2 - s);
    } else if (i < D::NVA + 3) e.qpos[i] += m.h * v;
  }
  g.sync();
  if (g.lane == 0) {
    T w[3] = {e.qvel[D::NVA + 3], e.qvel[D::NVA + 4], e.qvel[D::NVA + 5]};
    T* q = e.qpos + D::NVA + 3;
    const T angle = m.h * normalize3(w);
    T qrot[4] = {1, 0, 0, 0}, t[4];
    if (angle != T(0)) {
      T s, c;
      Num<T>::sincos(angle * T(0.5), &s, &c);
      qrot[0] = c; qrot[1] = w[0] * s; qrot[2] = w[1] * s; qrot[3] = w[2] * s;
    }
    qnormalize(q);
    qmul(t, q, qrot);
    q[0] = t[0]; q[1] = t[1]; q[2] = t[2]; q[3] = t[3];
    e.time += m.h;
  }
  g.sync();
}

KM_TPL KM_HD void step1(KM_ARGS) {
  fwd_position<S, T, G>(e, m, g);
  g.cta_sync();
  KM_CLK(CLK_BARRIER);
  fwd_velocity<S, T, G>(e, m, g);
  KM_CLK(CLK_VEL);
}
// KM_WARP_SOLVER = 0 keeps the generic lane-group solver for every env (A/B experiments)
#ifndef KM_WARP_SOLVER
#define KM_WARP_SOLVER 1
#endif
KM_TPL KM_HD void step2(KM_ARGS) {
  g.cta_sync();
  KM_CLK(CLK_BARRIER);
  fwd_actuation<S, T, G>(e, m, g);
  KM_CLK(CLK_ACC);
  g.cta_sync();
  KM_CLK(CLK_BARRIER);
#if KM_WARP_CODE && KM_WARP_SOLVER
  if constexpr (G == 32) {
    // register-resident solver (km_solver_warp.cuh); the coupled instantiation (a finger pad touches the cube) carries
    // two contact sets (up to eight contacts) and a dense Hessian
    if (e.coupled) fwd_acc_constraint_w<S, T, true>(e, m, g);
    else fwd_acc_constraint_w<S, T, false>(e, m, g);
    euler<S, T, G>(e, m, g);
    KM_CLK(CLK_EULER);
    return;
  }
#endif
  fwd_acceleration<S, T, G>(e, m, g);
  if constexpr (G == 1) { fwd_constraint_tpe<S, T>(e, m); KM_CLK(CLK_SOL_DIR); }
  else fwd_constraint<S, T, G>(e, m, g);
  euler<S, T, G>(e, m, g);
  KM_CLK(CLK_EULER);
}

// =========================================================================================== task: action decode + IK
// site pose of arm a from the current link frames
KM_TPL KM_HD void site_pose(const E& e, const Model<S, T>& m, int a, T* pos, T* mat) {
  const int l = m.arm_site_link[a];
  T t[3], q[4];
  mulv3(t, e.xmat[l], m.site_pos[a]);
  for (int i = 0; i < 3; i++) pos[i] = e.xpos[l][i] + t[i];
  qmul(q, e.xquat[l], m.site_quat[a]);
  q2mat(mat, q);
}

// scipy Rotation.from_matrix(R).as_euler("xyz") (extrinsic), then from_euler("xyz", e).as_quat()[[3,0,1,2]]
template <typename T> KM_HD void mat_to_euler_xyz_ext(T* eul, const T* R) {
  typedef Num<T> N;
  eul[1] = N::asin(tclip(-R[6], T(-1), T(1)));
  eul[0] = N::atan2(R[7], R[8]);
  eul[2] = N::atan2(R[3], R[0]);
}
template <typename T> KM_HD void euler_xyz_ext_to_quat(T* q, const T* eul) {
  typedef Num<T> N;
  T s0, c0, s1, c1, s2, c2;
  N::sincos(eul[0] * T(0.5), &s0, &c0);
  N::sincos(eul[1] * T(0.5), &s1, &c1);
  N::sincos(eul[2] * T(0.5), &s2, &c2);
  const T qx[4] = {c0, s0, 0, 0}, qy[4] = {c1, 0, s1, 0}, qz[4] = {c2, 0, 0, s2};
  T t[4];
  qmul(t, qy, qx);
  qmul(q, qz, t);
}

// ---- inverse kinematics.  All IK arithmetic is fp64 in both builds: the problem is regularised only by
// lam ~ 5e-5 along the arm's null space, so fp32 evaluation noise of the chain kinematics (1e-7 m) would move the
// solution by ~1e-4 rad, which the kp = 1000 servos then turn into visible velocity differences.  The chain is
// serial anyway (one lane sweeps it); the per-column work is spread over the lanes.

// forward kinematics of arm a's chain at joint values b.x (masked joints) / qpos (the others): per-link axis and
// anchor, site pose.  One lane.
template <class S, typename T, int G, class E, class B> KM_HD void ik_chain_fk(E& e, B& b, const Model<S, T>& m, int a, const double* x) {
  typedef Num<double> N;
  double q[4] = {1, 0, 0, 0}, pos[3] = {0, 0, 0}, mat[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int k = 0; k < m.arm_nchain[a]; k++) {
    const int l = m.arm_chain[a][k], mi = m.arm_chain_mask[a][k];
    double t[3], ql[4];
    mulv3(t, mat, m.dk_lpos[l]);
    for (int i = 0; i < 3; i++) pos[i] += t[i];
    qmul(ql, q, m.dk_lquat[l]);
    const double th = mi >= 0 ? x[mi] : (double)e.qpos[l];
    if (m.jtype[l] == JT_HINGE) {
      double s, c;
      N::sincos(th * 0.5, &s, &c);
      q[0] = ql[0] * c - ql[3] * s; q[1] = ql[1] * c + ql[2] * s; q[2] = ql[2] * c - ql[1] * s; q[3] = ql[3] * c + ql[0] * s;
    } else { q[0] = ql[0]; q[1] = ql[1]; q[2] = ql[2]; q[3] = ql[3]; }
    qnormalize(q);
    q2mat(mat, q);
    for (int i = 0; i < 3; i++) { b.an[k][i] = pos[i]; b.ax[k][i] = mat[3 * i + 2]; }
    if (m.jtype[l] == JT_SLIDE) { pos[0] += mat[2] * th; pos[1] += mat[5] * th; pos[2] += mat[8] * th; }
  }
  double t[3], sq[4];
  mulv3(t, mat, m.dk_site_pos[a]);
  for (int i = 0; i < 3; i++) b.spos[i] = pos[i] + t[i];
  qmul(sq, q, m.dk_site_quat[a]);
  q2mat(b.smat, sq);
}

// The same with one chain link per lane (warp-per-env kernels; every lane of the group calls it): each lane forms its
// link's transform relative to the parent (the fp64 sincos and quaternion products of all links at once), an inclusive
// prefix product over the chain composes them in four rounds of shuffles, and the quaternion is normalised once at the
// end instead of after every link -- results differ from the serial sweep by a few ulp of fp64.
template <class S, typename T, int G, class E, class B> KM_HD void ik_chain_fk_lanes(E& e, B& b, const Model<S, T>& m, const Grp<G>& g, int a, const double* x) {
  typedef Num<double> N;
  static_assert(Dim<S>::MAXLEVEL <= 16 && G >= Dim<S>::MAXLEVEL, "one chain link per lane, four prefix rounds");
  const int nc = m.arm_nchain[a];
  const bool on = g.lane < nc;
  const int k = on ? g.lane : nc - 1;            // surplus lanes shadow the last link (they take part in the shuffles)
  const int l = m.arm_chain[a][k], mi = m.arm_chain_mask[a][k];
  const double th = mi >= 0 ? x[mi] : (double)e.qpos[l];
  const bool hinge = m.jtype[l] == JT_HINGE;
  double q[4] = {m.dk_lquat[l][0], m.dk_lquat[l][1], m.dk_lquat[l][2], m.dk_lquat[l][3]};
  double p[3] = {m.dk_lpos[l][0], m.dk_lpos[l][1], m.dk_lpos[l][2]};
  if (hinge) {                                     // q <- q * (c, 0, 0, s)
    double sn, cs;
    N::sincos(th * 0.5, &sn, &cs);
    const double t0 = q[0] * cs - q[3] * sn, t1 = q[1] * cs + q[2] * sn, t2 = q[2] * cs - q[1] * sn, t3 = q[3] * cs + q[0] * sn;
    q[0] = t0; q[1] = t1; q[2] = t2; q[3] = t3;
  } else {                                         // slide along the link's own z: the offset belongs to the link's frame origin
    const double z[3] = {0, 0, th};
    double d[3];
    qrot(d, q, z);
    p[0] += d[0]; p[1] += d[1]; p[2] += d[2];
  }
#pragma unroll
  for (int d = 1; d < 16; d *= 2) {
    const bool has = g.lane >= d && on;
    const int src = has ? g.lane - d : g.lane;
    double qa[4], pa[3];
    for (int i = 0; i < 4; i++) qa[i] = g.shfl(q[i], src);
    for (int i = 0; i < 3; i++) pa[i] = g.shfl(p[i], src);
    if (has) {
      double t[3], qq[4];
      qrot(t, qa, p);
      p[0] = pa[0] + t[0]; p[1] = pa[1] + t[1]; p[2] = pa[2] + t[2];
      qmul(qq, qa, q);
      q[0] = qq[0]; q[1] = qq[1]; q[2] = qq[2]; q[3] = qq[3];
    }
  }
  qnormalize(q);
  double mat[9];
  q2mat(mat, q);
  if (on) {
    // anchor = the joint's position: the frame origin, minus the slide offset for sliders
    const double off = hinge ? 0.0 : th;
    for (int i = 0; i < 3; i++) { b.ax[k][i] = mat[3 * i + 2]; b.an[k][i] = p[i] - mat[3 * i + 2] * off; }
  }
  if (g.lane == nc - 1) {
    double t[3], sq[4];
    mulv3(t, mat, m.dk_site_pos[a]);
    for (int i = 0; i < 3; i++) b.spos[i] = p[i] + t[i];
    qmul(sq, q, m.dk_site_quat[a]);
    q2mat(b.smat, sq);
  }
  g.sync();
}

// ik_res (reference ik_mujoco.py:20-53) from the chain pose currently in e.b; pose rows by lane 0, regularisers by all
template <class S, typename T, int G, class E, class B> KM_HD void ik_residual(E& e, B& b, const Model<S, T>& m, const Grp<G>& g, int a, const double* x, double* res) {
  const int n = m.arm_nmask[a];
  if (g.lane == 0) {
    double cur[4], rq[3];
    for (int i = 0; i < 3; i++) res[i] = b.spos[i] - b.goal[i];
    mat2quat(cur, b.smat);
    subquat(rq, b.goal + 3, cur);
    for (int i = 0; i < 3; i++) res[3 + i] = rq[i] * 0.02;                        // IK_RES_RAD
  }
  KM_FOR(i, n) {
    res[6 + i] = 6e-3 * (x[i] - b.qprev[i]);                                     // IK_RES_REG_PREV
    res[6 + n + i] = 2e-6 * (x[i] - m.dk_qhome[m.arm_mask[a][i]]);                 // IK_RES_REG_HOME
  }
  g.sync();
}

// pose rows of ik_jac (reference ik_mujoco.py:56-97): [Jp ; rad * d subQuat(goal, cur) / dq], columns = mask.
// The orientation block is -rad * Jl^{-1}(phi) R_site^T Jr with phi = subQuat(goal, cur)  (DESIGN.md).
template <class S, typename T, int G, class E, class B> KM_HD void ik_jacobian(E& e, B& b, const Model<S, T>& m, const Grp<G>& g, int a) {
  typedef Num<double> N;
  const int n = m.arm_nmask[a];
  KM_FOR(c, n) {
    const double* R = b.smat;
    double cur[4], phi[3];
    mat2quat(cur, R);
    subquat(phi, b.goal + 3, cur);
    double u[3] = {phi[0], phi[1], phi[2]};
    const double half = 0.5 * normalize3(u);
    const double coef = 1.0 - (half < 6e-8 ? 1.0 : half / N::tan(half));
    const double K[9] = {0, -u[2], u[1], u[2], 0, -u[0], -u[1], u[0], 0};
    double Dm[9];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        double kk = 0;
        for (int k = 0; k < 3; k++) kk += K[3 * i + k] * K[3 * k + j];
        Dm[3 * i + j] = (i == j ? 1.0 : 0.0) - half * K[3 * i + j] + coef * kk;
      }
    const int k = m.arm_mask_chain[a][c];   // position of masked joint c in the chain
    const double* ax = b.ax[k];
    double jp[3], jr[3] = {0, 0, 0};
    if (m.jtype[m.arm_mask[a][c]] == JT_SLIDE) { jp[0] = ax[0]; jp[1] = ax[1]; jp[2] = ax[2]; }
    else {
      const double o[3] = {b.spos[0] - b.an[k][0], b.spos[1] - b.an[k][1], b.spos[2] - b.an[k][2]};
      cross3(jp, ax, o);
      jr[0] = ax[0]; jr[1] = ax[1]; jr[2] = ax[2];
    }
    double jl[3], o3[3];
    mulTv3(jl, R, jr);
    mulv3(o3, Dm, jl);
    for (int i = 0; i < 3; i++) { b.J[i][c] = jp[i]; b.J[3 + i][c] = -0.02 * o3[i]; }   // IK_JAC_RAD
  }
  g.sync();
}

}  // namespace km
#include "km_ik_trf.cuh"
namespace km {

// Device IK of one arm, fused in front of the sub-steps: end-effector target from the action (reference
// env_sim.py:60-70, 80-90), then projected Levenberg-Marquardt on the reference's stationarity condition J^T r = 0
// with J, r exactly as ik_jac / ik_res build them (including their mismatched regulariser weights),
// a fixed iteration count, the bound handling of scipy's TRF at its KKT point, and the reference's side effect of
// leaving qpos[mask] at the solution (B-1).  Skipped when x0 is out of bounds (B-4).
KM_TPL KM_FN void ik_solve(KM_ARGS, int a, const float* act) {
  typedef Dim<S> D;
  // the IK scratch is fp64; the thread-per-env working set holds no 8-byte members, so there it is a local
  typename E::StageB blocal;
  typename E::StageB& b = stage_b(e, blocal);
  const int n = m.arm_nmask[a], nr = 6 + 2 * n;
  bool bad = false;
  KM_FOR(i, n) {
    const int j = m.arm_mask[a][i];
    const double x = (double)e.qpos[j];
    b.x[i] = x; b.qprev[i] = x; b.lo[i] = m.dk_range[j][0]; b.hi[i] = m.dk_range[j][1];
    bad = bad || x < m.dk_range[j][0] || x > m.dk_range[j][1];
  }
  const bool feasible = !g.any(bad);
  g.sync();
  // chain kinematics: one link per lane in the warp-per-env kernels, one lane sweeping the chain otherwise
#if KM_WARP_CODE
  constexpr bool kFkLanes = G >= D::MAXLEVEL && G > 1;
#else
  constexpr bool kFkLanes = false;
#endif
  if constexpr (kFkLanes) ik_chain_fk_lanes<S, T, G>(e, b, m, g, a, b.x);
  if (g.lane == 0) {
    // goal = current site pose displaced by the action (EE_POS_DELTA 0.01, EE_ORN_DELTA 0.1, extrinsic xyz Euler)
    double eul[3];
    if constexpr (!kFkLanes) ik_chain_fk<S, T, G>(e, b, m, a, b.x);
    for (int i = 0; i < 3; i++) b.goal[i] = (double)act[m.off_pos[a] + i] * 0.01 + b.spos[i];
    mat_to_euler_xyz_ext(eul, b.smat);
    for (int i = 0; i < 3; i++) eul[i] = (double)act[m.off_orn[a] + i] * 0.1 + eul[i];
    euler_xyz_ext_to_quat(b.goal + 3, eul);
    for (int i = 0; i < 7; i++) e.mocap[7 * m.arm_mocap[a] + i] = (T)b.goal[i];
  }
  g.sync();
  // The exact-parity mode is compiled into the lane-group kernels only (and the host build): its serial fp64 code in a
  // thread-per-env kernel made ptxas halve the register budget of the whole kernel (255 -> 128 registers, 7x the spills
  // in the solver: DualArm 32768 envs 2.65 -> 2.27e6 env-steps/s).  km_api.cu routes ik_mode = 1 handles to the
  // lane-group mapping, so no mapping runs the other optimiser silently.
#if defined(__CUDA_ARCH__)
  constexpr bool kTrf = G > 1;
#else
  constexpr bool kTrf = true;
#endif
  bool trf = false;
  if constexpr (kTrf) trf = feasible && m.ik_mode == 1;
  if (trf) {
    // exact-parity mode: the reference's optimiser (scipy TRF) restated, km_ik_trf.cuh; one lane, fp64
    if constexpr (kTrf) {
      if (g.lane == 0) {
#if defined(__CUDA_ARCH__)
        // the solve's work arrays: this env's slice of the dynamic shared memory behind the env records (km_launch.cuh)
        extern __shared__ __align__(16) unsigned char km_dyn_smem[];
        constexpr size_t model_b = (sizeof(Model<S, T>) + 15) / 16 * 16, env_b = (sizeof(E) + 15) / 16 * 16;
        const int epb = blockDim.x / G, slot = threadIdx.x / G;
        trf::TrfWork& wk = *(trf::TrfWork*)(km_dyn_smem + model_b + (size_t)epb * env_b + (size_t)slot * sizeof(trf::TrfWork));
#else
        trf::TrfWork wk_local;
        trf::TrfWork& wk = wk_local;
#endif
        ik_trf_serial<S, T>(e, b, m, a, wk);
      }
    }
    g.sync();
  } else if (feasible) {
    const double lam = 9e-3 * (6e-3 + 2e-6), reg = 9e-3;   // IK_JAC_REG * (IK_RES_REG_PREV + IK_RES_REG_HOME)
    double mu = 0;
    ik_residual<S, T, G>(e, b, m, g, a, b.x, b.r);
    ik_jacobian<S, T, G>(e, b, m, g, a);
    double cs = 0;
    KM_FOR(k, nr) cs += 0.5 * b.r[k] * b.r[k];
    double cost = g.sum(cs);
    for (int it = 0; it < m.ik_iters; it++) {
      KM_FOR(i, n) {
        double gi = 0;
        for (int k = 0; k < 6; k++) gi += b.J[k][i] * b.r[k];
        gi += reg * b.r[6 + i] + reg * b.r[6 + n + i];
        b.gv[i] = -gi;
        b.active[i] = (b.x[i] <= b.lo[i] && gi > 0.0) || (b.x[i] >= b.hi[i] && gi < 0.0);
      }
      g.sync();
      KM_FOR(w, n * n) {
        const int i = w / n, j = w - i * n;
        if (j > i) continue;
        double sacc = 0;
        if (b.active[i] || b.active[j]) sacc = i == j ? 1.0 : 0.0;
        else {
          for (int k = 0; k < 6; k++) sacc += b.J[k][i] * b.J[k][j];
          if (i == j) sacc += lam + mu;
        }
        b.A[i][j] = sacc;
      }
      g.sync();
      if (g.lane == 0) {   // 7 x 7 Cholesky and the two triangular solves
        for (int i = 0; i < n; i++) if (b.active[i]) b.gv[i] = 0;
        for (int j = 0; j < n; j++) {
          double d = b.A[j][j];
          for (int k = 0; k < j; k++) d -= b.A[j][k] * b.A[j][k];
          d = Num<double>::sqrt(d);
          b.A[j][j] = d;
          for (int i = j + 1; i < n; i++) {
            double u = b.A[i][j];
            for (int k = 0; k < j; k++) u -= b.A[i][k] * b.A[j][k];
            b.A[i][j] = u / d;
          }
        }
        for (int i = 0; i < n; i++) { double t = b.gv[i]; for (int k = 0; k < i; k++) t -= b.A[i][k] * b.gv[k]; b.gv[i] = t / b.A[i][i]; }
        for (int i = n - 1; i >= 0; i--) { double t = b.gv[i]; for (int k = i + 1; k < n; k++) t -= b.A[k][i] * b.gv[k]; b.gv[i] = t / b.A[i][i]; }
        for (int i = 0; i < n; i++) b.xn[i] = tclip(b.x[i] + b.gv[i], b.lo[i], b.hi[i]);
        if constexpr (!kFkLanes) ik_chain_fk<S, T, G>(e, b, m, a, b.xn);
      }
      g.sync();
      if constexpr (kFkLanes) ik_chain_fk_lanes<S, T, G>(e, b, m, g, a, b.xn);
      ik_residual<S, T, G>(e, b, m, g, a, b.xn, b.rn);
      double cn = 0;
      KM_FOR(k, nr) cn += 0.5 * b.rn[k] * b.rn[k];
      const double costn = g.sum(cn);
      if (costn <= cost) {
        KM_FOR(i, n) b.x[i] = b.xn[i];
        KM_FOR(k, nr) b.r[k] = b.rn[k];
        g.sync();
        cost = costn;
        ik_jacobian<S, T, G>(e, b, m, g, a);
        mu = mu * 0.25;
        if (mu < 1e-6) mu = 0;
      } else {
        mu = mu == 0.0 ? 1e-4 : mu * 4.0;
      }
    }
  }
  KM_FOR(i, n) {
    const int j = m.arm_mask[a][i];
    const float q = (float)tclip(b.x[i], b.lo[i], b.hi[i]);   // ik_mujoco.py:147-152, then ctrl is float32
    e.ctrl[j] = (T)q;
    // the reference leaves qpos[mask] at the last point it evaluated (B-1): the solution, or TRF's last trial point
    if (feasible && m.ik_teleport) e.qpos[j] = (T)(m.ik_mode == 1 ? b.xn[i] : b.x[i]);
  }
  g.sync();
}

// KManipTask.before_step (reference env_sim.py:38-108); `act` points at this env's float32 action record
KM_TPL KM_FN void before_step(KM_ARGS, const float* act) {
  typedef Dim<S> D;
  // ctrl passes through float32 (env_sim.py:40); the state already holds float32-representable values
  KM_FOR(i, D::NU) e.ctrl[i] = (T)(float)e.ctrl[i];
  g.sync();
  // grippers (env_sim.py:41-59): both sliders follow slider 0, float32 arithmetic, clipped to EE_S_MIN/MAX
  if (g.lane == 0) {
    for (int a = 0; a < m.n_arm; a++)
      if (m.off_grip[a] >= 0) {
        float gv = act[m.off_grip[a]] * 0.0001f;
        gv = (float)((double)gv + (double)e.qpos[m.arm_grip[a][0]]);
        gv = gv < -0.029f ? -0.029f : (gv > 0.005f ? 0.005f : gv);
        e.ctrl[m.arm_grip[a][0]] = (T)gv;
        e.ctrl[m.arm_grip[a][1]] = (T)gv;
      }
  }
  if (m.act_mode == 0) {
    // end-effector targets + IK, right arm then left (env_sim.py:60-99); the arms' chains are independent
    for (int a = 0; a < m.n_arm; a++)
      if (m.off_pos[a] >= 0) ik_solve<S, T, G>(e, m, g, a, act);
  } else {
    // joint-position deltas (env_sim.py:100-103)
    for (int a = 0; a < m.n_arm; a++) {
      if (m.off_q[a] < 0) continue;
      KM_FOR(i, m.arm_nmask[a]) {
        const int j = m.arm_mask[a][i];
        const float da = act[m.off_q[a] + i] * 0.1f;
        e.ctrl[j] = (T)(float)((double)e.qpos[j] + (double)da);
      }
    }
  }
  g.sync();
}

// KManipTask.get_observation (reference env_sim.py:110-146): [q_pos, q_vel, cube_pos, cube_orn]
KM_TPL KM_FN void observation(KM_ARGS) {
  typedef Dim<S> D;
  const T pi = T(3.14159265358979323846);
  KM_FOR(i, D::OBS) {
    T v;
    if (i < D::QLEN) v = tclip((e.qpos[i] - m.range[i][0]) / (m.range[i][1] - m.range[i][0]), T(-1), T(1));
    else if (i < 2 * D::QLEN) v = tclip(e.qvel[i - D::QLEN] / pi, T(-1), T(1));
    else if (i < 2 * D::QLEN + 3) {
      const int k = i - 2 * D::QLEN;
      v = tclip(((e.qpos[D::NVA + k] - m.spawn_lo[k]) + e.cube_lo[k]) / (m.spawn_hi[k] - m.spawn_lo[k]), T(-1), T(1));
    } else v = e.qpos[D::NVA + 3 + (i - 2 * D::QLEN - 3)];
    e.obs[i] = v;
  }
  g.sync();
}

// KManipTask.get_reward (reference env_sim.py:148-179).  The touch/lift bonuses are unreachable in the shipped
// model; *flags reports what the contact scan saw: bit0 cube-table, bit1 right pads, bit2 left pads.
KM_TPL KM_FN T reward(KM_ARGS, int* flags) {
  typedef Dim<S> D;
  typedef Num<T> N;
  T vn = 0;
  KM_FOR(i, D::NV) vn += e.qvel[i] * e.qvel[i];
  T r = -T(0.01) * N::sqrt(g.sum(vn));
  for (int a = m.n_arm - 1; a >= 0; a--) {
    if (m.off_grip[a] < 0) continue;
    T pos[3], mat[9];
    site_pose<S, T, G>(e, m, a, pos, mat);
    const T d[3] = {e.qpos[D::NVA] - pos[0], e.qpos[D::NVA + 1] - pos[1], e.qpos[D::NVA + 2] - pos[2]};
    r += T(0.01) * (T(1) / (N::sqrt(dot3(d, d)) + T(1e-6)));
  }
  int f = 0;
  for (int c = 0; c < e.ncon; c++) {
    const int s = e.con_slot[c];
    f |= s >= D::NPAD ? 1 : (m.pad_arm[s] == 0 ? 2 : 4);
  }
  *flags = f;
  return r;
}

// KManipTask.initialize_episode (reference env_sim.py:23-36) with a counter-based device RNG for the cube spawn
KM_TPL KM_FN void reset_state(KM_ARGS, uint64_t seed, uint64_t env_id, const T* cube_xyz) {
  typedef Dim<S> D;
  KM_FOR(i, D::NV) { e.qvel[i] = 0; e.warm[i] = 0; }
  KM_FOR(i, D::NVA) {
    const T v = i < D::QLEN ? m.q_home[i] : T(0);
    e.qpos[i] = v; e.ctrl[i] = v;
  }
  KM_FOR(i, D::NMOCAP * 7) e.mocap[i] = m.mocap0[i];
  if (g.lane == 0) {
    if (cube_xyz) { for (int i = 0; i < 3; i++) { e.qpos[D::NVA + i] = cube_xyz[i]; e.cube_lo[i] = 0; } }
    else {
      double u[3];
      spawn_uniforms(seed, env_id, (uint32_t)e.episode, u);
      for (int i = 0; i < 3; i++) {
        const double x = m.spawn_lo_d[i] + u[i] * (m.spawn_hi_d[i] - m.spawn_lo_d[i]);
        const T hi = (T)x;
        e.qpos[D::NVA + i] = hi;
        e.cube_lo[i] = sizeof(T) == 4 ? (T)(x - (double)hi) : T(0);
      }
    }
    for (int i = 0; i < 4; i++) e.qpos[D::NVA + 3 + i] = m.cube_quat0[i];
    e.time = 0; e.step = 0;
  }
  g.sync();
}

// one-time set-up of the parts of the working set that never change
KM_TPL KM_FN void init_env(KM_ARGS) {
  typedef Dim<S> D;
  KM_FOR(w, D::NVA * D::NVA) (&e.M[0][0])[w] = 0;
  KM_FOR(r, D::NFRIC) {
    e.efc_desc[r] = efc_pack(EFC_FRICTION, m.fric_dof[r], 0, 0);
    e.efc_D[r] = m.fr_D[r];
  }
  if (g.lane == 0) { e.ls_evals = 0; e.ik_evals = 0; e.solver_niter = 0; e.ncon = 0; e.nlim = 0; e.nefc = D::NFRIC; }
  g.sync();
}

// Outputs of one env step (any pointer may be null)
template <typename T> struct StepOut {
  T* obs; T* final_obs; T* reward; unsigned char* truncated; unsigned char* terminated;
  int* con_flags; int* ncon; int* con_geoms; int con_cap;
  unsigned* clk;   // debug build (KM_PHASE_CLOCKS): [n][16] cycles per phase of this env step
  // episode bookkeeping (reference env_base.py:243-250 `info`, batched): running return of every env (persistent),
  // per-step info arrays, and the CTA's accumulators of the rollout totals {sum reward, env steps, finished episodes,
  // success steps} that the kernel adds to the handle's totals once per CTA
  T* ep_return; T* episode_return; T* final_return; T* sim_time; unsigned char* is_success; int* step_out; int* episode_out;
  double* cta_totals;
};
// adds to one of a CTA's accumulators (shared memory on the device; plain memory in the host build of the tests)
KM_HD void km_accumulate(double* dst, double v) {
#if defined(__CUDA_ARCH__)
  atomicAdd(dst, v);
#else
  *dst += v;
#endif
}

// One env step on the working set already holding the env's state.
KM_TPL KM_FN void env_step(KM_ARGS, const float* act, const StepOut<T>& o, long env, int autoreset, uint64_t seed,
                           uint64_t env0) {
  typedef Dim<S> D;
#if defined(KM_PHASE_CLOCKS) && defined(__CUDA_ARCH__)
  if (g.lane == 0) { for (int i = 0; i < 16; i++) e.clk[i] = 0; e.clk_last = (unsigned)clock(); }
#endif
  step1<S, T, G>(e, m, g);
  g.cta_sync();
  KM_CLK(CLK_BARRIER);
  before_step<S, T, G>(e, m, g, act);
  KM_CLK(CLK_BEFORE);
  step2<S, T, G>(e, m, g);
  for (int s = 1; s < m.nsub; s++) { step1<S, T, G>(e, m, g); step2<S, T, G>(e, m, g); }
  // closing mj_step1: only kinematics and collision feed the reward / contact report; the next env step
  // recomputes the full position and velocity stages from the stored state
  kinematics<S, T, G>(e, m, g);
  collision<S, T, G>(e, m, g);
  int fl;
  const T r = reward<S, T, G>(e, m, g, &fl);
  observation<S, T, G>(e, m, g);
  if (o.obs) KM_FOR(i, D::OBS) o.obs[env * D::OBS + i] = e.obs[i];
  if (g.lane == 0) {
    if (o.reward) o.reward[env] = r;
    if (o.con_flags) o.con_flags[env] = fl;
    if (o.ncon) o.ncon[env] = e.ncon;
    if (o.con_geoms)
      for (int c = 0; c < o.con_cap; c++) {
        int g1 = -1, g2 = -1;
        if (c < e.ncon) { const int s = e.con_slot[c]; g1 = s < D::NPAD ? m.pad_geom[s] : m.table_geom; g2 = m.cube_geom; }
        o.con_geoms[env * 2 * o.con_cap + 2 * c] = g1;
        o.con_geoms[env * 2 * o.con_cap + 2 * c + 1] = g2;
      }
    e.step += 1;
  }
  g.sync();
  const bool trunc = e.step >= m.max_episode_steps;
  if (g.lane == 0) {
    if (o.truncated) o.truncated[env] = trunc ? 1 : 0;
    if (o.terminated) o.terminated[env] = 0;   // the reference never terminates
    // info of this step (env_base.py:243-250): step index inside the episode, episode index, simulation time,
    // is_success = reward > REWARD_SUCCESS_THRESHOLD (2.0, __init__.py:204), the running and the finished return
    const bool success = r > T(2);
    if (o.is_success) o.is_success[env] = success ? 1 : 0;
    if (o.step_out) o.step_out[env] = e.step;
    if (o.episode_out) o.episode_out[env] = e.episode;
    if (o.sim_time) o.sim_time[env] = e.time;
    if (o.ep_return) {
      const T ret = o.ep_return[env] + r;
      if (o.episode_return) o.episode_return[env] = ret;
      if (o.final_return) o.final_return[env] = trunc ? ret : T(0);
      o.ep_return[env] = (autoreset && trunc) ? T(0) : ret;
    }
    if (o.cta_totals) {
      km_accumulate(o.cta_totals + 0, (double)r);
      km_accumulate(o.cta_totals + 1, 1.0);
      km_accumulate(o.cta_totals + 2, trunc ? 1.0 : 0.0);
      km_accumulate(o.cta_totals + 3, success ? 1.0 : 0.0);
    }
  }
  if (autoreset && trunc) {
    // same-step autoreset: final_obs keeps the last observation of the finished episode
    if (o.final_obs) KM_FOR(i, D::OBS) o.final_obs[env * D::OBS + i] = e.obs[i];
    if (g.lane == 0) e.episode += 1;
    g.sync();
    reset_state<S, T, G>(e, m, g, seed, env0 + (uint64_t)env, (const T*)0);
    observation<S, T, G>(e, m, g);
    if (o.obs) KM_FOR(i, D::OBS) o.obs[env * D::OBS + i] = e.obs[i];
  }
  g.sync();
  KM_CLK(CLK_EPILOGUE);
#if defined(KM_PHASE_CLOCKS) && defined(__CUDA_ARCH__)
  if (o.clk && g.lane == 0) for (int i = 0; i < 16; i++) o.clk[env * 16 + i] = e.clk[i];
#endif
}

}  // namespace km
