// km_ik_trf.cuh -- exact-parity IK mode: the reference's optimiser itself, restated for the device.
//
// The reference solves its IK with scipy.optimize.least_squares (reference ik_mujoco.py:129-135): method 'trf' with
// bounds, tr_solver 'exact', ftol = xtol = gtol = 1e-8, x_scale = 1, max_nfev = 100 n, analytic Jacobian.  scipy is a
// third-party dependency of the reference that is not vendored in it (undeclared in pyproject.toml; 1.18.1 is what the
// build image carries), so this file restates the published algorithm of that version:
//     scipy/optimize/_lsq/trf.py      trf_bounds, select_step
//     scipy/optimize/_lsq/common.py   CL_scaling_vector, solve_lsq_trust_region, update_tr_radius, check_termination,
//                                      step_size_to_bound, intersect_trust_region, build/minimize/evaluate_quadratic,
//                                      make_strictly_feasible, find_active_constraints
//     scipy/optimize/_lsq/least_squares.py  (x0 made strictly feasible, ValueError when x0 is out of bounds)
// with one substitution that is exact in exact arithmetic: instead of the SVD of the (m+n) x n augmented Jacobian it
// uses the n x n matrix A = J_h^T J_h + diag(diag_h): its Cholesky factor for the Gauss-Newton step (what the
// trust-region solver returns whenever that step lies inside the region), and otherwise its symmetric
// eigen-decomposition (cyclic Jacobi), whose eigenvalues are the squared singular values, whose eigenvectors are V,
// and for which s * (U^T f_aug) = V^T g_h.
// Parity against the real scipy is checked by tests/test_ik_trf.py (host build of this file, fp64) and on the GPU.
//
// Everything is fp64 and serial: one lane runs it (n <= 8 unknowns) on work arrays in shared memory; it is the parity mode, not the fast path
// (km_task.ik_mode = 1; the default mode 0 is the fixed-iteration projected LM of km_sim.cuh).
#pragma once

namespace km {

namespace trf {
constexpr int NMAX = 8;
constexpr double EPS = 2.220446049250313e-16;
constexpr double INF = __builtin_huge_val();

// Work arrays of one TRF solve.  On the device they live in the CTA's dynamic shared memory behind the env records
// (km_launch.cuh adds sizeof(TrfWork) per env when the handle runs ik_mode = 1): as locals of the one lane that runs the
// solve they are a 5 KB stack frame in local memory that thrashes the small L1 left beside the shared-memory carve-out,
// and every access goes to L2.
struct TrfWork {
  double x[NMAX], f[6 + 2 * NMAX], g[NMAX], JtJ[NMAX][NMAX], x0n[NMAX];
  double v[NMAX], dv[NMAX], d[NMAX], diag_h[NMAX], g_h[NMAX], Bh[NMAX][NMAX], Aw[NMAX][NMAX], V[NMAX][NMAX], w[NMAX], s[NMAX], suf[NMAX], t0[NMAX];
  double p_gn[NMAX], Lc[NMAX][NMAX], x_new[NMAX], f_new[6 + 2 * NMAX];
  double p_h[NMAX], p[NMAX], step[NMAX], step_h[NMAX], xp[NMAX], r_h[NMAX], r[NMAX], x_on_bound[NMAX], ag_h[NMAX], ag[NMAX];
  double steps[NMAX], t[NMAX];
  int hits[NMAX];
};

KM_HD double norm(const double* x, int n) {
  double s = 0;
  for (int i = 0; i < n; i++) s += x[i] * x[i];
  return Num<double>::sqrt(s);
}
KM_HD double dot(const double* a, const double* b, int n) {
  double s = 0;
  for (int i = 0; i < n; i++) s += a[i] * b[i];
  return s;
}
KM_HD double nextafter_toward(double x, double to) {
#if defined(__CUDA_ARCH__)
  return ::nextafter(x, to);
#else
  return std::nextafter(x, to);
#endif
}
// s^T B s for a symmetric n x n matrix (row stride NMAX), and s0^T B s
KM_HD double quad(const double (*Bm)[NMAX], const double* s, const double* s0, int n) {
  double q = 0;
  for (int i = 0; i < n; i++) {
    double r = 0;
    for (int j = 0; j < n; j++) r += Bm[i][j] * s[j];
    q += s0[i] * r;
  }
  return q;
}
KM_HD bool in_bounds(const double* x, const double* lb, const double* ub, int n) {
  for (int i = 0; i < n; i++) if (!(x[i] >= lb[i] && x[i] <= ub[i])) return false;
  return true;
}
// common.py: step_size_to_bound -- min over components of the step to the bound it moves towards; hits[i] = sign(s_i)
// where that minimum is attained
KM_HD double step_size_to_bound(const double* x, const double* s, const double* lb, const double* ub, int n, int* hits, double* steps) {
  double mn = INF;
  for (int i = 0; i < n; i++) {
    steps[i] = INF;
    if (s[i] != 0.0) {
      const double a = (lb[i] - x[i]) / s[i], b = (ub[i] - x[i]) / s[i];
      steps[i] = a > b ? a : b;
    }
    if (steps[i] < mn) mn = steps[i];
  }
  if (hits)
    for (int i = 0; i < n; i++) hits[i] = steps[i] == mn ? (s[i] > 0 ? 1 : (s[i] < 0 ? -1 : 0)) : 0;
  return mn;
}
// common.py: make_strictly_feasible
KM_HD void make_strictly_feasible(double* x, const double* lb, const double* ub, int n, double rstep) {
  for (int i = 0; i < n; i++) {
    int active = 0;
    if (rstep == 0.0) {
      if (x[i] <= lb[i]) active = -1;
      if (x[i] >= ub[i]) active = 1;
    } else {
      const double ld = x[i] - lb[i], ud = ub[i] - x[i];
      const double lt = rstep * tmax(1.0, Num<double>::abs(lb[i])), ut = rstep * tmax(1.0, Num<double>::abs(ub[i]));
      if (ld <= tmin(ud, lt)) active = -1;
      if (ud <= tmin(ld, ut)) active = 1;
    }
    if (active == -1) x[i] = rstep == 0.0 ? nextafter_toward(lb[i], ub[i]) : lb[i] + rstep * tmax(1.0, Num<double>::abs(lb[i]));
    if (active == 1) x[i] = rstep == 0.0 ? nextafter_toward(ub[i], lb[i]) : ub[i] - rstep * tmax(1.0, Num<double>::abs(ub[i]));
    if (x[i] < lb[i] || x[i] > ub[i]) x[i] = 0.5 * (lb[i] + ub[i]);
  }
}
// common.py: CL_scaling_vector
KM_HD void cl_scaling(const double* x, const double* g, const double* lb, const double* ub, int n, double* v, double* dv) {
  for (int i = 0; i < n; i++) {
    v[i] = 1.0; dv[i] = 0.0;
    if (g[i] < 0.0) { v[i] = ub[i] - x[i]; dv[i] = -1.0; }
    if (g[i] > 0.0) { v[i] = x[i] - lb[i]; dv[i] = 1.0; }
  }
}
// common.py: minimize_quadratic_1d -- min of a t^2 + b t + c on [lo, hi] (first minimum in the order lo, hi, extremum)
KM_HD void minimize_quadratic_1d(double a, double b, double lo, double hi, double c, double* t_out, double* y_out) {
  double tb = lo, yb = lo * (a * lo + b) + c;
  const double yh = hi * (a * hi + b) + c;
  if (yh < yb) { tb = hi; yb = yh; }
  if (a != 0.0) {
    const double ex = -0.5 * b / a;
    if (lo < ex && ex < hi) {
      const double ye = ex * (a * ex + b) + c;
      if (ye < yb) { tb = ex; yb = ye; }
    }
  }
  *t_out = tb; *y_out = yb;
}
// symmetric eigen-decomposition by cyclic Jacobi: A (destroyed) -> eigenvalues w (descending), eigenvectors in columns of V
KM_HD void eig_sym(double (*A)[NMAX], int n, double* w, double (*V)[NMAX]) {
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) V[i][j] = i == j ? 1.0 : 0.0;
  // Convergence: the off-diagonal mass falls quadratically until it reaches the rounding floor of the rotations
  // (off ~ 1e-32 diag); stop at 1e-30, or once a sweep no longer reduces an already negligible remainder.  (A threshold
  // below the floor never triggers and every decomposition runs all 30 sweeps: the slowest envs of a batch spent 90 % of
  // their step there.)
  double prev_off = INF;
  for (int sweep = 0; sweep < 30; sweep++) {
    double off = 0, diag = 0;
    for (int i = 0; i < n; i++) { diag += A[i][i] * A[i][i]; for (int j = 0; j < i; j++) off += A[i][j] * A[i][j]; }
    if (off <= 1e-30 * diag || off == 0.0 || (off <= 1e-24 * diag && off > 0.25 * prev_off)) break;
    prev_off = off;
    for (int p = 0; p < n - 1; p++)
      for (int q = p + 1; q < n; q++) {
        const double apq = A[p][q];
        if (apq == 0.0) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (Num<double>::abs(theta) + Num<double>::sqrt(theta * theta + 1.0));
        const double c = 1.0 / Num<double>::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; k++) {
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; k++) {
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; k++) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < n; i++) w[i] = A[i][i];
  for (int i = 0; i < n - 1; i++) {   // sort descending (selection sort, columns follow)
    int best = i;
    for (int j = i + 1; j < n; j++) if (w[j] > w[best]) best = j;
    if (best != i) {
      const double tw = w[i]; w[i] = w[best]; w[best] = tw;
      for (int k = 0; k < n; k++) { const double tv = V[k][i]; V[k][i] = V[k][best]; V[k][best] = tv; }
    }
  }
}
// common.py: solve_lsq_trust_region with s = singular values, suf = s * uf (= V^T g_h), V; m = number of residuals
KM_HD void solve_lsq_trust_region(int n, int mres, const double* suf, const double* s, const double (*V)[NMAX], double Delta,
                                  double* alpha_io, double* p, double* t) {
  bool full_rank = false;
  if (mres >= n) full_rank = s[n - 1] > EPS * mres * s[0];
  if (full_rank) {
    for (int i = 0; i < n; i++) t[i] = suf[i] / (s[i] * s[i]);
    for (int i = 0; i < n; i++) { double r = 0; for (int j = 0; j < n; j++) r += V[i][j] * t[j]; p[i] = -r; }
    if (norm(p, n) <= Delta) { *alpha_io = 0.0; return; }
  }
  double alpha_upper = norm(suf, n) / Delta, alpha_lower = 0.0;
  auto phi_and_derivative = [&](double alpha, double* phi, double* phi_prime) {
    double pn = 0, dsum = 0;
    for (int i = 0; i < n; i++) {
      const double den = s[i] * s[i] + alpha, r = suf[i] / den;
      pn += r * r;
      dsum += suf[i] * suf[i] / (den * den * den);
    }
    pn = Num<double>::sqrt(pn);
    *phi = pn - Delta;
    *phi_prime = -dsum / pn;
  };
  if (full_rank) {
    double phi, phi_prime;
    phi_and_derivative(0.0, &phi, &phi_prime);
    alpha_lower = -phi / phi_prime;
  }
  double alpha = *alpha_io;
  if (!full_rank && alpha == 0.0) alpha = tmax(0.001 * alpha_upper, Num<double>::sqrt(alpha_lower * alpha_upper));
  for (int it = 0; it < 10; it++) {
    if (alpha < alpha_lower || alpha > alpha_upper) alpha = tmax(0.001 * alpha_upper, Num<double>::sqrt(alpha_lower * alpha_upper));
    double phi, phi_prime;
    phi_and_derivative(alpha, &phi, &phi_prime);
    if (phi < 0) alpha_upper = alpha;
    const double ratio = phi / phi_prime;
    alpha_lower = tmax(alpha_lower, alpha - ratio);
    alpha -= (phi + Delta) * ratio / Delta;
    if (Num<double>::abs(phi) < 0.01 * Delta) break;
  }
  for (int i = 0; i < n; i++) t[i] = suf[i] / (s[i] * s[i] + alpha);
  for (int i = 0; i < n; i++) { double r = 0; for (int j = 0; j < n; j++) r += V[i][j] * t[j]; p[i] = -r; }
  const double sc = Delta / norm(p, n);
  for (int i = 0; i < n; i++) p[i] *= sc;
  *alpha_io = alpha;
}
}  // namespace trf

// reference ik() with the restated scipy TRF.  In: b.x = x0 (current masked joints), b.qprev, b.lo, b.hi, b.goal.
// Out: b.x = result.x, b.xn = the last point the residual / Jacobian were evaluated at (where the reference leaves
// qpos[mask]).  Serial: call from one lane.
template <class S, typename T, class E, class B> KM_HD void ik_trf_serial(E& e, B& b, const Model<S, T>& m, int a, trf::TrfWork& wk) {
  using namespace trf;
  typedef Num<double> N;
  Grp<1> g1;
  g1.lane = 0; g1.mask = 1u; g1.wmask = 1u;
  const int n = m.arm_nmask[a], mres = 6 + 2 * n;
  const double ftol = 1e-8, xtol = 1e-8, gtol = 1e-8, reg = 9e-3;   // IK_JAC_REG rows of ik_jac
  const int max_nfev = 100 * n;
  const double* lb = b.lo;
  const double* ub = b.hi;
  double (&x)[NMAX] = wk.x, (&f)[6 + 2 * NMAX] = wk.f, (&g)[NMAX] = wk.g, (&JtJ)[NMAX][NMAX] = wk.JtJ;
  auto eval_f = [&](const double* xx, double* ff) {
    ik_chain_fk<S, T, 1>(e, b, m, a, xx);
    ik_residual<S, T, 1>(e, b, m, g1, a, xx, ff);
    for (int i = 0; i < n; i++) b.xn[i] = xx[i];
  };
  auto eval_jac = [&](const double* xx, const double* ff) {   // b.J pose rows, g = J^T f, JtJ = J^T J
    ik_chain_fk<S, T, 1>(e, b, m, a, xx);
    ik_jacobian<S, T, 1>(e, b, m, g1, a);
    for (int i = 0; i < n; i++) b.xn[i] = xx[i];
    for (int i = 0; i < n; i++) {
      double gi = 0;
      for (int k = 0; k < 6; k++) gi += b.J[k][i] * ff[k];
      g[i] = gi + reg * ff[6 + i] + reg * ff[6 + n + i];
      for (int j = 0; j < n; j++) {
        double sacc = 0;
        for (int k = 0; k < 6; k++) sacc += b.J[k][i] * b.J[k][j];
        JtJ[i][j] = sacc + (i == j ? 2.0 * reg * reg : 0.0);
      }
    }
  };
  for (int i = 0; i < n; i++) x[i] = b.x[i];
  make_strictly_feasible(x, lb, ub, n, 1e-10);           // least_squares.py
  double (&x0n)[NMAX] = wk.x0n;
  for (int i = 0; i < n; i++) x0n[i] = x[i];
  eval_f(x, f);
  int nfev = 1;
  eval_jac(x, f);
  double cost = 0.5 * dot(f, f, mres);
  double (&v)[NMAX] = wk.v, (&dv)[NMAX] = wk.dv, (&d)[NMAX] = wk.d, (&diag_h)[NMAX] = wk.diag_h, (&g_h)[NMAX] = wk.g_h;
  double (&Bh)[NMAX][NMAX] = wk.Bh, (&Aw)[NMAX][NMAX] = wk.Aw, (&V)[NMAX][NMAX] = wk.V;
  double (&w)[NMAX] = wk.w, (&s)[NMAX] = wk.s, (&suf)[NMAX] = wk.suf, (&t0)[NMAX] = wk.t0;
  cl_scaling(x, g, lb, ub, n, v, dv);
  for (int i = 0; i < n; i++) t0[i] = x0n[i] / N::sqrt(v[i]);
  double Delta = norm(t0, n);
  if (Delta == 0.0) Delta = 1.0;
  double alpha = 0.0;
  int termination = 0;
  while (true) {
    cl_scaling(x, g, lb, ub, n, v, dv);
    double g_norm = 0;
    for (int i = 0; i < n; i++) g_norm = tmax(g_norm, N::abs(g[i] * v[i]));
    if (g_norm < gtol) termination = 1;
    if (termination != 0 || nfev == max_nfev) break;
    for (int i = 0; i < n; i++) { d[i] = N::sqrt(v[i]); diag_h[i] = g[i] * dv[i]; g_h[i] = d[i] * g[i]; }
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) {
        Bh[i][j] = d[i] * JtJ[i][j] * d[j];
        Aw[i][j] = Bh[i][j] + (i == j ? diag_h[i] : 0.0);
      }
    // Gauss-Newton step p = -A^{-1} g_h by Cholesky.  When A is safely positive definite and the step lies inside the
    // trust region this is what solve_lsq_trust_region returns (full-rank branch, alpha = 0) and the eigen-decomposition
    // is never needed; otherwise (trust region active, or A close to singular) the decomposition is computed lazily.
    double (&p_gn)[NMAX] = wk.p_gn, (&Lc)[NMAX][NMAX] = wk.Lc;
    bool gn_ok = true, have_eig = false;
    {
      double dmax = 0;
      for (int i = 0; i < n; i++) dmax = tmax(dmax, Aw[i][i]);
      for (int j = 0; j < n && gn_ok; j++) {
        double dj = Aw[j][j];
        for (int k = 0; k < j; k++) dj -= Lc[j][k] * Lc[j][k];
        if (!(dj > 1e-10 * dmax)) { gn_ok = false; break; }      // far above the EPS * m * s_max rank threshold of scipy
        Lc[j][j] = N::sqrt(dj);
        for (int i = j + 1; i < n; i++) {
          double sacc = Aw[i][j];
          for (int k = 0; k < j; k++) sacc -= Lc[i][k] * Lc[j][k];
          Lc[i][j] = sacc / Lc[j][j];
        }
      }
      if (gn_ok) {
        for (int i = 0; i < n; i++) { double t = -g_h[i]; for (int k = 0; k < i; k++) t -= Lc[i][k] * p_gn[k]; p_gn[i] = t / Lc[i][i]; }
        for (int i = n - 1; i >= 0; i--) { double t = p_gn[i]; for (int k = i + 1; k < n; k++) t -= Lc[k][i] * p_gn[k]; p_gn[i] = t / Lc[i][i]; }
      }
    }
    const double theta = tmax(0.995, 1.0 - g_norm);
    double actual_reduction = -1.0, cost_new = cost;
    double (&x_new)[NMAX] = wk.x_new, (&f_new)[6 + 2 * NMAX] = wk.f_new;
    while (actual_reduction <= 0.0 && nfev < max_nfev) {
      double (&p_h)[NMAX] = wk.p_h, (&p)[NMAX] = wk.p, (&step)[NMAX] = wk.step, (&step_h)[NMAX] = wk.step_h, predicted;
      if (gn_ok && norm(p_gn, n) <= Delta) {
        for (int i = 0; i < n; i++) p_h[i] = p_gn[i];
        alpha = 0.0;
      } else {
        if (!have_eig) {
          have_eig = true;
          eig_sym(Aw, n, w, V);
          for (int i = 0; i < n; i++) {
            s[i] = N::sqrt(tmax(w[i], 0.0));
            double r = 0;
            for (int k = 0; k < n; k++) r += V[k][i] * g_h[k];
            suf[i] = r;
          }
        }
        solve_lsq_trust_region(n, mres, suf, s, V, Delta, &alpha, p_h, wk.t);
      }
      for (int i = 0; i < n; i++) p[i] = d[i] * p_h[i];
      // ---- trf.py: select_step
      {
        double (&xp)[NMAX] = wk.xp;
        for (int i = 0; i < n; i++) xp[i] = x[i] + p[i];
        auto evalq = [&](const double* sv) {   // evaluate_quadratic(J_h, g_h, s, diag_h)
          double q = quad(Bh, sv, sv, n);
          for (int i = 0; i < n; i++) q += sv[i] * diag_h[i] * sv[i];
          return 0.5 * q + dot(sv, g_h, n);
        };
        if (in_bounds(xp, lb, ub, n)) {
          for (int i = 0; i < n; i++) { step[i] = p[i]; step_h[i] = p_h[i]; }
          predicted = -evalq(p_h);
        } else {
          int (&hits)[NMAX] = wk.hits;
          const double p_stride = step_size_to_bound(x, p, lb, ub, n, hits, wk.steps);
          double (&r_h)[NMAX] = wk.r_h, (&r)[NMAX] = wk.r, (&x_on_bound)[NMAX] = wk.x_on_bound;
          for (int i = 0; i < n; i++) { r_h[i] = hits[i] != 0 ? -p_h[i] : p_h[i]; r[i] = d[i] * r_h[i]; }
          for (int i = 0; i < n; i++) { p[i] *= p_stride; p_h[i] *= p_stride; x_on_bound[i] = x[i] + p[i]; }
          // intersect_trust_region(p_h, r_h, Delta): positive root
          double to_tr;
          {
            const double qa = dot(r_h, r_h, n), qb = dot(p_h, r_h, n), qc = dot(p_h, p_h, n) - Delta * Delta;
            const double dd = N::sqrt(qb * qb - qa * qc);
            const double q = -(qb + (qb >= 0 ? dd : -dd));   // copysign(d, b); b = +0.0 gives +d as in numpy
            const double t1 = q / qa, t2 = qc / q;
            to_tr = t1 < t2 ? t2 : t1;
          }
          const double to_bound = step_size_to_bound(x_on_bound, r, lb, ub, n, nullptr, wk.steps);
          double r_stride = tmin(to_bound, to_tr), r_stride_l, r_stride_u;
          if (r_stride > 0) {
            r_stride_l = (1.0 - theta) * p_stride / r_stride;
            r_stride_u = r_stride == to_bound ? theta * to_bound : to_tr;
          } else { r_stride_l = 0; r_stride_u = -1; }
          double r_value = INF;
          if (r_stride_l <= r_stride_u) {
            // build_quadratic_1d(J_h, g_h, r_h, s0 = p_h, diag = diag_h)
            double qa = quad(Bh, r_h, r_h, n), qb = dot(g_h, r_h, n) + quad(Bh, r_h, p_h, n);
            double qc = 0.5 * quad(Bh, p_h, p_h, n) + dot(g_h, p_h, n);
            for (int i = 0; i < n; i++) { qa += r_h[i] * diag_h[i] * r_h[i]; qb += p_h[i] * diag_h[i] * r_h[i]; qc += 0.5 * p_h[i] * diag_h[i] * p_h[i]; }
            qa *= 0.5;
            minimize_quadratic_1d(qa, qb, r_stride_l, r_stride_u, qc, &r_stride, &r_value);
            for (int i = 0; i < n; i++) { r_h[i] = r_h[i] * r_stride + p_h[i]; r[i] = r_h[i] * d[i]; }
          }
          for (int i = 0; i < n; i++) { p[i] *= theta; p_h[i] *= theta; }
          const double p_value = evalq(p_h);
          double (&ag_h)[NMAX] = wk.ag_h, (&ag)[NMAX] = wk.ag;
          for (int i = 0; i < n; i++) { ag_h[i] = -g_h[i]; ag[i] = d[i] * ag_h[i]; }
          const double to_tr2 = Delta / norm(ag_h, n);
          const double to_bound2 = step_size_to_bound(x, ag, lb, ub, n, nullptr, wk.steps);
          double ag_stride = to_bound2 < to_tr2 ? theta * to_bound2 : to_tr2, ag_value;
          double qa = quad(Bh, ag_h, ag_h, n), qb = dot(g_h, ag_h, n);
          for (int i = 0; i < n; i++) qa += ag_h[i] * diag_h[i] * ag_h[i];
          qa *= 0.5;
          minimize_quadratic_1d(qa, qb, 0.0, ag_stride, 0.0, &ag_stride, &ag_value);
          for (int i = 0; i < n; i++) { ag_h[i] *= ag_stride; ag[i] *= ag_stride; }
          if (p_value < r_value && p_value < ag_value) {
            for (int i = 0; i < n; i++) { step[i] = p[i]; step_h[i] = p_h[i]; }
            predicted = -p_value;
          } else if (r_value < p_value && r_value < ag_value) {
            for (int i = 0; i < n; i++) { step[i] = r[i]; step_h[i] = r_h[i]; }
            predicted = -r_value;
          } else {
            for (int i = 0; i < n; i++) { step[i] = ag[i]; step_h[i] = ag_h[i]; }
            predicted = -ag_value;
          }
        }
      }
      for (int i = 0; i < n; i++) x_new[i] = x[i] + step[i];
      make_strictly_feasible(x_new, lb, ub, n, 0.0);
      eval_f(x_new, f_new);
      nfev++;
      const double step_h_norm = norm(step_h, n);
      bool finite = true;
      for (int i = 0; i < mres; i++) finite = finite && (f_new[i] - f_new[i] == 0.0);
      if (!finite) { Delta = 0.25 * step_h_norm; continue; }
      cost_new = 0.5 * dot(f_new, f_new, mres);
      actual_reduction = cost - cost_new;
      // update_tr_radius
      double ratio;
      if (predicted > 0) ratio = actual_reduction / predicted;
      else if (predicted == 0.0 && actual_reduction == 0.0) ratio = 1;
      else ratio = 0;
      double Delta_new = Delta;
      if (ratio < 0.25) Delta_new = 0.25 * step_h_norm;
      else if (ratio > 0.75 && step_h_norm > 0.95 * Delta) Delta_new = Delta * 2.0;
      const double step_norm = norm(step, n);
      // check_termination
      const bool ftol_ok = actual_reduction < ftol * cost && ratio > 0.25;
      const bool xtol_ok = step_norm < xtol * (xtol + norm(x, n));
      termination = (ftol_ok && xtol_ok) ? 4 : (ftol_ok ? 2 : (xtol_ok ? 3 : 0));
      if (termination != 0) break;
      alpha *= Delta / Delta_new;
      Delta = Delta_new;
    }
    if (actual_reduction > 0.0) {
      for (int i = 0; i < n; i++) x[i] = x_new[i];
      for (int i = 0; i < mres; i++) f[i] = f_new[i];
      cost = cost_new;
      eval_jac(x, f);
    }
  }
  for (int i = 0; i < n; i++) b.x[i] = x[i];
  e.ik_evals += nfev;
}

}  // namespace km
