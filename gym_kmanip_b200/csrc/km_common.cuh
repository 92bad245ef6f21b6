// km_common.cuh -- scalar helpers, compile-time loops and the run-time parameter block.
//
// Everything here is usable from device code (nvcc, sm_100a) and from a host compiler (g++), the latter
// only for tests/hostsim (a CPU emulation of one CUDA thread used to debug the kernel body without a GPU;
// it is not part of the product path).
#pragma once
#include <cmath>
#include <cstdint>
#include <type_traits>

#if defined(__CUDACC__)
#define KM_HD __host__ __device__ __forceinline__
#define KM_HDN __host__ __device__ __noinline__
#else
#define KM_HD inline
#define KM_HDN inline
#endif

namespace km {

// ---------------------------------------------------------------------------------------- scalar traits
template <typename T> struct Num;
template <> struct Num<float> {
  static KM_HD float sqrt(float x) { return sqrtf(x); }
  static KM_HD float abs(float x) { return fabsf(x); }
  static KM_HD float atan2(float y, float x) { return atan2f(y, x); }
  static KM_HD float asin(float x) { return asinf(x); }
  static KM_HD float tan(float x) { return tanf(x); }
  static KM_HD float pow(float x, float y) { return powf(x, y); }
  static KM_HD void sincos(float x, float* s, float* c) {
#if defined(__CUDA_ARCH__)
    sincosf(x, s, c);
#else
    *s = sinf(x); *c = cosf(x);
#endif
  }
  static constexpr float eps = 1.1920929e-7f;
  static constexpr float minval = 1e-15f;
};
template <> struct Num<double> {
  static KM_HD double sqrt(double x) { return ::sqrt(x); }
  static KM_HD double abs(double x) { return ::fabs(x); }
  static KM_HD double atan2(double y, double x) { return ::atan2(y, x); }
  static KM_HD double asin(double x) { return ::asin(x); }
  static KM_HD double tan(double x) { return ::tan(x); }
  static KM_HD double pow(double x, double y) { return ::pow(x, y); }
  static KM_HD void sincos(double x, double* s, double* c) {
#if defined(__CUDA_ARCH__)
    ::sincos(x, s, c);
#else
    *s = ::sin(x); *c = ::cos(x);
#endif
  }
  static constexpr double eps = 2.220446049250313e-16;
  static constexpr double minval = 1e-15;
};
template <typename T> KM_HD T tmax(T a, T b) { return a > b ? a : b; }
template <typename T> KM_HD T tmin(T a, T b) { return a < b ? a : b; }
template <typename T> KM_HD T tclip(T x, T lo, T hi) { return x < lo ? lo : (x > hi ? hi : x); }

// ---------------------------------------------------------------------------------------- compile-time loops
template <int I> using IC = std::integral_constant<int, I>;
template <int I, int N, class F> KM_HD void sfor(F&& f) {
  if constexpr (I < N) { f(IC<I>{}); sfor<I + 1, N>(static_cast<F&&>(f)); }
}
// I = N-1 ... 0
template <int N, class F> KM_HD void sfor_rev(F&& f) {
  if constexpr (N > 0) { f(IC<N - 1>{}); sfor_rev<N - 1>(static_cast<F&&>(f)); }
}

// Topology queries evaluated at compile time.
template <class S> struct Topo {
  // dof j is i or an ancestor of i in the dof tree
  static constexpr bool is_anc(int j, int i) {
    while (i >= 0) { if (i == j) return true; i = S::dof_parent[i]; }
    return false;
  }
  // dof j moves body b
  static constexpr bool moves(int j, int b) { return S::body_lastdof[b] >= 0 && is_anc(j, S::body_lastdof[b]); }
  // articulated (non-cube) moving body
  static constexpr bool arm_body(int b) { return S::body_moving[b] && b != S::CUBE_BODY; }
  static constexpr bool subtree_has_mass(int b) {
    for (int c = b; c < S::NBODY; c++) {
      int a = c; bool in = false;
      while (a > 0) { if (a == b) { in = true; break; } a = S::body_parent[a]; }
      if (in && S::body_hasmass[c]) return true;
    }
    return false;
  }
  // a static body whose world pose some moving child needs
  static constexpr bool static_parent_needed(int b) {
    if (S::body_moving[b]) return false;
    for (int c = 1; c < S::NBODY; c++) if (S::body_parent[c] == b && S::body_moving[c] && c != S::CUBE_BODY) return true;
    return false;
  }
  static constexpr bool has_dof(int b) { return S::body_jtype[b] >= 0; }
  static constexpr bool fric_on(int d) {
    for (int k = 0; k < S::NFRIC; k++) if (S::fric_dof[k] == d) return true;
    return false;
  }
  static constexpr int fric_slot(int d) {
    for (int k = 0; k < S::NFRIC; k++) if (S::fric_dof[k] == d) return k;
    return -1;
  }
  // pad p's Jacobian touches articulated dof d
  static constexpr bool pad_dof(int p, int d) { return moves(d, S::pad_body[p]); }
  // dofs d and e can be coupled in the constraint Hessian (same chain ancestry, or linked through a pad to the cube)
  static constexpr int NCS = S::NPAD + 4;   // contact slots: one per pad, four for the table
};

// ---------------------------------------------------------------------------------------- run-time parameters
// Numeric model + task parameters of one scene, passed BY VALUE as the kernel argument (constant bank).
// Filled on the host by km_api (from the flat km_model) -- see fill_params() there.
template <class S, typename T> struct Params {
  // bodies: local frame in the parent for moving bodies, WORLD frame for static bodies
  T bpos[S::NBODY][3], brot[S::NBODY][9];
  T bmass[S::NBODY], bipos[S::NBODY][3], binertia[S::NBODY][3];
  // articulated joints (index == dof == qpos address)
  T jrange[S::NVA][2], lim_invw[S::NVA], lim_solref[S::NVA][2], lim_solimp[S::NVA][5];
  // position actuators
  T kp[S::NU], ctrl_lo[S::NU], ctrl_hi[S::NU], frc_lo[S::NU], frc_hi[S::NU];
  // friction-loss rows (constant: pos = 0): loss, R, D, B
  T fr_loss[S::NFRIC > 0 ? S::NFRIC : 1], fr_R[S::NFRIC > 0 ? S::NFRIC : 1], fr_D[S::NFRIC > 0 ? S::NFRIC : 1],
      fr_B[S::NFRIC > 0 ? S::NFRIC : 1];
  // finger pads (spheres) and their contact pairs with the cube
  T pad_pos[S::NPAD][3], pad_rad[S::NPAD];
  T pad_solref[S::NPAD][2], pad_solimp[S::NPAD][5], pad_mu[S::NPAD][3], pad_tran[S::NPAD], pad_rot[S::NPAD];
  // table plane z = tab_z and its pair with the cube
  T tab_z, tab_solref[2], tab_solimp[5], tab_mu[3], tab_tran, tab_rot;
  T cube_size[3];
  // options
  T h, grav[3], tol, ls_tol, meaninertia, impratio;
  int iterations, ls_iterations;
  // task
  T q_home[S::Q_LEN], spawn_lo[3], spawn_hi[3], cube_quat0[4], mocap0[S::NMOCAP * 7];
  int act_dim, off_pos[2], off_orn[2], off_grip[2], off_q[2], n_arm_act;
  int ik_iters, ik_teleport, max_episode_steps;
};

}  // namespace km
