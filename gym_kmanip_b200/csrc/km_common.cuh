// km_common.cuh -- scalar traits, small vector/quaternion algebra and the lane-group primitives.
//
// Execution model of the whole simulator: one *group* of G lanes (G = 16 or 32, all inside one warp, one lane per dof)
// owns one environment whose working set lives in shared memory.  Code is written as bulk-synchronous
// phases: `KM_FOR(i, n)` distributes n independent items over the lanes, `g.sync()` separates phases,
// `g.sum()` is a butterfly reduction that leaves the bit-identical result in every lane, so scalar
// control flow (Newton iterations, line search) stays uniform inside a group.
//
// Everything here also compiles with a host compiler (g++) with G = 1; tests/hostsim uses that to run the
// identical kernel body on the CPU against the oracle while no GPU is attached.  That host build is a
// debugging aid for tests only, never a product path.
#pragma once
#include <cmath>
#include <cstdint>

#include <type_traits>
#if defined(__CUDACC__)
#define KM_HD __host__ __device__ __forceinline__
#define KM_FN __host__ __device__ __noinline__
#define KM_DI __device__ __forceinline__
#define KM_DN __device__ __noinline__
#else
#define KM_HD inline
#define KM_FN
#define KM_DI inline
#define KM_DN
#endif
// Warp-collective code paths: the device, or the host emulation of one warp that tests/hostsim/warpemu.h provides (it
// defines KM_WARP_EMU and the __shfl_sync / __ballot_sync / __syncwarp ... intrinsics before including this file)
#if defined(__CUDA_ARCH__) || defined(KM_WARP_EMU)
#define KM_WARP_CODE 1
#else
#define KM_WARP_CODE 0
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
  // sin/cos for |x| up to a few turns (half joint angles, Euler half angles): quadrant reduction with a
  // two-term Cody-Waite pi/2 and the usual degree-7/8 minimax kernels; ~1 ulp, no large-argument slow path
  // (the library sincosf drags a Payne-Hanek path with a local-memory table into every call site).
  static KM_HD void sincos(float x, float* s, float* c) {
#if defined(__CUDA_ARCH__)
    const float q = rintf(x * 0.636619772367581343f);
    float r = fmaf(q, -1.57079601287841796875f, x);
    r = fmaf(q, -3.1391647326017846e-7f, r);
    r = fmaf(q, -5.3903025299577648e-15f, r);
    const int n = (int)q;
    const float r2 = r * r;
    float sp = fmaf(r2, -1.95152959e-4f, 8.33216087e-3f);
    sp = fmaf(sp, r2, -1.66666546e-1f);
    sp = fmaf(sp * r2, r, r);
    float cp = fmaf(r2, 2.44331571e-5f, -1.38873163e-3f);
    cp = fmaf(cp, r2, 4.16666457e-2f);
    cp = fmaf(cp, r2, -0.5f);
    cp = fmaf(cp, r2, 1.0f);
    const float ss = (n & 1) ? cp : sp, cc = (n & 1) ? sp : cp;
    *s = (n & 2) ? -ss : ss;
    *c = ((n + 1) & 2) ? -cc : cc;
#else
    *s = sinf(x); *c = cosf(x);
#endif
  }
  static KM_HD float rsqrt(float x) {
#if defined(__CUDA_ARCH__)
    return rsqrtf(x);
#else
    return 1.0f / sqrtf(x);
#endif
  }
  static KM_HD float eps() { return 1.1920929e-7f; }
  static KM_HD float minval() { return 1e-15f; }
  static KM_HD float huge() { return 3.0e38f; }
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
  static KM_HD double rsqrt(double x) { return 1.0 / ::sqrt(x); }
  static KM_HD double eps() { return 2.220446049250313e-16; }
  static KM_HD double minval() { return 1e-15; }
  static KM_HD double huge() { return 1.0e300; }
};
template <typename T> KM_HD T tmax(T a, T b) { return a > b ? a : b; }
template <typename T> KM_HD T tmin(T a, T b) { return a < b ? a : b; }
template <typename T> KM_HD T tclip(T x, T lo, T hi) { return x < lo ? lo : (x > hi ? hi : x); }

KM_HD int popc(unsigned x) {
#if defined(__CUDA_ARCH__)
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}
KM_HD int ffs1(unsigned x) {   // 1-based index of the lowest set bit, 0 if none
#if defined(__CUDA_ARCH__)
  return __ffs((int)x);
#else
  return __builtin_ffs((int)x);
#endif
}

// ---------------------------------------------------------------------------------------- lane group
template <int G> struct Grp {
  int lane;        // 0..G-1 inside the group
  unsigned mask;   // the group's lanes inside its warp
  unsigned wmask;  // lanes to reconverge with: the whole warp while every group of the warp runs an env, else `mask`
  // Groups sharing a warp run data-dependent loops (Newton, line search) and would otherwise stay diverged,
  // halving the issue efficiency of everything that follows; converge() is placed after such loops.
  // a whole-warp group names its lanes with a literal mask: the compiler then emits bare SHFL / no WARPSYNC
  KM_HD unsigned lanes() const { return G == 32 ? 0xffffffffu : mask; }
  // G == 1 is the thread-per-env mapping: every thread owns an env, so all group collectives are identities and
  // nothing may synchronise with other threads (their control flow is independent).
  KM_HD void converge() const {
#if KM_WARP_CODE
    if (G > 1 && G < 32) __syncwarp(wmask);
#endif
  }
  template <typename T> KM_HD T shfl(T v, int src) const {
#if KM_WARP_CODE
    if (G == 1) return v;
    return __shfl_sync(lanes(), v, src, G);
#else
    return v;
#endif
  }
  // CTA-wide phase alignment.  The envs of a CTA are independent, yet marching them through the big phases of a
  // sub-step together makes the warps of an SM fetch the same instructions at the same time: the hot loop is far
  // larger than the 32 KB L1.5 instruction cache, and unaligned warps spent over half their stall time waiting for
  // instruction fetch.  Only called from code every thread of the CTA executes the same number of times.
  // KM_LOCKSTEP: 2 = phases and Newton iterations CTA-wide (generic solver only: every warp of the CTA must run it),
  // 1 = phases only (default: the register-resident solver of km_solver_warp.cuh is small enough for the cache), 0 = warps run free
#ifndef KM_LOCKSTEP
#define KM_LOCKSTEP 1
#endif
  // (thread-per-env groups set wmask = ~0u to ask for the phase alignment: there every thread of the CTA runs the
  // same number of env steps, shadowing a valid env where the batch ends)
  KM_HD void cta_sync() const {
#if defined(__CUDA_ARCH__)
    if (G > 1 ? KM_LOCKSTEP >= 1 : wmask == 0xffffffffu) __syncthreads();
#elif defined(KM_WARP_EMU)
    if (G > 1) __syncwarp(0xffffffffu);   // the emulated CTA is one warp
#endif
  }
  KM_HD bool cta_any(bool p) const {
#if defined(__CUDA_ARCH__)
    if (G == 1) return p;
    if (KM_LOCKSTEP < 2) return (__ballot_sync(0xffffffffu, p) != 0u);
    return __syncthreads_or(p) != 0;
#elif defined(KM_WARP_EMU)
    if (G == 1) return p;
    return __ballot_sync(0xffffffffu, p) != 0u;
#else
    return p;
#endif
  }
  KM_HD void sync() const {
#if KM_WARP_CODE
    if (G > 1) __syncwarp(lanes());
#endif
  }
  template <typename T> KM_HD T sum(T v) const {
#if KM_WARP_CODE
    if (G > 1) {
#pragma unroll
      for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(lanes(), v, o, G);
    }
#endif
    return v;
  }
  // bit i set: lane i of the group voted true
  KM_HD unsigned ballot(bool p) const {
#if KM_WARP_CODE
    if (G == 1) return p ? 1u : 0u;
    if (G == 32) return __ballot_sync(0xffffffffu, p);
    const unsigned b = __ballot_sync(mask, p) & mask;
    return b >> (ffs1(mask) - 1);
#else
    return p ? 1u : 0u;
#endif
  }
  KM_HD bool any(bool p) const {
#if KM_WARP_CODE
    if (G == 1) return p;
    return (__ballot_sync(lanes(), p) & lanes()) != 0u;
#else
    return p;
#endif
  }
};
#define KM_FOR(i, n) for (int i = g.lane; i < (n); i += G)
// Debug build (-DKM_PHASE_CLOCKS): lane 0 charges the cycles since the previous mark to phase `id` (km_debug_phase_clocks)
#if defined(KM_PHASE_CLOCKS) && defined(__CUDA_ARCH__)
#define KM_CLK(id) do { if (g.lane == 0) { const unsigned t__ = (unsigned)clock(); e.clk[id] += t__ - e.clk_last; e.clk_last = t__; } } while (0)
#else
#define KM_CLK(id) do {} while (0)
#endif
enum { CLK_KIN = 0, CLK_CRB, CLK_COLL, CLK_VEL, CLK_ACC, CLK_SOL_SETUP, CLK_SOL_DIR, CLK_SOL_LS, CLK_SOL_UPD, CLK_SOL_VOTE, CLK_EULER,
       CLK_BARRIER, CLK_BEFORE, CLK_EPILOGUE, CLK_N };

// compile-time loops (bodies receive std::integral_constant so indices can select registers / if constexpr)
template <int I> using IC = std::integral_constant<int, I>;
template <int I, int N, class F> KM_HD void sfor(F&& f) {
  if constexpr (I < N) { f(IC<I>{}); sfor<I + 1, N>(static_cast<F&&>(f)); }
}
template <int N, class F> KM_HD void sfor_rev(F&& f) {
  if constexpr (N > 0) { f(IC<N - 1>{}); sfor_rev<N - 1>(static_cast<F&&>(f)); }
}

// ---------------------------------------------------------------------------------------- 3-vectors, quaternions (wxyz)
template <typename T> KM_HD T dot3(const T* a, const T* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <typename T> KM_HD void cross3(T* r, const T* a, const T* b) {
  T x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
// normalise in place, return the norm; degenerate vectors become (1,0,0) (mju_normalize3)
template <typename T> KM_HD T normalize3(T* a) {
  T n = Num<T>::sqrt(dot3(a, a));
  if (n < Num<T>::minval()) { a[0] = 1; a[1] = 0; a[2] = 0; }
  else { T s = T(1) / n; a[0] *= s; a[1] *= s; a[2] *= s; }
  return n;
}
template <typename T> KM_HD void qnormalize(T* q) {
  T n = Num<T>::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < Num<T>::minval()) { q[0] = 1; q[1] = 0; q[2] = 0; q[3] = 0; }
  else { T s = T(1) / n; q[0] *= s; q[1] *= s; q[2] *= s; q[3] *= s; }
}
template <typename T> KM_HD void qmul(T* r, const T* a, const T* b) {
  T t0 = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  T t1 = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  T t2 = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  T t3 = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = t0; r[1] = t1; r[2] = t2; r[3] = t3;
}
// r = q v q^-1 for a unit quaternion: v + 2 u x (u x v + w v)
template <typename T> KM_HD void qrot(T* r, const T* q, const T* v) {
  const T tx = q[2] * v[2] - q[3] * v[1] + q[0] * v[0], ty = q[3] * v[0] - q[1] * v[2] + q[0] * v[1], tz = q[1] * v[1] - q[2] * v[0] + q[0] * v[2];
  const T x = v[0] + T(2) * (q[2] * tz - q[3] * ty), y = v[1] + T(2) * (q[3] * tx - q[1] * tz), z = v[2] + T(2) * (q[1] * ty - q[2] * tx);
  r[0] = x; r[1] = y; r[2] = z;
}
// rotation matrix (row-major) of a unit quaternion (mju_quat2Mat)
template <typename T> KM_HD void q2mat(T* m, const T* q) {
  T q00 = q[0] * q[0], q01 = q[0] * q[1], q02 = q[0] * q[2], q03 = q[0] * q[3];
  T q11 = q[1] * q[1], q12 = q[1] * q[2], q13 = q[1] * q[3], q22 = q[2] * q[2], q23 = q[2] * q[3], q33 = q[3] * q[3];
  m[0] = q00 + q11 - q22 - q33; m[4] = q00 - q11 + q22 - q33; m[8] = q00 - q11 - q22 + q33;
  m[1] = T(2) * (q12 - q03); m[2] = T(2) * (q13 + q02);
  m[3] = T(2) * (q12 + q03); m[5] = T(2) * (q23 - q01);
  m[6] = T(2) * (q13 - q02); m[7] = T(2) * (q23 + q01);
}
template <typename T> KM_HD void mulv3(T* r, const T* m, const T* v) {
  T x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2], y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2],
    z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
template <typename T> KM_HD void mulTv3(T* r, const T* m, const T* v) {
  T x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2], y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2],
    z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
// mju_mat2Quat
template <typename T> KM_HD void mat2quat(T* q, const T* m) {
  typedef Num<T> N;
  if (m[0] + m[4] + m[8] > T(0)) {
    q[0] = T(0.5) * N::sqrt(T(1) + m[0] + m[4] + m[8]);
    q[1] = T(0.25) * (m[7] - m[5]) / q[0]; q[2] = T(0.25) * (m[2] - m[6]) / q[0]; q[3] = T(0.25) * (m[3] - m[1]) / q[0];
  } else if (m[0] > m[4] && m[0] > m[8]) {
    q[1] = T(0.5) * N::sqrt(T(1) + m[0] - m[4] - m[8]);
    q[0] = T(0.25) * (m[7] - m[5]) / q[1]; q[2] = T(0.25) * (m[1] + m[3]) / q[1]; q[3] = T(0.25) * (m[2] + m[6]) / q[1];
  } else if (m[4] > m[8]) {
    q[2] = T(0.5) * N::sqrt(T(1) - m[0] + m[4] - m[8]);
    q[0] = T(0.25) * (m[2] - m[6]) / q[2]; q[1] = T(0.25) * (m[1] + m[3]) / q[2]; q[3] = T(0.25) * (m[5] + m[7]) / q[2];
  } else {
    q[3] = T(0.5) * N::sqrt(T(1) - m[0] - m[4] + m[8]);
    q[0] = T(0.25) * (m[3] - m[1]) / q[3]; q[1] = T(0.25) * (m[2] + m[6]) / q[3]; q[2] = T(0.25) * (m[5] + m[7]) / q[3];
  }
  qnormalize(q);
}
// mju_subQuat: rotation vector taking qb to qa, in qb's frame
template <typename T> KM_HD void subquat(T* res, const T* qa, const T* qb) {
  T qneg[4] = {qb[0], -qb[1], -qb[2], -qb[3]}, qd[4];
  qmul(qd, qneg, qa);
  T axis[3] = {qd[1], qd[2], qd[3]};
  T s = normalize3(axis);
  T speed = T(2) * Num<T>::atan2(s, qd[0]);
  if (speed > T(3.14159265358979323846)) speed -= T(2.0 * 3.14159265358979323846);
  res[0] = axis[0] * speed; res[1] = axis[1] * speed; res[2] = axis[2] * speed;
}
// mju_makeFrame: complete an orthonormal frame whose first row (the contact normal) is given, second row zero
template <typename T> KM_HD void makeframe(T* f) {
  normalize3(f);
  T* y = f + 3;
  y[0] = 0; y[1] = 0; y[2] = 0;
  if (f[1] < T(0.5) && f[1] > T(-0.5)) y[1] = 1; else y[2] = 1;
  T t = dot3(f, y);
  y[0] -= t * f[0]; y[1] -= t * f[1]; y[2] -= t * f[2];
  normalize3(y);
  cross3(f + 6, f, y);
}
// spatial inertia about a point: 10-vector [Ixx Iyy Izz Ixy Ixz Iyz, m*off(3), m]; 6-vectors are [angular; linear]
template <typename T> KM_HD void inert_com(T* res, const T* inert, const T* mat, const T* dif, T mass) {
  T tmp[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      tmp[3 * i + j] = mat[3 * i] * inert[0] * mat[3 * j] + mat[3 * i + 1] * inert[1] * mat[3 * j + 1] +
                       mat[3 * i + 2] * inert[2] * mat[3 * j + 2];
  res[0] = tmp[0] + mass * (dif[1] * dif[1] + dif[2] * dif[2]);
  res[1] = tmp[4] + mass * (dif[0] * dif[0] + dif[2] * dif[2]);
  res[2] = tmp[8] + mass * (dif[0] * dif[0] + dif[1] * dif[1]);
  res[3] = tmp[1] - mass * dif[0] * dif[1];
  res[4] = tmp[2] - mass * dif[0] * dif[2];
  res[5] = tmp[5] - mass * dif[1] * dif[2];
  res[6] = mass * dif[0]; res[7] = mass * dif[1]; res[8] = mass * dif[2]; res[9] = mass;
}
template <typename T> KM_HD void mul_inert_vec(T* r, const T* i, const T* v) {
  T a0 = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
  T a1 = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
  T a2 = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
  T a3 = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
  T a4 = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
  T a5 = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
  r[0] = a0; r[1] = a1; r[2] = a2; r[3] = a3; r[4] = a4; r[5] = a5;
}
template <typename T> KM_HD void cross_motion(T* r, const T* vel, const T* v) {
  T a[3], b[3], c[3];
  cross3(a, vel, v);
  cross3(b, vel, v + 3);
  cross3(c, vel + 3, v);
  r[0] = a[0]; r[1] = a[1]; r[2] = a[2];
  r[3] = b[0] + c[0]; r[4] = b[1] + c[1]; r[5] = b[2] + c[2];
}
template <typename T> KM_HD void cross_force(T* r, const T* vel, const T* f) {
  T a[3], b[3], c[3];
  cross3(a, vel, f);
  cross3(b, vel + 3, f + 3);
  cross3(c, vel, f + 3);
  r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2];
  r[3] = c[0]; r[4] = c[1]; r[5] = c[2];
}

// Philox4x32-10: the cube-spawn generator, keyed by (seed, global env id, episode) so that results do not
// depend on how envs are sharded across GPUs.
KM_HD void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
KM_HD void spawn_uniforms(uint64_t seed, uint64_t env_id, uint32_t episode, double u[3]) {
  uint32_t c[4] = {(uint32_t)env_id, (uint32_t)(env_id >> 32), episode, 0x4b4d414eu};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  for (int i = 0; i < 3; i++) u[i] = (double)(c[i] >> 8) * (1.0 / 16777216.0);
}

}  // namespace km
