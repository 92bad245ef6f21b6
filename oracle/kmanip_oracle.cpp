/* kmanip_oracle.cpp -- CPU restatement of the gym-kmanip env-step hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see ko_model.h header: who may load this, and "PARITY UNPINNED").
 *
 * What is restated, and from where:
 *   - reference gym_kmanip/env_sim.py:23-36   initialize_episode      -> ko_env_reset
 *   - reference gym_kmanip/env_sim.py:38-108  before_step             -> ko_before_step
 *   - reference gym_kmanip/env_sim.py:110-146 get_observation         -> ko_observation
 *   - reference gym_kmanip/env_sim.py:148-179 get_reward              -> ko_reward
 *   - reference gym_kmanip/ik_mujoco.py:20-53 ik_res, :56-97 ik_jac   -> ko_ik_residual / ko_ik_jacobian
 *   - reference gym_kmanip/ik_mujoco.py:100-155 ik                    -> ko_ik_dls (fixed-count damped
 *       Gauss-Newton; the reference's optimiser is scipy TRF -- oracle/oracle.py can drive the same
 *       residual/Jacobian with the real scipy.optimize.least_squares for comparison)
 *   - dm_control Physics.step legacy ordering mj_step2; mj_step x9; mj_step1 -> ko_env_step
 *   - MuJoCo engine (third party, not vendored; restated from its published algorithm):
 *       mj_kinematics, mj_comPos, mj_crb, mj_factorM, mj_collision (plane-box, sphere-box), mj_makeConstraint,
 *       mj_comVel, mj_rne, mj_referenceConstraint, mj_fwdActuation, mj_fwdAcceleration,
 *       mj_fwdConstraint (Newton, pyramidal cones, friction loss, limits), mj_Euler.
 *
 * Style: deliberately generic (loops over flat model arrays, dense matrices), the opposite of the
 * topology-specialised CUDA path, so that agreement between the two is informative.
 *
 * Scalar type: double, or (with -DKO_COUNT_FLOPS) a counting wrapper that tallies algorithmic FLOPs.
 */
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "ko_model.h"

/* ------------------------------------------------------------------------------------------- scalar */
#ifdef KO_COUNT_FLOPS
static thread_local unsigned long long g_flops = 0;
struct real {
  double v;
  real() : v(0) {}
  real(double x) : v(x) {}
  real(int x) : v(x) {}
  explicit operator double() const { return v; }
  explicit operator float() const { return (float)v; }
};
static inline real operator+(real a, real b) { g_flops++; return real(a.v + b.v); }
static inline real operator-(real a, real b) { g_flops++; return real(a.v - b.v); }
static inline real operator*(real a, real b) { g_flops++; return real(a.v * b.v); }
static inline real operator/(real a, real b) { g_flops++; return real(a.v / b.v); }
static inline real operator-(real a) { return real(-a.v); }
static inline real& operator+=(real& a, real b) { g_flops++; a.v += b.v; return a; }
static inline real& operator-=(real& a, real b) { g_flops++; a.v -= b.v; return a; }
static inline real& operator*=(real& a, real b) { g_flops++; a.v *= b.v; return a; }
static inline real& operator/=(real& a, real b) { g_flops++; a.v /= b.v; return a; }
static inline bool operator<(real a, real b) { return a.v < b.v; }
static inline bool operator>(real a, real b) { return a.v > b.v; }
static inline bool operator<=(real a, real b) { return a.v <= b.v; }
static inline bool operator>=(real a, real b) { return a.v >= b.v; }
static inline bool operator==(real a, real b) { return a.v == b.v; }
static inline bool operator!=(real a, real b) { return a.v != b.v; }
static inline real rsqrt_(real a) { g_flops++; return real(std::sqrt(a.v)); }
static inline real rsin(real a) { g_flops++; return real(std::sin(a.v)); }
static inline real rcos(real a) { g_flops++; return real(std::cos(a.v)); }
static inline real ratan2(real a, real b) { g_flops++; return real(std::atan2(a.v, b.v)); }
static inline real rasin(real a) { g_flops++; return real(std::asin(a.v)); }
static inline real rtan(real a) { g_flops++; return real(std::tan(a.v)); }
static inline real rpow(real a, real b) { g_flops++; return real(std::pow(a.v, b.v)); }
static inline real rabs(real a) { return real(std::fabs(a.v)); }
static inline double D(real a) { return a.v; }
#else
typedef double real;
static inline real rsqrt_(real a) { return std::sqrt(a); }
static inline real rsin(real a) { return std::sin(a); }
static inline real rcos(real a) { return std::cos(a); }
static inline real ratan2(real a, real b) { return std::atan2(a, b); }
static inline real rasin(real a) { return std::asin(a); }
static inline real rtan(real a) { return std::tan(a); }
static inline real rpow(real a, real b) { return std::pow(a, b); }
static inline real rabs(real a) { return std::fabs(a); }
static inline double D(real a) { return a; }
#endif
static inline real rmax(real a, real b) { return a > b ? a : b; }
static inline real rmin(real a, real b) { return a < b ? a : b; }
static inline real rclip(real x, real lo, real hi) { return x < lo ? lo : (x > hi ? hi : x); }

#define KO_MINVAL 1e-15
#define KO_PI 3.14159265358979323846

/* ------------------------------------------------------------------------------------------- sizes */
enum { NB = 32, NJ = 32, NQ = 40, NV = 32, NU = 24, NSITE = 16, NGEOM = 16, MAXCON = 16, MAXEFC = 192 };
enum { EFC_FRICTION = 0, EFC_LIMIT = 1, EFC_CONTACT = 2 };

/* reference constants (gym_kmanip/__init__.py) */
static const double K_EE_POS_DELTA = 0.01;   /* :174-180 */
static const double K_EE_ORN_DELTA = 0.1;    /* :181-187 */
static const float K_Q_POS_DELTA_F = 0.1f;   /* :196 (python float times a float32 array stays float32) */
static const float K_EE_S_MIN_F = -0.029f;   /* :199 */
static const float K_EE_S_MAX_F = 0.005f;    /* :200 */
static const float K_EE_S_DELTA_F = 0.0001f; /* :201 */
static const double K_MAX_Q_VEL = KO_PI;     /* :31 */
static const double K_REWARD_VEL_PENALTY = 0.01, K_REWARD_GRIP_DIST = 0.01, K_EPSILON = 1e-6; /* :190,205-206 */
static const double K_IK_RES_RAD = 0.02, K_IK_RES_REG_PREV = 6e-3, K_IK_RES_REG_HOME = 2e-6; /* :37-39 */
static const double K_IK_JAC_RAD = 0.02, K_IK_JAC_REG = 9e-3;                              /* :40-41 */
static const double K_CONTROL_TIMESTEP = 0.02;                                             /* :30 */

struct ko_contact {
  real dist, pos[3], frame[9], friction[5];
  int geom1, geom2, pair, dim, efc;
};

struct ko_data {
  /* state */
  real qpos[NQ], qvel[NV], ctrl[NU], qacc_warmstart[NV], time;
  real mocap_pos[2][3], mocap_quat[2][4];
  /* derived tables */
  int body_lastdof[NB];
  /* position stage */
  real xpos[NB][3], xquat[NB][4], xmat[NB][9], xipos[NB][3], ximat[NB][9];
  real xanchor[NJ][3], xaxis[NJ][3];
  real site_xpos[NSITE][3], site_xmat[NSITE][9], geom_xpos[NGEOM][3], geom_xmat[NGEOM][9];
  real subtree_com[NB][3], cinert[NB][10], crb[NB][10], cdof[NV][6];
  real qM[NV][NV], qLD[NV][NV], qLDiagInv[NV];
  real actuator_length[NU];
  int ncon;
  int ncon_peak;   /* test diagnostic: most contacts seen by any collision pass since the caller cleared it */
  ko_contact contact[MAXCON];
  int nefc;
  int efc_type[MAXEFC], efc_id[MAXEFC];
  real efc_J[MAXEFC][NV], efc_pos[MAXEFC], efc_margin[MAXEFC], efc_frictionloss[MAXEFC];
  real efc_diagApprox[MAXEFC], efc_R[MAXEFC], efc_D[MAXEFC], efc_K[MAXEFC], efc_B[MAXEFC], efc_imp[MAXEFC];
  /* velocity stage */
  real cvel[NB][6], cdof_dot[NV][6], qfrc_bias[NV], actuator_velocity[NU];
  real efc_vel[MAXEFC], efc_aref[MAXEFC];
  /* acceleration stage */
  real actuator_force[NU], qfrc_actuator[NV], qfrc_smooth[NV], qacc_smooth[NV], qacc[NV];
  real efc_force[MAXEFC], qfrc_constraint[NV];
  int efc_state[MAXEFC];
  int solver_niter, ls_evals;
};

/* ------------------------------------------------------------------------------------------- small math */
static inline void v3set(real* r, real a, real b, real c) { r[0] = a; r[1] = b; r[2] = c; }
static inline void v3copy(real* r, const real* a) { r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; }
static inline real v3dot(const real* a, const real* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void v3cross(real* r, const real* a, const real* b) {
  real x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline real v3norm(const real* a) { return rsqrt_(v3dot(a, a)); }
/* mju_normalize3: returns the norm; degenerate vectors become (1,0,0) */
static inline real v3normalize(real* a) {
  real n = v3norm(a);
  if (n < real(KO_MINVAL)) { a[0] = 1; a[1] = 0; a[2] = 0; }
  else { real s = real(1.0) / n; a[0] *= s; a[1] *= s; a[2] *= s; }
  return n;
}
static inline void q4normalize(real* q) {
  real n = rsqrt_(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < real(KO_MINVAL)) { q[0] = 1; q[1] = 0; q[2] = 0; q[3] = 0; }
  else { real s = real(1.0) / n; for (int i = 0; i < 4; i++) q[i] *= s; }
}
static inline void q4mul(real* r, const real* a, const real* b) {
  real t0 = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  real t1 = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  real t2 = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  real t3 = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = t0; r[1] = t1; r[2] = t2; r[3] = t3;
}
static inline void q4mat(real* m, const real* q) { /* mju_quat2Mat, row-major */
  real q00 = q[0] * q[0], q01 = q[0] * q[1], q02 = q[0] * q[2], q03 = q[0] * q[3];
  real q11 = q[1] * q[1], q12 = q[1] * q[2], q13 = q[1] * q[3], q22 = q[2] * q[2], q23 = q[2] * q[3], q33 = q[3] * q[3];
  m[0] = q00 + q11 - q22 - q33; m[4] = q00 - q11 + q22 - q33; m[8] = q00 - q11 - q22 + q33;
  m[1] = real(2.0) * (q12 - q03); m[2] = real(2.0) * (q13 + q02);
  m[3] = real(2.0) * (q12 + q03); m[5] = real(2.0) * (q23 - q01);
  m[6] = real(2.0) * (q13 - q02); m[7] = real(2.0) * (q23 + q01);
}
static inline void m3mulv(real* r, const real* m, const real* v) {
  real x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2], y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2],
       z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void m3Tmulv(real* r, const real* m, const real* v) {
  real x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2], y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2],
       z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void axisangle2quat(real* q, const real* axis, real angle) {
  if (angle == real(0.0)) { q[0] = 1; q[1] = 0; q[2] = 0; q[3] = 0; return; }
  real s = rsin(angle * real(0.5));
  q[0] = rcos(angle * real(0.5)); q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}
/* mju_mat2Quat */
static void mat2quat(real* q, const real* m) {
  if (m[0] + m[4] + m[8] > real(0.0)) {
    q[0] = real(0.5) * rsqrt_(real(1.0) + m[0] + m[4] + m[8]);
    q[1] = real(0.25) * (m[7] - m[5]) / q[0]; q[2] = real(0.25) * (m[2] - m[6]) / q[0]; q[3] = real(0.25) * (m[3] - m[1]) / q[0];
  } else if (m[0] > m[4] && m[0] > m[8]) {
    q[1] = real(0.5) * rsqrt_(real(1.0) + m[0] - m[4] - m[8]);
    q[0] = real(0.25) * (m[7] - m[5]) / q[1]; q[2] = real(0.25) * (m[1] + m[3]) / q[1]; q[3] = real(0.25) * (m[2] + m[6]) / q[1];
  } else if (m[4] > m[8]) {
    q[2] = real(0.5) * rsqrt_(real(1.0) - m[0] + m[4] - m[8]);
    q[0] = real(0.25) * (m[2] - m[6]) / q[2]; q[1] = real(0.25) * (m[1] + m[3]) / q[2]; q[3] = real(0.25) * (m[5] + m[7]) / q[2];
  } else {
    q[3] = real(0.5) * rsqrt_(real(1.0) - m[0] - m[4] + m[8]);
    q[0] = real(0.25) * (m[3] - m[1]) / q[3]; q[1] = real(0.25) * (m[2] + m[6]) / q[3]; q[2] = real(0.25) * (m[5] + m[7]) / q[3];
  }
  q4normalize(q);
}
/* mju_subQuat: 3D velocity taking qb to qa (in qb's frame), via quat2Vel with dt = 1 */
static void subquat(real* res, const real* qa, const real* qb) {
  real qneg[4] = {qb[0], -qb[1], -qb[2], -qb[3]}, qd[4];
  q4mul(qd, qneg, qa);
  real axis[3] = {qd[1], qd[2], qd[3]};
  real s = v3normalize(axis);
  real speed = real(2.0) * ratan2(s, qd[0]);
  if (speed > real(KO_PI)) speed -= real(2.0 * KO_PI);
  res[0] = axis[0] * speed; res[1] = axis[1] * speed; res[2] = axis[2] * speed;
}
/* mju_makeFrame: complete a frame whose first row (normal) is given */
static void makeframe(real* f) {
  v3normalize(f);
  real* y = f + 3;
  if (v3norm(y) < real(0.5)) {
    y[0] = 0; y[1] = 0; y[2] = 0;
    if (f[1] < real(0.5) && f[1] > real(-0.5)) y[1] = 1; else y[2] = 1;
  }
  real t = v3dot(f, y);
  y[0] -= t * f[0]; y[1] -= t * f[1]; y[2] -= t * f[2];
  v3normalize(y);
  v3cross(f + 6, f, y);
}
/* spatial algebra, MuJoCo layout: 6-vectors are [angular(3); linear(3)], cinert is the 10-vector
   [Ixx Iyy Izz Ixy Ixz Iyz, m*off(3), m] about the reference point */
static void inert_com(real* res, const real* inert, const real* mat, const real* dif, real mass) {
  real tmp[9];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
    tmp[3 * i + j] = mat[3 * i] * inert[0] * mat[3 * j] + mat[3 * i + 1] * inert[1] * mat[3 * j + 1] + mat[3 * i + 2] * inert[2] * mat[3 * j + 2];
  res[0] = tmp[0] + mass * (dif[1] * dif[1] + dif[2] * dif[2]);
  res[1] = tmp[4] + mass * (dif[0] * dif[0] + dif[2] * dif[2]);
  res[2] = tmp[8] + mass * (dif[0] * dif[0] + dif[1] * dif[1]);
  res[3] = tmp[1] - mass * dif[0] * dif[1];
  res[4] = tmp[2] - mass * dif[0] * dif[2];
  res[5] = tmp[5] - mass * dif[1] * dif[2];
  res[6] = mass * dif[0]; res[7] = mass * dif[1]; res[8] = mass * dif[2]; res[9] = mass;
}
static void mul_inert_vec(real* res, const real* i, const real* v) {
  real r[6];
  r[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
  r[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
  r[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
  r[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
  r[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
  r[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
  for (int k = 0; k < 6; k++) res[k] = r[k];
}
static void cross_motion(real* res, const real* vel, const real* v) {
  real r[6];
  v3cross(r, vel, v);
  real a[3], b[3];
  v3cross(a, vel, v + 3);
  v3cross(b, vel + 3, v);
  r[3] = a[0] + b[0]; r[4] = a[1] + b[1]; r[5] = a[2] + b[2];
  for (int k = 0; k < 6; k++) res[k] = r[k];
}
static void cross_force(real* res, const real* vel, const real* f) {
  real r[6], a[3], b[3];
  v3cross(a, vel, f);
  v3cross(b, vel + 3, f + 3);
  r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2];
  v3cross(r + 3, vel, f + 3);
  for (int k = 0; k < 6; k++) res[k] = r[k];
}

/* ------------------------------------------------------------------------------------------- init */
static void ko_init(const ko_model* m, ko_data* d) {
  if (m->nbody > NB || m->njnt > NJ || m->nq > NQ || m->nv > NV || m->nu > NU || m->nsite > NSITE || m->ngeom > NGEOM) {
    fprintf(stderr, "kmanip_oracle: model exceeds the oracle's static sizes\n");
    abort();
  }
  memset((void*)d, 0, sizeof(ko_data));
  d->body_lastdof[0] = -1;
  for (int b = 1; b < m->nbody; b++) {
    d->body_lastdof[b] = d->body_lastdof[m->body_parent[b]];
    for (int k = 0; k < m->body_jntnum[b]; k++) {
      int j = m->body_jntadr[b] + k;
      d->body_lastdof[b] = m->jnt_dofadr[j] + (m->jnt_type[j] == KO_JNT_FREE ? 5 : 0);
    }
  }
  for (int i = 0; i < m->nq; i++) d->qpos[i] = m->qpos0[i];
  for (int k = 0; k < m->nmocap; k++) {
    for (int i = 0; i < 3; i++) d->mocap_pos[k][i] = m->mocap_pos0[3 * k + i];
    for (int i = 0; i < 4; i++) d->mocap_quat[k][i] = m->mocap_quat0[4 * k + i];
  }
}

/* ------------------------------------------------------------------------------------------- mj_kinematics */
static void ko_kinematics(const ko_model* m, ko_data* d) {
  v3set(d->xpos[0], 0, 0, 0);
  d->xquat[0][0] = 1; d->xquat[0][1] = 0; d->xquat[0][2] = 0; d->xquat[0][3] = 0;
  q4mat(d->xmat[0], d->xquat[0]);
  for (int b = 1; b < m->nbody; b++) {
    real* xp = d->xpos[b]; real* xq = d->xquat[b];
    int p = m->body_parent[b];
    int jadr = m->body_jntadr[b], jnum = m->body_jntnum[b];
    if (jnum == 1 && m->jnt_type[jadr] == KO_JNT_FREE) {
      int a = m->jnt_qposadr[jadr];
      q4normalize(d->qpos + a + 3);                    /* mj_kinematics normalises qpos quaternions in place */
      v3copy(xp, d->qpos + a);
      for (int i = 0; i < 4; i++) xq[i] = d->qpos[a + 3 + i];
      v3copy(d->xanchor[jadr], xp);
      for (int i = 0; i < 3; i++) d->xaxis[jadr][i] = m->jnt_axis[3 * jadr + i];
    } else {
      real bpos[3] = {m->body_pos[3 * b], m->body_pos[3 * b + 1], m->body_pos[3 * b + 2]};
      real bquat[4] = {m->body_quat[4 * b], m->body_quat[4 * b + 1], m->body_quat[4 * b + 2], m->body_quat[4 * b + 3]};
      if (m->body_mocapid[b] >= 0) {
        int k = m->body_mocapid[b];
        v3copy(bpos, d->mocap_pos[k]);
        for (int i = 0; i < 4; i++) bquat[i] = d->mocap_quat[k][i];
        q4normalize(bquat);
      }
      real t[3];
      m3mulv(t, d->xmat[p], bpos);
      xp[0] = d->xpos[p][0] + t[0]; xp[1] = d->xpos[p][1] + t[1]; xp[2] = d->xpos[p][2] + t[2];
      q4mul(xq, d->xquat[p], bquat);
      for (int k = 0; k < jnum; k++) {
        int j = jadr + k, a = m->jnt_qposadr[j];
        real jpos[3] = {m->jnt_pos[3 * j], m->jnt_pos[3 * j + 1], m->jnt_pos[3 * j + 2]};
        real jax[3] = {m->jnt_axis[3 * j], m->jnt_axis[3 * j + 1], m->jnt_axis[3 * j + 2]};
        real mat[9];
        q4mat(mat, xq);
        m3mulv(d->xanchor[j], mat, jpos);
        for (int i = 0; i < 3; i++) d->xanchor[j][i] += xp[i];
        m3mulv(d->xaxis[j], mat, jax);
        real q = d->qpos[a] - real(m->qpos0[a]);
        if (m->jnt_type[j] == KO_JNT_SLIDE) {
          for (int i = 0; i < 3; i++) xp[i] += d->xaxis[j][i] * q;
        } else if (m->jnt_type[j] == KO_JNT_HINGE) {
          real qloc[4];
          axisangle2quat(qloc, jax, q);
          q4mul(xq, xq, qloc);
          q4mat(mat, xq);
          m3mulv(t, mat, jpos);                        /* off-centre rotation correction */
          for (int i = 0; i < 3; i++) xp[i] = d->xanchor[j][i] - t[i];
        }
      }
    }
    q4normalize(xq);
    q4mat(d->xmat[b], xq);
    real ip[3] = {m->body_ipos[3 * b], m->body_ipos[3 * b + 1], m->body_ipos[3 * b + 2]}, t[3];
    m3mulv(t, d->xmat[b], ip);
    for (int i = 0; i < 3; i++) d->xipos[b][i] = xp[i] + t[i];
    for (int i = 0; i < 9; i++) d->ximat[b][i] = d->xmat[b][i];   /* body_iquat = identity in the completed model */
  }
  for (int s = 0; s < m->nsite; s++) {
    int b = m->site_bodyid[s];
    real sp[3] = {m->site_pos[3 * s], m->site_pos[3 * s + 1], m->site_pos[3 * s + 2]}, t[3];
    real sq[4] = {m->site_quat[4 * s], m->site_quat[4 * s + 1], m->site_quat[4 * s + 2], m->site_quat[4 * s + 3]}, q[4];
    m3mulv(t, d->xmat[b], sp);
    for (int i = 0; i < 3; i++) d->site_xpos[s][i] = d->xpos[b][i] + t[i];
    q4mul(q, d->xquat[b], sq);
    q4mat(d->site_xmat[s], q);
  }
  for (int g = 0; g < m->ngeom; g++) {
    int b = m->geom_bodyid[g];
    real gp[3] = {m->geom_pos[3 * g], m->geom_pos[3 * g + 1], m->geom_pos[3 * g + 2]}, t[3];
    real gq[4] = {m->geom_quat[4 * g], m->geom_quat[4 * g + 1], m->geom_quat[4 * g + 2], m->geom_quat[4 * g + 3]}, q[4];
    m3mulv(t, d->xmat[b], gp);
    for (int i = 0; i < 3; i++) d->geom_xpos[g][i] = d->xpos[b][i] + t[i];
    q4mul(q, d->xquat[b], gq);
    q4mat(d->geom_xmat[g], q);
  }
}

/* ------------------------------------------------------------------------------------------- mj_comPos */
static void ko_com_pos(const ko_model* m, ko_data* d) {
  real smass[NB];
  for (int b = 0; b < m->nbody; b++) {
    smass[b] = m->body_mass[b];
    for (int i = 0; i < 3; i++) d->subtree_com[b][i] = real(m->body_mass[b]) * d->xipos[b][i];
  }
  for (int b = m->nbody - 1; b > 0; b--) {
    int p = m->body_parent[b];
    smass[p] += smass[b];
    for (int i = 0; i < 3; i++) d->subtree_com[p][i] += d->subtree_com[b][i];
  }
  for (int b = 0; b < m->nbody; b++) {
    if (smass[b] < real(KO_MINVAL)) v3copy(d->subtree_com[b], d->xipos[b]);
    else for (int i = 0; i < 3; i++) d->subtree_com[b][i] /= smass[b];
  }
  for (int b = 1; b < m->nbody; b++) {
    real off[3], inert[3] = {m->body_inertia[3 * b], m->body_inertia[3 * b + 1], m->body_inertia[3 * b + 2]};
    for (int i = 0; i < 3; i++) off[i] = d->xipos[b][i] - d->subtree_com[m->body_rootid[b]][i];
    inert_com(d->cinert[b], inert, d->ximat[b], off, m->body_mass[b]);
  }
  for (int j = 0; j < m->njnt; j++) {
    int b = m->jnt_bodyid[j], da = m->jnt_dofadr[j];
    real off[3];
    for (int i = 0; i < 3; i++) off[i] = d->subtree_com[m->body_rootid[b]][i] - d->xanchor[j][i];
    if (m->jnt_type[j] == KO_JNT_FREE) {
      for (int k = 0; k < 3; k++) {
        for (int i = 0; i < 6; i++) d->cdof[da + k][i] = 0;
        d->cdof[da + k][3 + k] = 1;
        real ax[3] = {d->xmat[b][k], d->xmat[b][3 + k], d->xmat[b][6 + k]};   /* body axis k in world */
        v3copy(d->cdof[da + 3 + k], ax);
        v3cross(d->cdof[da + 3 + k] + 3, ax, off);
      }
    } else if (m->jnt_type[j] == KO_JNT_SLIDE) {
      v3set(d->cdof[da], 0, 0, 0);
      v3copy(d->cdof[da] + 3, d->xaxis[j]);
    } else {
      v3copy(d->cdof[da], d->xaxis[j]);
      v3cross(d->cdof[da] + 3, d->xaxis[j], off);
    }
  }
}

/* ------------------------------------------------------------------------------------------- mj_crb + mj_factorM */
static void ko_crb(const ko_model* m, ko_data* d) {
  int nv = m->nv;
  for (int b = 1; b < m->nbody; b++) for (int i = 0; i < 10; i++) d->crb[b][i] = d->cinert[b][i];
  for (int i = 0; i < 10; i++) d->crb[0][i] = 0;
  for (int b = m->nbody - 1; b > 0; b--) {
    int p = m->body_parent[b];
    if (p > 0) for (int i = 0; i < 10; i++) d->crb[p][i] += d->crb[b][i];
  }
  for (int i = 0; i < nv; i++) for (int j = 0; j < nv; j++) d->qM[i][j] = 0;
  for (int i = 0; i < nv; i++) {
    real buf[6];
    mul_inert_vec(buf, d->crb[m->dof_bodyid[i]], d->cdof[i]);
    for (int j = i; j >= 0; j = m->dof_parentid[j]) {
      real s = 0;
      for (int k = 0; k < 6; k++) s += d->cdof[j][k] * buf[k];
      d->qM[i][j] = s; d->qM[j][i] = s;
    }
  }
}
/* L^T D L with the dof-tree sparsity (mj_factorM): qLD[k][i] (i ancestor of k) holds L, diagonal holds D */
static void ko_factor_m(const ko_model* m, ko_data* d) {
  int nv = m->nv;
  for (int i = 0; i < nv; i++) for (int j = 0; j < nv; j++) d->qLD[i][j] = d->qM[i][j];
  for (int k = nv - 1; k >= 0; k--) {
    real Mkk = d->qLD[k][k];
    for (int i = m->dof_parentid[k]; i >= 0; i = m->dof_parentid[i]) {
      real tmp = d->qLD[k][i] / Mkk;
      for (int j = i; j >= 0; j = m->dof_parentid[j]) d->qLD[i][j] -= d->qLD[k][j] * tmp;
      d->qLD[k][i] = tmp;
    }
    d->qLDiagInv[k] = real(1.0) / Mkk;
  }
}
static void ko_solve_m(const ko_model* m, const ko_data* d, real* x) {
  int nv = m->nv;
  for (int i = nv - 1; i >= 0; i--)
    for (int j = m->dof_parentid[i]; j >= 0; j = m->dof_parentid[j]) x[j] -= d->qLD[i][j] * x[i];
  for (int i = 0; i < nv; i++) x[i] *= d->qLDiagInv[i];
  for (int i = 0; i < nv; i++)
    for (int j = m->dof_parentid[i]; j >= 0; j = m->dof_parentid[j]) x[i] -= d->qLD[i][j] * x[j];
}

/* ------------------------------------------------------------------------------------------- collision */
static void ko_add_contact(const ko_model* m, ko_data* d, int pair, real dist, const real* pos, const real* normal) {
  if (d->ncon >= MAXCON) return;
  ko_contact* c = &d->contact[d->ncon++];
  c->dist = dist;
  v3copy(c->pos, pos);
  v3copy(c->frame, normal);
  v3set(c->frame + 3, 0, 0, 0);
  makeframe(c->frame);
  c->geom1 = m->pair_geom1[pair]; c->geom2 = m->pair_geom2[pair]; c->pair = pair;
  c->dim = m->pair_condim[pair];
  for (int i = 0; i < 5; i++) c->friction[i] = m->pair_friction[5 * pair + i];
  c->efc = -1;
}
/* plane (geom1) vs box (geom2): corners below the plane, at most 4 (mjc_PlaneBox) */
static void ko_plane_box(const ko_model* m, ko_data* d, int pair) {
  int g1 = m->pair_geom1[pair], g2 = m->pair_geom2[pair];
  real margin = m->pair_margin[pair];
  const real* mat1 = d->geom_xmat[g1]; const real* pos1 = d->geom_xpos[g1];
  const real* mat2 = d->geom_xmat[g2]; const real* pos2 = d->geom_xpos[g2];
  real normal[3] = {mat1[2], mat1[5], mat1[8]};
  real dif[3] = {pos2[0] - pos1[0], pos2[1] - pos1[1], pos2[2] - pos1[2]};
  real dist = v3dot(dif, normal);
  int cnt = 0;
  for (int i = 0; i < 8; i++) {
    real vec[3] = {real(i & 1 ? 1.0 : -1.0) * real(m->geom_size[3 * g2]), real(i & 2 ? 1.0 : -1.0) * real(m->geom_size[3 * g2 + 1]),
                   real(i & 4 ? 1.0 : -1.0) * real(m->geom_size[3 * g2 + 2])};
    real corner[3];
    m3mulv(corner, mat2, vec);
    real ldist = v3dot(normal, corner);
    if (dist + ldist > margin || ldist > real(0.0)) continue;
    real cdist = dist + ldist, pos[3];
    for (int k = 0; k < 3; k++) pos[k] = corner[k] + pos2[k] - normal[k] * cdist * real(0.5);
    ko_add_contact(m, d, pair, cdist, pos, normal);
    if (++cnt >= 4) return;
  }
}
/* sphere (geom1) vs box (geom2) */
static void ko_sphere_box(const ko_model* m, ko_data* d, int pair) {
  int g1 = m->pair_geom1[pair], g2 = m->pair_geom2[pair];
  real margin = m->pair_margin[pair], radius = m->geom_size[3 * g1];
  const real* pos1 = d->geom_xpos[g1];
  const real* mat2 = d->geom_xmat[g2]; const real* pos2 = d->geom_xpos[g2];
  real size[3] = {m->geom_size[3 * g2], m->geom_size[3 * g2 + 1], m->geom_size[3 * g2 + 2]};
  real tmp[3] = {pos1[0] - pos2[0], pos1[1] - pos2[1], pos1[2] - pos2[2]}, center[3], clamped[3], n[3], pos_b[3];
  m3Tmulv(center, mat2, tmp);
  for (int i = 0; i < 3; i++) { clamped[i] = rclip(center[i], -size[i], size[i]); n[i] = clamped[i] - center[i]; }
  real dist = v3norm(n), cdist;
  if (dist - radius > margin) return;
  if (dist <= real(KO_MINVAL)) {
    /* centre inside the box: push out through the nearest face */
    real closest = real(2.0) * (size[0] + size[1] + size[2]);
    int k = 0;
    for (int i = 0; i < 6; i++) {
      real face = real(i % 2 ? 1.0 : -1.0) * size[i / 2];
      real t = rabs(face - center[i / 2]);
      if (t < closest) { closest = t; k = i; }
    }
    real fo[3] = {0, 0, 0};
    fo[k / 2] = real(k % 2 ? 1.0 : -1.0);
    for (int i = 0; i < 3; i++) { n[i] = -fo[i]; pos_b[i] = center[i] + fo[i] * (closest - radius) * real(0.5); }
    cdist = -closest - radius;
  } else {
    for (int i = 0; i < 3; i++) n[i] /= dist;
    cdist = dist - radius;
    for (int i = 0; i < 3; i++) pos_b[i] = center[i] + n[i] * (radius + real(0.5) * cdist);
  }
  real pos[3], normal[3];
  m3mulv(pos, mat2, pos_b);
  for (int i = 0; i < 3; i++) pos[i] += pos2[i];
  m3mulv(normal, mat2, n);
  ko_add_contact(m, d, pair, cdist, pos, normal);
}
static void ko_collision(const ko_model* m, ko_data* d) {
  d->ncon = 0;
  for (int p = 0; p < m->npair; p++) {
    int t1 = m->geom_type[m->pair_geom1[p]], t2 = m->geom_type[m->pair_geom2[p]];
    if (t1 == KO_GEOM_PLANE && t2 == KO_GEOM_BOX) ko_plane_box(m, d, p);
    else if (t1 == KO_GEOM_SPHERE && t2 == KO_GEOM_BOX) ko_sphere_box(m, d, p);
  }
  if (d->ncon > d->ncon_peak) d->ncon_peak = d->ncon;
}

/* ------------------------------------------------------------------------------------------- Jacobians */
/* mj_jac: translational/rotational Jacobian (3 x nv each) of a world point attached to `body` */
static void ko_jac(const ko_model* m, const ko_data* d, real jacp[3][NV], real jacr[3][NV], const real* point, int body) {
  for (int i = 0; i < 3; i++) for (int j = 0; j < m->nv; j++) { jacp[i][j] = 0; jacr[i][j] = 0; }
  if (d->body_lastdof[body] < 0) return;
  real off[3];
  for (int i = 0; i < 3; i++) off[i] = point[i] - d->subtree_com[m->body_rootid[body]][i];
  for (int i = d->body_lastdof[body]; i >= 0; i = m->dof_parentid[i]) {
    real t[3];
    v3cross(t, d->cdof[i], off);
    for (int k = 0; k < 3; k++) { jacr[k][i] = d->cdof[i][k]; jacp[k][i] = d->cdof[i][3 + k] + t[k]; }
  }
}

/* ------------------------------------------------------------------------------------------- mj_makeConstraint */
static void ko_impedance(const real* solimp, real pos, real margin, real* imp) {
  real dmin = solimp[0], dmax = solimp[1], width = solimp[2], mid = solimp[3], power = solimp[4];
  if (dmin == dmax || width <= real(KO_MINVAL)) { *imp = real(0.5) * (dmin + dmax); return; }
  real x = (pos - margin) / width;
  if (x < real(0.0)) x = -x;
  if (x >= real(1.0)) { *imp = dmax; return; }
  if (x == real(0.0)) { *imp = dmin; return; }
  real y;
  if (power == real(1.0)) y = x;
  else if (x <= mid) y = rpow(x, power) / rpow(mid, power - real(1.0));
  else y = real(1.0) - rpow(real(1.0) - x, power) / rpow(real(1.0) - mid, power - real(1.0));
  *imp = dmin + y * (dmax - dmin);
}
static int ko_add_efc(ko_data* d, int nv, int type, int id, real pos, real margin, real floss, real diag) {
  int r = d->nefc++;
  for (int j = 0; j < nv; j++) d->efc_J[r][j] = 0;
  d->efc_type[r] = type; d->efc_id[r] = id; d->efc_pos[r] = pos; d->efc_margin[r] = margin;
  d->efc_frictionloss[r] = floss; d->efc_diagApprox[r] = diag;
  return r;
}
static void ko_make_constraint(const ko_model* m, ko_data* d) {
  int nv = m->nv;
  d->nefc = 0;
  /* friction loss, dof order */
  for (int i = 0; i < nv; i++) if (m->dof_frictionloss[i] > 0) {
    int r = ko_add_efc(d, nv, EFC_FRICTION, i, 0, 0, m->dof_frictionloss[i], m->dof_invweight0[i]);
    d->efc_J[r][i] = 1;
  }
  /* joint limits, joint order; lower side first */
  for (int j = 0; j < m->njnt; j++) if (m->jnt_limited[j] && m->jnt_type[j] != KO_JNT_FREE) {
    real q = d->qpos[m->jnt_qposadr[j]];
    for (int side = -1; side <= 1; side += 2) {
      real dist = real((double)side) * (real(m->jnt_range[2 * j + (side + 1) / 2]) - q);
      if (dist < real(0.0)) {
        int r = ko_add_efc(d, nv, EFC_LIMIT, j, dist, 0, 0, m->dof_invweight0[m->jnt_dofadr[j]]);
        d->efc_J[r][m->jnt_dofadr[j]] = real((double)-side);
      }
    }
  }
  /* contacts: pyramidal friction cones */
  for (int c = 0; c < d->ncon; c++) {
    ko_contact* con = &d->contact[c];
    int b1 = m->geom_bodyid[con->geom1], b2 = m->geom_bodyid[con->geom2];
    static thread_local real jp1[3][NV], jr1[3][NV], jp2[3][NV], jr2[3][NV];
    ko_jac(m, d, jp1, jr1, con->pos, b1);
    ko_jac(m, d, jp2, jr2, con->pos, b2);
    real jc[6][NV];   /* contact-frame rows: 3 translational, 3 rotational */
    for (int r = 0; r < 3; r++) for (int j = 0; j < nv; j++) {
      real sp = 0, sr = 0;
      for (int k = 0; k < 3; k++) {
        sp += con->frame[3 * r + k] * (jp2[k][j] - jp1[k][j]);
        sr += con->frame[3 * r + k] * (jr2[k][j] - jr1[k][j]);
      }
      jc[r][j] = sp; jc[3 + r][j] = sr;
    }
    real tran = real(m->body_invweight0[2 * b1]) + real(m->body_invweight0[2 * b2]);
    real rot = real(m->body_invweight0[2 * b1 + 1]) + real(m->body_invweight0[2 * b2 + 1]);
    con->efc = d->nefc;
    for (int k = 1; k < con->dim; k++) {
      real mu = con->friction[k - 1];
      real diag = tran + mu * mu * (k < 3 ? tran : rot);
      for (int sgn = 1; sgn >= -1; sgn -= 2) {
        int r = ko_add_efc(d, nv, EFC_CONTACT, c, con->dist, 0, 0, diag);
        for (int j = 0; j < nv; j++) d->efc_J[r][j] = jc[0][j] + real((double)sgn) * mu * jc[k][j];
      }
    }
  }
  /* impedance, regulariser, reference-acceleration gains (mj_makeImpedance) */
  for (int r = 0; r < d->nefc; r++) {
    real solref[2], solimp[5];
    if (d->efc_type[r] == EFC_FRICTION) {
      for (int i = 0; i < 2; i++) solref[i] = m->dof_solref[2 * d->efc_id[r] + i];
      for (int i = 0; i < 5; i++) solimp[i] = m->dof_solimp[5 * d->efc_id[r] + i];
    } else if (d->efc_type[r] == EFC_LIMIT) {
      for (int i = 0; i < 2; i++) solref[i] = m->jnt_solref[2 * d->efc_id[r] + i];
      for (int i = 0; i < 5; i++) solimp[i] = m->jnt_solimp[5 * d->efc_id[r] + i];
    } else {
      int p = d->contact[d->efc_id[r]].pair;
      for (int i = 0; i < 2; i++) solref[i] = m->pair_solref[2 * p + i];
      for (int i = 0; i < 5; i++) solimp[i] = m->pair_solimp[5 * p + i];
    }
    real imp;
    ko_impedance(solimp, d->efc_pos[r], d->efc_margin[r], &imp);
    d->efc_imp[r] = imp;
    d->efc_R[r] = rmax(real(KO_MINVAL), (real(1.0) - imp) * d->efc_diagApprox[r] / imp);
    real tc = rmax(solref[0], real(2.0 * m->timestep)), dr = solref[1], dmax = solimp[1];
    d->efc_K[r] = d->efc_type[r] == EFC_FRICTION ? real(0.0) : real(1.0) / (dmax * dmax * tc * tc * dr * dr);
    d->efc_B[r] = real(2.0) / (dmax * tc);
  }
  /* pyramidal contacts: all rows of a contact share R = 2 mu^2 R_first, mu = friction[0]/sqrt(impratio) */
  for (int c = 0; c < d->ncon; c++) {
    ko_contact* con = &d->contact[c];
    real mu = con->friction[0] / rsqrt_(real(m->impratio));
    real Rpy = real(2.0) * mu * mu * d->efc_R[con->efc];
    for (int k = 0; k < 2 * (con->dim - 1); k++) d->efc_R[con->efc + k] = Rpy;
  }
  for (int r = 0; r < d->nefc; r++) d->efc_D[r] = real(1.0) / d->efc_R[r];
}

static void ko_transmission(const ko_model* m, ko_data* d) {
  for (int i = 0; i < m->nu; i++) d->actuator_length[i] = d->qpos[m->jnt_qposadr[m->act_jntid[i]]];
}

static void ko_fwd_position(const ko_model* m, ko_data* d) {
  ko_kinematics(m, d);
  ko_com_pos(m, d);
  ko_crb(m, d);
  ko_factor_m(m, d);
  ko_collision(m, d);
  ko_make_constraint(m, d);
  ko_transmission(m, d);
}

/* ------------------------------------------------------------------------------------------- velocity stage */
static void ko_com_vel(const ko_model* m, ko_data* d) {
  for (int i = 0; i < 6; i++) d->cvel[0][i] = 0;
  for (int b = 1; b < m->nbody; b++) {
    real cvel[6];
    for (int i = 0; i < 6; i++) cvel[i] = d->cvel[m->body_parent[b]][i];
    for (int k = 0; k < m->body_jntnum[b]; k++) {
      int j = m->body_jntadr[b] + k, da = m->jnt_dofadr[j];
      if (m->jnt_type[j] == KO_JNT_FREE) {
        for (int t = 0; t < 3; t++) {
          for (int i = 0; i < 6; i++) d->cdof_dot[da + t][i] = 0;
          for (int i = 0; i < 6; i++) cvel[i] += d->cdof[da + t][i] * d->qvel[da + t];
        }
        for (int t = 3; t < 6; t++) cross_motion(d->cdof_dot[da + t], cvel, d->cdof[da + t]);
        for (int t = 3; t < 6; t++) for (int i = 0; i < 6; i++) cvel[i] += d->cdof[da + t][i] * d->qvel[da + t];
      } else {
        cross_motion(d->cdof_dot[da], cvel, d->cdof[da]);
        for (int i = 0; i < 6; i++) cvel[i] += d->cdof[da][i] * d->qvel[da];
      }
    }
    for (int i = 0; i < 6; i++) d->cvel[b][i] = cvel[i];
  }
}
/* mj_rne with flg_acc = 0: Coriolis/centrifugal/gravity bias */
static void ko_rne(const ko_model* m, ko_data* d) {
  static thread_local real cacc[NB][6], cfrc[NB][6];
  v3set(cacc[0], 0, 0, 0);
  for (int i = 0; i < 3; i++) cacc[0][3 + i] = -real(m->gravity[i]);
  for (int i = 0; i < 6; i++) cfrc[0][i] = 0;
  for (int b = 1; b < m->nbody; b++) {
    for (int i = 0; i < 6; i++) cacc[b][i] = cacc[m->body_parent[b]][i];
    for (int k = 0; k < m->body_jntnum[b]; k++) {
      int j = m->body_jntadr[b] + k, da = m->jnt_dofadr[j], n = m->jnt_type[j] == KO_JNT_FREE ? 6 : 1;
      for (int t = 0; t < n; t++) for (int i = 0; i < 6; i++) cacc[b][i] += d->cdof_dot[da + t][i] * d->qvel[da + t];
    }
    real t1[6], t2[6];
    mul_inert_vec(cfrc[b], d->cinert[b], cacc[b]);
    mul_inert_vec(t1, d->cinert[b], d->cvel[b]);
    cross_force(t2, d->cvel[b], t1);
    for (int i = 0; i < 6; i++) cfrc[b][i] += t2[i];
  }
  for (int b = m->nbody - 1; b > 0; b--) {
    int p = m->body_parent[b];
    if (p > 0) for (int i = 0; i < 6; i++) cfrc[p][i] += cfrc[b][i];
  }
  for (int i = 0; i < m->nv; i++) {
    real s = 0;
    for (int k = 0; k < 6; k++) s += d->cdof[i][k] * cfrc[m->dof_bodyid[i]][k];
    d->qfrc_bias[i] = s;
  }
}
static void ko_reference_constraint(const ko_model* m, ko_data* d) {
  for (int r = 0; r < d->nefc; r++) {
    real v = 0;
    for (int j = 0; j < m->nv; j++) v += d->efc_J[r][j] * d->qvel[j];
    d->efc_vel[r] = v;
    d->efc_aref[r] = -d->efc_B[r] * v - d->efc_K[r] * d->efc_imp[r] * (d->efc_pos[r] - d->efc_margin[r]);
  }
}
static void ko_fwd_velocity(const ko_model* m, ko_data* d) {
  for (int i = 0; i < m->nu; i++) d->actuator_velocity[i] = d->qvel[m->jnt_dofadr[m->act_jntid[i]]];
  ko_com_vel(m, d);
  ko_rne(m, d);
  ko_reference_constraint(m, d);
}

/* ------------------------------------------------------------------------------------------- acceleration stage */
static void ko_fwd_actuation(const ko_model* m, ko_data* d, int disable) {
  for (int i = 0; i < m->nv; i++) d->qfrc_actuator[i] = 0;
  if (disable) { for (int i = 0; i < m->nu; i++) d->actuator_force[i] = 0; return; }
  for (int i = 0; i < m->nu; i++) {
    real c = d->ctrl[i];
    if (m->act_ctrllimited[i]) c = rclip(c, m->act_ctrlrange[2 * i], m->act_ctrlrange[2 * i + 1]);
    /* <position kp>: gain kp, bias (0, -kp, 0) */
    real f = real(m->act_kp[i]) * c - real(m->act_kp[i]) * d->actuator_length[i];
    if (m->act_forcelimited[i]) f = rclip(f, m->act_forcerange[2 * i], m->act_forcerange[2 * i + 1]);
    d->actuator_force[i] = f;
    d->qfrc_actuator[m->jnt_dofadr[m->act_jntid[i]]] += f;
  }
}
static void ko_fwd_acceleration(const ko_model* m, ko_data* d) {
  for (int i = 0; i < m->nv; i++) {
    d->qfrc_smooth[i] = d->qfrc_actuator[i] - d->qfrc_bias[i];
    d->qacc_smooth[i] = d->qfrc_smooth[i];
  }
  ko_solve_m(m, d, d->qacc_smooth);
}

/* ---- Newton solver (mj_solNewton / primal), SURVEY.md A5 */
struct ko_solver {
  int nv, nefc;
  real Ma[NV], jar[MAXEFC], grad[NV], Mgrad[NV], search[NV], Mv[NV], jv[MAXEFC];
  real gauss, cost;
  real H[NV][NV];
};
static void sol_update(const ko_model* m, ko_data* d, ko_solver* s) {
  int nv = s->nv;
  real cost = 0;
  for (int r = 0; r < s->nefc; r++) {
    real x = s->jar[r], Dr = d->efc_D[r];
    if (d->efc_type[r] == EFC_FRICTION) {
      real f = d->efc_frictionloss[r], rf = d->efc_R[r] * f;
      if (x <= -rf) { d->efc_state[r] = 2; d->efc_force[r] = f; cost += -f * (real(0.5) * rf + x); }
      else if (x >= rf) { d->efc_state[r] = 3; d->efc_force[r] = -f; cost += -f * (real(0.5) * rf - x); }
      else { d->efc_state[r] = 1; d->efc_force[r] = -Dr * x; cost += real(0.5) * Dr * x * x; }
    } else {
      if (x < real(0.0)) { d->efc_state[r] = 1; d->efc_force[r] = -Dr * x; cost += real(0.5) * Dr * x * x; }
      else { d->efc_state[r] = 0; d->efc_force[r] = 0; }
    }
  }
  for (int i = 0; i < nv; i++) {
    real q = 0;
    for (int r = 0; r < s->nefc; r++) q += d->efc_J[r][i] * d->efc_force[r];
    d->qfrc_constraint[i] = q;
  }
  real g = 0;
  for (int i = 0; i < nv; i++) g += (s->Ma[i] - d->qfrc_smooth[i]) * (d->qacc[i] - d->qacc_smooth[i]);
  s->gauss = real(0.5) * g;
  s->cost = s->gauss + cost;
  for (int i = 0; i < nv; i++) s->grad[i] = s->Ma[i] - d->qfrc_smooth[i] - d->qfrc_constraint[i];
}
static void sol_hessian_dir(const ko_model* m, ko_data* d, ko_solver* s) {
  int nv = s->nv;
  for (int i = 0; i < nv; i++) for (int j = 0; j < nv; j++) s->H[i][j] = d->qM[i][j];
  for (int r = 0; r < s->nefc; r++) if (d->efc_state[r] == 1) {
    real Dr = d->efc_D[r];
    for (int i = 0; i < nv; i++) if (d->efc_J[r][i] != real(0.0))
      for (int j = 0; j <= i; j++) s->H[i][j] += Dr * d->efc_J[r][i] * d->efc_J[r][j];
  }
  /* dense Cholesky H = L L^T in the lower triangle */
  for (int j = 0; j < nv; j++) {
    real t = s->H[j][j];
    for (int k = 0; k < j; k++) t -= s->H[j][k] * s->H[j][k];
    t = rsqrt_(rmax(t, real(KO_MINVAL)));
    s->H[j][j] = t;
    for (int i = j + 1; i < nv; i++) {
      real u = s->H[i][j];
      for (int k = 0; k < j; k++) u -= s->H[i][k] * s->H[j][k];
      s->H[i][j] = u / t;
    }
  }
  for (int i = 0; i < nv; i++) {
    real t = s->grad[i];
    for (int k = 0; k < i; k++) t -= s->H[i][k] * s->Mgrad[k];
    s->Mgrad[i] = t / s->H[i][i];
  }
  for (int i = nv - 1; i >= 0; i--) {
    real t = s->Mgrad[i];
    for (int k = i + 1; k < nv; k++) t -= s->H[k][i] * s->Mgrad[k];
    s->Mgrad[i] = t / s->H[i][i];
  }
}
/* 1-D cost along the search direction: value and first/second derivative at alpha */
static void ls_eval(const ko_data* d, const ko_solver* s, const real* qg, real alpha, real* d1, real* d2) {
  real q1 = qg[1], q2 = qg[2];
  for (int r = 0; r < s->nefc; r++) {
    real x = s->jar[r] + alpha * s->jv[r], Dr = d->efc_D[r];
    if (d->efc_type[r] == EFC_FRICTION) {
      real f = d->efc_frictionloss[r], rf = d->efc_R[r] * f;
      if (x <= -rf) { q1 += -f * s->jv[r]; continue; }
      if (x >= rf) { q1 += f * s->jv[r]; continue; }
    } else if (x >= real(0.0)) continue;
    q1 += Dr * s->jar[r] * s->jv[r];
    q2 += real(0.5) * Dr * s->jv[r] * s->jv[r];
  }
  *d1 = real(2.0) * alpha * q2 + q1;
  *d2 = real(2.0) * q2;
}
static real sol_linesearch(const ko_model* m, ko_data* d, ko_solver* s, real scale) {
  int nv = s->nv;
  real snorm = 0;
  for (int i = 0; i < nv; i++) snorm += s->search[i] * s->search[i];
  snorm = rsqrt_(snorm);
  if (snorm < real(KO_MINVAL)) return 0;
  for (int i = 0; i < nv; i++) {
    real t = 0;
    for (int j = 0; j < nv; j++) t += d->qM[i][j] * s->search[j];
    s->Mv[i] = t;
  }
  for (int r = 0; r < s->nefc; r++) {
    real t = 0;
    for (int j = 0; j < nv; j++) t += d->efc_J[r][j] * s->search[j];
    s->jv[r] = t;
  }
  real qg[3] = {s->gauss, 0, 0};
  for (int i = 0; i < nv; i++) {
    qg[1] += s->search[i] * (s->Ma[i] - d->qfrc_smooth[i]);
    qg[2] += real(0.5) * s->search[i] * s->Mv[i];
  }
  real d1, d2;
  ls_eval(d, s, qg, 0, &d1, &d2);
  d->ls_evals++;
  /* tolerance: MuJoCo's tolerance*ls_tolerance*snorm/scale, floored at rounding level of the slope at 0 */
  real gtol = rmax(real(m->tolerance * m->ls_tolerance) * snorm / scale, real(64.0 * 2.220446049250313e-16) * rabs(d1));
  if (rabs(d1) < gtol || d1 > real(0.0)) return 0;
  /* phase 1: Newton steps to the right until the slope changes sign */
  real lo = 0, lo_d1 = d1, lo_d2 = d2, hi = 0, hi_d1 = 0, hi_d2 = 0;
  int bracket = 0, it = 0;
  for (; it < m->ls_iterations; it++) {
    real a = lo - lo_d1 / lo_d2;
    ls_eval(d, s, qg, a, &d1, &d2);
    d->ls_evals++;
    if (rabs(d1) < gtol) return a;
    if (d1 > real(0.0)) { hi = a; hi_d1 = d1; hi_d2 = d2; bracket = 1; break; }
    lo = a; lo_d1 = d1; lo_d2 = d2;
  }
  if (!bracket) return lo;
  /* phase 2: safeguarded Newton inside [lo, hi] */
  for (; it < m->ls_iterations; it++) {
    real a = rabs(lo_d1) < rabs(hi_d1) ? lo - lo_d1 / lo_d2 : hi - hi_d1 / hi_d2;
    if (!(a > lo && a < hi)) a = real(0.5) * (lo + hi);
    if (a == lo || a == hi) break;
    ls_eval(d, s, qg, a, &d1, &d2);
    d->ls_evals++;
    if (rabs(d1) < gtol) return a;
    if (d1 > real(0.0)) { hi = a; hi_d1 = d1; hi_d2 = d2; } else { lo = a; lo_d1 = d1; lo_d2 = d2; }
  }
  return rabs(lo_d1) < rabs(hi_d1) ? lo : hi;
}
static real sol_total_cost(const ko_model* m, ko_data* d, ko_solver* s, const real* qacc) {
  int nv = s->nv;
  real cost = 0;
  for (int r = 0; r < s->nefc; r++) {
    real x = -d->efc_aref[r];
    for (int j = 0; j < nv; j++) x += d->efc_J[r][j] * qacc[j];
    real Dr = d->efc_D[r];
    if (d->efc_type[r] == EFC_FRICTION) {
      real f = d->efc_frictionloss[r], rf = d->efc_R[r] * f;
      if (x <= -rf) cost += -f * (real(0.5) * rf + x);
      else if (x >= rf) cost += -f * (real(0.5) * rf - x);
      else cost += real(0.5) * Dr * x * x;
    } else if (x < real(0.0)) cost += real(0.5) * Dr * x * x;
  }
  real g = 0;
  for (int i = 0; i < nv; i++) {
    real Ma = 0;
    for (int j = 0; j < nv; j++) Ma += d->qM[i][j] * qacc[j];
    g += (Ma - d->qfrc_smooth[i]) * (qacc[i] - d->qacc_smooth[i]);
  }
  return cost + real(0.5) * g;
}
static void ko_fwd_constraint(const ko_model* m, ko_data* d) {
  int nv = m->nv;
  static thread_local ko_solver S;
  ko_solver* s = &S;
  s->nv = nv; s->nefc = d->nefc;
  d->solver_niter = 0;
  if (d->nefc == 0) {
    for (int i = 0; i < nv; i++) { d->qacc[i] = d->qacc_smooth[i]; d->qacc_warmstart[i] = d->qacc_smooth[i]; d->qfrc_constraint[i] = 0; }
    return;
  }
  /* warm start: previous qacc if it is cheaper than the unconstrained acceleration */
  real cw = sol_total_cost(m, d, s, d->qacc_warmstart), cs = sol_total_cost(m, d, s, d->qacc_smooth);
  for (int i = 0; i < nv; i++) d->qacc[i] = cw > cs ? d->qacc_smooth[i] : d->qacc_warmstart[i];
  for (int i = 0; i < nv; i++) {
    real t = 0;
    for (int j = 0; j < nv; j++) t += d->qM[i][j] * d->qacc[j];
    s->Ma[i] = t;
  }
  for (int r = 0; r < s->nefc; r++) {
    real t = -d->efc_aref[r];
    for (int j = 0; j < nv; j++) t += d->efc_J[r][j] * d->qacc[j];
    s->jar[r] = t;
  }
  real scale = real(1.0) / (real(m->meaninertia) * real((double)(nv > 1 ? nv : 1)));
  sol_update(m, d, s);
  sol_hessian_dir(m, d, s);
  for (int i = 0; i < nv; i++) s->search[i] = -s->Mgrad[i];
  while (d->solver_niter < m->iterations) {
    real alpha = sol_linesearch(m, d, s, scale);
    if (alpha == real(0.0)) break;
    for (int i = 0; i < nv; i++) { d->qacc[i] += alpha * s->search[i]; s->Ma[i] += alpha * s->Mv[i]; }
    for (int r = 0; r < s->nefc; r++) s->jar[r] += alpha * s->jv[r];
    real oldcost = s->cost;
    sol_update(m, d, s);
    sol_hessian_dir(m, d, s);
    real gn = 0;
    for (int i = 0; i < nv; i++) gn += s->grad[i] * s->grad[i];
    real improvement = scale * (oldcost - s->cost), gradient = scale * rsqrt_(gn);
    d->solver_niter++;
    if (improvement < real(m->tolerance) || gradient < real(m->tolerance)) break;
    for (int i = 0; i < nv; i++) s->search[i] = -s->Mgrad[i];
  }
  for (int i = 0; i < nv; i++) d->qacc_warmstart[i] = d->qacc[i];
}

/* ------------------------------------------------------------------------------------------- mj_Euler */
static void ko_euler(const ko_model* m, ko_data* d) {
  real h = m->timestep;
  for (int i = 0; i < m->nv; i++) d->qvel[i] += h * d->qacc[i];
  for (int j = 0; j < m->njnt; j++) {
    int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
    if (m->jnt_type[j] == KO_JNT_FREE) {
      for (int i = 0; i < 3; i++) d->qpos[qa + i] += h * d->qvel[da + i];
      real w[3] = {d->qvel[da + 3], d->qvel[da + 4], d->qvel[da + 5]}, qrot[4], t[4];
      real angle = h * v3normalize(w);
      axisangle2quat(qrot, w, angle);
      q4normalize(d->qpos + qa + 3);
      q4mul(t, d->qpos + qa + 3, qrot);
      for (int i = 0; i < 4; i++) d->qpos[qa + 3 + i] = t[i];
    } else {
      d->qpos[qa] += h * d->qvel[da];
    }
  }
  d->time += h;
}

static void ko_step1(const ko_model* m, ko_data* d) { ko_fwd_position(m, d); ko_fwd_velocity(m, d); }
static void ko_step2(const ko_model* m, ko_data* d, int disable_actuation) {
  ko_fwd_actuation(m, d, disable_actuation);
  ko_fwd_acceleration(m, d);
  ko_fwd_constraint(m, d);
  ko_euler(m, d);
}
/* mj_forward with actuation disabled (dm_control reset_context) */
static void ko_forward_noact(const ko_model* m, ko_data* d) {
  ko_step1(m, d);
  ko_fwd_actuation(m, d, 1);
  ko_fwd_acceleration(m, d);
  ko_fwd_constraint(m, d);
}

/* ------------------------------------------------------------------------------------------- IK pieces */
/* ik_res (ik_mujoco.py:20-53): writes q into qpos[mask] (side effect kept) and runs mj_kinematics */
static void ko_ik_residual_(const ko_model* m, ko_data* d, const ko_task* t, int arm, const real* q, const real* goal_pos,
                            const real* goal_quat, const real* q_prev_mask, real* res) {
  int n = t->arm_nmask[arm], s = t->arm_site[arm];
  for (int i = 0; i < n; i++) d->qpos[t->arm_mask[arm][i]] = q[i];
  ko_kinematics(m, d);
  for (int i = 0; i < 3; i++) res[i] = d->site_xpos[s][i] - goal_pos[i];
  real cur[4], rq[3];
  mat2quat(cur, d->site_xmat[s]);
  subquat(rq, goal_quat, cur);
  for (int i = 0; i < 3; i++) res[3 + i] = rq[i] * real(K_IK_RES_RAD);
  for (int i = 0; i < n; i++) {
    res[6 + i] = real(K_IK_RES_REG_PREV) * (q[i] - q_prev_mask[i]);
    res[6 + n + i] = real(K_IK_RES_REG_HOME) * (q[i] - real(t->q_home[t->arm_mask[arm][i]]));
  }
}
/* ik_jac (ik_mujoco.py:56-97): rows [Jp; rad * D_ee^T R_site^T Jr; reg I; reg I], columns = mask.
   D_ee^T R_site^T Jr is taken as the true derivative of subQuat(goal, cur) w.r.t. q, i.e.
   -Jl^{-1}(phi) R_site^T Jr with phi = subQuat(goal, cur)  (DESIGN.md "mjd_subQuat convention";
   verified against finite differences of ko_ik_residual_ in tests). */
static void ko_ik_jacobian_(const ko_model* m, ko_data* d, const ko_task* t, int arm, const real* q, const real* goal_quat,
                            real* jac /* (6+2n) x n row-major */) {
  int n = t->arm_nmask[arm], s = t->arm_site[arm];
  for (int i = 0; i < n; i++) d->qpos[t->arm_mask[arm][i]] = q[i];
  ko_kinematics(m, d);
  ko_com_pos(m, d);
  static thread_local real jp[3][NV], jr[3][NV];
  ko_jac(m, d, jp, jr, d->site_xpos[s], m->site_bodyid[s]);
  real cur[4], phi[3];
  mat2quat(cur, d->site_xmat[s]);
  subquat(phi, goal_quat, cur);
  real u[3] = {phi[0], phi[1], phi[2]};
  real half = real(0.5) * v3normalize(u);
  /* Jl^{-1}(phi) = I - half*[u]x + coef*[u]x^2 */
  real coef = real(1.0) - (half < real(6e-8) ? real(1.0) : half / rtan(half));
  real K[9] = {0, -u[2], u[1], u[2], 0, -u[0], -u[1], u[0], 0}, Dm[9];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    real kk = 0;
    for (int k = 0; k < 3; k++) kk += K[3 * i + k] * K[3 * k + j];
    Dm[3 * i + j] = real(i == j ? 1.0 : 0.0) - half * K[3 * i + j] + coef * kk;
  }
  const real* R = d->site_xmat[s];
  for (int c = 0; c < n; c++) {
    int dof = m->jnt_dofadr[t->arm_mask[arm][c]];   /* hinge/slide joints: qpos index == joint index for the first q_len joints */
    real jr_w[3] = {jr[0][dof], jr[1][dof], jr[2][dof]}, jr_l[3], o[3];
    m3Tmulv(jr_l, R, jr_w);
    m3mulv(o, Dm, jr_l);
    for (int i = 0; i < 3; i++) {
      jac[i * n + c] = jp[i][dof];
      jac[(3 + i) * n + c] = -real(K_IK_JAC_RAD) * o[i];
    }
    for (int i = 0; i < n; i++) {
      jac[(6 + i) * n + c] = i == c ? real(K_IK_JAC_REG) : real(0.0);
      jac[(6 + n + i) * n + c] = i == c ? real(K_IK_JAC_REG) : real(0.0);
    }
  }
}
/* Device IK restated: fixed-count damped Gauss-Newton on the reference residual/Jacobian, clipped to
   the joint range each iteration.  Mirrors the reference's bounds check (ik_mujoco.py:128-138 with scipy
   raising on an infeasible x0: IK is skipped and qpos[mask] returned). */
static void ko_ik_dls(const ko_model* m, ko_data* d, const ko_task* t, int arm, const real* goal_pos, const real* goal_quat,
                      const real* q_prev_full, real* q_out) {
  int n = t->arm_nmask[arm], nr = 6 + 2 * n;
  real x[KO_MAXMASK], lo[KO_MAXMASK], hi[KO_MAXMASK], qprev[KO_MAXMASK], saved[KO_MAXMASK];
  int feasible = 1;
  for (int i = 0; i < n; i++) {
    int j = t->arm_mask[arm][i];
    x[i] = d->qpos[j]; saved[i] = x[i]; qprev[i] = q_prev_full[j];
    lo[i] = m->jnt_range[2 * j]; hi[i] = m->jnt_range[2 * j + 1];
    if (x[i] < lo[i] || x[i] > hi[i]) feasible = 0;
  }
  if (feasible) {
    /* Projected Levenberg-Marquardt on the reference's stationarity condition J^T r = 0 (J, r exactly as
       ik_jac / ik_res build them, including their mismatched regulariser weights).
       Iteration matrix: Jpose^T Jpose + (lam + mu) I with lam = IK_JAC_REG*(IK_RES_REG_PREV + IK_RES_REG_HOME),
       i.e. J^T (dr/dq); with the reference's own J^T J (lam = 2*IK_JAC_REG^2) the null-space error only
       contracts by 2/3 per iteration, the fixed point is the same.  Coordinates on a bound with the gradient
       pushing outward are frozen for the iteration (the KKT point scipy's bounded TRF converges to).  A step
       is kept only if 0.5|r|^2 does not increase (TRF's acceptance test); otherwise mu grows. */
    const real lam = real(K_IK_JAC_REG * (K_IK_RES_REG_PREV + K_IK_RES_REG_HOME));
    real r[6 + 2 * KO_MAXMASK], J[(6 + 2 * KO_MAXMASK) * KO_MAXMASK], rn[6 + 2 * KO_MAXMASK], xn[KO_MAXMASK];
    real A[KO_MAXMASK][KO_MAXMASK], g[KO_MAXMASK], mu = 0, cost = 0;
    ko_ik_residual_(m, d, t, arm, x, goal_pos, goal_quat, qprev, r);
    ko_ik_jacobian_(m, d, t, arm, x, goal_quat, J);
    for (int k = 0; k < nr; k++) cost += real(0.5) * r[k] * r[k];
    for (int it = 0; it < t->ik_iters; it++) {
      int active[KO_MAXMASK];
      for (int i = 0; i < n; i++) {
        real gi = 0;
        for (int k = 0; k < nr; k++) gi += J[k * n + i] * r[k];
        g[i] = -gi;
        active[i] = (x[i] <= lo[i] && gi > real(0.0)) || (x[i] >= hi[i] && gi < real(0.0));
        for (int j = 0; j <= i; j++) {
          real a = 0;
          for (int k = 0; k < 6; k++) a += J[k * n + i] * J[k * n + j];
          A[i][j] = a + (i == j ? lam + mu : real(0.0));
        }
      }
      for (int i = 0; i < n; i++) if (active[i]) {
        for (int j = 0; j < n; j++) { if (j <= i) A[i][j] = 0; else A[j][i] = 0; }
        A[i][i] = 1; g[i] = 0;
      }
      for (int j = 0; j < n; j++) {   /* Cholesky + solve */
        real tt = A[j][j];
        for (int k = 0; k < j; k++) tt -= A[j][k] * A[j][k];
        tt = rsqrt_(tt);
        A[j][j] = tt;
        for (int i = j + 1; i < n; i++) {
          real u = A[i][j];
          for (int k = 0; k < j; k++) u -= A[i][k] * A[j][k];
          A[i][j] = u / tt;
        }
      }
      for (int i = 0; i < n; i++) { real tt = g[i]; for (int k = 0; k < i; k++) tt -= A[i][k] * g[k]; g[i] = tt / A[i][i]; }
      for (int i = n - 1; i >= 0; i--) { real tt = g[i]; for (int k = i + 1; k < n; k++) tt -= A[k][i] * g[k]; g[i] = tt / A[i][i]; }
      for (int i = 0; i < n; i++) xn[i] = rclip(x[i] + g[i], lo[i], hi[i]);
      ko_ik_residual_(m, d, t, arm, xn, goal_pos, goal_quat, qprev, rn);
      real costn = 0;
      for (int k = 0; k < nr; k++) costn += real(0.5) * rn[k] * rn[k];
      if (costn <= cost) {
        for (int i = 0; i < n; i++) x[i] = xn[i];
        for (int k = 0; k < nr; k++) r[k] = rn[k];
        cost = costn;
        ko_ik_jacobian_(m, d, t, arm, x, goal_quat, J);
        mu = mu * real(0.25);
        if (mu < real(1e-6)) mu = 0;
      } else {
        mu = mu == real(0.0) ? real(1e-4) : mu * real(4.0);
      }
    }
    /* leave the model at the solution, as the reference's last evaluation normally does */
    for (int i = 0; i < n; i++) d->qpos[t->arm_mask[arm][i]] = t->ik_teleport ? x[i] : saved[i];
    ko_kinematics(m, d);
  }
  for (int i = 0; i < n; i++) q_out[i] = rclip(x[i], lo[i], hi[i]);   /* ik_mujoco.py:147-152 */
}

/* scipy Rotation.from_matrix(R).as_euler("xyz") (extrinsic) then from_euler("xyz", e).as_quat()[[3,0,1,2]] */
static void mat_to_euler_xyz_ext(real* e, const real* R) {
  /* R = Rz(c) Ry(b) Rx(a) */
  real sb = -R[6];
  sb = rclip(sb, -1.0, 1.0);
  e[1] = rasin(sb);
  e[0] = ratan2(R[7], R[8]);
  e[2] = ratan2(R[3], R[0]);
}
static void euler_xyz_ext_to_quat(real* q, const real* e) {
  real qx[4] = {rcos(e[0] * real(0.5)), rsin(e[0] * real(0.5)), 0, 0};
  real qy[4] = {rcos(e[1] * real(0.5)), 0, rsin(e[1] * real(0.5)), 0};
  real qz[4] = {rcos(e[2] * real(0.5)), 0, 0, rsin(e[2] * real(0.5))};
  real t[4];
  q4mul(t, qy, qx);
  q4mul(q, qz, t);
  /* scipy returns a canonical sign only when asked; as_quat() keeps the product's sign */
}

/* ------------------------------------------------------------------------------------------- task */
typedef void (*ko_ik_fn)(const ko_model*, ko_data*, const ko_task*, int arm, const real* goal_pos, const real* goal_quat,
                         const real* q_prev_full, real* q_out, void* user);
static void ko_ik_default(const ko_model* m, ko_data* d, const ko_task* t, int arm, const real* gp, const real* gq,
                          const real* qprev, real* q_out, void*) {
  ko_ik_dls(m, d, t, arm, gp, gq, qprev, q_out);
}

/* before_step (env_sim.py:38-108).  action is float32 as in the reference's action space. */
static void ko_before_step(const ko_model* m, ko_data* d, const ko_task* t, const float* action, ko_ik_fn ik, void* user) {
  real q_pos[NQ];
  float ctrl[NU];
  for (int i = 0; i < m->nq; i++) q_pos[i] = d->qpos[i];
  for (int i = 0; i < m->nu; i++) ctrl[i] = (float)D(d->ctrl[i]);            /* :40 astype(float32) */
  /* grippers first (:41-59); the reference tests grip_r before grip_l, the two are independent */
  for (int a = 0; a < t->n_arm; a++) if (t->off_grip[a] >= 0) {
    float g = action[t->off_grip[a]] * K_EE_S_DELTA_F;                       /* float32 product */
    g = (float)((double)g + D(d->qpos[t->arm_grip[a][0]]));                 /* += float64 scalar, stored float32 */
    g = g < K_EE_S_MIN_F ? K_EE_S_MIN_F : (g > K_EE_S_MAX_F ? K_EE_S_MAX_F : g);
    ctrl[t->arm_grip[a][0]] = g;
    ctrl[t->arm_grip[a][1]] = g;
  }
  for (int a = 0; a < t->n_arm; a++) if (t->act_mode == 0 && t->off_pos[a] >= 0) {
    int s = t->arm_site[a], n = t->arm_nmask[a];
    real goal_pos[3], eul[3], goal_quat[4], q_sol[KO_MAXMASK];
    for (int i = 0; i < 3; i++) goal_pos[i] = real((double)action[t->off_pos[a] + i] * K_EE_POS_DELTA) + d->site_xpos[s][i];
    mat_to_euler_xyz_ext(eul, d->site_xmat[s]);
    for (int i = 0; i < 3; i++) eul[i] = real((double)action[t->off_orn[a] + i] * K_EE_ORN_DELTA) + eul[i];
    euler_xyz_ext_to_quat(goal_quat, eul);
    int k = t->arm_mocap[a];
    for (int i = 0; i < 3; i++) d->mocap_pos[k][i] = goal_pos[i];
    for (int i = 0; i < 4; i++) d->mocap_quat[k][i] = goal_quat[i];
    ik(m, d, t, a, goal_pos, goal_quat, q_pos, q_sol, user);
    for (int i = 0; i < n; i++) ctrl[t->arm_mask[a][i]] = (float)D(q_sol[i]);
  }
  for (int a = 0; a < t->n_arm; a++) if (t->act_mode == 1 && t->off_q[a] >= 0) {
    for (int i = 0; i < t->arm_nmask[a]; i++) {
      float da = action[t->off_q[a] + i] * K_Q_POS_DELTA_F;
      ctrl[t->arm_mask[a][i]] = (float)(D(q_pos[t->arm_mask[a][i]]) + (double)da);
    }
  }
  for (int i = 0; i < m->nu; i++) d->ctrl[i] = real((double)ctrl[i]);         /* CTRL_ALPHA = 1 (:106) */
}

static int ko_obs_dim(const ko_task* t) { return 2 * t->q_len + 7; }
/* get_observation (env_sim.py:110-146): [q_pos(q_len), q_vel(q_len), cube_pos(3), cube_orn(4)] */
static void ko_observation(const ko_model* m, const ko_data* d, const ko_task* t, double* obs) {
  int n = t->q_len;
  for (int i = 0; i < n; i++) {
    real lo = m->jnt_range[2 * i], hi = m->jnt_range[2 * i + 1];
    obs[i] = D(rclip((d->qpos[i] - lo) / (hi - lo), -1.0, 1.0));
    obs[n + i] = D(rclip(d->qvel[i] / real(K_MAX_Q_VEL), -1.0, 1.0));
  }
  for (int i = 0; i < 3; i++) {
    real lo = t->cube_spawn_lo[i], hi = t->cube_spawn_hi[i];
    obs[2 * n + i] = D(rclip((d->qpos[m->nq - 7 + i] - lo) / (hi - lo), -1.0, 1.0));
  }
  for (int i = 0; i < 4; i++) obs[2 * n + 3 + i] = D(d->qpos[m->nq - 4 + i]);
}
/* get_reward (env_sim.py:148-179).  The touch/lift bonuses need geoms named *_gripper_finger as the
   contact's second geom; no such geom exists, so they never fire.  flags reports what the
   contact scan saw: bit0 cube-table, bit1 right pads, bit2 left pads. */
static double ko_reward(const ko_model* m, const ko_data* d, const ko_task* t, int* flags) {
  real r = 0, vn = 0;
  for (int i = 0; i < m->nv; i++) vn += d->qvel[i] * d->qvel[i];
  r -= real(K_REWARD_VEL_PENALTY) * rsqrt_(vn);
  for (int a = t->n_arm - 1; a >= 0; a--) {     /* grip_l term is added before grip_r (:155-161) */
    if (t->off_grip[a] < 0) continue;
    real dif[3];
    for (int i = 0; i < 3; i++) dif[i] = d->xpos[t->cube_body][i] - d->xpos[t->arm_eebody[a]][i];
    r += real(K_REWARD_GRIP_DIST) * (real(1.0) / (v3norm(dif) + real(K_EPSILON)));
  }
  int f = 0;
  for (int c = 0; c < d->ncon; c++) {
    int g1 = d->contact[c].geom1;
    if (m->geom_type[g1] == KO_GEOM_PLANE) f |= 1;
    else {
      int b = m->geom_bodyid[g1], arm = 0;
      for (int a = 0; a < t->n_arm; a++) {      /* pad belongs to the arm whose ee body shares its parent chain */
        int e = t->arm_eebody[a];
        for (int p = e; p > 0; p = m->body_parent[p]) if (p == m->body_parent[b]) arm = a;
      }
      f |= arm == 0 ? 2 : 4;
    }
  }
  if (flags) *flags = f;
  return D(r);
}

/* ------------------------------------------------------------------------------------------- RNG (Philox4x32-10) */
static inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
/* cube spawn: three uniforms in [0,1) with 24 random bits, keyed by (seed, global env id, episode) */
static void ko_spawn_uniforms(uint64_t seed, uint64_t env_id, uint32_t episode, double u[3]) {
  uint32_t c[4] = {(uint32_t)env_id, (uint32_t)(env_id >> 32), episode, 0x4b4d414eu};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  for (int i = 0; i < 3; i++) u[i] = (double)(c[i] >> 8) * (1.0 / 16777216.0);
}

/* initialize_episode (env_sim.py:23-36) + the two mj_forward(actuation disabled) of reset_context */
static void ko_env_reset_(const ko_model* m, ko_data* d, const ko_task* t, const double* cube_xyz) {
  for (int i = 0; i < m->nq; i++) d->qpos[i] = m->qpos0[i];
  for (int i = 0; i < m->nv; i++) { d->qvel[i] = 0; d->qacc_warmstart[i] = 0; }
  for (int i = 0; i < m->nu; i++) d->ctrl[i] = 0;
  d->time = 0;
  for (int k = 0; k < m->nmocap; k++) {
    for (int i = 0; i < 3; i++) d->mocap_pos[k][i] = m->mocap_pos0[3 * k + i];
    for (int i = 0; i < 4; i++) d->mocap_quat[k][i] = m->mocap_quat0[4 * k + i];
  }
  for (int i = 0; i < t->q_len; i++) { d->qpos[i] = t->q_home[i]; d->ctrl[i] = t->q_home[i]; }
  for (int i = 0; i < 3; i++) d->qpos[t->cube_qposadr + i] = cube_xyz[i];
  /* The reference leaves reset_context through mj_forward(actuation disabled), which also seeds
     qacc_warmstart with that forward pass's qacc.  Here the seed is left at zero (it only chooses the
     Newton starting point; the minimiser is unique) and only the position/velocity stages are refreshed. */
  ko_step1(m, d);
}

/* one env step: before_step; mj_step2; mj_step x (n_sub-1); mj_step1 */
static void ko_env_step_(const ko_model* m, ko_data* d, const ko_task* t, const float* action, ko_ik_fn ik, void* user) {
  int nsub = t->n_sub_steps > 0 ? t->n_sub_steps : (int)std::lround(K_CONTROL_TIMESTEP / m->timestep);
  ko_before_step(m, d, t, action, ik ? ik : ko_ik_default, user);
  ko_step2(m, d, 0);
  for (int i = 0; i < nsub - 1; i++) { ko_step1(m, d); ko_step2(m, d, 0); }
  ko_step1(m, d);
}

/* =========================================================================================== C entry points */
extern "C" {

size_t ko_sizeof_data(void) { return sizeof(ko_data); }
int ko_scalar_is_counting(void) {
#ifdef KO_COUNT_FLOPS
  return 1;
#else
  return 0;
#endif
}
unsigned long long ko_flops_get(void) {
#ifdef KO_COUNT_FLOPS
  return g_flops;
#else
  return 0;
#endif
}
void ko_flops_reset(void) {
#ifdef KO_COUNT_FLOPS
  g_flops = 0;
#endif
}

void* ko_data_new(const ko_model* m) {
  ko_data* d = (ko_data*)malloc(sizeof(ko_data));
  ko_init(m, d);
  return d;
}
void ko_data_free(void* d) { free(d); }

static void load_state(const ko_model* m, ko_data* d, const double* qpos, const double* qvel, const double* ctrl,
                       const double* warm, double time, const double* mocap) {
  for (int i = 0; i < m->nq; i++) d->qpos[i] = qpos[i];
  for (int i = 0; i < m->nv; i++) { d->qvel[i] = qvel[i]; d->qacc_warmstart[i] = warm ? warm[i] : 0.0; }
  for (int i = 0; i < m->nu; i++) d->ctrl[i] = ctrl[i];
  d->time = time;
  if (mocap) for (int k = 0; k < m->nmocap; k++) {
    for (int i = 0; i < 3; i++) d->mocap_pos[k][i] = mocap[7 * k + i];
    for (int i = 0; i < 4; i++) d->mocap_quat[k][i] = mocap[7 * k + 3 + i];
  }
}
/* set the state and refresh the position/velocity stages (what mj_step1 left behind in the reference) */
void ko_set_state(const ko_model* m, void* dv, const double* qpos, const double* qvel, const double* ctrl,
                  const double* warm, double time, const double* mocap) {
  ko_data* d = (ko_data*)dv;
  load_state(m, d, qpos, qvel, ctrl, warm, time, mocap);
  ko_step1(m, d);
}
void ko_get_state(const ko_model* m, const void* dv, double* qpos, double* qvel, double* ctrl, double* warm, double* time,
                  double* mocap) {
  const ko_data* d = (const ko_data*)dv;
  for (int i = 0; i < m->nq; i++) qpos[i] = D(d->qpos[i]);
  for (int i = 0; i < m->nv; i++) { qvel[i] = D(d->qvel[i]); if (warm) warm[i] = D(d->qacc_warmstart[i]); }
  for (int i = 0; i < m->nu; i++) ctrl[i] = D(d->ctrl[i]);
  if (time) *time = D(d->time);
  if (mocap) for (int k = 0; k < m->nmocap; k++) {
    for (int i = 0; i < 3; i++) mocap[7 * k + i] = D(d->mocap_pos[k][i]);
    for (int i = 0; i < 4; i++) mocap[7 * k + 3 + i] = D(d->mocap_quat[k][i]);
  }
}
void ko_reset(const ko_model* m, void* dv, const ko_task* t, const double* cube_xyz) { ko_env_reset_(m, (ko_data*)dv, t, cube_xyz); }
void ko_spawn(const ko_task* t, unsigned long long seed, unsigned long long env_id, unsigned int episode, double* xyz) {
  double u[3];
  ko_spawn_uniforms(seed, env_id, episode, u);
  for (int i = 0; i < 3; i++) xyz[i] = t->cube_spawn_lo[i] + u[i] * (t->cube_spawn_hi[i] - t->cube_spawn_lo[i]);
}

/* python-side IK hook (scipy TRF): callback(arm, goal_pos[3], goal_quat[4], q_prev[nq], q_out[n]) */
typedef void (*ko_py_ik)(int arm, const double* goal_pos, const double* goal_quat, const double* q_prev, double* q_out);
static void ik_trampoline(const ko_model* m, ko_data* d, const ko_task* t, int arm, const real* gp, const real* gq,
                          const real* qprev, real* q_out, void* user) {
  double a[3], b[4], c[NQ], o[KO_MAXMASK];
  for (int i = 0; i < 3; i++) a[i] = D(gp[i]);
  for (int i = 0; i < 4; i++) b[i] = D(gq[i]);
  for (int i = 0; i < m->nq; i++) c[i] = D(qprev[i]);
  ((ko_py_ik)user)(arm, a, b, c, o);
  for (int i = 0; i < t->arm_nmask[arm]; i++) q_out[i] = o[i];
}
void ko_env_step(const ko_model* m, void* dv, const ko_task* t, const float* action, ko_py_ik py_ik) {
  ko_env_step_(m, (ko_data*)dv, t, action, py_ik ? ik_trampoline : (ko_ik_fn)0, (void*)py_ik);
}
void ko_before_step_only(const ko_model* m, void* dv, const ko_task* t, const float* action, ko_py_ik py_ik) {
  ko_before_step(m, (ko_data*)dv, t, action, py_ik ? ik_trampoline : ko_ik_default, (void*)py_ik);
}
/* single physics sub-step pieces, for teacher-forced parity of the sub-step kernel stages */
void ko_mj_step(const ko_model* m, void* dv) { ko_step1(m, (ko_data*)dv); ko_step2(m, (ko_data*)dv, 0); }
void ko_mj_step1(const ko_model* m, void* dv) { ko_step1(m, (ko_data*)dv); }
void ko_mj_step2(const ko_model* m, void* dv) { ko_step2(m, (ko_data*)dv, 0); }
void ko_mj_forward(const ko_model* m, void* dv, int disable_actuation) {
  ko_data* d = (ko_data*)dv;
  ko_step1(m, d); ko_fwd_actuation(m, d, disable_actuation); ko_fwd_acceleration(m, d); ko_fwd_constraint(m, d);
}
int ko_obs_size(const ko_task* t) { return ko_obs_dim(t); }
void ko_get_obs(const ko_model* m, const void* dv, const ko_task* t, double* obs) { ko_observation(m, (const ko_data*)dv, t, obs); }
double ko_get_reward(const ko_model* m, const void* dv, const ko_task* t, int* flags) { return ko_reward(m, (const ko_data*)dv, t, flags); }

/* IK residual / Jacobian exposed so that the real scipy.optimize.least_squares can drive them */
void ko_ik_residual(const ko_model* m, void* dv, const ko_task* t, int arm, const double* q, const double* goal_pos,
                    const double* goal_quat, const double* q_prev_mask, double* res) {
  int n = t->arm_nmask[arm];
  real qq[KO_MAXMASK], gp[3], gq[4], qp[KO_MAXMASK], r[6 + 2 * KO_MAXMASK];
  for (int i = 0; i < n; i++) { qq[i] = q[i]; qp[i] = q_prev_mask[i]; }
  for (int i = 0; i < 3; i++) gp[i] = goal_pos[i];
  for (int i = 0; i < 4; i++) gq[i] = goal_quat[i];
  ko_ik_residual_(m, (ko_data*)dv, t, arm, qq, gp, gq, qp, r);
  for (int i = 0; i < 6 + 2 * n; i++) res[i] = D(r[i]);
}
void ko_ik_jacobian(const ko_model* m, void* dv, const ko_task* t, int arm, const double* q, const double* goal_quat, double* jac) {
  int n = t->arm_nmask[arm];
  real qq[KO_MAXMASK], gq[4], J[(6 + 2 * KO_MAXMASK) * KO_MAXMASK];
  for (int i = 0; i < n; i++) qq[i] = q[i];
  for (int i = 0; i < 4; i++) gq[i] = goal_quat[i];
  ko_ik_jacobian_(m, (ko_data*)dv, t, arm, qq, gq, J);
  for (int i = 0; i < (6 + 2 * n) * n; i++) jac[i] = D(J[i]);
}
void ko_ik_solve_dls(const ko_model* m, void* dv, const ko_task* t, int arm, const double* goal_pos, const double* goal_quat,
                     const double* q_prev_full, double* q_out) {
  real gp[3], gq[4], qp[NQ], qo[KO_MAXMASK];
  for (int i = 0; i < 3; i++) gp[i] = goal_pos[i];
  for (int i = 0; i < 4; i++) gq[i] = goal_quat[i];
  for (int i = 0; i < m->nq; i++) qp[i] = q_prev_full[i];
  ko_ik_dls(m, (ko_data*)dv, t, arm, gp, gq, qp, qo);
  for (int i = 0; i < t->arm_nmask[arm]; i++) q_out[i] = D(qo[i]);
}
void ko_euler_pieces(const double* mat9, double* eul3, double* quat4) {
  real R[9], e[3], q[4];
  for (int i = 0; i < 9; i++) R[i] = mat9[i];
  mat_to_euler_xyz_ext(e, R);
  euler_xyz_ext_to_quat(q, e);
  for (int i = 0; i < 3; i++) eul3[i] = D(e[i]);
  for (int i = 0; i < 4; i++) quat4[i] = D(q[i]);
}
void ko_quat_pieces(const double* mat9, const double* qa, const double* qb, double* quat_of_mat, double* sub3) {
  real R[9], a[4], b[4], q[4], s[3];
  for (int i = 0; i < 9; i++) R[i] = mat9[i];
  for (int i = 0; i < 4; i++) { a[i] = qa[i]; b[i] = qb[i]; }
  mat2quat(q, R);
  subquat(s, a, b);
  for (int i = 0; i < 4; i++) quat_of_mat[i] = D(q[i]);
  for (int i = 0; i < 3; i++) sub3[i] = D(s[i]);
}

/* field access for tests (names follow mjData) */
int ko_get_field(const ko_model* m, const void* dv, const char* name, double* out, int cap) {
  const ko_data* d = (const ko_data*)dv;
  int n = 0;
#define PUT(x) do { if (n < cap) out[n] = D(x); n++; } while (0)
  if (!strcmp(name, "xpos")) { for (int b = 0; b < m->nbody; b++) for (int i = 0; i < 3; i++) PUT(d->xpos[b][i]); }
  else if (!strcmp(name, "xquat")) { for (int b = 0; b < m->nbody; b++) for (int i = 0; i < 4; i++) PUT(d->xquat[b][i]); }
  else if (!strcmp(name, "xipos")) { for (int b = 0; b < m->nbody; b++) for (int i = 0; i < 3; i++) PUT(d->xipos[b][i]); }
  else if (!strcmp(name, "site_xpos")) { for (int b = 0; b < m->nsite; b++) for (int i = 0; i < 3; i++) PUT(d->site_xpos[b][i]); }
  else if (!strcmp(name, "site_xmat")) { for (int b = 0; b < m->nsite; b++) for (int i = 0; i < 9; i++) PUT(d->site_xmat[b][i]); }
  else if (!strcmp(name, "geom_xpos")) { for (int b = 0; b < m->ngeom; b++) for (int i = 0; i < 3; i++) PUT(d->geom_xpos[b][i]); }
  else if (!strcmp(name, "subtree_com")) { for (int b = 0; b < m->nbody; b++) for (int i = 0; i < 3; i++) PUT(d->subtree_com[b][i]); }
  else if (!strcmp(name, "qM")) { for (int i = 0; i < m->nv; i++) for (int j = 0; j < m->nv; j++) PUT(d->qM[i][j]); }
  else if (!strcmp(name, "qfrc_bias")) { for (int i = 0; i < m->nv; i++) PUT(d->qfrc_bias[i]); }
  else if (!strcmp(name, "qfrc_actuator")) { for (int i = 0; i < m->nv; i++) PUT(d->qfrc_actuator[i]); }
  else if (!strcmp(name, "qfrc_smooth")) { for (int i = 0; i < m->nv; i++) PUT(d->qfrc_smooth[i]); }
  else if (!strcmp(name, "qfrc_constraint")) { for (int i = 0; i < m->nv; i++) PUT(d->qfrc_constraint[i]); }
  else if (!strcmp(name, "qacc_smooth")) { for (int i = 0; i < m->nv; i++) PUT(d->qacc_smooth[i]); }
  else if (!strcmp(name, "qacc")) { for (int i = 0; i < m->nv; i++) PUT(d->qacc[i]); }
  else if (!strcmp(name, "cvel")) { for (int b = 0; b < m->nbody; b++) for (int i = 0; i < 6; i++) PUT(d->cvel[b][i]); }
  else if (!strcmp(name, "efc_J")) { for (int r = 0; r < d->nefc; r++) for (int j = 0; j < m->nv; j++) PUT(d->efc_J[r][j]); }
  else if (!strcmp(name, "efc_force")) { for (int r = 0; r < d->nefc; r++) PUT(d->efc_force[r]); }
  else if (!strcmp(name, "efc_aref")) { for (int r = 0; r < d->nefc; r++) PUT(d->efc_aref[r]); }
  else if (!strcmp(name, "efc_D")) { for (int r = 0; r < d->nefc; r++) PUT(d->efc_D[r]); }
  else if (!strcmp(name, "efc_R")) { for (int r = 0; r < d->nefc; r++) PUT(d->efc_R[r]); }
  else if (!strcmp(name, "efc_pos")) { for (int r = 0; r < d->nefc; r++) PUT(d->efc_pos[r]); }
  else if (!strcmp(name, "efc_type")) { for (int r = 0; r < d->nefc; r++) PUT(real((double)d->efc_type[r])); }
  else if (!strcmp(name, "efc_state")) { for (int r = 0; r < d->nefc; r++) PUT(real((double)d->efc_state[r])); }
  else if (!strcmp(name, "efc_frictionloss")) { for (int r = 0; r < d->nefc; r++) PUT(d->efc_frictionloss[r]); }
  else if (!strcmp(name, "contact_dist")) { for (int c = 0; c < d->ncon; c++) PUT(d->contact[c].dist); }
  else if (!strcmp(name, "contact_pos")) { for (int c = 0; c < d->ncon; c++) for (int i = 0; i < 3; i++) PUT(d->contact[c].pos[i]); }
  else if (!strcmp(name, "contact_frame")) { for (int c = 0; c < d->ncon; c++) for (int i = 0; i < 9; i++) PUT(d->contact[c].frame[i]); }
  else if (!strcmp(name, "contact_geoms")) { for (int c = 0; c < d->ncon; c++) { PUT(real((double)d->contact[c].geom1)); PUT(real((double)d->contact[c].geom2)); } }
  else if (!strcmp(name, "solver_niter")) { PUT(real((double)d->solver_niter)); }
  else if (!strcmp(name, "ls_evals")) { PUT(real((double)d->ls_evals)); }
  else if (!strcmp(name, "ncon")) { PUT(real((double)d->ncon)); }
  else if (!strcmp(name, "nefc")) { PUT(real((double)d->nefc)); }
  else return -1;
#undef PUT
  return n;
}

/* Batched rollout leg used as the CPU baseline and by batch parity tests.
   State arrays are [n][dim] row-major.  For each env: optional autoreset bookkeeping identical to the
   device path (truncate at max_episode_steps, respawn with Philox(seed, env0+i, episode)). */
int ko_batch_step(const ko_model* m, const ko_task* t, int n, double* qpos, double* qvel, double* ctrl, double* warm,
                  double* time, int* step_count, int* episode, double* mocap, const float* action, double* obs,
                  double* final_obs, double* reward, unsigned char* truncated, int* con_flags, int* ncon,
                  int* con_geoms /* [n][2*MAXCON] */, int autoreset, unsigned long long seed, unsigned long long env0, int nthreads,
                  int* ncon_peak /* optional [n]: most contacts in any collision pass of the step, incl. the opening one */) {
  int od = ko_obs_dim(t);
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
  {
    ko_data* d = (ko_data*)malloc(sizeof(ko_data));
    ko_init(m, d);
#pragma omp for schedule(static)
    for (int e = 0; e < n; e++) {
      load_state(m, d, qpos + (size_t)e * m->nq, qvel + (size_t)e * m->nv, ctrl + (size_t)e * m->nu, warm + (size_t)e * m->nv,
                 time[e], mocap ? mocap + (size_t)e * 7 * m->nmocap : 0);
      d->ncon_peak = 0;
      ko_step1(m, d);
      ko_env_step_(m, d, t, action + (size_t)e * t->act_dim, 0, 0);
      int fl;
      reward[e] = ko_reward(m, d, t, &fl);
      if (con_flags) con_flags[e] = fl;
      if (ncon) ncon[e] = d->ncon;
      if (ncon_peak) ncon_peak[e] = d->ncon_peak;
      if (con_geoms) for (int c = 0; c < MAXCON; c++) {
        con_geoms[(size_t)e * 2 * MAXCON + 2 * c] = c < d->ncon ? d->contact[c].geom1 : -1;
        con_geoms[(size_t)e * 2 * MAXCON + 2 * c + 1] = c < d->ncon ? d->contact[c].geom2 : -1;
      }
      ko_observation(m, d, t, obs + (size_t)e * od);
      step_count[e] += 1;
      truncated[e] = step_count[e] >= t->max_episode_steps;
      if (autoreset && truncated[e]) {
        episode[e] += 1;
        double xyz[3];
        ko_spawn(t, seed, env0 + (unsigned long long)e, (unsigned)episode[e], xyz);
        ko_env_reset_(m, d, t, xyz);
        step_count[e] = 0;
        /* same-step autoreset: obs carries the first observation of the new episode */
        if (final_obs) memcpy(final_obs + (size_t)e * od, obs + (size_t)e * od, sizeof(double) * od);
        ko_observation(m, d, t, obs + (size_t)e * od);
      }
      double tm;
      ko_get_state(m, d, qpos + (size_t)e * m->nq, qvel + (size_t)e * m->nv, ctrl + (size_t)e * m->nu, warm + (size_t)e * m->nv, &tm,
                   mocap ? mocap + (size_t)e * 7 * m->nmocap : 0);
      time[e] = tm;
    }
    free(d);
  }
  return 0;
}
int ko_max_contacts(void) { return MAXCON; }

} /* extern "C" */
