// km_model.cuh -- device-side model tables and the per-environment working set.
//
// The articulated part of every KManip scene is a set of 1-dof links (hinge or slide about the link's
// local z through its origin, one joint per body) numbered so that link index == dof index == qpos
// address, parents before children, depth-first (so a link's subtree is the contiguous index range
// [l, sub_end[l])).  The cube is the only free body and always owns the last 7 qpos / 6 dofs.  These
// structural facts are validated when the flat model is converted (km_fill.h).  Static bodies are folded
// into the world-frame pose of the first moving link under them.
#pragma once
#include "km_common.cuh"
#include "scenes/scene_solo_arm.h"
#include "scenes/scene_dual_arm.h"
#include "scenes/scene_torso.h"

namespace km {

enum { JT_SLIDE = 2, JT_HINGE = 3 };
enum { EFC_FRICTION = 0, EFC_LIMIT = 1, EFC_CONTACT = 2 };
enum { ST_SATISFIED = 0, ST_QUADRATIC = 1, ST_LINEARNEG = 2, ST_LINEARPOS = 3 };

template <class S> struct Dim {
  static constexpr int NVA = S::NVA;          // articulated dofs (== links)
  static constexpr int NV = S::NVA + 6;
  static constexpr int NQ = S::NVA + 7;
  static constexpr int NU = S::NU;
  static constexpr int NPAD = S::NPAD;
  static constexpr int NARM = S::NARM;
  static constexpr int NFRIC = S::NFRIC;
  static constexpr int NMOCAP = S::NMOCAP;
  static constexpr int QLEN = S::Q_LEN;
  static constexpr int NSLOT = S::NPAD + 8;   // collision slots: one per pad, eight cube corners
  static constexpr int MAXCON = S::NPAD + 4;  // active contacts: every pad + at most four corners
  static constexpr int MAXEFC = S::NFRIC + S::NVA + 6 * MAXCON;
  static constexpr int HS = NV + 1;           // row stride of the solver Hessian (bank-conflict padding)
  static constexpr int MAXLEVEL = 12;
  static constexpr int OBS = 2 * S::Q_LEN + 7;
  static constexpr int MAXMASK = 8;
  static constexpr int NRES = 6 + 2 * MAXMASK;
  static constexpr int STATE = NQ + 2 * NV + NU + 7 * NMOCAP + 1 + 3;   // scalars per state record
  static_assert(NU == NVA, "one position actuator per articulated joint");
  static_assert(NV <= 32, "dof support masks are 32-bit");
};

// Diagonal block (kinematic chain, or the cube) of dof j in the mass matrix: [blk_begin, blk_end).  Known at compile
// time from the scene header, so the factorisation can skip structurally zero blocks without run-time tests.
template <class S> constexpr int dof_root(int j) {
  while (S::dof_parent[j] >= 0) j = S::dof_parent[j];
  return j;
}
template <class S> constexpr int blk_end(int j) {
  if (j >= S::NVA) return S::NVA + 6;
  int e = j + 1;
  while (e < S::NVA && dof_root<S>(e) == dof_root<S>(j)) e++;
  return e;
}

template <class S> constexpr int blk_begin(int j) { return j >= S::NVA ? S::NVA : dof_root<S>(j); }
// largest diagonal block (longest kinematic chain, at least the cube's 6)
template <class S> constexpr int max_block() {
  int mx = 6;
  for (int j = 0; j < S::NVA; j++) mx = blk_end<S>(j) - blk_begin<S>(j) > mx ? blk_end<S>(j) - blk_begin<S>(j) : mx;
  return mx;
}

// -------------------------------------------------------------------------------------------- model
template <class S, typename T> struct Model {
  typedef Dim<S> D;
  // links
  int parent[D::NVA], jtype[D::NVA], sub_end[D::NVA];
  unsigned ancmask[D::NVA];                       // bit j set: dof j is l or an ancestor of l
  unsigned short pair_ij[D::NV * (D::NV + 1) / 2]; // lower-triangle work list: i << 8 | j
  int blk0[D::NV], blkn[D::NV];                   // diagonal block of the mass matrix that dof j lies in: first dof, size
  int nlevel, level_adr[D::MAXLEVEL + 1], level_link[D::NVA];
  T lpos[D::NVA][3], lquat[D::NVA][4];            // pose in the parent link, or in the world when parent < 0
  T mass[D::NVA], ipos[D::NVA][3], inertia[D::NVA][3];
  T total_mass_inv;
  T range[D::NVA][2], lim_invw[D::NVA], lim_solref[D::NVA][2], lim_solimp[D::NVA][7];
  T kp[D::NVA], ctrl_lo[D::NVA], ctrl_hi[D::NVA], frc_lo[D::NVA], frc_hi[D::NVA];
  // friction-loss rows (constant: pos = 0)
  int fric_dof[D::NFRIC], dof_fric[D::NV];         // row <-> dof of the friction-loss rows (-1: none)
  T fr_loss[D::NFRIC], fr_R[D::NFRIC], fr_Rf[D::NFRIC], fr_D[D::NFRIC], fr_B[D::NFRIC];   // fr_Rf = R * loss
  // finger pads (spheres) vs cube
  int pad_link[D::NPAD], pad_geom[D::NPAD], pad_arm[D::NPAD];
  T pad_pos[D::NPAD][3], pad_rad[D::NPAD], pad_mu[D::NPAD][3], pad_solref[D::NPAD][2], pad_solimp[D::NPAD][7];
  T pad_tran[D::NPAD], pad_rot[D::NPAD];
  // table plane z = tab_z vs cube
  int table_geom, cube_geom;
  T tab_z, tab_mu[3], tab_solref[2], tab_solimp[7], tab_tran, tab_rot;
  T cube_size[3], cube_mass, cube_inertia[3];
  // options
  T h, grav[3], tol, ls_tol, meaninertia, impratio;
  int iterations, ls_iterations, nsub;
  // task
  int act_dim, act_mode, n_arm;
  int arm_nmask[2], arm_mask[2][D::MAXMASK], arm_grip[2][2], arm_site_link[2], arm_mocap[2];
  int off_pos[2], off_orn[2], off_grip[2], off_q[2];
  T site_pos[2][3], site_quat[2][4];              // end-effector site frame in its link
  // fp64 copies for the IK (kinematic chain of each arm from its base to the site link)
  int arm_nchain[2], arm_chain[2][D::MAXLEVEL], arm_chain_mask[2][D::MAXLEVEL], arm_mask_chain[2][D::MAXMASK];
  double dk_lpos[D::NVA][3], dk_lquat[D::NVA][4], dk_site_pos[2][3], dk_site_quat[2][4], dk_range[D::NVA][2], dk_qhome[D::QLEN];
  int ik_iters, ik_teleport, max_episode_steps, ik_mode;
  T q_home[D::QLEN], spawn_lo[3], spawn_hi[3], cube_quat0[4], mocap0[D::NMOCAP * 7];
  double spawn_lo_d[3], spawn_hi_d[3];
};

// -------------------------------------------------------------------------------------------- working set
struct EmptyStage {};
// TPE = thread-per-env working set: one contiguous record per thread in shared memory whose size in words is odd
// (lane t's field i sits in bank (t * words + i) mod 32, so equal fields of the 32 lanes never conflict).  It must not
// contain 8-byte members when T is float (odd strides break their alignment): the fp64 IK scratch is left out.
template <class S, typename T, bool TPE_ = false> struct Env {
  typedef Dim<S> D;
  static constexpr bool TPE = TPE_;
  // persistent state (what km_get_state / km_set_state expose)
  T qpos[D::NQ], qvel[D::NV], ctrl[D::NU], warm[D::NV], mocap[D::NMOCAP * 7], time;
  // low-order part of the cube's position: qpos[NVA..NVA+2] + cube_lo is the position.  The cube rests ~1e-7 m deep in
  // the table (solimp 0.9999), two float32 ulps of its height, so the fp32 build integrates the cube's translation in
  // float-float arithmetic; always zero in the fp64 build.
  T cube_lo[3];
  int step, episode;
  // position stage: link frames, cube rotation, mass matrix (articulated block; the cube block is a constant diagonal)
  T xpos[D::NVA][3], xquat[D::NVA][4], xmat[D::NVA][9];
  T cmat[9], com[3];
  T M[D::NVA][D::NVA];
  T actlen[D::NU];
  // collision
  int sl_on[D::NSLOT];
  T sl_dist[D::NSLOT], sl_pos[D::NSLOT][3], sl_frame[D::NSLOT][9];
  int ncon, con_slot[D::MAXCON];
  unsigned con_sup[D::MAXCON];                    // dof support of the contact's Jacobian rows
  T con_mu[D::MAXCON][3], con_D[D::MAXCON], con_B[D::MAXCON], con_Kip[D::MAXCON], con_W[D::MAXCON][3];
  // contact Jacobian base rows (normal, tangent1, tangent2, torsion): cube columns of every contact,
  // articulated columns of the finger-pad slots only (the table touches nothing but the cube)
  T Jq[D::MAXCON][4][6], Ja[D::NPAD][4][D::NVA];
  T cb[D::MAXCON][4];                             // base-row products / base-row forces / Hessian weights
  // constraint rows: friction loss (constant, first NFRIC rows), active joint limits, 6 pyramid rows per contact
  int nefc, nlim, coupled;                        // coupled: a finger pad touches the cube (arm and cube blocks interact)
  int dof_lim[D::NVA];                            // active limit row of each joint, or -1
  int efc_desc[D::MAXEFC];
  T efc_D[D::MAXEFC], efc_aref[D::MAXEFC], efc_jv[D::MAXEFC];
  T lim_B[D::NVA], lim_Kip[D::NVA];
  T bias[D::NV], qfrc_smooth[D::NV], qacc_smooth[D::NV], qacc[D::NV];
  T obs[D::OBS];
  int solver_niter, ls_evals;                     // diagnostics of the last sub-step
  int ik_evals;                                   // exact-parity IK: function evaluations of this env step (ordering key)
#ifdef KM_PHASE_CLOCKS
  unsigned clk[16], clk_last;                     // debug build: cycles per phase of the current env step (KM_CLK)
#endif
  // Scratch that is live in disjoint phases shares storage:
  //   a: position/velocity stage (step1)   b: IK between step1 and step2   c: Newton solver (step2)
  struct StageA {
    T cdof[D::NVA][6], cinert[D::NVA][10], cvel[D::NVA][6], cdof_dot[D::NVA][6], cfrc[D::NVA][6];
  };
  struct StageB {   // IK runs in fp64 in both builds (km_sim.cuh)
    double ax[D::MAXLEVEL][3], an[D::MAXLEVEL][3], spos[3], smat[9], goal[7];
    double J[6][D::MAXMASK], r[D::NRES], rn[D::NRES], x[D::MAXMASK], xn[D::MAXMASK], lo[D::MAXMASK], hi[D::MAXMASK],
        qprev[D::MAXMASK], A[D::MAXMASK][D::MAXMASK], gv[D::MAXMASK];
    int active[D::MAXMASK];
  };
  struct StageC {
    T efc_jar[D::MAXEFC], efc_force[D::MAXEFC];
    int efc_state[D::MAXEFC];
    T Ma[D::NV], grad[D::NV], Mgrad[D::NV], search[D::NV], Mv[D::NV], qfc[D::NV], hdiag[D::NV];
    T H[D::NV][D::HS], Hd[D::NV];
  };
  struct StageT {   // Newton solver of the thread-per-env mapping (km_solver_tpe.cuh)
    T Ma[D::NV], grad[D::NV], search[D::NV], Mv[D::NV];
    T jar_f[D::NFRIC], jv_f[D::NFRIC], jar_l[D::NVA], jv_l[D::NVA], jarb[D::MAXCON][4], jvb[D::MAXCON][4];
    T Z[D::NVA][6], Bm[6][D::NVA];   // coupled case: A^{-1} B and B^T of the pad-touched chain block (Schur complement)
  };
  union {
    StageA a;
    StageT t;
    typename std::conditional<TPE_ && sizeof(T) == 4, EmptyStage, StageB>::type b;
    typename std::conditional<TPE_, EmptyStage, StageC>::type c;   // the thread-per-env mapping has its own solver scratch (t)
  };
};
template <class E> KM_HD typename E::StageB& stage_b(E& e, typename E::StageB& local) {
  if constexpr (E::TPE && sizeof(e.qpos[0]) == 4) return local;
  else return e.b;
}

// efc row descriptor: type | id << 2 | k << 10 | neg << 12   (id = dof for friction/limit rows, contact for contact rows;
// k = 1..3 pyramid edge; neg = row uses the negative edge / limit row has J = -e)
KM_HD int efc_pack(int type, int id, int k, int neg) { return type | (id << 2) | (k << 10) | (neg << 12); }
KM_HD int efc_type(int d) { return d & 3; }
KM_HD int efc_id(int d) { return (d >> 2) & 255; }
KM_HD int efc_k(int d) { return (d >> 10) & 3; }
KM_HD int efc_neg(int d) { return (d >> 12) & 1; }

}  // namespace km
