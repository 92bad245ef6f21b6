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
  static_assert(NU == NVA, "one position actuator per articulated joint");
  static_assert(NV <= 32, "dof support masks are 32-bit");
};

// -------------------------------------------------------------------------------------------- model
template <class S, typename T> struct Model {
  typedef Dim<S> D;
  // links
  int parent[D::NVA], jtype[D::NVA], sub_end[D::NVA];
  unsigned ancmask[D::NVA];                       // bit j set: dof j is l or an ancestor of l
  int nlevel, level_adr[D::MAXLEVEL + 1], level_link[D::NVA];
  T lpos[D::NVA][3], lquat[D::NVA][4];            // pose in the parent link, or in the world when parent < 0
  T mass[D::NVA], ipos[D::NVA][3], inertia[D::NVA][3];
  T total_mass_inv;
  T range[D::NVA][2], lim_invw[D::NVA], lim_solref[D::NVA][2], lim_solimp[D::NVA][7];
  T kp[D::NVA], ctrl_lo[D::NVA], ctrl_hi[D::NVA], frc_lo[D::NVA], frc_hi[D::NVA];
  // friction-loss rows (constant: pos = 0)
  int fric_dof[D::NFRIC];
  T fr_loss[D::NFRIC], fr_R[D::NFRIC], fr_D[D::NFRIC], fr_B[D::NFRIC];
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
  int ik_iters, ik_teleport, max_episode_steps;
  T q_home[D::QLEN], spawn_lo[3], spawn_hi[3], cube_quat0[4], mocap0[D::NMOCAP * 7];
  double spawn_lo_d[3], spawn_hi_d[3];
};

// -------------------------------------------------------------------------------------------- working set
template <class S, typename T> struct Env {
  typedef Dim<S> D;
  // persistent state (what km_get_state / km_set_state expose)
  T qpos[D::NQ], qvel[D::NV], ctrl[D::NU], warm[D::NV], mocap[D::NMOCAP * 7], time;
  int step, episode;
  // position stage
  T xpos[D::NVA][3], xquat[D::NVA][4], xmat[D::NVA][9], xipos[D::NVA][3];
  T cmat[9], com[3];
  T cdof[D::NVA][6], cinert[D::NVA][10];
  T M[D::NVA][D::NVA];                            // articulated block (cube block is constant diagonal)
  T Lm[D::NVA][D::NVA + 1], Lmd[D::NVA];          // its Cholesky factor
  T actlen[D::NU];
  // collision
  int sl_on[D::NSLOT];
  T sl_dist[D::NSLOT], sl_pos[D::NSLOT][3], sl_frame[D::NSLOT][9];
  int ncon, con_slot[D::MAXCON];
  unsigned con_sup[D::MAXCON];                    // dof support of the contact's Jacobian rows
  T con_mu[D::MAXCON][3], con_D[D::MAXCON], con_W[D::MAXCON][3];
  T Jc[D::MAXCON][4][D::NV];                      // base rows: normal, tangent1, tangent2, torsion
  T cb[D::MAXCON][4];                             // base-row products / base-row forces
  // constraint rows
  int nefc, nlim;
  int efc_desc[D::MAXEFC];
  T efc_D[D::MAXEFC], efc_R[D::MAXEFC], efc_B[D::MAXEFC], efc_Kip[D::MAXEFC], efc_aref[D::MAXEFC], efc_floss[D::MAXEFC];
  T efc_jar[D::MAXEFC], efc_jv[D::MAXEFC], efc_force[D::MAXEFC];
  int efc_state[D::MAXEFC];
  // velocity stage
  T cvel[D::NVA][6], cdof_dot[D::NVA][6], cfrc[D::NVA][6], bias[D::NV];
  // acceleration stage
  T qfrc_smooth[D::NV], qacc_smooth[D::NV], qacc[D::NV];
  // solver / IK scratch
  T Ma[D::NV], grad[D::NV], Mgrad[D::NV], search[D::NV], Mv[D::NV], qfc[D::NV];
  T H[D::NV][D::HS], Hd[D::NV];
  T ik_J[6][D::MAXMASK], ik_r[D::NRES], ik_rn[D::NRES], ik_x[D::MAXMASK], ik_xn[D::MAXMASK], ik_lo[D::MAXMASK],
      ik_hi[D::MAXMASK], ik_qprev[D::MAXMASK], ik_goal[2][7];
  int ik_active[D::MAXMASK];
  T obs[D::OBS];
  // diagnostics of the last sub-step
  int solver_niter, ls_evals;
};

// efc row descriptor: type | id << 2 | k << 10 | neg << 12   (id = dof for friction/limit rows, contact for contact rows;
// k = 1..3 pyramid edge; neg = row uses the negative edge / limit row has J = -e)
KM_HD int efc_pack(int type, int id, int k, int neg) { return type | (id << 2) | (k << 10) | (neg << 12); }
KM_HD int efc_type(int d) { return d & 3; }
KM_HD int efc_id(int d) { return (d >> 2) & 255; }
KM_HD int efc_k(int d) { return (d >> 10) & 3; }
KM_HD int efc_neg(int d) { return (d >> 12) & 1; }

}  // namespace km
