/* ko_model.h -- flat model + task structs shared by the oracle's C entry points.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * PARITY UNPINNED: the arithmetic of the reference's hot path lives in third-party MuJoCo
 * (pyproject.toml:15 `mujoco>=2.3.7`, README.md:75 uses 3.1.5), dm_control>=1.0.14 and scipy, none of
 * which is vendored under /root/reference; MuJoCo/dm_control are not installable here and the scenes
 * cannot compile without the git-ignored STL meshes.  This oracle restates the published MuJoCo
 * algorithms on the completed model (assets/completion_spec.json).  It is
 * anchored on the reference call sites env_sim.py:23-179 and ik_mujoco.py:20-155, on scipy itself
 * (Rotation / least_squares are importable and are used to pin the Euler/quaternion and IK pieces,
 * tests/golden/), and on physical self-checks (tests/test_oracle_physics.py).
 *
 * The arrays mirror mjModel field names so each use can be checked against MuJoCo's documentation.
 */
#ifndef KO_MODEL_H
#define KO_MODEL_H

#ifdef __cplusplus
extern "C" {
#endif

enum { KO_JNT_FREE = 0, KO_JNT_BALL = 1, KO_JNT_SLIDE = 2, KO_JNT_HINGE = 3 };
enum { KO_GEOM_PLANE = 0, KO_GEOM_SPHERE = 2, KO_GEOM_BOX = 6 };

typedef struct ko_model {
  int nbody, njnt, nq, nv, nu, nsite, ngeom, npair, nmocap;
  int iterations, ls_iterations;
  double timestep, gravity[3], tolerance, ls_tolerance, impratio, meaninertia;
  /* bodies */
  const int *body_parent, *body_rootid, *body_mocapid, *body_jntadr, *body_jntnum;
  const double *body_pos, *body_quat, *body_mass, *body_ipos, *body_inertia, *body_invweight0;
  /* joints */
  const int *jnt_type, *jnt_bodyid, *jnt_qposadr, *jnt_dofadr, *jnt_limited;
  const double *jnt_pos, *jnt_axis, *jnt_range, *jnt_solref, *jnt_solimp, *qpos0;
  /* dofs */
  const int *dof_bodyid, *dof_jntid, *dof_parentid;
  const double *dof_frictionloss, *dof_solref, *dof_solimp, *dof_invweight0;
  /* actuators (<position> on joints) */
  const int *act_jntid, *act_ctrllimited, *act_forcelimited;
  const double *act_kp, *act_ctrlrange, *act_forcerange;
  /* sites */
  const int *site_bodyid;
  const double *site_pos, *site_quat;
  /* collision primitives and the explicit pair list */
  const int *geom_type, *geom_bodyid;
  const double *geom_pos, *geom_quat, *geom_size;
  const int *pair_geom1, *pair_geom2, *pair_condim;
  const double *pair_friction, *pair_solref, *pair_solimp, *pair_margin;
  const double *mocap_pos0, *mocap_quat0;
} ko_model;

/* Task = the reference's KManipTask configuration (env_sim.py:18-179, __init__.py:28-208). */
#define KO_MAXARM 2
#define KO_MAXMASK 8
typedef struct ko_task {
  int q_len;                 /* env_base.py:66 */
  int n_arm;                 /* arms that appear in act_list, processed right then left (env_sim.py:60,80) */
  int act_dim;               /* flat action length */
  int act_mode;              /* 0: end-effector pos/orn -> IK (env_sim.py:60-99); 1: joint deltas (env_sim.py:100-103) */
  int arm_nmask[KO_MAXARM];
  int arm_mask[KO_MAXARM][KO_MAXMASK];   /* q_id_{r,l}_mask, __init__.py:125-136 */
  int arm_grip[KO_MAXARM][2];            /* ctrl_id_{r,l}_grip */
  int arm_site[KO_MAXARM];               /* site id of ee{r,l}_site_pos */
  int arm_eebody[KO_MAXARM];             /* body id of ee{r,l}_site (reward, env_sim.py:153-161) */
  int arm_mocap[KO_MAXARM];              /* MOCAP_ID_R/L, __init__.py:139-140 */
  int off_pos[KO_MAXARM], off_orn[KO_MAXARM], off_grip[KO_MAXARM], off_q[KO_MAXARM]; /* offsets in the flat action, -1 if absent */
  int cube_body, cube_qposadr;
  int ik_iters;              /* damped Gauss-Newton iterations of the device IK */
  int ik_teleport;           /* reproduce ik_mujoco.py:34,67 leaving qpos[mask] at the solution */
  int max_episode_steps;     /* __init__.py:28 */
  int ik_mode;               /* device IK: 0 fixed-iteration projected LM, 1 restated scipy TRF (the oracle's own "trf" mode calls the real scipy) */
  int n_sub_steps;           /* physics sub-steps per env step; 0 = round(CONTROL_TIMESTEP / timestep) (dm_control) */
  int reserved0;
  double q_home[32];         /* float32-rounded home pose, __init__.py:53-122 */
  double cube_spawn_lo[3], cube_spawn_hi[3];   /* __init__.py:164-170 */
} ko_task;

#ifdef __cplusplus
}
#endif
#endif
