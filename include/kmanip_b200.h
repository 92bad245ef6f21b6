/* kmanip_b200.h -- C-ABI of the B200 batched simulator for the gym-kmanip env-step hot path.
 *
 * This is the drop-in boundary.  The reference selects a simulation backend through the seam
 *     gym_kmanip/env_base.py:192-200   self.env = new(self)
 * and talks to it through
 *     gym_kmanip/env_sim.py:187-203    k_reset() / k_step(action) / k_render(cam) / k_close()
 * (each returning (terminated, reward, discount, observation, sim_time)).  The entry points below are what a
 * replacement backend binds for that seam, batched over n_envs independent environments:
 *
 *   km_create / km_destroy   <- env_sim.py:206-211  new(): Physics.from_xml_path + KManipTask + control.Environment
 *                               (the MJCF is flattened on the host; the library receives plain arrays)
 *   km_reset                 <- env_sim.py:190-194  k_reset -> KManipTask.initialize_episode (env_sim.py:23-36)
 *   km_step                  <- env_sim.py:196-200  k_step  -> before_step (:38-108, incl. ik_mujoco.py:100-155),
 *                               physics.step(10), get_reward (:148-179), get_observation (:110-146);
 *                               truncation after max_episode_steps = gymnasium TimeLimit of __init__.py:28,247
 *   km_get_state/km_set_state<- callers reaching through env.unwrapped.env.physics.data.{qpos,qvel,ctrl,mocap_pos,mocap_quat,time}
 *                               (examples/1_control.py:26, examples/2_synthetic_data.py:33-34, examples/4_teleop.py:77-83)
 *   km_contacts              <- the contact scan of get_reward (env_sim.py:166-174)
 *   km_step_host / km_reset_host: the same calls with HOST buffers (copies inside), what a ctypes/cgo-style
 *                               binding that owns no device memory would call.
 *
 * Conventions: every call returns 0 on success or a negative error code, with a thread-local message from
 * km_last_error().  The library never owns caller buffers.  Device pointers must live on the handle's device.
 * Calls are asynchronous on `stream` (a cudaStream_t passed as void*; NULL = the legacy default stream) except the
 * *_host variants, which synchronise before returning.  One handle per device; calls on one handle must be
 * serialised by the caller; handles on different devices are independent (env sharding needs no collective).
 *
 * Record layouts (env-major, one contiguous record per env -- the coalesced layout for lane-group-per-env kernels):
 *   state    qpos[nq] qvel[nv] ctrl[nu] qacc_warmstart[nv] mocap[7*nmocap] time cube_lo[3]   (dtype of the handle)
 *            cube_lo: low-order part of the cube's position in the f32 build (position = qpos[nq-7..nq-5] + cube_lo;
 *            the cube rests ~1e-7 m deep in the table, two float32 ulps of its height); always 0 in the f64 build
 *   action   float32[act_dim], keys concatenated in the order of env_base.py:149-190
 *   obs      [q_pos(q_len) q_vel(q_len) cube_pos(3) cube_orn(4)]                    (dtype of the handle)
 */
#ifndef KMANIP_B200_H
#define KMANIP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Flat model: mjModel-named arrays produced by gym_kmanip_b200/mjcf.py (host pointers, copied at km_create). */
typedef struct km_model {
  int nbody, njnt, nq, nv, nu, nsite, ngeom, npair, nmocap;
  int iterations, ls_iterations;
  double timestep, gravity[3], tolerance, ls_tolerance, impratio, meaninertia;
  const int *body_parent, *body_rootid, *body_mocapid, *body_jntadr, *body_jntnum;
  const double *body_pos, *body_quat, *body_mass, *body_ipos, *body_inertia, *body_invweight0;
  const int *jnt_type, *jnt_bodyid, *jnt_qposadr, *jnt_dofadr, *jnt_limited;
  const double *jnt_pos, *jnt_axis, *jnt_range, *jnt_solref, *jnt_solimp, *qpos0;
  const int *dof_bodyid, *dof_jntid, *dof_parentid;
  const double *dof_frictionloss, *dof_solref, *dof_solimp, *dof_invweight0;
  const int *act_jntid, *act_ctrllimited, *act_forcelimited;
  const double *act_kp, *act_ctrlrange, *act_forcerange;
  const int *site_bodyid;
  const double *site_pos, *site_quat;
  const int *geom_type, *geom_bodyid;
  const double *geom_pos, *geom_quat, *geom_size;
  const int *pair_geom1, *pair_geom2, *pair_condim;
  const double *pair_friction, *pair_solref, *pair_solimp, *pair_margin;
  const double *mocap_pos0, *mocap_quat0;
} km_model;

/* Task: the KManipTask configuration of one registered env id (env_sim.py:18-179, __init__.py:28-208). */
#define KM_MAXARM 2
#define KM_MAXMASK 8
typedef struct km_task {
  int q_len, n_arm, act_dim;
  int act_mode;                          /* 0: end-effector targets -> IK; 1: joint-position deltas */
  int arm_nmask[KM_MAXARM];
  int arm_mask[KM_MAXARM][KM_MAXMASK];   /* q_id_{r,l}_mask */
  int arm_grip[KM_MAXARM][2];            /* ctrl_id_{r,l}_grip */
  int arm_site[KM_MAXARM], arm_eebody[KM_MAXARM], arm_mocap[KM_MAXARM];
  int off_pos[KM_MAXARM], off_orn[KM_MAXARM], off_grip[KM_MAXARM], off_q[KM_MAXARM];
  int cube_body, cube_qposadr;
  int ik_iters, ik_teleport, max_episode_steps;
  int ik_mode;                           /* 0: fixed-iteration projected LM (fast path); 1: restated scipy TRF (exact-parity mode) */
  int n_sub_steps;                       /* physics sub-steps per env step; 0 = round(CONTROL_TIMESTEP / timestep) = 10 (__init__.py:30).
                                            Tests set 1 to compare single mj_step's (per-sub-step parity) */
  int reserved0;
  double q_home[32], cube_spawn_lo[3], cube_spawn_hi[3];
} km_task;

typedef struct km_sim* km_handle;

enum { KM_F32 = 32, KM_F64 = 64 };
enum { KM_SCENE_SOLO_ARM = 0, KM_SCENE_DUAL_ARM = 1, KM_SCENE_TORSO = 2 };
enum { KM_OK = 0, KM_ERR_ARG = -1, KM_ERR_MODEL = -2, KM_ERR_CUDA = -3, KM_ERR_NODEVICE = -4 };

/* Optional outputs of km_step (device pointers; any may be NULL). */
typedef struct km_step_out {
  void* obs;                 /* [n][obs_dim]  observation after the step (first obs of the new episode on autoreset) */
  void* final_obs;           /* [n][obs_dim]  last observation of the finished episode, written only where truncated */
  void* reward;              /* [n] */
  unsigned char* truncated;  /* [n] 1 on the max_episode_steps-th step since reset */
  unsigned char* terminated; /* [n] always 0 (the reference never terminates) */
  int* con_flags;            /* [n] bit0 cube-table, bit1 right finger pads, bit2 left finger pads */
  int* ncon;                 /* [n] */
  int* con_geoms;            /* [n][2*km_max_contacts] (geom1, geom2) per contact, -1 padded */
  /* the per-step `info` of reference env_base.py:243-250, as arrays: */
  unsigned char* is_success; /* [n] reward > REWARD_SUCCESS_THRESHOLD (2.0) */
  void* episode_return;      /* [n] return of the running episode including this step (dtype of the handle) */
  void* final_return;        /* [n] return of the episode this step finished (truncated), else 0 */
  void* sim_time;            /* [n] simulation time after the step (before an autoreset) */
  int* step_count;           /* [n] step index inside the episode after this step (max_episode_steps where truncated) */
  int* episode;              /* [n] index of the episode this step belongs to */
} km_step_out;

const char* km_last_error(void);
const char* km_version(void);
/* Measured FMA throughput of the CUDA-core pipe this path is bound by (dtype KM_F32 / KM_F64), in TFLOP/s: the
   denominator of the FP roofline that bench.py reports (MEASURED_PEAKS.json carries no CUDA-core figure). */
int km_measure_fma_peak(int device, int dtype, double* tflops);

/* scene: KM_SCENE_*; dtype: KM_F32 / KM_F64; seed/env0: cube-spawn RNG key and the global id of env 0 of this shard. */
int km_create(const km_model* model, const km_task* task, int scene, int n_envs, int device, int dtype,
              uint64_t seed, uint64_t env0, km_handle* out);
void km_destroy(km_handle h);

int km_nq(km_handle h);
int km_nv(km_handle h);
int km_nu(km_handle h);
int km_nmocap(km_handle h);
int km_obs_dim(km_handle h);
int km_act_dim(km_handle h);
int km_state_dim(km_handle h);        /* scalars per state record */
int km_max_contacts(km_handle h);
int km_num_envs(km_handle h);
int km_dtype(km_handle h);
/* launch configuration.  lanes_per_env: 32 / 16 = lane group per env (at least the dof count); 1 = thread per env with the
   env records in shared memory; 2 = thread per env with the records in local memory; 0 keeps the current mapping.
   envs_per_block: envs per CTA, 0 = choose.  km_create picks a default from the batch size.
   Handles created with km_task.ik_mode = 1 (exact-parity TRF IK) run in the lane-group mapping only: 1 / 2 are refused. */
int km_configure(km_handle h, int lanes_per_env, int envs_per_block);
/* Cost-ordered walk of the step kernel.  The envs of a CTA march through the sub-step in phase, so a tile of envs costs what
   its slowest env costs, and an env's cost (Newton iterations: its contact state) persists from step to step.  When on,
   km_step first buckets the envs by the line-search evaluations of their previous step (one small counting-sort launch,
   most expensive first) and the CTAs fetch the tiles dynamically.  Results do not depend on it (every env is stepped
   exactly as before; only the rollout totals' atomics change order).  mode: -1 = automatic (default: on when the
   lane-group mapping walks more tiles than there are CTAs), 0 = off, 1 = on.  No counterpart in the reference. */
int km_set_env_ordering(km_handle h, int mode);
long long km_launch_count(km_handle h);   /* kernels launched by this handle so far */
int km_launch_config(km_handle h, int* lanes_per_env, int* envs_per_block, int* grid, int* ctas_per_sm, int* smem_bytes);

/* Episode reset.  mask_dev: [n] uint8, NULL = all envs.  cube_xyz_dev: [n][3] (dtype of the handle), NULL = device
   RNG Philox4x32-10(seed, env0 + i, episode).  obs_dev: optional [n][obs_dim]. */
int km_reset(km_handle h, const unsigned char* mask_dev, const void* cube_xyz_dev, void* obs_dev, void* stream);

/* One env step of every env (10 physics sub-steps).  action_dev: [n][act_dim] float32.  autoreset != 0 resets
   truncated envs inside the same launch. */
int km_step(km_handle h, const float* action_dev, const km_step_out* out, int autoreset, void* stream);

/* Teacher-forced access to the full state (device buffers of the handle's dtype, record layout above).
   step_count / episode: optional [n] int32. */
int km_get_state(km_handle h, void* state_dev, int* step_count_dev, int* episode_dev, void* stream);
int km_set_state(km_handle h, const void* state_dev, const int* step_count_dev, const int* episode_dev, void* stream);
/* Pointer to the library-owned state buffer ([n][km_state_dim], dtype of the handle) for zero-copy views. */
void* km_state_ptr(km_handle h);

/* Contacts of the most recent step: runs the position stage on the stored state. */
int km_contacts(km_handle h, int* ncon_dev, int* con_geoms_dev, void* stream);

/* End-effector site frames of the stored state (position stage only): xpos_dev [n][n_arm][3], xmat_dev [n][n_arm][9]
   row-major, dtype of the handle; either may be NULL.  Arms in task order (right, then left).
   <- callers reading physics.data.site("eer_site_pos").xpos / .xmat (examples/2_synthetic_data.py:34, env_sim.py:62-64). */
int km_site_poses(km_handle h, void* xpos_dev, void* xmat_dev, void* stream);
int km_n_arm(km_handle h);

/* Camera observations of the Vision ids  <- gym_kmanip/env_sim.py:140-145 (get_observation: physics.render(height=cam.h,
   width=cam.w, camera_id=cam.name) for every camera of obs_list) and env_sim.py:187-188 (k_render(cam)); camera sizes are
   the Cam table of __init__.py:157-161.  Renders the stored state of every env: rgb_dev [n][height][width][3] uint8,
   row 0 on top.  What is drawn is the completed model's primitives (table plane, cube, finger pads, one capsule per
   moving link) under MuJoCo's camera / fixed-function lighting conventions; see csrc/km_render.cuh.
   km_camera: an MJCF <camera mode="targetbody"> (reference _env_solo_arm.xml:14-15, arm_r_body.xml:68) after the host
   folded static bodies away: `link` is the joint index of the moving body the camera rides on (-1: fixed in the world),
   `pos` its position in that body's frame (or the world); the tracked body's origin likewise. */
typedef struct km_camera {
  int width, height;
  double fovy;               /* vertical field of view, degrees */
  int link;
  double pos[3];
  int target_link;
  double target_pos[3];
} km_camera;
/* Lights and colours: <visual><headlight>, the directional <light>s and geom rgba of scene.xml:5-20, and the appearance
   the completion spec gives the link proxies.  light_dir is the direction the light travels (MJCF `dir`). */
typedef struct km_visual {
  int nlight;                /* directional lights, at most 4 */
  double light_dir[4][3], light_diffuse[4][3], light_specular[4][3], light_ambient[4][3];
  double head_ambient[3], head_diffuse[3], head_specular[3];
  double rgb_table[3], rgb_cube[3], rgb_link[3], rgb_pad[3];
  double specular, shininess; /* material defaults of MuJoCo: 0.5, 0.5 (exponent shininess * 128) */
  double link_radius;
} km_visual;
int km_render(km_handle h, const km_camera* cam, const km_visual* vis, unsigned char* rgb_dev, void* stream);
int km_render_host(km_handle h, const km_camera* cam, const km_visual* vis, unsigned char* rgb_host);
/* Copy of the per-env render records of the last km_render ([n][km_render_record_floats] float32: camera origin and
   axes, then 16 floats per primitive) -- exposed so that tests can check the pixel stage against the oracle exactly. */
int km_render_record_floats(km_handle h);
int km_get_render_records(km_handle h, float* recs_dev, void* stream);

/* Diagnostics of the most recent step's last sub-step: [n] Newton iterations, [n] line-search evaluations (cumulative). */
int km_solver_stats(km_handle h, int* niter_dev, int* ls_evals_dev, void* stream);
/* Rollout statistics accumulated by the step kernel itself (no extra launches): totals_dev[4] = {sum of rewards, env
   steps, finished (truncated) episodes, success steps} over every km_step since the last reset of the totals, and/or the
   running return of every env (episode_return_dev [n], dtype of the handle).  Either pointer may be NULL.  reset != 0
   zeroes the totals after they were copied.  The sums are formed with floating-point atomics (one per CTA): their last
   bits depend on the order in which CTAs finish. */
int km_episode_stats(km_handle h, double* totals_dev, void* episode_return_dev, int reset, void* stream);
/* Stream used by the *_host entry points (default: the legacy default stream).  Callers that mix the host-buffer calls
   with stream-ordered device-pointer calls pass the same stream to both. */
int km_set_host_stream(km_handle h, void* stream);

/* Profiling aid: after this call every km_step also writes [n][16] uint32 cycle counts per phase of the env step
   (position stage, velocity stage, Newton solver pieces, barrier waits ...; enum CLK_* in csrc/km_common.cuh) into the
   caller's device buffer; NULL switches it off.  Only libraries built with -DKM_PHASE_CLOCKS record anything; the
   production build returns KM_ERR_ARG. */
int km_debug_phase_clocks(km_handle h, unsigned* clk_dev);

/* Host-buffer variants (pageable or pinned host memory; copies and a stream synchronise inside). */
int km_reset_host(km_handle h, const unsigned char* mask, const void* cube_xyz, void* obs);
int km_step_host(km_handle h, const float* action, void* obs, void* reward, unsigned char* truncated, int autoreset);

#ifdef __cplusplus
}
#endif
#endif
