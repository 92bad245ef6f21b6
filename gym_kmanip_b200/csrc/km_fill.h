// km_fill.h -- host-side conversion of the flat mjModel-style arrays (include/kmanip_b200.h: km_model, km_task)
// into the device model tables (km_model.cuh: Model<S,T>), validating the structural facts the kernels rely on.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/kmanip_b200.h"
#include "km_sim.cuh"

namespace km {

namespace fill_detail {
inline void qmul_d(double* r, const double* a, const double* b) { qmul<double>(r, a, b); }
struct Pose { double p[3]; double q[4]; };
inline Pose compose(const Pose& a, const double* pos, const double* quat) {
  Pose r;
  double m[9], t[3];
  q2mat<double>(m, a.q);
  mulv3<double>(t, m, pos);
  for (int i = 0; i < 3; i++) r.p[i] = a.p[i] + t[i];
  qmul<double>(r.q, a.q, quat);
  return r;
}
}  // namespace fill_detail

#define KM_FILL_CHECK(cond, msg) do { if (!(cond)) { err = std::string("model check failed: ") + msg; return -2; } } while (0)

template <class S, typename T> int fill_model(const km_model* fm, const km_task* tk, Model<S, T>* out, std::string& err) {
  typedef Dim<S> D;
  using fill_detail::Pose;
  Model<S, T>& m = *out;
  std::memset((void*)&m, 0, sizeof(m));
  const int nb = fm->nbody;
  KM_FILL_CHECK(fm->nv == D::NV && fm->nq == D::NQ && fm->nu == D::NU && fm->nmocap == D::NMOCAP, "sizes do not match the scene");
  KM_FILL_CHECK(tk->q_len == D::QLEN, "q_len does not match the scene");
  // ---- bodies: which move, world pose of the static ones, link index of the jointed ones
  std::vector<int> link(nb, -1), moving(nb, 0);
  std::vector<Pose> world(nb);
  world[0] = Pose{{0, 0, 0}, {1, 0, 0, 0}};
  int cube_b = -1;
  for (int b = 1; b < nb; b++) {
    const int p = fm->body_parent[b];
    KM_FILL_CHECK(p < b, "bodies must be numbered parents first");
    KM_FILL_CHECK(fm->body_jntnum[b] <= 1, "at most one joint per body");
    if (fm->body_jntnum[b] == 1) {
      const int j = fm->body_jntadr[b];
      if (fm->jnt_type[j] == 0) {
        KM_FILL_CHECK(cube_b < 0 && p == 0 && fm->jnt_qposadr[j] == D::NVA && fm->jnt_dofadr[j] == D::NVA, "one free body, last in qpos");
        cube_b = b;
      } else {
        KM_FILL_CHECK(fm->jnt_type[j] == JT_HINGE || fm->jnt_type[j] == JT_SLIDE, "hinge/slide joints only");
        KM_FILL_CHECK(fm->jnt_qposadr[j] == j && fm->jnt_dofadr[j] == j && j < D::NVA, "joint index == qpos == dof address");
        KM_FILL_CHECK(fm->jnt_axis[3 * j] == 0 && fm->jnt_axis[3 * j + 1] == 0 && fm->jnt_axis[3 * j + 2] == 1, "joint axis must be local z");
        KM_FILL_CHECK(fm->jnt_pos[3 * j] == 0 && fm->jnt_pos[3 * j + 1] == 0 && fm->jnt_pos[3 * j + 2] == 0, "joint anchor at the body origin");
        KM_FILL_CHECK(fm->qpos0[j] == 0 && fm->jnt_limited[j], "qpos0 = 0 and limited joints");
        link[b] = j;
      }
      moving[b] = 1;
    } else {
      moving[b] = moving[p];
      if (!moving[b]) world[b] = fill_detail::compose(world[p], fm->body_pos + 3 * b, fm->body_quat + 4 * b);
      else KM_FILL_CHECK(fm->body_mass[b] == 0, "moving bodies without a joint must be massless");
    }
  }
  KM_FILL_CHECK(cube_b >= 0, "no free cube body");
  // ---- links
  int nlink = 0;
  double total_mass = 0;
  std::vector<int> depth(D::NVA, 0);
  for (int b = 1; b < nb; b++) {
    const int l = link[b];
    if (l < 0) continue;
    KM_FILL_CHECK(l == nlink, "links must appear in dof order");
    nlink++;
    const int p = fm->body_parent[b];
    m.jtype[l] = fm->jnt_type[l];
    if (link[p] >= 0) {
      m.parent[l] = link[p];
      for (int i = 0; i < 3; i++) { m.lpos[l][i] = (T)fm->body_pos[3 * b + i]; m.dk_lpos[l][i] = fm->body_pos[3 * b + i]; }
      for (int i = 0; i < 4; i++) { m.lquat[l][i] = (T)fm->body_quat[4 * b + i]; m.dk_lquat[l][i] = fm->body_quat[4 * b + i]; }
      depth[l] = depth[link[p]] + 1;
    } else {
      KM_FILL_CHECK(!moving[p], "a link's parent is another link or a static body");
      const Pose w = fill_detail::compose(world[p], fm->body_pos + 3 * b, fm->body_quat + 4 * b);
      m.parent[l] = -1;
      for (int i = 0; i < 3; i++) { m.lpos[l][i] = (T)w.p[i]; m.dk_lpos[l][i] = w.p[i]; }
      for (int i = 0; i < 4; i++) { m.lquat[l][i] = (T)w.q[i]; m.dk_lquat[l][i] = w.q[i]; }
    }
    KM_FILL_CHECK(fm->dof_parentid[l] == m.parent[l], "dof tree must mirror the link tree");
    m.mass[l] = (T)fm->body_mass[b];
    total_mass += fm->body_mass[b];
    for (int i = 0; i < 3; i++) { m.ipos[l][i] = (T)fm->body_ipos[3 * b + i]; m.inertia[l][i] = (T)fm->body_inertia[3 * b + i]; }
    m.range[l][0] = (T)fm->jnt_range[2 * l]; m.range[l][1] = (T)fm->jnt_range[2 * l + 1];
    KM_FILL_CHECK(fm->jnt_range[2 * l] < fm->jnt_range[2 * l + 1], "joint range must be lo < hi (a joint violates at most one side)");
    m.dk_range[l][0] = fm->jnt_range[2 * l]; m.dk_range[l][1] = fm->jnt_range[2 * l + 1];
    m.lim_invw[l] = (T)fm->dof_invweight0[l];
    for (int i = 0; i < 2; i++) m.lim_solref[l][i] = (T)fm->jnt_solref[2 * l + i];
    for (int i = 0; i < 5; i++) m.lim_solimp[l][i] = (T)fm->jnt_solimp[5 * l + i];
    KM_FILL_CHECK(fm->jnt_solimp[5 * l + 4] == 2 || fm->jnt_solimp[5 * l + 4] == 1, "solimp power must be 1 or 2");
    m.lim_solimp[l][5] = (T)(1.0 - fm->jnt_solimp[5 * l]); m.lim_solimp[l][6] = (T)(1.0 - fm->jnt_solimp[5 * l + 1]);
    KM_FILL_CHECK(fm->body_rootid[b] == fm->body_rootid[fm->jnt_bodyid[0]], "all links share one tree root");
  }
  KM_FILL_CHECK(nlink == D::NVA, "link count");
  m.total_mass_inv = (T)(1.0 / total_mass);
  for (int l = 0; l < D::NVA; l++) {
    unsigned mask = 0;
    for (int j = l; j >= 0; j = m.parent[j]) mask |= 1u << j;
    m.ancmask[l] = mask;
  }
  for (int l = 0; l < D::NVA; l++) {
    int e = l + 1;
    while (e < D::NVA && ((m.ancmask[e] >> l) & 1u)) e++;
    for (int c = e; c < D::NVA; c++) KM_FILL_CHECK(!((m.ancmask[c] >> l) & 1u), "subtrees must be contiguous (depth-first numbering)");
    m.sub_end[l] = e;
  }
  for (int i = 0, w = 0; i < D::NV; i++)
    for (int j = 0; j <= i; j++) m.pair_ij[w++] = (unsigned short)(i << 8 | j);
  for (int j = 0; j < D::NV; j++) {   // blocks = kinematic chains (links sharing a root), then the cube
    int r = j;
    while (r < D::NVA && m.parent[r] >= 0) r = m.parent[r];
    int e = j + 1;
    if (j >= D::NVA) { r = D::NVA; e = D::NV; }
    else while (e < D::NVA) { int q = e; while (m.parent[q] >= 0) q = m.parent[q]; if (q != r) break; e++; }
    m.blk0[j] = r; m.blkn[j] = e - r;
  }
  int maxd = 0;
  for (int l = 0; l < D::NVA; l++) maxd = depth[l] > maxd ? depth[l] : maxd;
  KM_FILL_CHECK(maxd + 1 <= D::MAXLEVEL, "kinematic tree too deep");
  m.nlevel = maxd + 1;
  int k = 0;
  for (int d = 0; d <= maxd; d++) {
    m.level_adr[d] = k;
    for (int l = 0; l < D::NVA; l++) if (depth[l] == d) m.level_link[k++] = l;
  }
  m.level_adr[maxd + 1] = k;
  // ---- actuators
  for (int i = 0; i < D::NU; i++) {
    KM_FILL_CHECK(fm->act_jntid[i] == i, "actuator i drives joint i");
    m.kp[i] = (T)fm->act_kp[i];
    m.ctrl_lo[i] = fm->act_ctrllimited[i] ? (T)fm->act_ctrlrange[2 * i] : -Num<T>::huge();
    m.ctrl_hi[i] = fm->act_ctrllimited[i] ? (T)fm->act_ctrlrange[2 * i + 1] : Num<T>::huge();
    m.frc_lo[i] = fm->act_forcelimited[i] ? (T)fm->act_forcerange[2 * i] : -Num<T>::huge();
    m.frc_hi[i] = fm->act_forcelimited[i] ? (T)fm->act_forcerange[2 * i + 1] : Num<T>::huge();
  }
  // ---- friction-loss rows (pos = 0: impedance, regulariser and damping gain are constants)
  int nf = 0;
  for (int d = 0; d < D::NV; d++) m.dof_fric[d] = -1;
  for (int d = 0; d < fm->nv; d++) {
    if (!(fm->dof_frictionloss[d] > 0)) continue;
    KM_FILL_CHECK(nf < D::NFRIC, "friction-loss row count");
    const double* si0 = fm->dof_solimp + 5 * d;
    const double si[7] = {si0[0], si0[1], si0[2], si0[3], si0[4], 1.0 - si0[0], 1.0 - si0[1]};
    double omi;
    const double imp = impedance<double>(si, 0.0, &omi);
    const double R = std::fmax(1e-15, omi * fm->dof_invweight0[d] / imp);
    const double tc = std::fmax(fm->dof_solref[2 * d], 2.0 * fm->timestep);
    m.fric_dof[nf] = d; m.dof_fric[d] = nf; m.fr_loss[nf] = (T)fm->dof_frictionloss[d];
    m.fr_R[nf] = (T)R; m.fr_Rf[nf] = (T)(R * fm->dof_frictionloss[d]); m.fr_D[nf] = (T)(1.0 / R); m.fr_B[nf] = (T)(2.0 / (si[1] * tc));
    nf++;
  }
  KM_FILL_CHECK(nf == D::NFRIC, "friction-loss row count");
  // ---- collision pairs: pads first, the table last; geom2 is always the cube
  KM_FILL_CHECK(fm->npair == D::NPAD + 1, "pair count");
  for (int p = 0; p < fm->npair; p++) {
    const int g1 = fm->pair_geom1[p], g2 = fm->pair_geom2[p];
    KM_FILL_CHECK(fm->geom_bodyid[g2] == cube_b && fm->geom_type[g2] == 6, "geom2 of every pair is the cube box");
    KM_FILL_CHECK(fm->pair_condim[p] == 4 && fm->pair_margin[p] == 0, "condim 4, margin 0");
    KM_FILL_CHECK(fm->pair_solimp[5 * p + 4] == 2 || fm->pair_solimp[5 * p + 4] == 1, "solimp power must be 1 or 2");
    const int b1 = fm->geom_bodyid[g1];
    const double tran = fm->body_invweight0[2 * b1] + fm->body_invweight0[2 * cube_b];
    const double rot = fm->body_invweight0[2 * b1 + 1] + fm->body_invweight0[2 * cube_b + 1];
    if (p < D::NPAD) {
      KM_FILL_CHECK(fm->geom_type[g1] == 2 && link[b1] >= 0, "pads are spheres on links");
      m.pad_link[p] = link[b1]; m.pad_geom[p] = g1; m.pad_rad[p] = (T)fm->geom_size[3 * g1];
      for (int i = 0; i < 3; i++) { m.pad_pos[p][i] = (T)fm->geom_pos[3 * g1 + i]; m.pad_mu[p][i] = (T)fm->pair_friction[5 * p + i]; }
      for (int i = 0; i < 2; i++) m.pad_solref[p][i] = (T)fm->pair_solref[2 * p + i];
      for (int i = 0; i < 5; i++) m.pad_solimp[p][i] = (T)fm->pair_solimp[5 * p + i];
      m.pad_solimp[p][5] = (T)(1.0 - fm->pair_solimp[5 * p]); m.pad_solimp[p][6] = (T)(1.0 - fm->pair_solimp[5 * p + 1]);
      m.pad_tran[p] = (T)tran; m.pad_rot[p] = (T)rot;
      int arm = 0;   // the arm whose end-effector body hangs off the pad's hand (oracle ko_reward)
      for (int a = 0; a < tk->n_arm; a++)
        for (int x = tk->arm_eebody[a]; x > 0; x = fm->body_parent[x]) if (x == fm->body_parent[b1]) arm = a;
      m.pad_arm[p] = arm;
    } else {
      KM_FILL_CHECK(fm->geom_type[g1] == 0 && !moving[b1], "the last pair is the static table plane");
      const Pose w = fill_detail::compose(world[b1], fm->geom_pos + 3 * g1, fm->geom_quat + 4 * g1);
      KM_FILL_CHECK(std::fabs(w.q[0]) > 1 - 1e-12, "table plane must be horizontal");
      m.table_geom = g1; m.cube_geom = g2; m.tab_z = (T)w.p[2];
      for (int i = 0; i < 3; i++) m.tab_mu[i] = (T)fm->pair_friction[5 * p + i];
      for (int i = 0; i < 2; i++) m.tab_solref[i] = (T)fm->pair_solref[2 * p + i];
      for (int i = 0; i < 5; i++) m.tab_solimp[i] = (T)fm->pair_solimp[5 * p + i];
      m.tab_solimp[5] = (T)(1.0 - fm->pair_solimp[5 * p]); m.tab_solimp[6] = (T)(1.0 - fm->pair_solimp[5 * p + 1]);
      m.tab_tran = (T)tran; m.tab_rot = (T)rot;
      KM_FILL_CHECK(fm->geom_pos[3 * g2] == 0 && fm->geom_pos[3 * g2 + 1] == 0 && fm->geom_pos[3 * g2 + 2] == 0, "cube geom at the body origin");
      for (int i = 0; i < 3; i++) m.cube_size[i] = (T)fm->geom_size[3 * g2 + i];
    }
  }
  m.cube_mass = (T)fm->body_mass[cube_b];
  for (int i = 0; i < 3; i++) {
    m.cube_inertia[i] = (T)fm->body_inertia[3 * cube_b + i];
    KM_FILL_CHECK(fm->body_ipos[3 * cube_b + i] == 0, "cube COM at its origin");
  }
  // ---- options
  m.h = (T)fm->timestep;
  for (int i = 0; i < 3; i++) m.grav[i] = (T)fm->gravity[i];
  m.tol = (T)fm->tolerance; m.ls_tol = (T)fm->ls_tolerance; m.meaninertia = (T)fm->meaninertia; m.impratio = (T)fm->impratio;
  m.iterations = fm->iterations; m.ls_iterations = fm->ls_iterations;
  m.nsub = tk->n_sub_steps > 0 ? tk->n_sub_steps : (int)std::lround(0.02 / fm->timestep);   // CONTROL_TIMESTEP, reference __init__.py:30
  // ---- task
  m.act_dim = tk->act_dim; m.act_mode = tk->act_mode; m.n_arm = tk->n_arm;
  KM_FILL_CHECK(tk->n_arm <= D::NARM && tk->cube_body == cube_b && tk->cube_qposadr == D::NVA, "task does not match the scene");
  for (int a = 0; a < 2; a++) { m.off_pos[a] = m.off_orn[a] = m.off_grip[a] = m.off_q[a] = -1; }
  for (int a = 0; a < tk->n_arm; a++) {
    m.arm_nmask[a] = tk->arm_nmask[a];
    KM_FILL_CHECK(tk->arm_nmask[a] <= D::MAXMASK, "mask length");
    const int sb = fm->site_bodyid[tk->arm_site[a]];
    KM_FILL_CHECK(sb == tk->arm_eebody[a], "end-effector site must sit on the ee body");
    const double* sp = fm->site_pos + 3 * tk->arm_site[a];
    const double* sq = fm->site_quat + 4 * tk->arm_site[a];
    KM_FILL_CHECK(sp[0] == 0 && sp[1] == 0 && sp[2] == 0, "end-effector site at its body origin");
    Pose off = Pose{{0, 0, 0}, {1, 0, 0, 0}};
    int hb = sb;
    if (link[sb] < 0) { off = fill_detail::compose(off, fm->body_pos + 3 * sb, fm->body_quat + 4 * sb); hb = fm->body_parent[sb]; }
    KM_FILL_CHECK(link[hb] >= 0, "end-effector body must hang off a link");
    off = fill_detail::compose(off, sp, sq);
    m.arm_site_link[a] = link[hb];
    for (int i = 0; i < 3; i++) m.site_pos[a][i] = (T)off.p[i];
    for (int i = 0; i < 4; i++) m.site_quat[a][i] = (T)off.q[i];
    for (int i = 0; i < tk->arm_nmask[a]; i++) {
      m.arm_mask[a][i] = tk->arm_mask[a][i];
      KM_FILL_CHECK((m.ancmask[link[hb]] >> tk->arm_mask[a][i]) & 1u, "masked joints must lie on the site's chain");
    }
    // chain from the base to the site link, for the fp64 IK
    {
      int chain[D::MAXLEVEL], nc = 0;
      for (int l = link[hb]; l >= 0; l = m.parent[l]) chain[nc++] = l;
      m.arm_nchain[a] = nc;
      for (int k = 0; k < nc; k++) {
        const int l = chain[nc - 1 - k];
        m.arm_chain[a][k] = l; m.arm_chain_mask[a][k] = -1;
        for (int i = 0; i < tk->arm_nmask[a]; i++) if (tk->arm_mask[a][i] == l) { m.arm_chain_mask[a][k] = i; m.arm_mask_chain[a][i] = k; }
      }
      for (int i = 0; i < 3; i++) m.dk_site_pos[a][i] = off.p[i];
      for (int i = 0; i < 4; i++) m.dk_site_quat[a][i] = off.q[i];
    }
    m.arm_grip[a][0] = tk->arm_grip[a][0]; m.arm_grip[a][1] = tk->arm_grip[a][1];
    m.arm_mocap[a] = tk->arm_mocap[a];
    m.off_pos[a] = tk->off_pos[a]; m.off_orn[a] = tk->off_orn[a]; m.off_grip[a] = tk->off_grip[a]; m.off_q[a] = tk->off_q[a];
  }
  m.ik_iters = tk->ik_iters; m.ik_teleport = tk->ik_teleport; m.ik_mode = tk->ik_mode; m.max_episode_steps = tk->max_episode_steps;
  for (int i = 0; i < D::QLEN; i++) { m.q_home[i] = (T)tk->q_home[i]; m.dk_qhome[i] = tk->q_home[i]; }
  for (int i = 0; i < 3; i++) {
    m.spawn_lo[i] = (T)tk->cube_spawn_lo[i]; m.spawn_hi[i] = (T)tk->cube_spawn_hi[i];
    m.spawn_lo_d[i] = tk->cube_spawn_lo[i]; m.spawn_hi_d[i] = tk->cube_spawn_hi[i];
  }
  for (int i = 0; i < 4; i++) m.cube_quat0[i] = (T)fm->qpos0[D::NVA + 3 + i];
  for (int k2 = 0; k2 < D::NMOCAP; k2++) {
    for (int i = 0; i < 3; i++) m.mocap0[7 * k2 + i] = (T)fm->mocap_pos0[3 * k2 + i];
    for (int i = 0; i < 4; i++) m.mocap0[7 * k2 + 3 + i] = (T)fm->mocap_quat0[4 * k2 + i];
  }
  return 0;
}

}  // namespace km
