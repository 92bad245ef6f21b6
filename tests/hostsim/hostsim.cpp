// hostsim.cpp -- host build (g++, one lane per env) of the kernel body in gym_kmanip_b200/csrc/km_sim.cuh.
//
// TEST INFRASTRUCTURE ONLY: lets tests run the identical simulator source on the CPU against the oracle while
// no GPU is attached (stage-by-stage parity, fp32 and fp64).  It is not a product path: the package never
// loads it, and the product fails loudly without the CUDA library.
//
// Two mappings: lanes = 1 runs the generic code with one lane per env (plain loops); lanes = 32 runs the warp-per-env
// code -- including the register-resident solver of km_solver_warp.cuh with its shuffles and votes -- on the emulated
// warp of warpemu.h, which also checks that the 32 lanes execute the same collectives in the same order.
#include <cstring>
#include <string>
#define KM_WARP_EMU 1
#include "warpemu.h"
#include "../../gym_kmanip_b200/csrc/km_fill.h"

using namespace km;

struct HostSimBase {
  virtual ~HostSimBase() {}
  virtual int dims(int* out) = 0;
  virtual void set_state(const double* rec, int step, int episode) = 0;
  virtual void get_state(double* rec, int* step, int* episode) = 0;
  virtual void step1() = 0;
  virtual void step2() = 0;
  virtual void before_step(const float* act) = 0;
  virtual void env_step(const float* act, double* obs, double* final_obs, double* reward, unsigned char* trunc, int* flags,
                        int* ncon, int* geoms, int autoreset, uint64_t seed, uint64_t env_id) = 0;
  virtual void reset(uint64_t seed, uint64_t env_id, const double* xyz, double* obs) = 0;
  virtual int field(const char* name, double* out, int cap) = 0;
  std::string fault;   // set when the emulated warp's lanes diverged at a collective
};

template <class S, typename T, int G> struct HostSim : HostSimBase {
  typedef Dim<S> D;
  Model<S, T> m;
  Env<S, T> e;
  km_emu::Warp* warp;
  HostSim() : warp(G == 32 ? new km_emu::Warp() : nullptr) { std::memset((void*)&e, 0, sizeof(e)); }
  ~HostSim() override { delete warp; }
  // run f(group) on the one lane (G = 1) or on the 32 lanes of the emulated warp
  template <class F> void par(F f) {
    if constexpr (G == 1) {
      Grp<1> g;
      g.lane = 0; g.mask = 1u; g.wmask = 1u;
      f(g);
    } else {
      const bool ok = km_emu::run(*warp, [&](int lane) {
        Grp<G> g;
        g.lane = lane; g.mask = 0xffffffffu; g.wmask = 0xffffffffu;
        f(g);
      });
      if (!ok && fault.empty()) fault = warp->error;
    }
  }
  void init() { par([&](const Grp<G>& g) { init_env<S, T, G>(e, m, g); }); }
  int dims(int* out) override {
    out[0] = D::NQ; out[1] = D::NV; out[2] = D::NU; out[3] = D::NMOCAP; out[4] = D::OBS; out[5] = m.act_dim; out[6] = D::MAXCON;
    return 0;
  }
  void set_state(const double* r, int step, int episode) override {
    int k = 0;
    for (int i = 0; i < D::NQ; i++) e.qpos[i] = (T)r[k++];
    for (int i = 0; i < D::NV; i++) e.qvel[i] = (T)r[k++];
    for (int i = 0; i < D::NU; i++) e.ctrl[i] = (T)r[k++];
    for (int i = 0; i < D::NV; i++) e.warm[i] = (T)r[k++];
    for (int i = 0; i < D::NMOCAP * 7; i++) e.mocap[i] = (T)r[k++];
    e.time = (T)r[k++];
    for (int i = 0; i < 3; i++) e.cube_lo[i] = (T)r[k++];
    e.step = step; e.episode = episode;
  }
  void get_state(double* r, int* step, int* episode) override {
    int k = 0;
    for (int i = 0; i < D::NQ; i++) r[k++] = (double)e.qpos[i];
    for (int i = 0; i < D::NV; i++) r[k++] = (double)e.qvel[i];
    for (int i = 0; i < D::NU; i++) r[k++] = (double)e.ctrl[i];
    for (int i = 0; i < D::NV; i++) r[k++] = (double)e.warm[i];
    for (int i = 0; i < D::NMOCAP * 7; i++) r[k++] = (double)e.mocap[i];
    r[k++] = (double)e.time;
    for (int i = 0; i < 3; i++) r[k++] = (double)e.cube_lo[i];
    if (step) *step = e.step;
    if (episode) *episode = e.episode;
  }
  void step1() override { par([&](const Grp<G>& g) { km::step1<S, T, G>(e, m, g); }); }
  void step2() override { par([&](const Grp<G>& g) { km::step2<S, T, G>(e, m, g); }); }
  void before_step(const float* act) override { par([&](const Grp<G>& g) { km::before_step<S, T, G>(e, m, g, act); }); }
  void env_step(const float* act, double* obs, double* final_obs, double* reward, unsigned char* trunc, int* flags, int* ncon,
                int* geoms, int autoreset, uint64_t seed, uint64_t env_id) override {
    T o[D::OBS], fo[D::OBS], r = 0;
    for (int i = 0; i < D::OBS; i++) fo[i] = 0;
    unsigned char term = 0;
    StepOut<T> so = {o, fo, &r, trunc, &term, flags, ncon, geoms, D::MAXCON};
    par([&](const Grp<G>& g) { km::env_step<S, T, G>(e, m, g, act, so, 0, autoreset, seed, env_id); });
    for (int i = 0; i < D::OBS; i++) { obs[i] = (double)o[i]; if (final_obs) final_obs[i] = (double)fo[i]; }
    *reward = (double)r;
  }
  void reset(uint64_t seed, uint64_t env_id, const double* xyz, double* obs) override {
    T c[3];
    if (xyz) for (int i = 0; i < 3; i++) c[i] = (T)xyz[i];
    par([&](const Grp<G>& g) {
      reset_state<S, T, G>(e, m, g, seed, env_id, xyz ? c : (const T*)0);
      observation<S, T, G>(e, m, g);
    });
    if (obs) for (int i = 0; i < D::OBS; i++) obs[i] = (double)e.obs[i];
  }
  int field(const char* name, double* out, int cap) override {
    int n = 0;
#define PUT(x) do { if (n < cap) out[n] = (double)(x); n++; } while (0)
    std::string s(name);
    if (s == "xpos") { for (int l = 0; l < D::NVA; l++) for (int i = 0; i < 3; i++) PUT(e.xpos[l][i]); }
    else if (s == "xquat") { for (int l = 0; l < D::NVA; l++) for (int i = 0; i < 4; i++) PUT(e.xquat[l][i]); }
    else if (s == "com") { for (int i = 0; i < 3; i++) PUT(e.com[i]); }
    else if (s == "qM") {
      for (int i = 0; i < D::NV; i++) for (int j = 0; j < D::NV; j++) {
        if (i < D::NVA && j < D::NVA) PUT(e.M[i][j]);
        else if (i == j) PUT(i - D::NVA < 3 ? m.cube_mass : m.cube_inertia[i - D::NVA - 3]);
        else PUT(0);
      }
    }
    else if (s == "qfrc_bias") { for (int i = 0; i < D::NV; i++) PUT(e.bias[i]); }
    else if (s == "qfrc_smooth") { for (int i = 0; i < D::NV; i++) PUT(e.qfrc_smooth[i]); }
    else if (s == "qacc_smooth") { for (int i = 0; i < D::NV; i++) PUT(e.qacc_smooth[i]); }
    else if (s == "qacc") { for (int i = 0; i < D::NV; i++) PUT(e.qacc[i]); }
    else if (s == "qfrc_constraint") { for (int i = 0; i < D::NV; i++) PUT(e.c.qfc[i]); }
    else if (s == "efc_aref") { for (int r = 0; r < e.nefc; r++) PUT(e.efc_aref[r]); }
    else if (s == "efc_D") { for (int r = 0; r < e.nefc; r++) PUT(e.efc_D[r]); }
    else if (s == "efc_force") { for (int r = 0; r < e.nefc; r++) PUT(e.c.efc_force[r]); }
    else if (s == "efc_state") { for (int r = 0; r < e.nefc; r++) PUT(e.c.efc_state[r]); }
    else if (s == "efc_J") {
      for (int r = 0; r < e.nefc; r++) {
        T x[D::NV];
        const int d = e.efc_desc[r], id = efc_id(d);
        for (int j = 0; j < D::NV; j++) {
          if (efc_type(d) == EFC_CONTACT) {
            bool on = (e.con_sup[id] >> j) & 1u;
            T jn = on ? jc<S, T>(e, id, 0, j) : T(0), jk = on ? jc<S, T>(e, id, efc_k(d), j) : T(0);
            x[j] = jn + (efc_neg(d) ? -1 : 1) * e.con_mu[id][efc_k(d) - 1] * jk;
          } else x[j] = j == id ? (efc_neg(d) ? T(-1) : T(1)) : T(0);
          PUT(x[j]);
        }
      }
    }
    else if (s == "contact_dist") { for (int c = 0; c < e.ncon; c++) PUT(e.sl_dist[e.con_slot[c]]); }
    else if (s == "contact_pos") { for (int c = 0; c < e.ncon; c++) for (int i = 0; i < 3; i++) PUT(e.sl_pos[e.con_slot[c]][i]); }
    else if (s == "contact_frame") { for (int c = 0; c < e.ncon; c++) for (int i = 0; i < 9; i++) PUT(e.sl_frame[e.con_slot[c]][i]); }
    else if (s == "solver_niter") PUT(e.solver_niter);
    else if (s == "ls_evals") PUT(e.ls_evals);
    else if (s == "ncon") PUT(e.ncon);
    else if (s == "nefc") PUT(e.nefc);
    else if (s == "site_xpos") {
      for (int a = 0; a < m.n_arm; a++) { T p[3], R[9]; site_pose<S, T, G>(e, m, a, p, R); for (int i = 0; i < 3; i++) PUT(p[i]); }
    }
    else if (s == "site_xmat") {
      for (int a = 0; a < m.n_arm; a++) { T p[3], R[9]; site_pose<S, T, G>(e, m, a, p, R); for (int i = 0; i < 9; i++) PUT(R[i]); }
    }
    else if (s == "sizeof_env") PUT(sizeof(e));
    else if (s == "sizeof_model") PUT(sizeof(m));
    else return -1;
#undef PUT
    return n;
  }
};

static thread_local std::string g_err;

template <class S, typename T, int G> static HostSimBase* make2(const km_model* fm, const km_task* tk) {
  auto* h = new HostSim<S, T, G>();
  if (fill_model<S, T>(fm, tk, &h->m, g_err)) { delete h; return nullptr; }
  h->init();
  return h;
}
template <class S> static HostSimBase* make(const km_model* fm, const km_task* tk, int dtype, int lanes) {
  if (lanes == 32) return dtype == 32 ? make2<S, float, 32>(fm, tk) : make2<S, double, 32>(fm, tk);
  return dtype == 32 ? make2<S, float, 1>(fm, tk) : make2<S, double, 1>(fm, tk);
}

extern "C" {
const char* hs_last_error() { return g_err.c_str(); }
void* hs_create(const km_model* fm, const km_task* tk, int scene, int dtype, int lanes) {
  if (lanes != 1 && lanes != 32) { g_err = "lanes must be 1 or 32"; return nullptr; }
  switch (scene) {
    case 0: return make<SceneSoloArm>(fm, tk, dtype, lanes);
    case 1: return make<SceneDualArm>(fm, tk, dtype, lanes);
    case 2: return make<SceneTorso>(fm, tk, dtype, lanes);
  }
  g_err = "unknown scene";
  return nullptr;
}
void hs_destroy(void* h) { delete (HostSimBase*)h; }
int hs_dims(void* h, int* out) { return ((HostSimBase*)h)->dims(out); }
void hs_set_state(void* h, const double* rec, int step, int episode) { ((HostSimBase*)h)->set_state(rec, step, episode); }
void hs_get_state(void* h, double* rec, int* step, int* episode) { ((HostSimBase*)h)->get_state(rec, step, episode); }
void hs_step1(void* h) { ((HostSimBase*)h)->step1(); }
void hs_step2(void* h) { ((HostSimBase*)h)->step2(); }
void hs_before_step(void* h, const float* act) { ((HostSimBase*)h)->before_step(act); }
void hs_env_step(void* h, const float* act, double* obs, double* final_obs, double* reward, unsigned char* trunc, int* flags,
                 int* ncon, int* geoms, int autoreset, unsigned long long seed, unsigned long long env_id) {
  ((HostSimBase*)h)->env_step(act, obs, final_obs, reward, trunc, flags, ncon, geoms, autoreset, seed, env_id);
}
void hs_reset(void* h, unsigned long long seed, unsigned long long env_id, const double* xyz, double* obs) {
  ((HostSimBase*)h)->reset(seed, env_id, xyz, obs);
}
int hs_field(void* h, const char* name, double* out, int cap) { return ((HostSimBase*)h)->field(name, out, cap); }
// non-empty when the lanes of the emulated warp diverged at a collective (a hang or undefined behaviour on the GPU)
const char* hs_fault(void* h) { return ((HostSimBase*)h)->fault.c_str(); }
// self-test of the warp emulator: mode 0 runs uniform collectives, mode 1 lets half of the lanes skip a shuffle;
// returns 1 when the emulator reported divergent collectives
int hs_emu_selftest(int mode) {
  km_emu::Warp w;
  const bool ok = km_emu::run(w, [&](int lane) {
    float v = (float)lane;
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    if (mode == 0 || lane < 16) v += __shfl_sync(0xffffffffu, v, 0);
    __syncwarp();
    (void)v;
  });
  return ok ? 0 : 1;
}
}
