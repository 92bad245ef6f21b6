// km_launch.cuh -- the CUDA kernels around km_sim.cuh and their launchers, instantiated once per
// (scene, scalar type) translation unit (km_kernels_*.cu) and reached from km_api.cu through KmVtable.
//
// Mapping: a CTA holds `epb` environments, each owned by a group of G lanes of one warp; the working set of
// every env (Env<S,T>) and one copy of the model tables live in dynamic shared memory; the grid is persistent
// (a multiple of the SM count) and strides over env tiles.  HBM sees one contiguous state record per env in,
// one out, plus the action / observation / reward records -- everything else stays on chip for all 10 sub-steps.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include "km_fill.h"
#include "km_render.cuh"

namespace km {

struct KmArgs {
  const void* model;
  void* state;
  int* step;
  int* episode;
  const float* act;
  void* obs;
  void* final_obs;
  void* reward;
  unsigned char* trunc;
  unsigned char* term;
  int* con_flags;
  int* ncon;
  int* con_geoms;
  int* niter;
  int* ls;
  unsigned* clk;   // debug build (KM_PHASE_CLOCKS): per-env phase cycles of the step
  // episode bookkeeping: persistent running return [n], optional per-step info outputs, rollout totals double[4]
  void* ep_return; void* episode_return; void* final_return; void* sim_time;
  unsigned char* is_success; int* step_out; int* episode_out;
  double* totals;
  const unsigned char* mask;
  const void* cube_xyz;
  void* site_pos;
  void* site_mat;
  int n, autoreset;
  unsigned long long seed, env0;
  int G, epb, grid;
  // Cost-ordered walk (km_api.cu: order_envs): order[i] = env processed at position i (descending cost of the previous
  // step, so that the envs sharing a CTA / warp need similar numbers of Newton iterations); tile_counter = dynamic tile
  // fetch (CTAs take the next tile when they finish one: longest tiles first, balanced finish).  Both may be null.
  const int* order;
  int* tile_counter;
  int* cost;       // ordering key of the next step: line-search evaluations + 64 x IK function evaluations of this one
  int trf_bytes;   // extra dynamic shared memory per env: work arrays of the exact-parity IK (ik_mode = 1), else 0
  int lpw;   // thread-per-env (local) mapping: active lanes per warp (envs of a CTA are spread over its warps)
  int tpl_small_regs;   // thread-per-env (local): use the 128-register instantiation even for CTAs of <= 256 threads (several CTAs per SM)
  cudaStream_t stream;
};

struct KmVtable {
  size_t model_bytes, env_bytes, scalar_bytes, tpe_env_bytes;
  int nq, nv, nu, nmocap, obs_dim, state_dim, maxcon, nlanes_min, max_threads, tpe_max_envs;
  int (*fill)(const km_model*, const km_task*, void* dst, std::string& err);
  cudaError_t (*step)(const KmArgs&);
  cudaError_t (*reset)(const KmArgs&);
  cudaError_t (*contacts)(const KmArgs&);
  // opt in to the dynamic shared memory of (G, epb); returns resident CTAs per SM through *ctas_per_sm
  cudaError_t (*prepare)(int G, int epb, int extra_per_env, int* ctas_per_sm);
  // camera observations: one record of floats per env (camera frame + primitive list) from the stored state
  int render_rec_floats;
  double (*table_z)(const void* host_model);
  cudaError_t (*render_setup)(const KmArgs&, const KmRenderParams&, float* recs);
};

#if defined(__CUDACC__)

// opt-in shared memory of one CTA on sm_100a (227 KB) minus the kernels' few bytes of static shared memory (CTA totals)
constexpr size_t KM_SMEM_STATIC = 64;
constexpr size_t KM_SMEM_OPTIN = 232448 - KM_SMEM_STATIC;
template <class S, typename T> constexpr size_t model_smem() { return (sizeof(Model<S, T>) + 15) / 16 * 16; }
template <class S, typename T> constexpr size_t env_smem() { return (sizeof(Env<S, T>) + 15) / 16 * 16; }
template <class S, typename T> size_t smem_bytes(int epb) { return model_smem<S, T>() + (size_t)epb * env_smem<S, T>(); }
// most envs one CTA can hold in the 227 KB of opt-in shared memory (one warp each at G = 32), and the thread bound
// the kernels are compiled for (it caps registers so that such a CTA is resident: 65536 / threads)
template <class S, typename T> constexpr int max_epb() {
  return (int)((KM_SMEM_OPTIN - model_smem<S, T>()) / env_smem<S, T>()) > 32 ? 32 : (int)((KM_SMEM_OPTIN - model_smem<S, T>()) / env_smem<S, T>());
}
template <class S, typename T> constexpr int max_threads() { return 32 * max_epb<S, T>(); }

template <int G> __device__ __forceinline__ Grp<G> make_group() {
  Grp<G> g;
  g.lane = threadIdx.x % G;
  g.mask = G == 32 ? 0xffffffffu : (((1u << (G & 31)) - 1u) << (((threadIdx.x & 31) / G) * G));
  g.wmask = g.mask;
  return g;
}

template <class S, typename T> __device__ __forceinline__ const Model<S, T>& stage_model(unsigned char* smem, const void* gm) {
  const uint32_t* src = (const uint32_t*)gm;
  uint32_t* dst = (uint32_t*)smem;
  for (int i = threadIdx.x; i < (int)(sizeof(Model<S, T>) / 4); i += blockDim.x) dst[i] = src[i];
  __syncthreads();
  return *(const Model<S, T>*)smem;
}

// state record <-> the leading members of Env (qpos, qvel, ctrl, warm, mocap, time, cube_lo are laid out contiguously)
template <class S, typename T, int G, class E_> __device__ __forceinline__ void load_state(E_& e, const KmArgs& a, long env, const Grp<G>& g) {
  constexpr int SD = Dim<S>::STATE;
  static_assert(offsetof(E_, time) == (SD - 4) * sizeof(T) && offsetof(E_, cube_lo) == (SD - 3) * sizeof(T), "state members of Env must be contiguous");
  const T* src = (const T*)a.state + env * SD;
  T* dst = (T*)&e;
  KM_FOR(i, SD) dst[i] = src[i];
  if (g.lane == 0) { e.step = a.step[env]; e.episode = a.episode[env]; }
  g.sync();
}
template <class S, typename T, int G, class E_> __device__ __forceinline__ void store_state(const E_& e, const KmArgs& a, long env, const Grp<G>& g) {
  constexpr int SD = Dim<S>::STATE;
  T* dst = (T*)a.state + env * SD;
  const T* src = (const T*)&e;
  KM_FOR(i, SD) dst[i] = src[i];
  if (g.lane == 0) { a.step[env] = e.step; a.episode[env] = e.episode; }
}

// one add per CTA to the handle's rollout totals (after every env of the CTA has stepped)
__device__ __forceinline__ void flush_totals(const KmArgs& a, const double* cta_totals) {
  __syncthreads();
  if (a.totals && threadIdx.x < 4 && cta_totals[threadIdx.x] != 0.0) atomicAdd(a.totals + threadIdx.x, cta_totals[threadIdx.x]);
}

template <class S, typename T, int G> __global__ void __launch_bounds__(max_threads<S, T>()) k_env_step(KmArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const Model<S, T>& m = stage_model<S, T>(smem, a.model);
  Grp<G> g = make_group<G>();
  const int slot = threadIdx.x / G;
  Env<S, T>& e = *(Env<S, T>*)(smem + model_smem<S, T>() + (size_t)slot * env_smem<S, T>());
  init_env<S, T, G>(e, m, g);
  __shared__ double cta_totals[4];
  if (threadIdx.x < 4) cta_totals[threadIdx.x] = 0;
  __syncthreads();
  StepOut<T> o = {(T*)a.obs, (T*)a.final_obs, (T*)a.reward, a.trunc, a.term, a.con_flags, a.ncon, a.con_geoms, Dim<S>::MAXCON, a.clk,
                  (T*)a.ep_return, (T*)a.episode_return, (T*)a.final_return, (T*)a.sim_time, a.is_success, a.step_out, a.episode_out,
                  a.totals ? cta_totals : nullptr};
  // every warp walks the same tiles so that the groups sharing a warp can reconverge
  __shared__ long next_tile;
  const long ntiles = ((long)a.n + a.epb - 1) / a.epb;
  for (long t = blockIdx.x; t < ntiles;) {
    const long tile = t * a.epb, idx = tile + slot;
    const bool valid = idx < a.n;
    if (G < 32) {
      __syncwarp();
      g.wmask = __all_sync(0xffffffffu, valid) ? 0xffffffffu : g.mask;
    }
    // env slots past the end of the batch shadow the tile's first env (the CTA marches in phase) and store nothing
    const long idxc = valid ? idx : tile;
    const long envc = a.order ? (long)a.order[idxc] : idxc;
    const StepOut<T> none = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, Dim<S>::MAXCON, nullptr,
                              nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    load_state<S, T, G>(e, a, envc, g);
    env_step<S, T, G>(e, m, g, a.act + envc * m.act_dim, valid ? o : none, envc, a.autoreset, a.seed, a.env0);
    if (valid) {
      store_state<S, T, G>(e, a, envc, g);
      if (g.lane == 0) {
        if (a.niter) a.niter[envc] = e.solver_niter;
        if (a.ls) a.ls[envc] = e.ls_evals;
        if (a.cost) a.cost[envc] = e.ls_evals + 64 * e.ik_evals;
      }
    }
    g.sync();
    if (a.tile_counter) {
      __syncthreads();
      if (threadIdx.x == 0) next_tile = (long)gridDim.x + atomicAdd(a.tile_counter, 1);
      __syncthreads();
      t = next_tile;
    } else t += gridDim.x;
  }
  flush_totals(a, cta_totals);
}

// Thread-per-env mapping (G = 1): every thread owns one env; its working set is one contiguous record in shared
// memory with an odd word stride (conflict-free across lanes, see Env), the model tables sit in front of the records.
template <class S, typename T> struct Tpe {
  typedef Env<S, T, true> E;
  static constexpr size_t unit = sizeof(T) == 4 ? 4 : 8;
  static constexpr size_t words = (sizeof(E) + unit - 1) / unit;
  static constexpr size_t stride = (words | 1) * unit;                 // odd number of bank units per record
  static constexpr int max_envs() {
    const int fit = (int)((KM_SMEM_OPTIN - model_smem<S, T>()) / stride);
    return fit > 128 ? 128 : fit;
  }
  static constexpr int threads() { return (max_envs() + 31) / 32 * 32; }
  static size_t smem(int epb) { return model_smem<S, T>() + (size_t)epb * stride; }
};
// LOCAL: the env record is a local variable (local memory: interleaved across lanes by the hardware, cached in L1/L2)
// instead of a shared-memory record -- no shared-memory limit on resident warps, at the price of cache-latency accesses.
// MAXT: largest CTA the instantiation is compiled for (it caps the registers per thread: 255 up to 256 threads, 128 up to 512)
template <class S, typename T, bool LOCAL, int MAXT = 256> __global__ void __launch_bounds__(LOCAL ? MAXT : Tpe<S, T>::threads()) k_env_step_tpe(KmArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  typedef typename Tpe<S, T>::E E;
  const Model<S, T>& m = stage_model<S, T>(smem, a.model);
  Grp<1> g;
  g.lane = 0; g.mask = 1u; g.wmask = 1u;
  E e_local;
  const bool has_slot = LOCAL || (int)threadIdx.x < a.epb;   // shared-memory records: a CTA may hold fewer envs than threads
  E& e = LOCAL ? e_local : *(E*)(smem + model_smem<S, T>() + (size_t)(has_slot ? threadIdx.x : 0) * Tpe<S, T>::stride);
  if (has_slot) init_env<S, T, 1>(e, m, g);
  __shared__ double cta_totals[4];
  if (threadIdx.x < 4) cta_totals[threadIdx.x] = 0;
  __syncthreads();
  StepOut<T> o = {(T*)a.obs, (T*)a.final_obs, (T*)a.reward, a.trunc, a.term, a.con_flags, a.ncon, a.con_geoms, Dim<S>::MAXCON, a.clk,
                  (T*)a.ep_return, (T*)a.episode_return, (T*)a.final_return, (T*)a.sim_time, a.is_success, a.step_out, a.episode_out,
                  a.totals ? cta_totals : nullptr};
  if (LOCAL) {
    // Every thread of the CTA walks the same number of tiles, and the warps of the CTA are kept in the same phase of
    // the sub-step by CTA barriers (the sub-step body is far larger than the instruction cache); threads past the end
    // of the batch shadow the tile's first env and store nothing.
    g.wmask = 0xffffffffu;
    const StepOut<T> none = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, Dim<S>::MAXCON, nullptr,
                              nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    const long tiles = ((long)a.n + a.epb - 1) / a.epb;
    __shared__ long next_tile;
    for (long tile = blockIdx.x; tile < tiles;) {
      // envs of the tile are dealt to the warps lpw at a time: with few envs per SM every scheduler still gets a
      // warp, and a warp only waits for the slowest of its own lpw envs
      const int lane = threadIdx.x & 31, slot = (threadIdx.x >> 5) * a.lpw + lane;
      const long idx = tile * a.epb + slot;
      const bool valid = lane < a.lpw && slot < a.epb && idx < a.n;
      const long idxc = valid ? idx : tile * a.epb;
      const long envc = a.order ? (long)a.order[idxc] : idxc;
      load_state<S, T, 1>(e, a, envc, g);
      env_step<S, T, 1>(e, m, g, a.act + envc * m.act_dim, valid ? o : none, envc, a.autoreset, a.seed, a.env0);
      if (valid) {
        store_state<S, T, 1>(e, a, envc, g);
        if (a.niter) a.niter[envc] = e.solver_niter;
        if (a.ls) a.ls[envc] = e.ls_evals;
        if (a.cost) a.cost[envc] = e.ls_evals;
      }
      if (a.tile_counter) {
        __syncthreads();
        if (threadIdx.x == 0) next_tile = (long)gridDim.x + atomicAdd(a.tile_counter, 1);
        __syncthreads();
        tile = next_tile;
      } else tile += gridDim.x;
    }
    flush_totals(a, cta_totals);
    return;
  }
  // (a CTA may hold fewer envs than its rounded-up warp: the surplus threads step nothing)
  for (long env = (long)blockIdx.x * a.epb + threadIdx.x; has_slot && env < a.n; env += (long)gridDim.x * a.epb) {
    load_state<S, T, 1>(e, a, env, g);
#ifdef KM_TPE_DEBUG
    const long long t0 = clock64();
    e.solver_niter = 0;
#endif
    env_step<S, T, 1>(e, m, g, a.act + env * m.act_dim, o, env, a.autoreset, a.seed, a.env0);
    store_state<S, T, 1>(e, a, env, g);
    if (a.niter) a.niter[env] = e.solver_niter;
    if (a.ls) a.ls[env] = e.ls_evals;
#ifdef KM_TPE_DEBUG
    if (a.ls) a.ls[env] = (int)((clock64() - t0) >> 10);   // debug build: kilo-cycles of this thread's env step
#endif
  }
  flush_totals(a, cta_totals);
}

template <class S, typename T, int G> __global__ void __launch_bounds__(max_threads<S, T>()) k_reset(KmArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const Model<S, T>& m = stage_model<S, T>(smem, a.model);
  const Grp<G> g = make_group<G>();
  const int slot = threadIdx.x / G;
  Env<S, T>& e = *(Env<S, T>*)(smem + model_smem<S, T>() + (size_t)slot * env_smem<S, T>());
  for (long env = (long)blockIdx.x * a.epb + slot; env < a.n; env += (long)gridDim.x * a.epb) {
    if (a.mask && !a.mask[env]) continue;
    if (g.lane == 0) { e.episode = a.episode[env] + 1; if (a.ep_return) ((T*)a.ep_return)[env] = 0; }
    g.sync();
    reset_state<S, T, G>(e, m, g, a.seed, a.env0 + (unsigned long long)env, a.cube_xyz ? (const T*)a.cube_xyz + 3 * env : (const T*)0);
    observation<S, T, G>(e, m, g);
    if (a.obs) KM_FOR(i, Dim<S>::OBS) ((T*)a.obs)[env * Dim<S>::OBS + i] = e.obs[i];
    store_state<S, T, G>(e, a, env, g);
    g.sync();
  }
}

template <class S, typename T, int G> __global__ void __launch_bounds__(max_threads<S, T>()) k_contacts(KmArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const Model<S, T>& m = stage_model<S, T>(smem, a.model);
  const Grp<G> g = make_group<G>();
  const int slot = threadIdx.x / G;
  Env<S, T>& e = *(Env<S, T>*)(smem + model_smem<S, T>() + (size_t)slot * env_smem<S, T>());
  constexpr int MC = Dim<S>::MAXCON;
  for (long env = (long)blockIdx.x * a.epb + slot; env < a.n; env += (long)gridDim.x * a.epb) {
    load_state<S, T, G>(e, a, env, g);
    kinematics<S, T, G>(e, m, g);
    collision<S, T, G>(e, m, g);
    if (g.lane == 0) {
      if (a.ncon) a.ncon[env] = e.ncon;
      if (a.con_geoms)
        for (int c = 0; c < MC; c++) {
          int g1 = -1, g2 = -1;
          if (c < e.ncon) { const int s = e.con_slot[c]; g1 = s < Dim<S>::NPAD ? m.pad_geom[s] : m.table_geom; g2 = m.cube_geom; }
          a.con_geoms[env * 2 * MC + 2 * c] = g1;
          a.con_geoms[env * 2 * MC + 2 * c + 1] = g2;
        }
      // end-effector site frames (callers read physics.data.site("eer_site_pos").xpos, reference examples/2_synthetic_data.py:34)
      if (a.site_pos || a.site_mat)
        for (int arm = 0; arm < m.n_arm; arm++) {
          T pos[3], mat[9];
          site_pose<S, T, G>(e, m, arm, pos, mat);
          if (a.site_pos) for (int i = 0; i < 3; i++) ((T*)a.site_pos)[(env * m.n_arm + arm) * 3 + i] = pos[i];
          if (a.site_mat) for (int i = 0; i < 9; i++) ((T*)a.site_mat)[(env * m.n_arm + arm) * 9 + i] = mat[i];
        }
    }
    g.sync();
  }
}

// Camera observations, stage 1: position stage on the stored state -> render record (km_render.cuh).  Not hot (the pixel
// kernel is): one env per warp, four per CTA.
template <class S, typename T, int G> __global__ void __launch_bounds__(max_threads<S, T>()) k_render_setup(KmArgs a, KmRenderParams P, float* recs) {
  extern __shared__ __align__(16) unsigned char smem[];
  const Model<S, T>& m = stage_model<S, T>(smem, a.model);
  const Grp<G> g = make_group<G>();
  const int slot = threadIdx.x / G;
  Env<S, T>& e = *(Env<S, T>*)(smem + model_smem<S, T>() + (size_t)slot * env_smem<S, T>());
  for (long env = (long)blockIdx.x * a.epb + slot; env < a.n; env += (long)gridDim.x * a.epb) {
    load_state<S, T, G>(e, a, env, g);
    kinematics<S, T, G>(e, m, g);
    render_record<S, T, G>(e, m, g, P, recs + env * render_rec_floats<S>());
    g.sync();
  }
}

template <class S, typename T> struct Launch {
  typedef Dim<S> D;
  template <int G> static cudaError_t run(int which, const KmArgs& a) {
    const size_t sm = smem_bytes<S, T>(a.epb) + (size_t)a.epb * a.trf_bytes;
    const dim3 block(a.epb * G), grid(a.grid);
    if (which == 0) k_env_step<S, T, G><<<grid, block, sm, a.stream>>>(a);
    else if (which == 1) k_reset<S, T, G><<<grid, block, sm, a.stream>>>(a);
    else k_contacts<S, T, G><<<grid, block, sm, a.stream>>>(a);
    return cudaGetLastError();
  }
  static cudaError_t dispatch(int which, const KmArgs& a) {
    if (a.G == 1 && which == 0) {
      k_env_step_tpe<S, T, false><<<dim3(a.grid), dim3((a.epb + 31) / 32 * 32), Tpe<S, T>::smem(a.epb), a.stream>>>(a);
      return cudaGetLastError();
    }
    if (a.G == 2 && which == 0) {   // thread per env, record in local memory
      const int threads = (a.epb + a.lpw - 1) / a.lpw * 32;   // warps needed to hold epb envs at lpw active lanes each
      if (threads <= 256 && !a.tpl_small_regs) k_env_step_tpe<S, T, true, 256><<<dim3(a.grid), dim3(threads), model_smem<S, T>(), a.stream>>>(a);
      else k_env_step_tpe<S, T, true, 512><<<dim3(a.grid), dim3(threads), model_smem<S, T>(), a.stream>>>(a);
      return cudaGetLastError();
    }
    if (a.G == 1 || a.G == 2) {   // reset / contacts are not hot: one env per warp with a small CTA
      KmArgs b = a;
      b.G = 32; b.epb = 4; b.grid = (a.n + 3) / 4 < 148 * 8 ? (a.n + 3) / 4 : 148 * 8;
      return run<32>(which, b);
    }
    if (a.G == 32) return run<32>(which, a);
    if constexpr (D::NV <= 16) { if (a.G == 16) return run<16>(which, a); }
    return cudaErrorInvalidValue;
  }
  static cudaError_t step(const KmArgs& a) { return dispatch(0, a); }
  static cudaError_t reset(const KmArgs& a) { return dispatch(1, a); }
  static cudaError_t contacts(const KmArgs& a) { return dispatch(2, a); }
  static double table_z(const void* host_model) { return (double)((const Model<S, T>*)host_model)->tab_z; }
  static cudaError_t render_setup(const KmArgs& a, const KmRenderParams& P, float* recs) {
    cudaError_t err = cudaFuncSetAttribute(k_render_setup<S, T, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<S, T>(4));
    if (err != cudaSuccess) return err;
    const int grid = (a.n + 3) / 4 < 148 * 8 ? (a.n + 3) / 4 : 148 * 8;
    KmArgs b = a;
    b.epb = 4;
    k_render_setup<S, T, 32><<<dim3(grid), dim3(128), smem_bytes<S, T>(4), a.stream>>>(b, P, recs);
    return cudaGetLastError();
  }
  template <int G> static cudaError_t prep(int epb, int extra, int* ctas) {
    // the attribute is per function, not per handle: always opt in to the device maximum so that handles with
    // different envs-per-CTA can coexist in one process
    int dev = 0, optin = 0;
    cudaError_t err;
    if ((err = cudaGetDevice(&dev)) != cudaSuccess) return err;
    if ((err = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess) return err;
    if ((err = cudaFuncSetAttribute(k_env_step<S, T, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)KM_SMEM_STATIC)) != cudaSuccess) return err;
    if ((err = cudaFuncSetAttribute(k_reset<S, T, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)KM_SMEM_STATIC)) != cudaSuccess) return err;
    if ((err = cudaFuncSetAttribute(k_contacts<S, T, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)KM_SMEM_STATIC)) != cudaSuccess) return err;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas, k_env_step<S, T, G>, epb * G, smem_bytes<S, T>(epb) + (size_t)epb * extra);
  }
  static cudaError_t prepare(int G, int epb, int extra, int* ctas) {
    if (G == 1) {
      cudaError_t err = prep<32>(4, 0, ctas);
      if (err != cudaSuccess) return err;
      int dev = 0, optin = 0;
      if ((err = cudaGetDevice(&dev)) != cudaSuccess) return err;
      if ((err = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess) return err;
      if ((err = cudaFuncSetAttribute(k_env_step_tpe<S, T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)KM_SMEM_STATIC)) != cudaSuccess) return err;
      return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas, k_env_step_tpe<S, T, false>, (epb + 31) / 32 * 32, Tpe<S, T>::smem(epb));
    }
    if (G == 2) {
      cudaError_t err = prep<32>(4, 0, ctas);
      if (err != cudaSuccess) return err;
      // the env records live in local memory: give the unified L1 / shared-memory array to the cache
      if ((err = cudaFuncSetAttribute(k_env_step_tpe<S, T, true, 256>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1)) != cudaSuccess) return err;
      if ((err = cudaFuncSetAttribute(k_env_step_tpe<S, T, true, 512>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1)) != cudaSuccess) return err;
      if (epb <= 256) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas, k_env_step_tpe<S, T, true, 256>, (epb + 31) / 32 * 32, model_smem<S, T>());
      return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas, k_env_step_tpe<S, T, true, 512>, (epb + 31) / 32 * 32, model_smem<S, T>());
    }
    if (G == 32) return prep<32>(epb, extra, ctas);
    if constexpr (D::NV <= 16) { if (G == 16) return prep<16>(epb, extra, ctas); }
    return cudaErrorInvalidValue;
  }
  static int fill(const km_model* fm, const km_task* tk, void* dst, std::string& err) {
    return fill_model<S, T>(fm, tk, (Model<S, T>*)dst, err);
  }
  static KmVtable vtable() {
    KmVtable v;
    v.model_bytes = sizeof(Model<S, T>); v.env_bytes = env_smem<S, T>(); v.scalar_bytes = sizeof(T);
    v.nq = D::NQ; v.nv = D::NV; v.nu = D::NU; v.nmocap = D::NMOCAP; v.obs_dim = D::OBS;
    v.state_dim = D::STATE; v.maxcon = D::MAXCON; v.nlanes_min = D::NV <= 16 ? 16 : 32; v.max_threads = max_threads<S, T>(); v.tpe_max_envs = Tpe<S, T>::max_envs(); v.tpe_env_bytes = Tpe<S, T>::stride;
    v.render_rec_floats = render_rec_floats<S>(); v.render_setup = &render_setup; v.table_z = &table_z;
    v.fill = &fill; v.step = &step; v.reset = &reset; v.contacts = &contacts; v.prepare = &prepare;
    return v;
  }
};

#endif  // __CUDACC__

}  // namespace km
