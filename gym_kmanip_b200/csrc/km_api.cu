// km_api.cu -- the C-ABI of include/kmanip_b200.h: handle management, buffers, launches.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#define KM_RENDER_PIXELS_IMPL
#include "km_launch.cuh"

namespace km {
KmVtable vtable_solo_arm_f32();
KmVtable vtable_solo_arm_f64();
KmVtable vtable_dual_arm_f32();
KmVtable vtable_dual_arm_f64();
KmVtable vtable_torso_f32();
KmVtable vtable_torso_f64();
}  // namespace km

using km::KmArgs;
using km::KmVtable;

static thread_local std::string g_err;

struct km_sim {
  KmVtable vt;
  int scene, dtype, n, device, act_dim, n_arm, ik_mode;
  unsigned long long seed, env0;
  int G, epb, grid, ctas_per_sm, num_sms, lpw, tpl_ctas;
  void* d_model;
  void* d_state;
  int *d_step, *d_episode, *d_niter, *d_ls;
  int *d_order, *d_tile_counter, *d_cost;   // cost-ordered walk of the step kernels (order_envs)
  int order_mode;   // -1 = automatic (on when the lane-group mapping walks more tiles than CTAs), 0 = off, 1 = on
  unsigned* d_clk;   // caller-owned buffer of km_debug_phase_clocks (debug builds)
  void* d_ep_return;  // running return of every env (dtype of the handle)
  double* d_totals;   // rollout totals {sum reward, env steps, finished episodes, success steps}
  cudaStream_t host_stream;   // stream of the *_host entry points
  // staging for the host-buffer entry points
  float* d_act;
  void *d_obs, *d_reward, *d_xyz;
  unsigned char *d_trunc, *d_mask;
  // camera observations: per-env render records, staging image for km_render_host
  float* d_recs;
  unsigned char* d_rgb;
  size_t rgb_bytes;
  double tab_z;
  long long launches;
};

static int fail(int code, const std::string& msg) { g_err = msg; return code; }
static int cuda_fail(cudaError_t e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return KM_ERR_CUDA;
}
#define KM_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return cuda_fail(e__, #call); } while (0)

struct DeviceGuard {
  int prev;
  bool ok;
  explicit DeviceGuard(int dev) : prev(-1), ok(false) {
    if (cudaGetDevice(&prev) != cudaSuccess) return;
    ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// envs per CTA of the thread-per-env (local memory) mapping: one CTA per SM holding its share of the batch, equal
// CTAs in as few full waves as possible (measured: 65536 solo-arm envs as 147 CTAs of 446 envs 6.6e6 env-steps/s, as
// 256 tiles of 256 over 148 CTAs 5.4e6)
static int tpe_local_envs(const km_sim* h, int ctas) {
  const long cap = 512 / ctas, slots = (long)h->num_sms * ctas, per_wave = slots * cap;
  const long waves = ((long)h->n + per_wave - 1) / per_wave;
  const long per_cta = ((long)h->n + slots * waves - 1) / (slots * waves);
  return (int)(per_cta > cap ? cap : per_cta);
}

static int configure(km_sim* h, int G, int epb) {
  if (G == 0) G = h->G;
  if (h->ik_mode == 1 && (G == 1 || G == 2))
    return fail(KM_ERR_ARG, "ik_mode = 1 (exact-parity TRF IK) runs in the lane-group mapping only (lanes_per_env 16 / 32)");
  if (G == 2) {   // thread-per-env with the env record in local memory: epb = threads per CTA (32..256)
    // experiment knob: KM_TPL_CTAS = CTAs per SM of this mapping (default 1; more CTAs = smaller barrier domains)
    const char* ev = std::getenv("KM_TPL_CTAS");
    const int want = ev ? std::atoi(ev) : 1;
    h->tpl_ctas = want >= 1 && want <= 4 ? want : 1;
    if (epb == 0) epb = tpe_local_envs(h, h->tpl_ctas);
    // Active lanes per warp: all 32.  Dealing the envs of a CTA to more warps with fewer active lanes each was measured
    // and is far worse (4096 solo-arm envs as 14 warps x 2 lanes per SM: 6.7 ms per launch against 2.7 ms): every
    // warp executes the whole instruction stream, so issue slots, not latency, become the limit.
    h->lpw = 32;
    const int threads = (epb + h->lpw - 1) / h->lpw * 32;
    if (epb < 1 || threads > 512) return fail(KM_ERR_ARG, "envs_per_block out of range for thread-per-env (local) CTAs");
    int ctas = 0;
    KM_CUDA(h->vt.prepare(2, threads, 0, &ctas));
    if (ctas < 1) return fail(KM_ERR_CUDA, "kernel does not fit on an SM with this configuration");
    // one CTA per SM by default: the records of the resident envs (6.9 KB each for the solo arm) should stay in L2
    ctas = h->tpl_ctas;
    h->G = 2; h->epb = epb; h->ctas_per_sm = ctas;
    const long tiles = ((long)h->n + epb - 1) / epb, resident = (long)h->num_sms * ctas;
    h->grid = (int)(tiles < resident ? tiles : resident);
    return KM_OK;
  }
  if (G == 1) {   // thread-per-env: epb = envs (threads) per CTA, bounded by the shared memory one CTA can hold
    const int cap = h->vt.tpe_max_envs;
    if (cap < 1) return fail(KM_ERR_ARG, "an env does not fit in shared memory");
    if (epb == 0) {   // balance the waves over the SMs (4096 envs on 148 SMs -> 28 envs per CTA, one wave)
      const long per_wave = (long)h->num_sms * cap, waves = ((long)h->n + per_wave - 1) / per_wave;
      epb = (int)(((long)h->n + h->num_sms * waves - 1) / (h->num_sms * waves));
      if (epb > cap) epb = cap;
    }
    if (epb < 1 || epb > cap) return fail(KM_ERR_ARG, "envs_per_block out of range for thread-per-env CTAs");
    int ctas = 0;
    KM_CUDA(h->vt.prepare(1, epb, 0, &ctas));
    if (ctas < 1) return fail(KM_ERR_CUDA, "kernel does not fit on an SM with this configuration");
    h->G = 1; h->epb = epb; h->ctas_per_sm = ctas;
    const long tiles = ((long)h->n + epb - 1) / epb, resident = (long)h->num_sms * ctas;
    h->grid = (int)(tiles < resident ? tiles : resident);
    return KM_OK;
  }
  if ((G != 16 && G != 32) || G < h->vt.nlanes_min) return fail(KM_ERR_ARG, "lanes_per_env must be 1, 16 or 32 and (for 16 / 32) at least the number of dofs");
  int dev_smem = 0;
  KM_CUDA(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
  dev_smem -= (int)km::KM_SMEM_STATIC;   // the kernels' static shared memory (CTA totals) comes out of the same budget
  const size_t model_b = (h->vt.model_bytes + 15) / 16 * 16;
  // the exact-parity IK keeps its work arrays in shared memory behind the env records (km_ik_trf.cuh)
  const size_t env_b = h->vt.env_bytes + (h->ik_mode == 1 ? sizeof(km::trf::TrfWork) : 0);
  const int fit = (int)(((size_t)dev_smem - model_b) / env_b);
  const int cap = fit * G > h->vt.max_threads ? h->vt.max_threads / G : fit;
  if (epb == 0) {
    // default: fill the shared memory of every SM with envs, shrunk so that the waves are balanced (4096 envs on 148 SMs
    // -> 28 per SM, a single wave); when an SM holds 16 or more, split them over two CTAs -- the CTA-wide lockstep then
    // couples fewer envs (measured +6..8 % for the solo-arm scene, profiles/r01_notes.md)
    if (cap < 1) return fail(KM_ERR_ARG, "an env does not fit in shared memory");
    const long per_wave = (long)h->num_sms * cap;
    const long waves = ((long)h->n + per_wave - 1) / per_wave;
    int per_sm = (int)(((long)h->n + h->num_sms * waves - 1) / (h->num_sms * waves));
    if (per_sm > cap) per_sm = cap;
    epb = per_sm >= 16 ? (per_sm + 1) / 2 : per_sm;
  }
  if (epb < 1 || epb > cap) return fail(KM_ERR_ARG, "envs_per_block out of range for this scene / precision");
  if (model_b + (size_t)epb * env_b > (size_t)dev_smem)
    return fail(KM_ERR_ARG, "envs_per_block needs more shared memory than a CTA can opt in to");
  int ctas = 0;
  KM_CUDA(h->vt.prepare(G, epb, (int)(env_b - h->vt.env_bytes), &ctas));
  if (ctas < 1) return fail(KM_ERR_CUDA, "kernel does not fit on an SM with this configuration");
  h->G = G; h->epb = epb; h->ctas_per_sm = ctas;
  const long tiles = ((long)h->n + epb - 1) / epb;
  const long resident = (long)h->num_sms * ctas;
  h->grid = (int)(tiles < resident ? tiles : resident);
  return KM_OK;
}

static KmArgs base_args(km_sim* h, void* stream) {
  KmArgs a;
  std::memset(&a, 0, sizeof(a));
  a.model = h->d_model; a.state = h->d_state; a.step = h->d_step; a.episode = h->d_episode;
  a.niter = h->d_niter; a.ls = h->d_ls; a.cost = h->d_cost; a.clk = h->d_clk;
  a.ep_return = h->d_ep_return; a.totals = h->d_totals;
  a.n = h->n; a.seed = h->seed; a.env0 = h->env0; a.G = h->G; a.epb = h->epb; a.grid = h->grid; a.lpw = h->lpw > 0 ? h->lpw : 32; a.tpl_small_regs = h->tpl_ctas > 1;
  a.trf_bytes = h->ik_mode == 1 ? (int)sizeof(km::trf::TrfWork) : 0;
  a.stream = (cudaStream_t)stream;
  return a;
}

// FMA-pipe micro-benchmark: 8 independent dependent-FMA chains per thread, enough threads to fill every SM
template <typename T> __global__ void k_fma_peak(T* out, int iters) {
  T a0 = (T)threadIdx.x * (T)1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const T b = (T)0.999, c = (T)1e-4;
  for (int i = 0; i < iters; i++) {
    a0 = a0 * b + c; a1 = a1 * b + c; a2 = a2 * b + c; a3 = a3 * b + c;
    a4 = a4 * b + c; a5 = a5 * b + c; a6 = a6 * b + c; a7 = a7 * b + c;
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
template <typename T> static int measure_fma(int device, double* tflops) {
  DeviceGuard guard(device);
  if (!guard.ok) return fail(KM_ERR_CUDA, "cudaSetDevice failed");
  int sms = 0;
  KM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  const int threads = 1024, blocks = sms * 2, iters = 1 << 16;
  T* out = nullptr;
  KM_CUDA(cudaMalloc((void**)&out, (size_t)threads * blocks * sizeof(T)));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    k_fma_peak<T><<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) break;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double tf = 2.0 * 8.0 * (double)iters * threads * blocks / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(out);
  KM_CUDA(cudaGetLastError());
  *tflops = best;
  return KM_OK;
}

// Cost-ordered walk.  The step kernels march the envs of a CTA (warp-per-env mapping) or of a warp (thread-per-env
// mapping) in lockstep, so a tile costs what its slowest env costs -- and the cost of an env (Newton iterations, line-search
// evaluations: contact state) persists from one step to the next.  Before every step the envs are therefore bucketed by
// the line-search evaluations of their previous step (counting sort, 256 buckets, most expensive first): tiles become
// homogeneous, and the CTAs fetch them dynamically, longest first.  One CTA: the batch is at most a few 100 000 envs.
__global__ void __launch_bounds__(1024) k_order_envs(const int* __restrict__ cost, int n, int shift, int* __restrict__ order, int* tile_counter) {
  __shared__ int hist[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int c = cost[i] >> shift;
    atomicAdd(&hist[255 - (c > 255 ? 255 : c)], 1);
  }
  __syncthreads();
  if (threadIdx.x < 32) {   // exclusive scan of the 256 counts by one warp (8 per lane)
    int v[8], s = 0;
    for (int k = 0; k < 8; k++) { v[k] = hist[threadIdx.x * 8 + k]; s += v[k]; }
    int incl = s;
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if ((int)threadIdx.x >= o) incl += t; }
    int run = incl - s;
    for (int k = 0; k < 8; k++) { hist[threadIdx.x * 8 + k] = run; run += v[k]; }
    if (threadIdx.x == 0) *tile_counter = 0;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int c = cost[i] >> shift;
    order[atomicAdd(&hist[255 - (c > 255 ? 255 : c)], 1)] = i;
  }
}

extern "C" {

int km_measure_fma_peak(int device, int dtype, double* tflops) {
  if (!tflops) return fail(KM_ERR_ARG, "null output");
  return dtype == KM_F64 ? measure_fma<double>(device, tflops) : measure_fma<float>(device, tflops);
}

const char* km_last_error(void) { return g_err.c_str(); }
const char* km_version(void) { return "kmanip_b200 0.1 (sm_100a)"; }

int km_create(const km_model* model, const km_task* task, int scene, int n_envs, int device, int dtype,
              uint64_t seed, uint64_t env0, km_handle* out) {
  if (!model || !task || !out || n_envs < 1) return fail(KM_ERR_ARG, "km_create: bad argument");
  if (dtype != KM_F32 && dtype != KM_F64) return fail(KM_ERR_ARG, "km_create: dtype must be 32 or 64");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1 || device < 0 || device >= ndev) {
    cudaGetLastError();
    return fail(KM_ERR_NODEVICE, "km_create: no usable CUDA device (this library has no CPU path)");
  }
  km_sim* h = new km_sim();
  std::memset((void*)h, 0, sizeof(*h));
  switch (scene * 2 + (dtype == KM_F64)) {
    case 0: h->vt = km::vtable_solo_arm_f32(); break;
    case 1: h->vt = km::vtable_solo_arm_f64(); break;
    case 2: h->vt = km::vtable_dual_arm_f32(); break;
    case 3: h->vt = km::vtable_dual_arm_f64(); break;
    case 4: h->vt = km::vtable_torso_f32(); break;
    case 5: h->vt = km::vtable_torso_f64(); break;
    default: delete h; return fail(KM_ERR_ARG, "km_create: unknown scene");
  }
  h->scene = scene; h->dtype = dtype; h->n = n_envs; h->device = device; h->seed = seed; h->env0 = env0; h->act_dim = task->act_dim; h->n_arm = task->n_arm; h->ik_mode = task->ik_mode;
  std::vector<unsigned char> host_model(h->vt.model_bytes);
  std::string err;
  if (h->vt.fill(model, task, host_model.data(), err) != 0) { delete h; return fail(KM_ERR_MODEL, err); }
  h->tab_z = h->vt.table_z(host_model.data());
  DeviceGuard guard(device);
  if (!guard.ok) { delete h; return fail(KM_ERR_CUDA, "km_create: cudaSetDevice failed"); }
  cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device);
  cudaError_t e;
  const size_t sb = h->vt.scalar_bytes, n = (size_t)n_envs;
#define KM_ALLOC(ptr, bytes) if ((e = cudaMalloc((void**)&(ptr), (bytes))) != cudaSuccess) { km_destroy(h); return cuda_fail(e, "cudaMalloc"); }
  KM_ALLOC(h->d_model, h->vt.model_bytes);
  KM_ALLOC(h->d_state, n * h->vt.state_dim * sb);
  KM_ALLOC(h->d_step, n * sizeof(int));
  KM_ALLOC(h->d_episode, n * sizeof(int));
  KM_ALLOC(h->d_niter, n * sizeof(int));
  KM_ALLOC(h->d_ls, n * sizeof(int));
  KM_ALLOC(h->d_order, n * sizeof(int));
  KM_ALLOC(h->d_cost, n * sizeof(int));
  KM_ALLOC(h->d_tile_counter, sizeof(int));
  KM_ALLOC(h->d_act, n * task->act_dim * sizeof(float));
  KM_ALLOC(h->d_obs, n * h->vt.obs_dim * sb);
  KM_ALLOC(h->d_reward, n * sb);
  KM_ALLOC(h->d_xyz, n * 3 * sb);
  KM_ALLOC(h->d_trunc, n);
  KM_ALLOC(h->d_mask, n);
  KM_ALLOC(h->d_ep_return, n * sb);
  KM_ALLOC(h->d_totals, 4 * sizeof(double));
#undef KM_ALLOC
  if ((e = cudaMemcpy(h->d_model, host_model.data(), h->vt.model_bytes, cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMemset(h->d_state, 0, n * h->vt.state_dim * sb)) != cudaSuccess ||
      (e = cudaMemset(h->d_step, 0, n * sizeof(int))) != cudaSuccess ||
      (e = cudaMemset(h->d_episode, 0xff, n * sizeof(int))) != cudaSuccess ||   // -1: the first reset starts episode 0
      (e = cudaMemset(h->d_niter, 0, n * sizeof(int))) != cudaSuccess ||
      (e = cudaMemset(h->d_ep_return, 0, n * sb)) != cudaSuccess ||
      (e = cudaMemset(h->d_totals, 0, 4 * sizeof(double))) != cudaSuccess ||
      (e = cudaMemset(h->d_cost, 0, n * sizeof(int))) != cudaSuccess ||
      (e = cudaMemset(h->d_ls, 0, n * sizeof(int))) != cudaSuccess) {
    km_destroy(h);
    return cuda_fail(e, "km_create: initialisation");
  }
  // Default mapping, from the measured grid of profiles/r01_notes.md: the lane-group kernel (one env per warp) while
  // the batch gives an SM fewer envs than a few warps of threads, the thread-per-env kernel with records in local
  // memory beyond that (its launch time no longer hangs on single slow envs, and it issues ~7x fewer instructions).
  const int per_sm = (n_envs + h->num_sms - 1) / h->num_sms;
  int rc;
  // envs per SM from which the thread-per-env kernel wins (measured crossovers with the register-resident warp-per-env
  // solver and the cost-ordered walk, profiles/r02_mapping_sweep.log: solo arm, joint actions: 6.75 vs 6.2e6 env-steps/s at
  // 221 envs per SM, 7.0 vs 7.8e6 at 443; solo arm + IK: 3.9 vs 3.6e6 at 111; dual arm fp32: 2.05 vs 1.34e6 at 55, 2.2 vs
  // 3.2e6 at 221; torso fp64: 1.69 vs 1.37e6 at 28, 1.65 vs 2.16e6 at 55; torso fp32: 3.6 vs 2.9e6 at 55, 3.7 vs 6.5e6 at 221)
  int tpe_from;
  if (h->vt.nv <= 16) tpe_from = task->act_mode == 1 ? 300 : 150;
  else tpe_from = dtype == KM_F64 ? 40 : 100;
  if (per_sm >= tpe_from && h->ik_mode != 1) rc = configure(h, 2, 0);   // the exact-parity IK mode lives in the lane-group kernels
  else {
    h->G = 32;
    rc = configure(h, 32, 0);
  }
  if (rc != KM_OK) { km_destroy(h); return rc; }
  { const char* ev = std::getenv("KM_ORDER"); h->order_mode = ev ? std::atoi(ev) : -1; }   // experiment knob; km_set_env_ordering
  *out = h;
  return KM_OK;
}

void km_destroy(km_handle h) {
  if (!h) return;
  DeviceGuard guard(h->device);
  void* ptrs[] = {h->d_model, h->d_state, h->d_step, h->d_episode, h->d_niter, h->d_ls, h->d_act, h->d_obs, h->d_reward,
                  h->d_xyz, h->d_trunc, h->d_mask, h->d_recs, h->d_rgb, h->d_ep_return, h->d_totals, h->d_order, h->d_tile_counter, h->d_cost};
  for (void* p : ptrs) if (p) cudaFree(p);
  delete h;
}

int km_nq(km_handle h) { return h->vt.nq; }
int km_nv(km_handle h) { return h->vt.nv; }
int km_nu(km_handle h) { return h->vt.nu; }
int km_nmocap(km_handle h) { return h->vt.nmocap; }
int km_obs_dim(km_handle h) { return h->vt.obs_dim; }
int km_act_dim(km_handle h) { return h->act_dim; }
int km_state_dim(km_handle h) { return h->vt.state_dim; }
int km_max_contacts(km_handle h) { return h->vt.maxcon; }
int km_num_envs(km_handle h) { return h->n; }
int km_dtype(km_handle h) { return h->dtype; }
long long km_launch_count(km_handle h) { return h->launches; }
void* km_state_ptr(km_handle h) { return h->d_state; }

int km_configure(km_handle h, int lanes_per_env, int envs_per_block) {
  if (!h) return fail(KM_ERR_ARG, "null handle");
  DeviceGuard guard(h->device);
  return configure(h, lanes_per_env, envs_per_block);
}

int km_set_env_ordering(km_handle h, int mode) {
  if (!h || mode < -1 || mode > 1) return fail(KM_ERR_ARG, "km_set_env_ordering: mode must be -1 (automatic), 0 or 1");
  h->order_mode = mode;
  return KM_OK;
}

int km_launch_config(km_handle h, int* lanes_per_env, int* envs_per_block, int* grid, int* ctas_per_sm, int* smem_bytes) {
  if (!h) return fail(KM_ERR_ARG, "null handle");
  if (lanes_per_env) *lanes_per_env = h->G;
  if (envs_per_block) *envs_per_block = h->epb;
  if (grid) *grid = h->grid;
  if (ctas_per_sm) *ctas_per_sm = h->ctas_per_sm;
  if (smem_bytes)   // the local-memory mapping (G == 2) keeps only the model tables in shared memory
    *smem_bytes = (int)((h->vt.model_bytes + 15) / 16 * 16 + (h->G == 2 ? 0 : (size_t)h->epb * (h->G == 1 ? h->vt.tpe_env_bytes : h->vt.env_bytes + (h->ik_mode == 1 ? sizeof(km::trf::TrfWork) : 0))));
  return KM_OK;
}

int km_reset(km_handle h, const unsigned char* mask_dev, const void* cube_xyz_dev, void* obs_dev, void* stream) {
  if (!h) return fail(KM_ERR_ARG, "null handle");
  DeviceGuard guard(h->device);
  KmArgs a = base_args(h, stream);
  a.mask = mask_dev; a.cube_xyz = cube_xyz_dev; a.obs = obs_dev;
  KM_CUDA(h->vt.reset(a));
  h->launches++;
  return KM_OK;
}

int km_step(km_handle h, const float* action_dev, const km_step_out* out, int autoreset, void* stream) {
  if (!h || !action_dev) return fail(KM_ERR_ARG, "km_step: null handle or action");
  DeviceGuard guard(h->device);
  KmArgs a = base_args(h, stream);
  a.act = action_dev; a.autoreset = autoreset;
  if (out) {
    a.obs = out->obs; a.final_obs = out->final_obs; a.reward = out->reward; a.trunc = out->truncated; a.term = out->terminated;
    a.con_flags = out->con_flags; a.ncon = out->ncon; a.con_geoms = out->con_geoms;
    a.is_success = out->is_success; a.episode_return = out->episode_return; a.final_return = out->final_return;
    a.sim_time = out->sim_time; a.step_out = out->step_count; a.episode_out = out->episode;
  }
  const long tiles_ = ((long)h->n + h->epb - 1) / h->epb;
  if (h->order_mode == 1 || (h->order_mode < 0 && h->G >= 16 && tiles_ > h->grid)) {
    // 256 buckets: two line-search evaluations wide, or (exact-parity IK, whose evaluations weigh 64) an eighth of an IK evaluation
    k_order_envs<<<1, 1024, 0, a.stream>>>(h->d_cost, h->n, h->ik_mode == 1 ? 3 : 1, h->d_order, h->d_tile_counter);
    KM_CUDA(cudaGetLastError());
    a.order = h->d_order; a.tile_counter = h->d_tile_counter;
    h->launches++;
  }
  KM_CUDA(h->vt.step(a));
  h->launches++;
  return KM_OK;
}

int km_get_state(km_handle h, void* state_dev, int* step_count_dev, int* episode_dev, void* stream) {
  if (!h) return fail(KM_ERR_ARG, "null handle");
  DeviceGuard guard(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = (size_t)h->n;
  if (state_dev) KM_CUDA(cudaMemcpyAsync(state_dev, h->d_state, n * h->vt.state_dim * h->vt.scalar_bytes, cudaMemcpyDeviceToDevice, s));
  if (step_count_dev) KM_CUDA(cudaMemcpyAsync(step_count_dev, h->d_step, n * sizeof(int), cudaMemcpyDeviceToDevice, s));
  if (episode_dev) KM_CUDA(cudaMemcpyAsync(episode_dev, h->d_episode, n * sizeof(int), cudaMemcpyDeviceToDevice, s));
  return KM_OK;
}

int km_set_state(km_handle h, const void* state_dev, const int* step_count_dev, const int* episode_dev, void* stream) {
  if (!h) return fail(KM_ERR_ARG, "null handle");
  DeviceGuard guard(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = (size_t)h->n;
  if (state_dev) KM_CUDA(cudaMemcpyAsync(h->d_state, state_dev, n * h->vt.state_dim * h->vt.scalar_bytes, cudaMemcpyDeviceToDevice, s));
  if (step_count_dev) KM_CUDA(cudaMemcpyAsync(h->d_step, step_count_dev, n * sizeof(int), cudaMemcpyDeviceToDevice, s));
  if (episode_dev) KM_CUDA(cudaMemcpyAsync(h->d_episode, episode_dev, n * sizeof(int), cudaMemcpyDeviceToDevice, s));
  return KM_OK;
}

int km_contacts(km_handle h, int* ncon_dev, int* con_geoms_dev, void* stream) {
  if (!h) return fail(KM_ERR_ARG, "null handle");
  DeviceGuard guard(h->device);
  KmArgs a = base_args(h, stream);
  a.ncon = ncon_dev; a.con_geoms = con_geoms_dev;
  KM_CUDA(h->vt.contacts(a));
  h->launches++;
  return KM_OK;
}

int km_site_poses(km_handle h, void* xpos_dev, void* xmat_dev, void* stream) {
  if (!h) return fail(KM_ERR_ARG, "null handle");
  DeviceGuard guard(h->device);
  KmArgs a = base_args(h, stream);
  a.site_pos = xpos_dev; a.site_mat = xmat_dev;
  KM_CUDA(h->vt.contacts(a));
  h->launches++;
  return KM_OK;
}

int km_n_arm(km_handle h) { return h->n_arm; }

int km_solver_stats(km_handle h, int* niter_dev, int* ls_evals_dev, void* stream) {
  if (!h) return fail(KM_ERR_ARG, "null handle");
  DeviceGuard guard(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  if (niter_dev) KM_CUDA(cudaMemcpyAsync(niter_dev, h->d_niter, (size_t)h->n * sizeof(int), cudaMemcpyDeviceToDevice, s));
  if (ls_evals_dev) KM_CUDA(cudaMemcpyAsync(ls_evals_dev, h->d_ls, (size_t)h->n * sizeof(int), cudaMemcpyDeviceToDevice, s));
  return KM_OK;
}

int km_episode_stats(km_handle h, double* totals_dev, void* episode_return_dev, int reset, void* stream) {
  if (!h) return fail(KM_ERR_ARG, "null handle");
  DeviceGuard guard(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  if (totals_dev) KM_CUDA(cudaMemcpyAsync(totals_dev, h->d_totals, 4 * sizeof(double), cudaMemcpyDeviceToDevice, s));
  if (episode_return_dev) KM_CUDA(cudaMemcpyAsync(episode_return_dev, h->d_ep_return, (size_t)h->n * h->vt.scalar_bytes, cudaMemcpyDeviceToDevice, s));
  if (reset) KM_CUDA(cudaMemsetAsync(h->d_totals, 0, 4 * sizeof(double), s));
  return KM_OK;
}

int km_set_host_stream(km_handle h, void* stream) {
  if (!h) return fail(KM_ERR_ARG, "null handle");
  h->host_stream = (cudaStream_t)stream;
  return KM_OK;
}

int km_debug_phase_clocks(km_handle h, unsigned* clk_dev) {
  if (!h) return fail(KM_ERR_ARG, "null handle");
#ifdef KM_PHASE_CLOCKS
  h->d_clk = clk_dev;
  return KM_OK;
#else
  (void)clk_dev;
  return fail(KM_ERR_ARG, "km_debug_phase_clocks: the library was built without -DKM_PHASE_CLOCKS");
#endif
}

int km_render(km_handle h, const km_camera* cam, const km_visual* vis, unsigned char* rgb_dev, void* stream) {
  if (!h || !cam || !vis || !rgb_dev) return fail(KM_ERR_ARG, "km_render: null argument");
  if (cam->width < 1 || cam->height < 1 || !(cam->fovy > 0 && cam->fovy < 180)) return fail(KM_ERR_ARG, "km_render: bad camera");
  if (vis->nlight < 0 || vis->nlight > 4) return fail(KM_ERR_ARG, "km_render: at most 4 directional lights");
  const int nlinks = h->vt.nv - 6;
  if (cam->link >= nlinks || cam->target_link >= nlinks) return fail(KM_ERR_ARG, "km_render: camera link out of range");
  DeviceGuard guard(h->device);
  if (!h->d_recs) KM_CUDA(cudaMalloc((void**)&h->d_recs, (size_t)h->n * h->vt.render_rec_floats * sizeof(float)));
  km::KmRenderParams P;
  std::memset(&P, 0, sizeof(P));
  P.W = cam->width; P.H = cam->height;
  P.tiles_x = (P.W + km::KM_RENDER_TILE - 1) / km::KM_RENDER_TILE; P.tiles_y = (P.H + km::KM_RENDER_TILE - 1) / km::KM_RENDER_TILE;
  P.rec_floats = h->vt.render_rec_floats; P.nlight = vis->nlight;
  P.focal = (float)(0.5 * cam->height / std::tan(0.5 * cam->fovy * 3.14159265358979323846 / 180.0));
  P.tab_z = (float)h->tab_z; P.link_radius = (float)vis->link_radius;
  for (int c = 0; c < 3; c++) {
    P.ambient[c] = (float)vis->head_ambient[c]; P.head_diffuse[c] = (float)vis->head_diffuse[c]; P.head_specular[c] = (float)vis->head_specular[c];
    P.mat[km::KM_MAT_TABLE][c] = (float)vis->rgb_table[c]; P.mat[km::KM_MAT_CUBE][c] = (float)vis->rgb_cube[c];
    P.mat[km::KM_MAT_LINK][c] = (float)vis->rgb_link[c]; P.mat[km::KM_MAT_PAD][c] = (float)vis->rgb_pad[c];
    P.cam_pos[c] = (float)cam->pos[c]; P.tgt_pos[c] = (float)cam->target_pos[c];
  }
  for (int l = 0; l < vis->nlight; l++) {
    const double* d = vis->light_dir[l];
    const double nn = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    if (!(nn > 0)) return fail(KM_ERR_ARG, "km_render: zero light direction");
    for (int c = 0; c < 3; c++) {
      P.ldir[l][c] = (float)(-d[c] / nn); P.ldiffuse[l][c] = (float)vis->light_diffuse[l][c]; P.lspecular[l][c] = (float)vis->light_specular[l][c];
      P.ambient[c] += (float)vis->light_ambient[l][c];
    }
  }
  P.mat_specular = (float)vis->specular;
  int k = 0;
  while ((1 << (k + 1)) <= (int)(vis->shininess * 128.0 + 0.5)) k++;   // exponent rounded down to a power of two (0.5 -> 64)
  P.shin_squarings = k;
  P.spec_cut = (float)std::exp2(-20.0 / (double)(1 << k));
  P.cam_link = cam->link; P.tgt_link = cam->target_link;
  KmArgs a = base_args(h, stream);
  KM_CUDA(h->vt.render_setup(a, P, h->d_recs));
  // grid = (tile column, tile row, env): at most 65535 envs per launch, larger batches in chunks
  for (int e0 = 0; e0 < h->n; e0 += 65535) {
    const int ne = h->n - e0 < 65535 ? h->n - e0 : 65535;
    if (P.shin_squarings == 6) km::k_render_pixels<6><<<dim3(P.tiles_x, P.tiles_y, ne), km::KM_RENDER_THREADS, 0, (cudaStream_t)stream>>>(h->d_recs, rgb_dev, P, e0);
    else km::k_render_pixels<-1><<<dim3(P.tiles_x, P.tiles_y, ne), km::KM_RENDER_THREADS, 0, (cudaStream_t)stream>>>(h->d_recs, rgb_dev, P, e0);
    KM_CUDA(cudaGetLastError());
  }
  h->launches += 2;
  return KM_OK;
}

int km_render_host(km_handle h, const km_camera* cam, const km_visual* vis, unsigned char* rgb_host) {
  if (!h || !cam || !vis || !rgb_host) return fail(KM_ERR_ARG, "km_render_host: null argument");
  if (cam->width < 1 || cam->height < 1 || cam->width > 16384 || cam->height > 16384) return fail(KM_ERR_ARG, "km_render_host: bad camera size");
  DeviceGuard guard(h->device);
  const size_t bytes = (size_t)h->n * cam->width * cam->height * 3;
  if (bytes > h->rgb_bytes) {
    if (h->d_rgb) cudaFree(h->d_rgb);
    h->d_rgb = nullptr; h->rgb_bytes = 0;
    KM_CUDA(cudaMalloc((void**)&h->d_rgb, bytes));
    h->rgb_bytes = bytes;
  }
  int rc = km_render(h, cam, vis, h->d_rgb, (void*)h->host_stream);
  if (rc != KM_OK) return rc;
  KM_CUDA(cudaMemcpyAsync(rgb_host, h->d_rgb, bytes, cudaMemcpyDeviceToHost, h->host_stream));
  KM_CUDA(cudaStreamSynchronize(h->host_stream));
  return KM_OK;
}

int km_render_record_floats(km_handle h) { return h ? h->vt.render_rec_floats : 0; }
int km_get_render_records(km_handle h, float* recs_dev, void* stream) {
  if (!h || !recs_dev) return fail(KM_ERR_ARG, "km_get_render_records: null argument");
  if (!h->d_recs) return fail(KM_ERR_ARG, "km_get_render_records: km_render has not run on this handle");
  DeviceGuard guard(h->device);
  KM_CUDA(cudaMemcpyAsync(recs_dev, h->d_recs, (size_t)h->n * h->vt.render_rec_floats * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return KM_OK;
}

int km_reset_host(km_handle h, const unsigned char* mask, const void* cube_xyz, void* obs) {
  if (!h) return fail(KM_ERR_ARG, "null handle");
  DeviceGuard guard(h->device);
  const size_t n = (size_t)h->n, sb = h->vt.scalar_bytes;
  cudaStream_t hs = h->host_stream;
  if (mask) KM_CUDA(cudaMemcpyAsync(h->d_mask, mask, n, cudaMemcpyHostToDevice, hs));
  if (cube_xyz) KM_CUDA(cudaMemcpyAsync(h->d_xyz, cube_xyz, n * 3 * sb, cudaMemcpyHostToDevice, hs));
  int rc = km_reset(h, mask ? h->d_mask : nullptr, cube_xyz ? h->d_xyz : nullptr, obs ? h->d_obs : nullptr, (void*)hs);
  if (rc != KM_OK) return rc;
  if (obs) KM_CUDA(cudaMemcpyAsync(obs, h->d_obs, n * h->vt.obs_dim * sb, cudaMemcpyDeviceToHost, hs));
  KM_CUDA(cudaStreamSynchronize(hs));
  return KM_OK;
}

int km_step_host(km_handle h, const float* action, void* obs, void* reward, unsigned char* truncated, int autoreset) {
  if (!h || !action) return fail(KM_ERR_ARG, "km_step_host: null handle or action");
  DeviceGuard guard(h->device);
  const size_t n = (size_t)h->n, sb = h->vt.scalar_bytes;
  cudaStream_t hs = h->host_stream;
  int act_dim = km_act_dim(h);
  KM_CUDA(cudaMemcpyAsync(h->d_act, action, n * act_dim * sizeof(float), cudaMemcpyHostToDevice, hs));
  km_step_out o;
  std::memset(&o, 0, sizeof(o));
  o.obs = obs ? h->d_obs : nullptr; o.reward = reward ? h->d_reward : nullptr; o.truncated = truncated ? h->d_trunc : nullptr;
  int rc = km_step(h, h->d_act, &o, autoreset, (void*)hs);
  if (rc != KM_OK) return rc;
  if (obs) KM_CUDA(cudaMemcpyAsync(obs, h->d_obs, n * h->vt.obs_dim * sb, cudaMemcpyDeviceToHost, hs));
  if (reward) KM_CUDA(cudaMemcpyAsync(reward, h->d_reward, n * sb, cudaMemcpyDeviceToHost, hs));
  if (truncated) KM_CUDA(cudaMemcpyAsync(truncated, h->d_trunc, n, cudaMemcpyDeviceToHost, hs));
  KM_CUDA(cudaStreamSynchronize(hs));
  return KM_OK;
}

}  // extern "C"
