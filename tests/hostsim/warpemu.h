// warpemu.h -- deterministic host emulation of ONE 32-lane warp, for running the warp-per-env device code on the CPU.
//
// TEST INFRASTRUCTURE ONLY (see hostsim.cpp).  The 32 lanes are coroutines (ucontext) stepped round-robin by a
// scheduler: a lane runs until its next warp collective (__shfl_sync, __shfl_xor_sync, __ballot_sync, __any_sync,
// __all_sync, __syncwarp), publishes its operand and yields; once every lane has arrived, each one resumes and reads
// the others' operands.  Besides giving the same results as the hardware, the emulator CHECKS what the hardware
// assumes: all 32 lanes must reach the same kind of collective in the same order and return together -- a mismatch
// (which on a GPU is a hang or undefined behaviour) aborts the run with a message instead.
#pragma once
#include <dlfcn.h>
#include <ucontext.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>

namespace km_emu {

constexpr int W = 32;
enum Kind { K_SHFL = 1, K_SHFL_XOR, K_BALLOT, K_ANY, K_ALL, K_SYNC };

struct Warp {
  ucontext_t sched, ctx[W];
  char* stacks[W];
  int cur = -1;
  bool done[W];
  uint64_t buf[2][W];
  int kind[2][W];
  void* site[2][W];   // call site of the collective (for the divergence report)
  long ngen[W];
  long collectives = 0;
  std::string error;
  std::function<void(int)> body;
  static constexpr size_t STACK = 1 << 20;
  Warp() { for (int i = 0; i < W; i++) stacks[i] = (char*)malloc(STACK); }
  ~Warp() { for (int i = 0; i < W; i++) free(stacks[i]); }
};

inline Warp*& current() { static thread_local Warp* w = nullptr; return w; }

inline void trampoline(int lane) {
  Warp* w = current();
  w->body(lane);
  w->done[lane] = true;
  swapcontext(&w->ctx[lane], &w->sched);
}

// run body(lane) for the 32 lanes; returns false (and sets w.error) when the lanes' collectives do not match
inline bool run(Warp& w, std::function<void(int)> body) {
  current() = &w;
  w.body = body;
  w.error.clear();
  for (int i = 0; i < W; i++) {
    w.done[i] = false; w.ngen[i] = 0;
    getcontext(&w.ctx[i]);
    w.ctx[i].uc_stack.ss_sp = w.stacks[i];
    w.ctx[i].uc_stack.ss_size = Warp::STACK;
    w.ctx[i].uc_link = &w.sched;
    makecontext(&w.ctx[i], (void (*)())trampoline, 1, i);
  }
  while (true) {
    int ndone = 0;
    for (int i = 0; i < W; i++) {
      if (w.done[i]) { ndone++; continue; }
      w.cur = i;
      swapcontext(&w.sched, &w.ctx[i]);
      if (w.done[i]) ndone++;
    }
    if (!w.error.empty()) return false;
    if (ndone == W) return true;
    if (ndone != 0) { w.error = "lanes left the warp-synchronous code at different times (divergent exit)"; return false; }
    // every live lane is now parked at its ngen-th collective: they must agree
    for (int i = 1; i < W; i++)
      if (w.ngen[i] != w.ngen[0] || w.kind[(w.ngen[i] - 1) & 1][i] != w.kind[(w.ngen[0] - 1) & 1][0]) {
        char msg[320];
        Dl_info d0, d1;
        void *s0 = w.site[(w.ngen[0] - 1) & 1][0], *s1 = w.site[(w.ngen[i] - 1) & 1][i];
        dladdr(s0, &d0); dladdr(s1, &d1);
        snprintf(msg, sizeof msg, "warp collectives diverge: lane 0 at #%ld kind %d (addr2line -e <so> 0x%lx), lane %d at #%ld kind %d (0x%lx)",
                 w.ngen[0], w.kind[(w.ngen[0] - 1) & 1][0], (unsigned long)((char*)s0 - (char*)d0.dli_fbase), i, w.ngen[i],
                 w.kind[(w.ngen[i] - 1) & 1][i], (unsigned long)((char*)s1 - (char*)d1.dli_fbase));
        w.error = msg;
        return false;
      }
    w.collectives++;
  }
}

// publish `payload` for a collective of `kind`, wait for the other lanes, return the exchange buffer of this collective
inline const uint64_t* exchange(uint64_t payload, int kind, void* site = nullptr) {
  Warp* w = current();
  const int l = w->cur;
  const int b = (int)(w->ngen[l] & 1);
  w->buf[b][l] = payload;
  w->kind[b][l] = kind;
  w->site[b][l] = site;
  w->ngen[l]++;
  swapcontext(&w->ctx[l], &w->sched);
  w->cur = l;
  return w->buf[b];
}
inline int lane_id() { return current()->cur; }

template <typename T> inline uint64_t bits(T v) { uint64_t u = 0; static_assert(sizeof(T) <= 8, "payload"); memcpy(&u, &v, sizeof(T)); return u; }
template <typename T> inline T unbits(uint64_t u) { T v; memcpy(&v, &u, sizeof(T)); return v; }

}  // namespace km_emu

// the CUDA warp intrinsics the device code uses (full-warp masks only)
#define KM_EMU_SITE __builtin_extract_return_addr(__builtin_return_address(0))
#define KM_EMU_NI __attribute__((noinline))
template <typename T> KM_EMU_NI T __shfl_sync(unsigned, T v, int src, int = 32) { return km_emu::unbits<T>(km_emu::exchange(km_emu::bits(v), km_emu::K_SHFL, KM_EMU_SITE)[src & 31]); }
template <typename T> KM_EMU_NI T __shfl_xor_sync(unsigned, T v, int o, int = 32) { const int l = km_emu::lane_id(); return km_emu::unbits<T>(km_emu::exchange(km_emu::bits(v), km_emu::K_SHFL_XOR, KM_EMU_SITE)[(l ^ o) & 31]); }
KM_EMU_NI inline unsigned __ballot_sync(unsigned, bool p) {
  const uint64_t* b = km_emu::exchange(p ? 1u : 0u, km_emu::K_BALLOT, KM_EMU_SITE);
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r |= (unsigned)(b[i] & 1u) << i;
  return r;
}
KM_EMU_NI inline bool __any_sync(unsigned, bool p) {
  const uint64_t* b = km_emu::exchange(p ? 1u : 0u, km_emu::K_ANY, KM_EMU_SITE);
  bool r = false;
  for (int i = 0; i < 32; i++) r = r || (b[i] & 1u);
  return r;
}
KM_EMU_NI inline bool __all_sync(unsigned, bool p) {
  const uint64_t* b = km_emu::exchange(p ? 1u : 0u, km_emu::K_ALL, KM_EMU_SITE);
  bool r = true;
  for (int i = 0; i < 32; i++) r = r && (b[i] & 1u);
  return r;
}
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
KM_EMU_NI inline void __syncwarp(unsigned = 0xffffffffu) { km_emu::exchange(0, km_emu::K_SYNC, KM_EMU_SITE); }
