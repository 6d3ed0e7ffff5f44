// TEST-ONLY: run warp-synchronous CUDA-style code on the CPU, one fiber per thread of a block.
//
// The streaming kernels (baseboostdepth_b200/csrc/bbd_stream.cuh) exchange values between the lanes
// of a warp with shuffles.  To step the very same source on the CPU, every thread of a block runs as
// a ucontext fiber; a warp collective (shuffle, ballot, warp barrier) parks the calling lane until all
// 32 lanes of its warp have arrived, a block barrier until all threads have.  Lanes are resumed
// round-robin, so any schedule-independent (i.e. correct) CUDA code gives the same result as on the
// device.  Never part of the product: only tests/emu/bbd_emu.cpp includes it.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>
#include <vector>

namespace simt {

struct WarpState {
  uint32_t slot[32];
  uint32_t snap[32];
  int arrived = 0;
  unsigned gen = 0;
};

struct Block {
  int nthreads = 0;
  std::vector<ucontext_t> ctx;
  std::vector<std::vector<char>> stacks;
  std::vector<char> done;
  std::vector<WarpState> warps;
  ucontext_t main_ctx;
  int cur = -1;
  int bar_arrived = 0;
  unsigned bar_gen = 0;
  std::function<void(int)> body;
};

inline Block*& current() {
  static thread_local Block* b = nullptr;
  return b;
}

inline int tid() { return current()->cur; }

inline void yield() {
  Block* b = current();
  swapcontext(&b->ctx[b->cur], &b->main_ctx);
}

inline void trampoline() {
  Block* b = current();
  b->body(b->cur);
  b->done[b->cur] = 1;
  // falls back to uc_link = main_ctx
}

// Run `body(tid)` for tid = 0..nthreads-1 as cooperating fibers until all have returned.
inline void run_block(int nthreads, const std::function<void(int)>& body, size_t stack_bytes = 256 * 1024) {
  static thread_local Block blk;
  Block* b = &blk;
  current() = b;
  b->nthreads = nthreads;
  b->body = body;
  b->ctx.resize(nthreads);
  if ((int)b->stacks.size() < nthreads) b->stacks.resize(nthreads);
  b->done.assign(nthreads, 0);
  b->warps.assign((nthreads + 31) / 32, WarpState());
  b->bar_arrived = 0;
  for (int t = 0; t < nthreads; ++t) {
    if (b->stacks[t].size() < stack_bytes) b->stacks[t].resize(stack_bytes);
    getcontext(&b->ctx[t]);
    b->ctx[t].uc_stack.ss_sp = b->stacks[t].data();
    b->ctx[t].uc_stack.ss_size = b->stacks[t].size();
    b->ctx[t].uc_link = &b->main_ctx;
    makecontext(&b->ctx[t], (void (*)())trampoline, 0);
  }
  for (;;) {
    bool any = false;
    for (int t = 0; t < nthreads; ++t) {
      if (b->done[t]) continue;
      any = true;
      b->cur = t;
      swapcontext(&b->main_ctx, &b->ctx[t]);
    }
    if (!any) break;
  }
  b->cur = -1;
}

// every lane deposits v; returns once all 32 lanes of the warp have; snap[] then holds the values
inline const uint32_t* warp_exchange(uint32_t v) {
  Block* b = current();
  WarpState& w = b->warps[b->cur >> 5];
  const int lane = b->cur & 31;
  const int lanes = (b->nthreads - (b->cur & ~31)) < 32 ? (b->nthreads - (b->cur & ~31)) : 32;
  w.slot[lane] = v;
  const unsigned g = w.gen;
  if (++w.arrived == lanes) {
    w.arrived = 0;
    memcpy(w.snap, w.slot, sizeof(w.snap));
    ++w.gen;
  } else {
    while (w.gen == g) yield();
  }
  return w.snap;  // valid until the next collective of this warp completes, which needs this lane too
}

inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

inline float shfl_up(float v, int d) {
  const int lane = tid() & 31;
  const uint32_t* s = warp_exchange(f2u(v));
  return lane >= d ? u2f(s[lane - d]) : v;
}
inline float shfl_down(float v, int d) {
  const int lane = tid() & 31;
  const uint32_t* s = warp_exchange(f2u(v));
  return lane + d < 32 ? u2f(s[lane + d]) : v;
}
inline float shfl_xor(float v, int m) {
  const int lane = tid() & 31;
  const uint32_t* s = warp_exchange(f2u(v));
  return u2f(s[lane ^ m]);
}
inline float shfl_idx(float v, int src) {
  const uint32_t* s = warp_exchange(f2u(v));
  return u2f(s[src & 31]);
}
inline int shfl_up_i(int v, int d) {
  const int lane = tid() & 31;
  const uint32_t* s = warp_exchange((uint32_t)v);
  return lane >= d ? (int)s[lane - d] : v;
}
inline int shfl_down_i(int v, int d) {
  const int lane = tid() & 31;
  const uint32_t* s = warp_exchange((uint32_t)v);
  return lane + d < 32 ? (int)s[lane + d] : v;
}
inline unsigned ballot(bool p) {
  const uint32_t* s = warp_exchange(p ? 1u : 0u);
  unsigned m = 0;
  for (int i = 0; i < 32; ++i) m |= (s[i] & 1u) << i;
  return m;
}
inline void syncwarp() { warp_exchange(0); }

inline void syncthreads() {
  Block* b = current();
  const unsigned g = b->bar_gen;
  if (++b->bar_arrived == b->nthreads) {
    b->bar_arrived = 0;
    ++b->bar_gen;
  } else {
    while (b->bar_gen == g) yield();
  }
}

}  // namespace simt
