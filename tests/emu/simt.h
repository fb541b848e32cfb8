// Cooperative SIMT emulation for whole kernels (tests/test_host_simt_emulation.py): every CUDA thread of a block is a fiber
// (ucontext); __syncthreads / __syncwarp and the warp collectives switch fibers, so warp-synchronous code -- rows staged in
// shared memory by one phase and read by all lanes in the next -- runs with its real data flow.  One block at a time, blocks
// in order; atomics are plain updates (fibers never preempt each other).  Test infrastructure only.
#pragma once
#include <ucontext.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

namespace simt {

enum State { READY = 0, AT_WARP = 1, AT_BLOCK = 2, DONE = 3 };

struct Fiber {
  ucontext_t ctx;
  std::vector<char> stack;
  unsigned tid = 0;
  int state = READY;
};

inline std::vector<Fiber>& fibers() { static std::vector<Fiber> f; return f; }
inline ucontext_t& sched_ctx() { static ucontext_t c; return c; }
inline Fiber*& current() { static Fiber* c = nullptr; return c; }
inline std::function<void()>& body() { static std::function<void()> b; return b; }
inline unsigned long long (*exchange())[32] { static unsigned long long x[64][32]; return x; }

inline void yield(int state) {
  Fiber* f = current();
  f->state = state;
  swapcontext(&f->ctx, &sched_ctx());
}

inline void trampoline() {
  body()();
  current()->state = DONE;
  swapcontext(&current()->ctx, &sched_ctx());
}

// run `kernel` for every thread of every block of the grid
inline void launch(unsigned grid, unsigned block, const std::function<void()>& kernel) {
  body() = kernel;
  gridDim = {grid, 1, 1};
  blockDim = {block, 1, 1};
  auto& fs = fibers();
  if (fs.size() < block) fs.resize(block);
  for (unsigned b = 0; b < grid; ++b) {
    blockIdx = {b, 0, 0};
    for (unsigned t = 0; t < block; ++t) {
      Fiber& f = fs[t];
      if (f.stack.empty()) f.stack.resize(256 * 1024);
      getcontext(&f.ctx);
      f.ctx.uc_stack.ss_sp = f.stack.data();
      f.ctx.uc_stack.ss_size = f.stack.size();
      f.ctx.uc_link = nullptr;
      makecontext(&f.ctx, trampoline, 0);
      f.tid = t;
      f.state = READY;
    }
    for (;;) {
      bool ran = false, all_done = true;
      for (unsigned t = 0; t < block; ++t) {
        Fiber& f = fs[t];
        if (f.state == READY) {
          current() = &f;
          threadIdx = {t, 0, 0};
          swapcontext(&sched_ctx(), &f.ctx);
          ran = true;
        }
        all_done &= f.state == DONE;
      }
      if (all_done) break;
      bool released = false;
      for (unsigned w = 0; w * 32 < block; ++w) {  // a warp barrier opens when every live lane of the warp waits at it
        bool any = false, open = true;
        for (unsigned t = w * 32; t < block && t < (w + 1) * 32; ++t) {
          if (fs[t].state == AT_WARP) any = true;
          else if (fs[t].state != DONE) open = false;
        }
        if (any && open) {
          for (unsigned t = w * 32; t < block && t < (w + 1) * 32; ++t)
            if (fs[t].state == AT_WARP) fs[t].state = READY;
          released = true;
        }
      }
      bool any = false, open = true;  // the block barrier opens when every live thread waits at it
      for (unsigned t = 0; t < block; ++t) {
        if (fs[t].state == AT_BLOCK) any = true;
        else if (fs[t].state != DONE) open = false;
      }
      if (any && open) {
        for (unsigned t = 0; t < block; ++t)
          if (fs[t].state == AT_BLOCK) fs[t].state = READY;
        released = true;
      }
      if (!ran && !released) {
        std::fprintf(stderr, "simt: deadlock in block %u (threads wait at different barriers)\n", b);
        std::abort();
      }
    }
  }
}

// value exchange of a warp collective: every live lane deposits, all meet, every lane reads what it needs, all meet again
template <typename T, typename Reduce>
inline T collective(T mine, Reduce reduce) {
  const unsigned tid = current()->tid, lane = tid & 31u, warp = tid >> 5;
  unsigned long long slot = 0;
  std::memcpy(&slot, &mine, sizeof(T));
  exchange()[warp][lane] = slot;
  yield(AT_WARP);
  T vals[32];
  bool live[32];
  for (unsigned l = 0; l < 32; ++l) {
    const unsigned t = warp * 32 + l;
    live[l] = t < blockDim.x && fibers()[t].state != DONE;
    std::memcpy(&vals[l], &exchange()[warp][l], sizeof(T));
  }
  const T r = reduce(vals, live, lane);
  yield(AT_WARP);
  return r;
}

}  // namespace simt

static inline void __syncthreads() { simt::yield(simt::AT_BLOCK); }
static inline void __syncwarp(unsigned = 0xffffffffu) { simt::yield(simt::AT_WARP); }
static inline int __any_sync(unsigned, int p) {
  return simt::collective<int>(p, [](const int* v, const bool* live, unsigned) { int r = 0; for (int l = 0; l < 32; ++l) if (live[l] && v[l]) r = 1; return r; });
}
static inline unsigned __reduce_add_sync(unsigned, unsigned x) {
  return simt::collective<unsigned>(x, [](const unsigned* v, const bool* live, unsigned) { unsigned r = 0; for (int l = 0; l < 32; ++l) if (live[l]) r += v[l]; return r; });
}
template <typename T>
static inline T __shfl_xor_sync(unsigned, T x, int mask) {
  return simt::collective<T>(x, [mask](const T* v, const bool* live, unsigned lane) { const unsigned src = lane ^ (unsigned)mask; return live[src] ? v[src] : v[lane]; });
}
