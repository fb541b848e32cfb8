// Host stand-in for <cuda_runtime.h>: just enough to compile the device code of this library with g++.  Default mode: one
// emulated thread at a time (tests/test_host_kernel_emulation.py) -- warp collectives have the semantics of a warp whose other
// lanes contribute nothing (any = own predicate, reduce = own value, shuffles from other lanes = 0), atomics are plain updates:
// valid for code in which lanes only meet in order-independent accumulations.  With EMU_SIMT the threads of a block are fibers
// and barriers / collectives are real (simt.h, tests/test_host_simt_emulation.py).  Test infrastructure only.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __restrict__
#define __shared__ static
#define __align__(n) alignas(n)
struct emu_dim3 { unsigned x, y, z; };
static thread_local emu_dim3 blockIdx, threadIdx, blockDim, gridDim;
struct alignas(16) float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct alignas(16) uint4 { unsigned x, y, z, w; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
typedef void* cudaStream_t;

static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
// position of the offset-th set bit of mask at or above bit `base` (offset >= 1), 0xffffffff if there is none
static inline unsigned __fns(unsigned mask, unsigned base, int offset) {
  for (unsigned b = base; b < 32; ++b)
    if ((mask >> b) & 1u) { if (--offset == 0) return b; }
  return 0xffffffffu;
}
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
  const unsigned long long v = ((unsigned long long)y << 32) | x;
  unsigned r = 0;
  for (int i = 0; i < 4; ++i) r |= (unsigned)((v >> (8 * ((s >> (4 * i)) & 7u))) & 0xffu) << (8 * i);
  return r;
}
static inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline void __threadfence() {}
#ifdef EMU_SIMT  // whole kernels: threads are fibers, barriers and warp collectives switch between them (simt.h)
#include "simt.h"
#else            // one thread at a time: barriers are no-ops, a warp is its calling lane
static inline void __syncthreads() {}
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline int __any_sync(unsigned, int p) { return p; }
static inline unsigned __reduce_add_sync(unsigned, unsigned v) { return v; }
template <typename T> static inline T __shfl_xor_sync(unsigned, T, int) { return T(0); }
#endif
template <typename T> static inline T atomicAdd(T* p, T v) { const T old = *p; *p = old + v; return old; }
static inline unsigned atomicAdd(volatile unsigned* p, unsigned v) { const unsigned old = *p; *p = old + v; return old; }
