// Host stand-in for <cuda_runtime.h>: just enough to compile the index logic of a simple __global__ function with g++ and run it
// thread by thread (tests/test_host_kernel_emulation.py).  Test infrastructure only; never part of the product build.
#pragma once
#include <cstdint>
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __restrict__
struct emu_dim3 { unsigned x, y, z; };
static thread_local emu_dim3 blockIdx, threadIdx, blockDim, gridDim;
struct alignas(16) float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
typedef void* cudaStream_t;
