// Micro-benchmark: how fast can an SM push shared-memory tiles to HBM with cp.async.bulk (UBLKCP S2G) versus
// plain 128-bit stores?  Used to size the staging tiles of susnet_tile.cuh.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_store_bench tma_store_bench.cu && ./tma_store_bench
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// each warp owns `depth` tiles of `tile_bytes`; loops: issue a bulk store per tile, commit, wait until <= depth-1 pending reads
template <int DEPTH>
__global__ void k_bulk(uint8_t* out, size_t total_bytes, int tile_bytes, int warps) {
  extern __shared__ __align__(128) uint8_t sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t* mine = sm + (size_t)warp * DEPTH * tile_bytes;
  for (int i = lane * 16; i < DEPTH * tile_bytes; i += 512) *reinterpret_cast<uint4*>(mine + i) = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const size_t n_tiles = total_bytes / tile_bytes;
  const size_t stride = (size_t)gridDim.x * warps;
  int slot = 0;
  if (lane == 0) {
    for (size_t t = (size_t)blockIdx.x * warps + warp; t < n_tiles; t += stride) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + t * tile_bytes),
                   "r"(smem_u32(mine + (size_t)slot * tile_bytes)), "r"(tile_bytes)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(DEPTH - 1) : "memory");
      slot = (slot + 1) % DEPTH;
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

__global__ void k_stg(uint4* out, size_t n16) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) out[i] = z;
}

// The fused kernel's output pattern without any compute: per group of 32 envs, 4 spatial tiles of 18 144 B, 5
// non-spatial rows blocks of 1 920 B, one rewards block of 640 B (all bulk), plus 40 B/env of plain state stores and
// 48 B/env of state loads.  mode bit0: dense blocks, bit1: state stores, bit2: state loads.
__global__ void k_pattern(uint8_t* sp, uint8_t* ns, uint8_t* rew, uint4* st_a, uint4* st_b, uint64_t* st_c, size_t n_envs,
                          int warps, int mode, unsigned long long* sink) {
  extern __shared__ __align__(128) uint8_t sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per_warp = 18176 + 9600 + 640;
  uint8_t* mine = sm + (size_t)warp * per_warp;
  for (int i = lane * 16; i < per_warp; i += 512) *reinterpret_cast<uint4*>(mine + i) = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const size_t n_groups = n_envs / 32, stride = (size_t)gridDim.x * warps;
  unsigned long long acc = 0;
  for (size_t g = (size_t)blockIdx.x * warps + warp; g < n_groups; g += stride) {
    const size_t e = g * 32 + lane;
    if (mode & 4) { uint4 a = st_a[e], b = st_b[e]; acc += a.x + b.y + st_c[e]; }
    if (mode & 2) { st_a[e] = make_uint4(1, 2, 3, 4); st_b[e] = make_uint4(5, 6, 7, 8); st_c[e] = 9; }
    if (lane == 0) {
      if (mode & 1) {
        for (int k = 0; k < 5; ++k)
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(ns + ((size_t)k * n_envs + g * 32) * 60),
                       "r"(smem_u32(mine + 18176 + k * 1920)), "r"(1920) : "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(rew + g * 640),
                     "r"(smem_u32(mine + 18176 + 9600)), "r"(640) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      for (int j = 0; j < 4; ++j) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(sp + (g * 32 + j * 8) * 2268),
                     "r"(smem_u32(mine)), "r"(18144) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    __syncwarp();
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  if (acc == 0x123456789ull) *sink = acc;
}

// copy-out of a shared-memory tile with ordinary 128-bit loads/stores by `warps` warps per CTA (one CTA per SM)
__global__ void k_lsu_tile(uint4* out, size_t n16_total, int tile16, int warps) {
  extern __shared__ __align__(128) uint8_t sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint4* mine = reinterpret_cast<uint4*>(sm) + (size_t)warp * tile16;
  for (int i = lane; i < tile16; i += 32) mine[i] = make_uint4(0, 0, 0, 0);
  __syncwarp();
  const size_t n_tiles = n16_total / tile16, stride = (size_t)gridDim.x * warps;
  for (size_t t = (size_t)blockIdx.x * warps + warp; t < n_tiles; t += stride) {
    uint4* dst = out + t * tile16;
#pragma unroll 8
    for (int i = lane; i < tile16; i += 32) dst[i] = mine[i];
  }
}

template <typename F>
float time_ms(F f, int reps = 5) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f();
  cudaDeviceSynchronize();
  float best = 1e9f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  const size_t total = (size_t)2688 << 20;  // ~2.8 GB, about what one fused step writes for 1M envs
  uint8_t* out;
  cudaMalloc(&out, total);
  int sms;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("SMs %d, bytes %zu\n", sms, total);
  float ms = time_ms([&] { cudaMemsetAsync(out, 0, total); });
  printf("cudaMemset            %8.3f ms %8.1f GB/s\n", ms, total / ms / 1e6);
  for (int mult : {2, 4, 8, 16}) {
    ms = time_ms([&] { k_stg<<<sms * mult, 512>>>((uint4*)out, total / 16); });
    printf("STG.128 grid %2dx SMs   %8.3f ms %8.1f GB/s\n", mult, ms, total / ms / 1e6);
  }
  cudaFuncSetAttribute(k_bulk<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(k_bulk<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(k_bulk<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  for (int tile : {18144}) {
    for (int warps : {1, 2, 4, 8}) {
      for (int depth : {1, 2, 4}) {
        const size_t smem = (size_t)warps * depth * tile;
        if (smem > 220 * 1024) continue;
        const size_t usable = total / tile * tile;
        auto launch = [&] {
          if (depth == 1) k_bulk<1><<<sms, warps * 32, smem>>>(out, usable, tile, warps);
          else if (depth == 2) k_bulk<2><<<sms, warps * 32, smem>>>(out, usable, tile, warps);
          else k_bulk<4><<<sms, warps * 32, smem>>>(out, usable, tile, warps);
        };
        ms = time_ms(launch, 3);
        cudaError_t e = cudaGetLastError();
        printf("bulk tile %6d warps %d depth %d (%3zu KB in flight/SM) %8.3f ms %8.1f GB/s %s\n", tile, warps, depth,
               smem / 1024, ms, usable / ms / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
    }
  }
  // two CTAs per SM
  for (int tile : {9072, 18144}) {
    const int warps = 4, depth = 1;
    const size_t smem = (size_t)warps * depth * tile;
    const size_t usable = total / tile * tile;
    ms = time_ms([&] { k_bulk<1><<<sms * 2, warps * 32, smem>>>(out, usable, tile, warps); }, 3);
    printf("bulk tile %6d warps %d depth %d, 2 CTAs/SM %8.3f ms %8.1f GB/s\n", tile, warps, depth, ms, usable / ms / 1e6);
  }
  cudaFuncSetAttribute(k_lsu_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  for (int warps : {1, 2, 4, 8}) {
    const int tile16 = 18144 / 16;
    const size_t n16 = total / 16 / tile16 * tile16;
    ms = time_ms([&] { k_lsu_tile<<<sms, warps * 32, (size_t)warps * 18144>>>((uint4*)out, n16, tile16, warps); }, 3);
    printf("LDS.128+STG.128 tile copy-out, %d warps/SM: %8.3f ms %8.1f GB/s\n", warps, ms, n16 * 16.0 / ms / 1e6);
  }
  {
    const size_t n_envs = 1 << 20;
    uint8_t *sp, *ns, *rew; uint4 *sa, *sb; uint64_t* sc; unsigned long long* sink;
    cudaMalloc(&sp, n_envs * 2268); cudaMalloc(&ns, n_envs * 300); cudaMalloc(&rew, n_envs * 20);
    cudaMalloc(&sa, n_envs * 16); cudaMalloc(&sb, n_envs * 16); cudaMalloc(&sc, n_envs * 8); cudaMalloc(&sink, 8);
    cudaMemset(sa, 0, n_envs * 16); cudaMemset(sb, 0, n_envs * 16); cudaMemset(sc, 0, n_envs * 8);
    cudaFuncSetAttribute(k_pattern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    const int warps = 8;
    const size_t smem = (size_t)warps * (18176 + 9600 + 640);
    for (int mode : {0, 1, 3, 7}) {
      ms = time_ms([&] { k_pattern<<<sms, warps * 32, smem>>>(sp, ns, rew, sa, sb, sc, n_envs, warps, mode, sink); }, 5);
      const double bytes = (double)n_envs * (2268 + ((mode & 1) ? 320 : 0) + ((mode & 2) ? 40 : 0) + ((mode & 4) ? 40 : 0));
      printf("pattern mode %d: %8.3f ms %8.1f GB/s (%s)\n", mode, ms, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
  }
  cudaFree(out);
  return 0;
}
