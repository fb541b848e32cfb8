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
  for (int tile : {2048, 4096, 9072, 18144, 36288}) {
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
  cudaFree(out);
  return 0;
}
