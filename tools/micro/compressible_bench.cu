// Micro-benchmark: does L2 "generic compression" (cuMemCreate + CU_MEM_ALLOCATION_COMP_GENERIC) help a write stream of
// almost-all-zero feature tiles?  Same kernels on a cudaMalloc buffer and on a compressible VMM allocation.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o compressible_bench compressible_bench.cu -lcuda && ./compressible_bench
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char* s_; cuGetErrorString(r_, &s_); printf("%s failed: %s\n", #x, s_); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// every warp owns two tiles of `tile_bytes` (zero except one float per `row_bytes`), alternates bulk stores of them
__global__ void k_bulk(uint8_t* out, size_t total_bytes, int tile_bytes, int row_bytes, int warps, int ones) {
  extern __shared__ __align__(128) uint8_t sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t* mine = sm + (size_t)warp * 2 * tile_bytes;
  for (int i = lane * 16; i < 2 * tile_bytes; i += 512) *reinterpret_cast<uint4*>(mine + i) = make_uint4(0, 0, 0, 0);
  __syncwarp();
  if (ones)
    for (int r = lane; r < 2 * tile_bytes / row_bytes; r += 32)
      for (int k = 0; k < ones; ++k) *reinterpret_cast<float*>(mine + r * row_bytes + ((r * 37 + k * 211) % (row_bytes / 4)) * 4) = 1.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const size_t n_tiles = total_bytes / tile_bytes, stride = (size_t)gridDim.x * warps;
  int slot = 0;
  if (lane == 0) {
    for (size_t t = (size_t)blockIdx.x * warps + warp; t < n_tiles; t += stride) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + t * tile_bytes),
                   "r"(smem_u32(mine + (size_t)slot * tile_bytes)), "r"(tile_bytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      slot ^= 1;
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

__global__ void k_stg(uint4* out, size_t n16) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) out[i] = z;
}

__global__ void k_read(const uint4* in, size_t n16, unsigned long long* sink) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  unsigned long long acc = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) { uint4 v = in[i]; acc += v.x + v.y + v.z + v.w; }
  if (acc == 0x123456789ull) *sink = acc;
}

template <typename F>
float time_ms(F f, int reps = 5) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int i = 0; i < reps; ++i) {
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaFree(0);
  CUdevice dev; CK(cuDeviceGet(&dev, 0));
  int comp = 0; CK(cuDeviceGetAttribute(&comp, CU_DEVICE_ATTRIBUTE_GENERIC_COMPRESSION_SUPPORTED, dev));
  printf("GENERIC_COMPRESSION_SUPPORTED = %d\n", comp);
  const size_t want = (size_t)2688 << 20;
  uint8_t* plain; cudaMalloc(&plain, want);
  uint8_t* cbuf = nullptr;
  if (comp) {
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = 0;
    prop.allocFlags.compressionType = CU_MEM_ALLOCATION_COMP_GENERIC;
    size_t gran = 0; CK(cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    const size_t size = (want + gran - 1) / gran * gran;
    CUmemGenericAllocationHandle h; CK(cuMemCreate(&h, size, &prop, 0));
    CUmemAllocationProp got = {}; CK(cuMemGetAllocationPropertiesFromHandle(&got, h));
    printf("granularity %zu, allocation compressionType = %d (1 = generic)\n", gran, (int)got.allocFlags.compressionType);
    CUdeviceptr p; CK(cuMemAddressReserve(&p, size, 0, 0, 0));
    CK(cuMemMap(p, size, 0, h, 0));
    CUmemAccessDesc acc = {}; acc.location = prop.location; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    CK(cuMemSetAccess(p, size, &acc, 1));
    cbuf = reinterpret_cast<uint8_t*>(p);
  }
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned long long* sink; cudaMalloc(&sink, 8);
  cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  const int tile = 16 * 2268, row = 2268, warps = 3;  // 36 KB tiles like the emitter's, 2 per warp
  const size_t usable = want / tile * tile;
  for (int which = 0; which < 2; ++which) {
    uint8_t* buf = which ? cbuf : plain;
    if (!buf) break;
    printf("---- %s\n", which ? "compressible (cuMemCreate, COMP_GENERIC)" : "cudaMalloc");
    float ms = time_ms([&] { cudaMemsetAsync(buf, 0, want); });
    printf("cudaMemset 0                    %8.3f ms %8.1f GB/s\n", ms, want / ms / 1e6);
    ms = time_ms([&] { k_stg<<<sms * 8, 512>>>((uint4*)buf, want / 16); });
    printf("STG.128 zeros                   %8.3f ms %8.1f GB/s\n", ms, want / ms / 1e6);
    for (int ones : {0, 1, 10}) {
      ms = time_ms([&] { k_bulk<<<sms, warps * 32, (size_t)warps * 2 * tile>>>(buf, usable, tile, row, warps, ones); });
      printf("bulk tiles, %2d ones / 2268 B row %8.3f ms %8.1f GB/s\n", ones, ms, usable / ms / 1e6);
      float rd = time_ms([&] { k_read<<<sms * 8, 512>>>((const uint4*)buf, want / 16, sink); });
      printf("   read back                    %8.3f ms %8.1f GB/s\n", rd, want / rd / 1e6);
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
