// Micro-benchmark: the SM -> L2 store ceiling of a B200 measured INDEPENDENTLY of the fused kernel's own store pattern
// (round-1 verdict: the 7.56 TB/s "ceiling" came from a benchmark shaped like the kernel -- one lane issuing, two tiles in
// flight per warp -- so it was the ceiling of that pattern, not of the fabric).  Every variant writes the same 2.6 GB of
// feature-like data (zeros with ~10 ones per 2268-byte row) into (a) cudaMalloc memory and (b) an L2-compressible allocation:
//   stg128       plain STG.128 grid-stride stores, occupancy swept up to 64 warps / SM
//   stg128_ones  the same with the ones computed arithmetically (the values the features would hold)
//   bulk         cp.async.bulk shared -> global of persistently staged tiles; swept: warps / SM issuing, tile size, tiles in
//                flight per warp (1, 2, 4)
//   bulk_lanes   every LANE of a warp issues its own bulk store of 1/32 of the tile (multi-lane issue)
//   memset       cudaMemsetAsync, for reference
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_ceiling_bench store_ceiling_bench.cu -lcuda && ./store_ceiling_bench
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char* s_; cuGetErrorString(r_, &s_); printf("%s failed: %s\n", #x, s_); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k_stg(uint4* out, size_t n16, int ones) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (ones && (i * 2654435761ull >> 7) % 57 == 0) v.y = 0x3f800000u;  // ~10 ones per 567 floats
    out[i] = v;
  }
}

// `warps` warps per CTA each own `depth` tiles of `tile_bytes`; lane 0 (or every lane, LANES) keeps `depth` bulk stores in flight
template <bool LANES>
__global__ void k_bulk(uint8_t* out, size_t total_bytes, int tile_bytes, int row_bytes, int warps, int depth, int ones) {
  extern __shared__ __align__(128) uint8_t sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t* mine = sm + (size_t)warp * depth * tile_bytes;
  for (int i = lane * 16; i < depth * tile_bytes; i += 512) *reinterpret_cast<uint4*>(mine + i) = make_uint4(0, 0, 0, 0);
  __syncwarp();
  if (ones)
    for (int r = lane; r < depth * tile_bytes / row_bytes; r += 32)
      for (int k = 0; k < ones; ++k) *reinterpret_cast<float*>(mine + r * row_bytes + ((r * 37 + k * 211) % (row_bytes / 4)) * 4) = 1.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const size_t n_tiles = total_bytes / tile_bytes, stride = (size_t)gridDim.x * warps;
  int slot = 0;
  const int chunk = tile_bytes / 32;
  if (LANES || lane == 0) {
    for (size_t t = (size_t)blockIdx.x * warps + warp; t < n_tiles; t += stride) {
      if (LANES)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + t * tile_bytes + (size_t)lane * chunk),
                     "r"(smem_u32(mine + (size_t)slot * tile_bytes + (size_t)lane * chunk)), "r"(chunk) : "memory");
      else
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + t * tile_bytes),
                     "r"(smem_u32(mine + (size_t)slot * tile_bytes)), "r"(tile_bytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (depth == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      else if (depth == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      else asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
      slot = slot + 1 == depth ? 0 : slot + 1;
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

template <typename F>
float time_ms(F f, int reps = 7) {
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
  const size_t want = (size_t)2592 << 20;  // a multiple of 36 KiB and 18 KiB and 9 KiB
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
    CUdeviceptr p; CK(cuMemAddressReserve(&p, size, 0, 0, 0));
    CK(cuMemMap(p, size, 0, h, 0));
    CUmemAccessDesc acc = {}; acc.location = prop.location; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    CK(cuMemSetAccess(p, size, &acc, 1));
    cbuf = reinterpret_cast<uint8_t*>(p);
  }
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaFuncSetAttribute(k_bulk<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(k_bulk<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  printf("{\"sms\": %d, \"bytes\": %zu, \"generic_compression\": %d, \"results\": [\n", sms, want, comp);
  bool first = true;
  auto emit = [&](const char* mem, const char* what, const char* cfg, float ms, size_t bytes) {
    printf("%s {\"memory\": \"%s\", \"variant\": \"%s\", \"config\": \"%s\", \"ms\": %.4f, \"gbs\": %.1f}", first ? "" : ",\n", mem, what, cfg, ms,
           bytes / ms / 1e6);
    first = false;
  };
  char cfg[160];
  for (int which = 0; which < 2; ++which) {
    uint8_t* buf = which ? cbuf : plain;
    if (!buf) break;
    const char* mem = which ? "compressible" : "cudaMalloc";
    float ms = time_ms([&] { cudaMemsetAsync(buf, 0, want); });
    emit(mem, "memset", "cudaMemsetAsync 0", ms, want);
    for (int ones = 0; ones < 2; ++ones)
      for (int threads : {256, 512, 1024})
        for (int ctas_per_sm : {1, 2, 4, 8}) {
          if (threads * ctas_per_sm > 2048) continue;
          ms = time_ms([&] { k_stg<<<sms * ctas_per_sm, threads>>>((uint4*)buf, want / 16, ones); });
          snprintf(cfg, sizeof(cfg), "%d CTAs/SM x %d threads = %d warps/SM", ctas_per_sm, threads, ctas_per_sm * threads / 32);
          emit(mem, ones ? "stg128_ones" : "stg128", cfg, ms, want);
        }
    const int row = 2268;
    for (int tile_rows : {4, 8, 16})
      for (int depth : {1, 2, 4})
        for (int warps : {1, 2, 3, 4, 6, 8, 12}) {
          const int tile = tile_rows * row;
          if ((size_t)warps * depth * tile > 220 * 1024) continue;
          const size_t usable = want / tile * tile;
          ms = time_ms([&] { k_bulk<false><<<sms, warps * 32, (size_t)warps * depth * tile>>>(buf, usable, tile, row, warps, depth, 10); });
          snprintf(cfg, sizeof(cfg), "%d-row tiles (%d B), %d in flight per warp, %d issuing warps/SM", tile_rows, tile, depth, warps);
          emit(mem, "bulk", cfg, ms, usable);
        }
    for (int depth : {2, 4})
      for (int warps : {1, 2, 3, 6}) {
        const int tile = 36864;  // 32 lanes x 1152 B
        if ((size_t)warps * depth * tile > 220 * 1024) continue;
        const size_t usable = want / tile * tile;
        ms = time_ms([&] { k_bulk<true><<<sms, warps * 32, (size_t)warps * depth * tile>>>(buf, usable, tile, 2304, warps, depth, 10); });
        snprintf(cfg, sizeof(cfg), "36864 B tiles as 32 x 1152 B lane stores, %d in flight per warp, %d issuing warps/SM", depth, warps);
        emit(mem, "bulk_lanes", cfg, ms, usable);
      }
  }
  printf("\n]}\n");
  cudaError_t e = cudaDeviceSynchronize();
  fprintf(stderr, "%s\n", cudaGetErrorString(e));
  return 0;
}
