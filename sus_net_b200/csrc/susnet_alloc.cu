// susnet_alloc.cu -- L2-compressible device memory for the feature tensors.
//
// The feature tensors are > 97 % of the bytes of a fused step and almost all zeros (<= A + J ones in (A + 2) * 81
// floats).  Blackwell's L2 can keep such lines compressed on their way to and from HBM ("generic compression"), but only
// for allocations created with cuMemCreate + CU_MEM_ALLOCATION_COMP_GENERIC -- cudaMalloc / the torch allocator never
// are.  Measured on B200 (tools/micro/compressible_bench.cu): the emitter's bulk-store stream of 36 KB tiles with 10 ones
// per 2 268-byte row runs at 7.5 TB/s into a compressible allocation against 6.4 TB/s into a cudaMalloc one, and reading
// the tiles back at 9.3 TB/s against 6.9 TB/s.  The Python featurizers therefore place their output buffers here.
//
// The driver entry points are resolved through cudaGetDriverEntryPoint, so the library does not link against libcuda
// (it must load on machines without a driver for the CPU-side tests).
#include <cstdint>
#include <map>
#include <mutex>
#include <string>

#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/susnet_b200.h"

extern "C" int sus_internal_fail(int code, const char* msg);

namespace {

struct Driver {
  decltype(&cuDeviceGetAttribute) deviceGetAttribute = nullptr;
  decltype(&cuMemGetAllocationGranularity) memGetAllocationGranularity = nullptr;
  decltype(&cuMemCreate) memCreate = nullptr;
  decltype(&cuMemGetAllocationPropertiesFromHandle) memGetAllocationPropertiesFromHandle = nullptr;
  decltype(&cuMemAddressReserve) memAddressReserve = nullptr;
  decltype(&cuMemMap) memMap = nullptr;
  decltype(&cuMemSetAccess) memSetAccess = nullptr;
  decltype(&cuMemUnmap) memUnmap = nullptr;
  decltype(&cuMemRelease) memRelease = nullptr;
  decltype(&cuMemAddressFree) memAddressFree = nullptr;
  bool ok = false;
};

template <typename F>
bool resolve(const char* name, F& fn) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult st;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess || !p) return false;
  fn = reinterpret_cast<F>(p);
  return true;
}

const Driver& driver() {
  static Driver d;
  static std::once_flag once;
  std::call_once(once, [] {
    d.ok = resolve("cuDeviceGetAttribute", d.deviceGetAttribute) &&
           resolve("cuMemGetAllocationGranularity", d.memGetAllocationGranularity) && resolve("cuMemCreate", d.memCreate) &&
           resolve("cuMemGetAllocationPropertiesFromHandle", d.memGetAllocationPropertiesFromHandle) &&
           resolve("cuMemAddressReserve", d.memAddressReserve) && resolve("cuMemMap", d.memMap) &&
           resolve("cuMemSetAccess", d.memSetAccess) && resolve("cuMemUnmap", d.memUnmap) &&
           resolve("cuMemRelease", d.memRelease) && resolve("cuMemAddressFree", d.memAddressFree);
  });
  return d;
}

struct Block {
  CUmemGenericAllocationHandle handle;
  size_t size;
  int device;
};
std::mutex g_mu;
std::map<uintptr_t, Block> g_blocks;

int drv_fail(const char* what, CUresult r) {
  return sus_internal_fail(SUS_ERR_CUDA, (std::string(what) + " failed with CUresult " + std::to_string((int)r)).c_str());
}

}  // namespace

extern "C" {

int sus_alloc_compressible(int device, uint64_t bytes, void** ptr, uint64_t* allocated) {
  if (!ptr || bytes == 0) return sus_internal_fail(SUS_ERR_INVALID_ARGUMENT, "sus_alloc_compressible: NULL ptr or zero size");
  *ptr = nullptr;
  struct Restore {  // leave the caller's current device as it was
    int prev = -1;
    Restore() { cudaGetDevice(&prev); }
    ~Restore() { if (prev >= 0) cudaSetDevice(prev); }
  } restore;
  if (cudaSetDevice(device) != cudaSuccess || cudaFree(nullptr) != cudaSuccess)
    return sus_internal_fail(SUS_ERR_CUDA, "sus_alloc_compressible: cannot select the CUDA device");
  const Driver& d = driver();
  if (!d.ok) return sus_internal_fail(SUS_ERR_UNSUPPORTED, "the CUDA driver does not export the virtual-memory API");
  int supported = 0;
  if (d.deviceGetAttribute(&supported, CU_DEVICE_ATTRIBUTE_GENERIC_COMPRESSION_SUPPORTED, device) != CUDA_SUCCESS || !supported)
    return sus_internal_fail(SUS_ERR_UNSUPPORTED, "this device does not support generic (L2) compression");
  CUmemAllocationProp prop = {};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  prop.allocFlags.compressionType = CU_MEM_ALLOCATION_COMP_GENERIC;
  size_t gran = 0;
  CUresult r = d.memGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
  if (r != CUDA_SUCCESS || gran == 0) return drv_fail("cuMemGetAllocationGranularity", r);
  const size_t size = ((size_t)bytes + gran - 1) / gran * gran;
  CUmemGenericAllocationHandle h;
  if ((r = d.memCreate(&h, size, &prop, 0)) != CUDA_SUCCESS) return drv_fail("cuMemCreate", r);
  CUmemAllocationProp got = {};
  if (d.memGetAllocationPropertiesFromHandle(&got, h) != CUDA_SUCCESS || got.allocFlags.compressionType != CU_MEM_ALLOCATION_COMP_GENERIC) {
    d.memRelease(h);  // the driver may silently fall back to an uncompressed allocation: report that instead
    return sus_internal_fail(SUS_ERR_UNSUPPORTED, "the driver did not grant a compressible allocation");
  }
  CUdeviceptr p = 0;
  if ((r = d.memAddressReserve(&p, size, 0, 0, 0)) != CUDA_SUCCESS) { d.memRelease(h); return drv_fail("cuMemAddressReserve", r); }
  if ((r = d.memMap(p, size, 0, h, 0)) != CUDA_SUCCESS) { d.memAddressFree(p, size); d.memRelease(h); return drv_fail("cuMemMap", r); }
  CUmemAccessDesc acc = {};
  acc.location = prop.location;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  if ((r = d.memSetAccess(p, size, &acc, 1)) != CUDA_SUCCESS) {
    d.memUnmap(p, size); d.memAddressFree(p, size); d.memRelease(h);
    return drv_fail("cuMemSetAccess", r);
  }
  {
    std::lock_guard<std::mutex> lock(g_mu);
    g_blocks[(uintptr_t)p] = Block{h, size, device};
  }
  *ptr = reinterpret_cast<void*>(p);
  if (allocated) *allocated = size;
  return SUS_OK;
}

// (library-internal) does `p` point into a live compressible block?  The launch heuristics differ: stores into such a
// block drain faster, which moves the balance between the compute warps and the emitter warp of k_step_ws.
int sus_internal_is_compressible(const void* p) {
  if (!p) return 0;
  std::lock_guard<std::mutex> lock(g_mu);
  auto it = g_blocks.upper_bound((uintptr_t)p);
  if (it == g_blocks.begin()) return 0;
  --it;
  return (uintptr_t)p < it->first + it->second.size ? 1 : 0;
}

int sus_free_compressible(void* ptr) {
  if (!ptr) return SUS_OK;
  Block b;
  {
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_blocks.find((uintptr_t)ptr);
    if (it == g_blocks.end()) return sus_internal_fail(SUS_ERR_INVALID_ARGUMENT, "sus_free_compressible: not a live compressible allocation");
    b = it->second;
    g_blocks.erase(it);
  }
  const Driver& d = driver();
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(b.device);
  cudaDeviceSynchronize();  // nothing in flight may still touch the range
  const CUdeviceptr p = (CUdeviceptr)(uintptr_t)ptr;
  CUresult r = d.memUnmap(p, b.size);
  if (r == CUDA_SUCCESS) r = d.memAddressFree(p, b.size);
  const CUresult r2 = d.memRelease(b.handle);
  cudaSetDevice(prev);
  if (r != CUDA_SUCCESS) return drv_fail("cuMemUnmap / cuMemAddressFree", r);
  if (r2 != CUDA_SUCCESS) return drv_fail("cuMemRelease", r2);
  return SUS_OK;
}

}  // extern "C"
