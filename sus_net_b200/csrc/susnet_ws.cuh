// susnet_ws.cuh -- warp-specialised fused step + encode kernel (the headline path when planes are written).
//
// The non-specialised TMA kernel (k_step_tma) lets every warp compute a group of 32 envs and then push that
// group's tiles itself, so the TMA engine idles whenever all warps happen to be computing (measured: 0.498 ms
// against 0.442 ms for the same store pattern without compute, tools/micro/tma_store_bench.cu).  Here the two
// jobs are decoupled:
//   * CW compute warps run the step of their groups and leave, per group, a small record in a shared-memory
//     slot: 16 x 10-bit plane offsets per env, the dense non-spatial / reward / next_flat rows (which they bulk-
//     store themselves), then signal an mbarrier;
//   * ONE emitter warp turns the records into plane tiles: it sets the <= A+J ones of 8 envs in one of two
//     persistently-zero 18 KB tiles with all 32 lanes, issues the cp.async.bulk store, and clears the ones of
//     the tile it used two stores ago.  Two tile stores are always in flight, so the bulk-store stream never
//     waits for step arithmetic.
// Producer/consumer hand-off: full[w][s] / empty[w][s] mbarriers, two slots per compute warp.
#pragma once
#include "susnet_tile.cuh"

namespace susnet {

struct WsLayout {
  int32_t compute_warps;            // CW; the CTA has CW + 1 warps
  int32_t tile_envs;                // envs per plane tile of the emitter: 16 (default) or 8
  int32_t tile_bytes;               // one plane tile: tile_envs x sp_floats x 4, rounded to 128
  int32_t slot_bytes;               // one group record
  int32_t po_off, ns_off, rew_off, nf_off;  // inside a slot
  int32_t slots_off, bars_off, total_bytes; // inside dynamic shared memory (tiles first)
};

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// 16 plane offsets of one env as 16 x u16 (0x3ff = none), Global channel order
__device__ __forceinline__ void write_plane_record(const DevConfig& c, const ObsState& o, bool have, uint16_t* rec) {
  uint4* dst = reinterpret_cast<uint4*>(rec);
  dst[0] = dst[1] = make_uint4(0x03ff03ffu, 0x03ff03ffu, 0x03ff03ffu, 0x03ff03ffu);
  if (!have) return;
  const PlaneOffsets po = plane_offsets(c, o, [](int i) { return i; });
  uint64_t x = po.w[0];
  for (int q = 0; q < po.n; ++q) {
    if (q == 6) x = po.w[1];
    if (q == 12) x = po.w[2];
    rec[q] = (uint16_t)((uint32_t)x & 0x3ffu);
    x >>= 10;
  }
}

}  // namespace susnet
