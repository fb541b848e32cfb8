// susnet_tile.cuh -- shared-memory staging + TMA bulk stores for everything a warp's 32 envs emit.
//
// Why: the feature tensors are >97 % of the bytes of a step (2568 of 2650 B per env at the headline config)
// and are almost all zeros (<= A+J ones in (A+2)*81 floats).  Writing them from registers costs either a
// 128-bit store instruction per 16 bytes plus the arithmetic to build the floats, or zero-fill + scattered
// partial-sector stores (the v1 path: 2.35x sector amplification into L2, LSU-throttled).  Instead each warp
// keeps a persistently ZERO tile of G envs x spatial floats in shared memory, sets the handful of ones, hands
// the whole tile to the TMA engine with one `cp.async.bulk.global.shared::cta` (SASS: UBLKCP), waits until
// the engine has read the tile, and clears the same handful of ones.  The dense per-env rows (non-spatial
// features, rewards, replay-layout state row) are transposed through a second shared-memory region so that
// they too leave the SM as full-line bulk stores.  The LSU never sees the bulk bytes.
#pragma once
#include "susnet_encode.cuh"

namespace susnet {

// per-warp shared-memory layout, computed on the host (bytes; every offset is a multiple of 128)
struct TileLayout {
  int32_t G;          // envs per spatial sub-tile (4 or 8; 4 | G keeps G*row_bytes a multiple of 16)
  int32_t sp_off, sp_bytes;    // spatial sub-tile  [G][sp_floats]
  int32_t ns_off, ns_bytes;    // non-spatial rows  [ns_views][32][ns_floats]
  int32_t rew_off, rew_bytes;  // rewards           [32][A] f32 or f64
  int32_t nf_off, nf_bytes;    // next_flat         [32][S] f32
  int32_t per_warp;            // total bytes per warp
  int32_t warps;               // warps per CTA
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Drain `bytes` of shared memory to global memory.  Lane 0 issues a TMA bulk store when source, destination and
// size are 16-byte multiples (the caller commits/waits); otherwise (ragged tail, odd caller pointers) the warp
// copies 4-byte words itself.  Must be called by all 32 lanes after fence_proxy_async_smem() + __syncwarp().
__device__ __forceinline__ bool drain(void* gdst, const void* ssrc, uint32_t bytes, int lane) {
  if (bytes == 0) return false;
  const bool bulk = (((uint32_t)reinterpret_cast<uintptr_t>(gdst) | bytes) & 15u) == 0;
  if (bulk) {
    if (lane == 0) bulk_store(gdst, ssrc, bytes);
    return true;
  }
  const uint32_t* s = static_cast<const uint32_t*>(ssrc);
  uint32_t* g = static_cast<uint32_t*>(gdst);
  for (uint32_t i = lane; i < (bytes >> 2); i += 32) g[i] = s[i];
  return false;
}

// Everything a warp emits for its 32 items besides the per-env scalars.  `sm` is the warp's staging block.
struct WarpEmitter {
  uint8_t* sm;
  const TileLayout* L;
  int lane;
  bool pending;  // lane 0 has uncommitted/unwaited bulk stores reading this warp's staging block

  __device__ __forceinline__ float* sp() const { return reinterpret_cast<float*>(sm + L->sp_off); }
  __device__ __forceinline__ float* ns() const { return reinterpret_cast<float*>(sm + L->ns_off); }
  __device__ __forceinline__ uint8_t* rew() const { return sm + L->rew_off; }
  __device__ __forceinline__ float* nf() const { return reinterpret_cast<float*>(sm + L->nf_off); }

  // the TMA engine must be done READING the staging block before the warp overwrites it
  __device__ __forceinline__ void acquire() {
    if (pending) {
      if (lane == 0) bulk_wait_read_all();
      pending = false;
    }
    __syncwarp();
  }

  __device__ __forceinline__ void zero_spatial() {  // once per kernel: the spatial sub-tile is zero between uses
    float4* p = reinterpret_cast<float4*>(sm + L->sp_off);
    for (int i = lane; i < (L->sp_bytes >> 4); i += 32) p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
  }

  __device__ __forceinline__ void finish() {  // before the CTA exits: all bulk stores complete
    if (lane == 0) bulk_wait_all();
    __syncwarp();
  }
};

// Feature encode of the warp's 32 items through the staging block (same outputs as warp_encode()).
// Call with all 32 lanes; `cnt` items exist starting at item0; lane's item exists iff `have`.
__device__ __forceinline__ void warp_encode_tma(const DevConfig& c, const DevEncode& enc, const GridTables& tb,
                                                WarpEmitter& em, const ObsState& o, int64_t item0, int cnt, bool have,
                                                int64_t n_items, float* __restrict__ spatial,
                                                float* __restrict__ non_spatial, bool ns_already_acquired) {
  const int lane = em.lane, A = c.A;
  const TileLayout& L = *em.L;
  const int F = enc.ns_floats, R = enc.sp_floats;
  // ---- dense rows: every lane writes its item's rows [view][lane][F]; one bulk store per view
  if (!ns_already_acquired) em.acquire();
  float* ns = em.ns();
  if (have) {
    if (enc.kind == SUS_ENCODE_GLOBAL) {
      for (int k = 0; k < A; ++k) global_ns_row(c, o, k, ns + (k * 32 + lane) * F);
    } else if (enc.kind == SUS_ENCODE_PERSPECTIVE) {
      for (int k = 0; k < A; ++k) persp_ns_row(c, o, k, ns + (k * 32 + lane) * F);
    } else {
      float* r = ns + lane * F;
      for (int q = 0; q < enc.n_components; ++q) r += flat_component(c, tb, o, enc.components[q], r);
    }
  }
  fence_proxy_async_smem();
  __syncwarp();
  const int views = enc.kind == SUS_ENCODE_FLAT ? 1 : A;
  bool any = false;
  for (int k = 0; k < views; ++k)
    any |= drain(non_spatial + ((int64_t)k * n_items + item0) * F, ns + k * 32 * F, (uint32_t)(cnt * F * 4), lane);
  if (any) {
    if (lane == 0) bulk_commit();
    em.pending = true;
  }
  if (R == 0) return;
  // ---- sparse planes: G items at a time through the persistently-zero sub-tile
  const int G = L.G;
  const int sp_views = enc.kind == SUS_ENCODE_GLOBAL ? 1 : A;
  float* sp = em.sp();
  for (int g0 = 0; g0 < cnt; g0 += G) {
    const int gc = cnt - g0 < G ? cnt - g0 : G;
    const bool mine = have && lane >= g0 && lane < g0 + G;
    for (int k = 0; k < sp_views; ++k) {
      if (mine) {
        if (enc.kind == SUS_ENCODE_GLOBAL) scatter_planes(c, o, sp + (lane - g0) * R, [](int i) { return i; }, 1.0f);
        else scatter_planes(c, o, sp + (lane - g0) * R, [k](int i) { return persp_channel_of_agent(k, i); }, 1.0f);
      }
      fence_proxy_async_smem();
      __syncwarp();
      const bool b = drain(spatial + ((int64_t)k * n_items + item0 + g0) * R, sp, (uint32_t)(gc * R * 4), lane);
      if (b) {
        if (lane == 0) { bulk_commit(); bulk_wait_read_all(); }
        em.pending = false;  // the wait covered every outstanding group of this lane
      }
      __syncwarp();
      if (mine) {
        if (enc.kind == SUS_ENCODE_GLOBAL) scatter_planes(c, o, sp + (lane - g0) * R, [](int i) { return i; }, 0.0f);
        else scatter_planes(c, o, sp + (lane - g0) * R, [k](int i) { return persp_channel_of_agent(k, i); }, 0.0f);
      }
    }
  }
  __syncwarp();
}

}  // namespace susnet
