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
__device__ __forceinline__ void bulk_wait_read_all_but_newest() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Drain `bytes` of shared memory to global memory.  Lane 0 issues a TMA bulk store when source, destination and
// size are 16-byte multiples (the caller commits/waits); otherwise (ragged tail, odd caller pointers) the warp
// copies 4-byte words (single bytes where the destination or the size is not a multiple of 4) itself.  Must be called by all 32 lanes after fence_proxy_async_smem() + __syncwarp().
__device__ __forceinline__ bool drain(void* gdst, const void* ssrc, uint32_t bytes, int lane) {
  if (bytes == 0) return false;
  const bool bulk = (((uint32_t)reinterpret_cast<uintptr_t>(gdst) | bytes) & 15u) == 0;
  if (bulk) {
    if (lane == 0) bulk_store(gdst, ssrc, bytes);
    return true;
  }
  if (((uint32_t)reinterpret_cast<uintptr_t>(gdst) & 3u) == 0) {
    const uint32_t* s = static_cast<const uint32_t*>(ssrc);
    uint32_t* g = static_cast<uint32_t*>(gdst);
    for (uint32_t i = lane; i < (bytes >> 2); i += 32) g[i] = s[i];
    for (uint32_t i = (bytes & ~3u) + lane; i < bytes; i += 32)  // byte tail (packed result records of a ragged batch)
      static_cast<uint8_t*>(gdst)[i] = static_cast<const uint8_t*>(ssrc)[i];
  } else {
    for (uint32_t i = lane; i < bytes; i += 32) static_cast<uint8_t*>(gdst)[i] = static_cast<const uint8_t*>(ssrc)[i];
  }
  return false;
}

// Everything a warp emits for its 32 items besides the per-env scalars.  `sm` is the warp's staging block.
//
// Bulk-group protocol of lane 0 (groups complete in commit order): per group of 32 items the warp commits one
// DENSE group (rewards + next_flat + non-spatial rows) and then one SPATIAL group per sub-tile.  A sub-tile store is
// retired (waited for, its ones cleared) only right before the sub-tile buffer is needed again, so the last
// spatial store of a group of items stays in flight while the warp computes the step of its next group.
struct WarpEmitter {
  uint8_t* sm;
  const TileLayout* L;
  int lane;
  bool dense_inflight;  // a DENSE group may still be reading the dense staging rows
  bool sp_inflight;     // a SPATIAL group may still be reading the sub-tile buffer
  bool dense_newer;     // the newest committed group is a DENSE one committed after the in-flight SPATIAL one
  bool dirty;           // this lane has ones set in the sub-tile buffer at dirty_row (described by dirty_po)
  float* dirty_row;
  PlaneOffsets dirty_po;

  __device__ __forceinline__ void init(uint8_t* block, const TileLayout* layout, int ln) {
    sm = block; L = layout; lane = ln;
    dense_inflight = sp_inflight = dense_newer = dirty = false;
    dirty_row = nullptr;
    dirty_po.w[0] = dirty_po.w[1] = dirty_po.w[2] = 0; dirty_po.n = 0;
  }
  __device__ __forceinline__ float* sp() const { return reinterpret_cast<float*>(sm + L->sp_off); }
  __device__ __forceinline__ float* ns() const { return reinterpret_cast<float*>(sm + L->ns_off); }
  __device__ __forceinline__ uint8_t* rew() const { return sm + L->rew_off; }
  __device__ __forceinline__ float* nf() const { return reinterpret_cast<float*>(sm + L->nf_off); }

  // before the dense staging rows are overwritten: the previous DENSE group has been read by the TMA engine
  __device__ __forceinline__ void acquire_dense() {
    if (dense_inflight) {
      if (lane == 0) {
        if (sp_inflight && !dense_newer) bulk_wait_read_all_but_newest();  // the spatial group is newer: let it fly
        else bulk_wait_read_all();
      }
      if (!(sp_inflight && !dense_newer)) sp_inflight = false;
      dense_inflight = false;
      dense_newer = false;
    }
    __syncwarp();
  }
  __device__ __forceinline__ void committed_dense() { dense_inflight = true; dense_newer = sp_inflight; }

  // before the sub-tile buffer is reused: its previous store has been read, then the ones it carried are cleared
  __device__ __forceinline__ void retire_spatial() {
    if (sp_inflight) {
      if (lane == 0) {
        if (dense_newer) bulk_wait_read_all_but_newest();  // the DENSE group committed after it may keep flying
        else bulk_wait_read_all();
      }
      if (!dense_newer) dense_inflight = false;
      sp_inflight = false;
    }
    __syncwarp();
    if (dirty) put_planes(dirty_po, dirty_row, 0.0f);
    dirty = false;
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
// `dense_open`: the caller already acquired the dense rows and issued (uncommitted) rewards / next_flat stores.
__device__ __forceinline__ void warp_encode_tma(const DevConfig& c, const DevEncode& enc, const GridTables& tb,
                                                WarpEmitter& em, const ObsState& o, int64_t item0, int cnt, bool have,
                                                int64_t n_items, void* __restrict__ spatial_any,
                                                float* __restrict__ non_spatial, bool dense_open, bool dense_any) {
  float* __restrict__ spatial = static_cast<float*>(spatial_any);  // (uint8 planes never take this path: make_layout refuses)
  const int lane = em.lane, A = c.A;
  const TileLayout& L = *em.L;
  const int F = enc.ns_floats, R = enc.sp_floats;
  // ---- dense rows: every lane writes its item's rows [view][lane][F]; one bulk store per view
  if (!dense_open) em.acquire_dense();
  float* ns = em.ns();
  if (have) {
    if (enc.kind == SUS_ENCODE_GLOBAL) {
      global_ns_rows(c, o, ns + lane * F, 32 * F);
    } else if (enc.kind == SUS_ENCODE_PERSPECTIVE) {
      for (int k = 0; k < A; ++k) persp_ns_row(c, o, k, ns + (k * 32 + lane) * F);
    } else {
      flat_row<FloatRow>(c, enc, tb, o, ns + lane * F);
    }
  }
  fence_proxy_async_smem();
  __syncwarp();
  const int views = enc.kind == SUS_ENCODE_FLAT ? 1 : A;
  bool any = dense_any;
  for (int k = 0; k < views; ++k)
    any |= drain(non_spatial + ((int64_t)k * n_items + item0) * F, ns + k * 32 * F, (uint32_t)(cnt * F * 4), lane);
  if (any) {
    if (lane == 0) bulk_commit();
    em.committed_dense();
  }
  if (R == 0) return;
  // ---- sparse planes: G items at a time through the persistently-zero sub-tile
  const int G = L.G;
  const int sp_views = enc.kind == SUS_ENCODE_GLOBAL ? 1 : A;
  float* sp = em.sp();
  PlaneOffsets po = plane_offsets(c, o, [](int i) { return i; });  // Global planes (= view 0 of Perspective)
  for (int k = 0; k < sp_views; ++k) {
    if (k > 0) po = plane_offsets(c, o, [k](int i) { return persp_channel_of_agent(k, i); });
    for (int g0 = 0; g0 < cnt; g0 += G) {
      const int gc = cnt - g0 < G ? cnt - g0 : G;
      em.retire_spatial();
      if (have && lane >= g0 && lane < g0 + G) {
        em.dirty = true; em.dirty_po = po; em.dirty_row = sp + (lane - g0) * R;
        put_planes(po, em.dirty_row, 1.0f);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (drain(spatial + ((int64_t)k * n_items + item0 + g0) * R, sp, (uint32_t)(gc * R * 4), lane)) {
        if (lane == 0) bulk_commit();
        em.sp_inflight = true;
        em.dense_newer = false;
      }
    }
  }
}

}  // namespace susnet
