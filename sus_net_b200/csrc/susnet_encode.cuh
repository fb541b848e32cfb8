// susnet_encode.cuh -- observation featurizers (src/features/component.py, model_ready.py) on the GPU.
//
// Output tensors (float32, item-major so that (B, T, ...) views of B*T items are free):
//   GLOBAL       spatial [n][A+2][9][9] (one copy, shared by all agent views)   non_spatial [A][n][F]
//   PERSPECTIVE  spatial [A][n][A+2][9][9]                                      non_spatial [A][n][F]
//   FLAT         (no spatial tensor; the reference returns zeros(B,T,1))        non_spatial [n][F]
// Plane index order is [channel][x][y] (component.py:47-49,98,125).
#pragma once
#include "susnet_device.cuh"

namespace susnet {

struct DevEncode {
  int32_t kind, n_components, sp_floats, ns_floats;
  int32_t planes_u8;  // SUS_ENCODE_PLANES_U8: the plane tensor holds one byte per cell instead of one float
  int32_t components[SUS_MAX_FLAT_COMPONENTS];
};

// Zero `n_floats` floats starting at `base` with the whole warp: scalar stores up to the first 16-byte
// boundary, 128-bit stores over the aligned body, scalar stores over the (at most 3-float) tail.
__device__ __forceinline__ void warp_zero_fill(float* __restrict__ base, int64_t n_floats, int lane) {
  int64_t head = (int64_t)((16u - (uint32_t)(reinterpret_cast<uintptr_t>(base) & 15u)) & 15u) >> 2;
  if (head > n_floats) head = n_floats;
  if (lane < head) base[lane] = 0.f;
  float* body = base + head;
  const int64_t n = n_floats - head, n4 = n >> 2;
  float4* b4 = reinterpret_cast<float4*>(body);
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int64_t i = lane; i < n4; i += 32) b4[i] = z;
  const int64_t tail = (n4 << 2) + lane;
  if (tail < n) body[tail] = 0.f;
}

__device__ __forceinline__ bool code_ok(uint32_t c) { return (c >> 4) <= 8u && (c & 15u) <= 8u; }

// agent / job planes of one item; `ch_of(i)` gives the channel agent i is shown in
// (AgentPositionsFeaturizer component.py:90-100, JobFeaturizer component.py:116-127).
template <typename ChannelOf, typename T>
__device__ __forceinline__ void scatter_planes(const DevConfig& c, const ObsState& o, T* __restrict__ sp,
                                               ChannelOf ch_of, T v) {
  const int A = c.A, J = c.J;
  for (int i = 0; i < A; ++i) {
    const uint32_t b = get_byte(o.pos, i);
    if (((o.alive >> i) & 1u) && code_ok(b)) sp[ch_of(i) * 81 + code_cell(b)] = v;
  }
  for (int j = 0; j < J; ++j) {
    const uint32_t b = get_byte(o.jobpos, j);
    if (code_ok(b)) sp[(A + (int)((o.jobdone >> j) & 1u)) * 81 + code_cell(b)] = v;
  }
}

// The same planes as a packed list of float offsets inside one item's (A+2)*81 row, computed once per item so that
// setting and clearing the ones in a staging tile is a handful of shifts and stores.  Up to 16 entries of 10 bits
// (offset < 1024 > (8+2)*81); 0x3ff marks "no entry" (dead agent, coordinate outside the grid).
struct PlaneOffsets {
  uint64_t w[3];  // entries 0-5, 6-11, 12-15
  int n;
  __device__ __forceinline__ uint32_t get(int q) const {
    const uint64_t x = q < 6 ? w[0] : (q < 12 ? w[1] : w[2]);
    const int sh = 10 * (q < 6 ? q : (q < 12 ? q - 6 : q - 12));
    return (uint32_t)(x >> sh) & 0x3ffu;
  }
};

template <typename ChannelOf>
__device__ __forceinline__ PlaneOffsets plane_offsets(const DevConfig& c, const ObsState& o, ChannelOf ch_of) {
  const int A = c.A, J = c.J;
  PlaneOffsets po;
  po.w[0] = po.w[1] = po.w[2] = 0;
  po.n = A + J;
  for (int q = 0; q < A + J; ++q) {
    uint32_t off = 0x3ffu;
    if (q < A) {
      const uint32_t b = get_byte(o.pos, q);
      if (((o.alive >> q) & 1u) && code_ok(b)) off = (uint32_t)ch_of(q) * 81u + code_cell(b);
    } else {
      const int j = q - A;
      const uint32_t b = get_byte(o.jobpos, j);
      if (code_ok(b)) off = (uint32_t)(A + (int)((o.jobdone >> j) & 1u)) * 81u + code_cell(b);
    }
    const int wi = q < 6 ? 0 : (q < 12 ? 1 : 2);
    const int sh = 10 * (q - 6 * wi);
    const uint64_t v = (uint64_t)off << sh;
    if (wi == 0) po.w[0] |= v; else if (wi == 1) po.w[1] |= v; else po.w[2] |= v;
  }
  return po;
}

__device__ __forceinline__ void put_planes(const PlaneOffsets& po, float* __restrict__ row, float v) {
  uint64_t x = po.w[0];
  for (int q = 0; q < po.n; ++q) {  // walk the 10-bit fields by shifting (no dynamic field index)
    if (q == 6) x = po.w[1];
    if (q == 12) x = po.w[2];
    const uint32_t off = (uint32_t)x & 0x3ffu;
    x >>= 10;
    if (off != 0x3ffu) row[off] = v;
  }
}

// GlobalFeaturizer non-spatial row of view k: alive, [tag counts], job status, one-hot(k)
// (model_ready.py:237-247,293-303).  Tagging fields are read in TUPLE order (documented deviation from
// the reference's inconsistent state_fields map).
__device__ __forceinline__ void global_ns_row(const DevConfig& c, const ObsState& o, int k, float* __restrict__ r) {
  const int A = c.A, J = c.J;
  int p = 0;
  for (int i = 0; i < A; ++i) r[p++] = (float)((o.alive >> i) & 1u);
  if (c.variant == SUS_VARIANT_TAGGING)
    for (int i = 0; i < A; ++i) r[p++] = (float)((o.tagcnt >> (4 * i)) & 15u);
  for (int j = 0; j < J; ++j) r[p++] = (float)((o.jobdone >> j) & 1u);
  for (int i = 0; i < A; ++i) r[p++] = i == k ? 1.0f : 0.0f;
}

// PerspectiveFeaturizer: view k shows agents in the order [k, 0..k-1, k+1..A-1] (model_ready.py:184-193)
__device__ __forceinline__ int persp_agent_of_channel(int k, int ch) { return ch == 0 ? k : (ch <= k ? ch - 1 : ch); }
__device__ __forceinline__ int persp_channel_of_agent(int k, int i) { return i == k ? 0 : (i < k ? i + 1 : i); }

__device__ __forceinline__ void persp_ns_row(const DevConfig& c, const ObsState& o, int k, float* __restrict__ r) {
  const int A = c.A, J = c.J;
  int p = 0;  // per-agent fields, field-major, permuted like the channels; then job status (model_ready.py:195-204)
  for (int ch = 0; ch < A; ++ch) r[p++] = (float)((o.alive >> persp_agent_of_channel(k, ch)) & 1u);
  if (c.variant == SUS_VARIANT_TAGGING)
    for (int ch = 0; ch < A; ++ch) r[p++] = (float)((o.tagcnt >> (4 * persp_agent_of_channel(k, ch))) & 15u);
  for (int j = 0; j < J; ++j) r[p++] = (float)((o.jobdone >> j) & 1u);
}

// All A view rows of one item in one pass (row k at base + k * stride): the alive / tag / job-status prefix is the
// same in every Global view, so each value is converted once and stored A times.
__device__ __forceinline__ void global_ns_rows(const DevConfig& c, const ObsState& o, float* __restrict__ base, int stride) {
  const int A = c.A, J = c.J;
  float* col = base;  // column p of view 0; view k is k * stride further
  for (int i = 0; i < A; ++i, ++col) {
    const float v = (float)((o.alive >> i) & 1u);
    float* w = col;
    for (int k = 0; k < A; ++k, w += stride) *w = v;
  }
  if (c.variant == SUS_VARIANT_TAGGING)
    for (int i = 0; i < A; ++i, ++col) {
      const float v = (float)((o.tagcnt >> (4 * i)) & 15u);
      float* w = col;
      for (int k = 0; k < A; ++k, w += stride) *w = v;
    }
  for (int j = 0; j < J; ++j, ++col) {
    const float v = (float)((o.jobdone >> j) & 1u);
    float* w = col;
    for (int k = 0; k < A; ++k, w += stride) *w = v;
  }
  for (int k = 0; k < A; ++k, col += stride)  // one-hot of the view index
    for (int i = 0; i < A; ++i) col[i] = i == k ? 1.0f : 0.0f;
}

__device__ __forceinline__ int iabs(int v) { return v < 0 ? -v : v; }

// Where a flat row's values go.  FloatRow writes the float32 feature row itself (global memory or a staging row);
// ByteRow writes one biased byte per value (v + 128) into a row that was PREFILLED with the byte of 0, for the
// staged-bytes kernel (k_step_flat), which expands the bytes to floats with coalesced 128-bit stores.  Every flat
// component but the scent one is integer-valued with |v| <= 30, so the byte row is exact.
struct FloatRow {
  static constexpr bool kPrefilledZero = false;
  float* __restrict__ r;
  __device__ __forceinline__ void put(int p, int v) const { r[p] = (float)v; }
  __device__ __forceinline__ void putf(int p, float v) const { r[p] = v; }
};
constexpr uint32_t kByteRowBias = 128u;
struct ByteRow {
  static constexpr bool kPrefilledZero = true;
  uint8_t* __restrict__ r;
  __device__ __forceinline__ void put(int p, int v) const { r[p] = (uint8_t)(v + (int)kByteRowBias); }
  __device__ __forceinline__ void putf(int, float) const {}  // no float-valued component is ever staged as bytes
};
// RecordRow: a flat row as the short list of its NON-ZERO values, for the warp-specialised Flat kernel (k_step_flat_ws): the
// compute warp leaves K 16-bit entries per env -- float offset inside the row (10 bits) | value as 6-bit two's complement
// (every flat component but the scent one is an integer with |v| <= 30) -- prefilled with 0xffff = "no entry"; an emitter
// warp sets those values in a persistently-zero float tile, bulk-stores the tile and clears them again.
struct RecordRow {
  static constexpr bool kPrefilledZero = true;
  uint16_t* __restrict__ rec;
  int* cnt;
  int base, K;
  __device__ __forceinline__ void put(int p, int v) const {
    if (v != 0 && *cnt < K) rec[(*cnt)++] = (uint16_t)((uint32_t)(base + p) | (((uint32_t)v & 63u) << 10));
  }
  __device__ __forceinline__ void putf(int, float) const {}
};
constexpr uint32_t kRecordNone = 0xffffu;
__device__ __forceinline__ float record_value(uint32_t ent) { return (float)((int)(((ent >> 10) & 63u) ^ 32u) - 32); }

// float value of byte k of a word of a ByteRow: the byte becomes the low mantissa bits of 2^23, minus (2^23 + bias)
__device__ __forceinline__ float byte_row_value(uint32_t w, int k) {
  return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650u + (uint32_t)k)) - (8388608.0f + (float)kByteRowBias);
}

// One component of a CompositeFeaturizer (component.py) written at r[0..n); returns n.
// TA: compile-time agent count of a specialised kernel instantiation (0: read it from the config).  With it the per-agent
// loops unroll, the byte extractions shift by constants and the row offsets fold into the store instructions: the row builder
// was 550 of the 1 910 warp instructions per 32 envs of the Flat-98 step kernel (ncu source counters), most of them loop control,
// variable 64-bit shifts and divergence bookkeeping.
template <typename Row, int TA = 0>
__device__ __forceinline__ int flat_component(const DevConfig& c, const GridTables& tb, const ObsState& o, int comp,
                                              const Row& r) {
  const int A = c.A, J = c.J;
  const int AH = TA ? TA : c.A;  // the components of the reference's training recipes (one-hot positions, crew flags, closest
                                 // crew) unroll; the rest keep runtime loops (unrolling all of them spilled at 64 registers)
  const uint32_t b0 = get_byte(o.pos, 0);  // "imposter is agent 0" (component.py:262,289,355,440,467)
  const int ix = (int)code_x(b0), iy = (int)code_y(b0);
  int p = 0;
  switch (comp) {
    case SUS_FC_ONEHOT_POS:  // component.py:226-240
      for (int i = 0; i < AH; ++i) {
        const uint32_t b = get_byte(o.pos, i);
        const bool al = (o.alive >> i) & 1u;
        if (Row::kPrefilledZero) {  // only the (at most two) ones of an alive agent
          if (al && code_x(b) < 9u) r.put(p + (int)code_x(b), 1);
          if (al && code_y(b) < 9u) r.put(p + 9 + (int)code_y(b), 1);
          p += 18;
        } else {
          for (int q = 0; q < 9; ++q) r.put(p++, (al && code_x(b) == (uint32_t)q) ? 1 : 0);
          for (int q = 0; q < 9; ++q) r.put(p++, (al && code_y(b) == (uint32_t)q) ? 1 : 0);
        }
      }
      break;
    case SUS_FC_COORDS:  // component.py:389-399 (dead agents included)
      for (int i = 0; i < A; ++i) {
        const uint32_t b = get_byte(o.pos, i);
        r.put(p++, (int)code_x(b)); r.put(p++, (int)code_y(b));
      }
      break;
    case SUS_FC_ALIVE_CREW:  // component.py:411-421
      for (int i = 1; i < AH; ++i) r.put(p++, (int)((o.alive >> i) & 1u));
      break;
    case SUS_FC_CLOSEST_CREW: {  // component.py:460-478: default distance 18, first argmin
      int best = 0, best_d = 1 << 20;
      for (int i = 1; i < AH; ++i) {
        const uint32_t b = get_byte(o.pos, i);
        const int d = ((o.alive >> i) & 1u) ? iabs(ix - (int)code_x(b)) + iabs(iy - (int)code_y(b)) : 18;
        if (d < best_d) { best_d = d; best = i - 1; }
      }
      if (Row::kPrefilledZero) {  // only the one
        r.put(best, 1);
        p = AH - 1;
      } else {
        for (int i = 0; i < AH - 1; ++i) r.put(p++, i == best ? 1 : 0);
      }
    } break;
    case SUS_FC_L1_CREW:  // component.py:433-448
      for (int i = 1; i < A; ++i) {
        const uint32_t b = get_byte(o.pos, i);
        r.put(p++, ((o.alive >> i) & 1u) ? iabs(ix - (int)code_x(b)) + iabs(iy - (int)code_y(b)) : -1);
      }
      break;
    case SUS_FC_DIST_TO_IMPOSTER: {  // component.py:255-273: alive others compacted, trailing zeros
      const int n = 2 * (A - 1);
      for (int i = 1; i < A; ++i) {
        const uint32_t b = get_byte(o.pos, i);
        if ((o.alive >> i) & 1u) { r.put(p++, ix - (int)code_x(b)); r.put(p++, iy - (int)code_y(b)); }
      }
      while (p < n) r.put(p++, 0);
    } break;
    case SUS_FC_WALLS:  // component.py:286-296: 3x3 patch of the zero-padded grid around agent 0
      for (int dx = -1; dx <= 1; ++dx)
        for (int dy = -1; dy <= 1; ++dy) {
          const uint32_t nb = ((uint32_t)(ix + dx) << 4 | ((uint32_t)(iy + dy) & 15u)) & 0xffu;
          const bool in = (uint32_t)(ix + dx) <= 8u && (uint32_t)(iy + dy) <= 8u;
          r.put(p++, (in && ((tb.valid_bits[nb >> 5] >> (nb & 31u)) & 1u)) ? 1 : 0);
        }
      break;
    case SUS_FC_ROOMS: {  // component.py:308-329, ROOM_MASKS component.py:8-17
      int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int i = 0; i < A; ++i) {
        if (!((o.alive >> i) & 1u)) continue;
        const uint32_t b = get_byte(o.pos, i);
        const bool xl = code_x(b) < 5u, yl = code_y(b) < 5u;
        const int room = xl ? (yl ? 0 : 1) : (yl ? 3 : 2);
        const int at = (i == 0 ? 0 : 4) + room;
#pragma unroll
        for (int q = 0; q < 8; ++q) cnt[q] += q == at ? 1 : 0;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) r.put(p++, cnt[q]);
    } break;
    case SUS_FC_SCENT: {  // component.py:344-375: float32 accumulation of (9 - d) / 9 computed in double
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      for (int i = 1; i < A; ++i) {
        if (!((o.alive >> i) & 1u)) continue;
        const uint32_t b = get_byte(o.pos, i);
        const double xs = (9.0 - (double)((int)code_x(b) - ix)) / 9.0, ys = (9.0 - (double)((int)code_y(b) - iy)) / 9.0;
        if (xs > 0) s0 += (float)xs; else s1 += (float)xs;
        if (ys > 0) s2 += (float)ys; else s3 += (float)ys;
      }
      r.putf(p++, s0); r.putf(p++, s1); r.putf(p++, s2); r.putf(p++, s3);
    } break;
    case SUS_FC_STATE_ALIVE:  // StateFieldFeaturizer: component.py:210-214
      for (int i = 0; i < A; ++i) r.put(p++, (int)((o.alive >> i) & 1u));
      break;
    case SUS_FC_STATE_JOB_STATUS:
      for (int j = 0; j < J; ++j) r.put(p++, (int)((o.jobdone >> j) & 1u));
      break;
    case SUS_FC_STATE_USED_TAGS:
      for (int i = 0; i < A; ++i) r.put(p++, (int)((o.used >> i) & 1u));
      break;
    case SUS_FC_STATE_TAG_COUNTS:
      for (int i = 0; i < A; ++i) r.put(p++, (int)((o.tagcnt >> (4 * i)) & 15u));
      break;
  }
  return p;
}

// all components of a flat row, back to back
template <typename RowT, int TA = 0, typename Elem>
__device__ __forceinline__ void flat_row(const DevConfig& c, const DevEncode& enc, const GridTables& tb, const ObsState& o,
                                         Elem* __restrict__ row) {
  for (int q = 0; q < enc.n_components; ++q) row += flat_component<RowT, TA>(c, tb, o, enc.components[q], RowT{row});
}

// all components of a flat row as a RecordRow (entries in component order); returns the number of entries written
__device__ __forceinline__ int flat_row_record(const DevConfig& c, const DevEncode& enc, const GridTables& tb, const ObsState& o,
                                               uint16_t* __restrict__ rec, int K) {
  int cnt = 0, base = 0;
  for (int q = 0; q < enc.n_components; ++q) base += flat_component(c, tb, o, enc.components[q], RecordRow{rec, &cnt, base, K});
  return cnt;
}

// Encode the 32 items a warp owns (item = item0 + lane; `cnt` of them exist, `have` says whether this
// lane's item exists).  All 32 lanes must call this together.
__device__ __forceinline__ void warp_encode(const DevConfig& c, const DevEncode& enc, const GridTables& tb,
                                            const ObsState& o, int64_t item0, int cnt, bool have, int64_t n_items,
                                            void* __restrict__ spatial_any, float* __restrict__ non_spatial) {
  const int lane = threadIdx.x & 31;
  const int64_t item = item0 + lane;
  const int A = c.A;
  if (enc.planes_u8 && (enc.kind == SUS_ENCODE_GLOBAL || enc.kind == SUS_ENCODE_PERSPECTIVE)) {
    // one byte per cell (opt-in compact planes): zero the warp's rows bytewise, then scatter the ones
    uint8_t* sp8 = static_cast<uint8_t*>(spatial_any);
    const int R = enc.sp_floats;
    const int views = enc.kind == SUS_ENCODE_GLOBAL ? 1 : A;
    for (int k = 0; k < views; ++k) {
      uint8_t* base = sp8 + ((int64_t)k * n_items + item0) * R;
      for (int64_t i = lane; i < (int64_t)cnt * R; i += 32) base[i] = 0;
    }
    __syncwarp();
    if (have) {
      for (int k = 0; k < views; ++k)
        scatter_planes(c, o, sp8 + ((int64_t)k * n_items + item) * R, [k](int i) { return persp_channel_of_agent(k, i); }, (uint8_t)1);
      for (int k = 0; k < A; ++k) {
        float* row = non_spatial + ((int64_t)k * n_items + item) * enc.ns_floats;
        if (enc.kind == SUS_ENCODE_GLOBAL) global_ns_row(c, o, k, row); else persp_ns_row(c, o, k, row);
      }
    }
    return;
  }
  float* __restrict__ spatial = static_cast<float*>(spatial_any);
  if (enc.kind == SUS_ENCODE_GLOBAL) {
    const int R = enc.sp_floats;
    warp_zero_fill(spatial + item0 * R, (int64_t)cnt * R, lane);
    __syncwarp();
    if (have) {
      scatter_planes(c, o, spatial + item * R, [](int i) { return i; }, 1.0f);
      for (int k = 0; k < A; ++k) global_ns_row(c, o, k, non_spatial + ((int64_t)k * n_items + item) * enc.ns_floats);
    }
  } else if (enc.kind == SUS_ENCODE_PERSPECTIVE) {
    const int R = enc.sp_floats;
    for (int k = 0; k < A; ++k) warp_zero_fill(spatial + ((int64_t)k * n_items + item0) * R, (int64_t)cnt * R, lane);
    __syncwarp();
    if (have) {
      for (int k = 0; k < A; ++k) {
        scatter_planes(c, o, spatial + ((int64_t)k * n_items + item) * R,
                       [k](int i) { return persp_channel_of_agent(k, i); }, 1.0f);
        persp_ns_row(c, o, k, non_spatial + ((int64_t)k * n_items + item) * enc.ns_floats);
      }
    }
  } else if (enc.kind == SUS_ENCODE_FLAT) {
    if (have) flat_row<FloatRow>(c, enc, tb, o, non_spatial + item * enc.ns_floats);
  }
}

}  // namespace susnet
