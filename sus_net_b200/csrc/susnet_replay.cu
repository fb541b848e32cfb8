// susnet_replay.cu -- row (f1): the reference's replay layout (src/replay_memory.py:33-44) filled on the GPU.
//
// A transition block is three float streams that are CONTIGUOUS in the flattened (env, t, s) index x -- `states` <- the running
// sequence, `next_states` <- the sequence rolled by one row with the new state appended, and the sequence the next step starts
// from -- plus a few scalars per agent.  The N ring slots (idx + e) mod M are contiguous too, except at one wrap point.
// k_replay_push_v: the first part of the grid moves the streams four floats per thread (128-bit loads and stores wherever the
// bases are 16-byte aligned, per-element accesses for the one vector that straddles the ring's wrap point, a ragged tail or
// unaligned buffers); the second part has one thread per (env, agent) for the int64 actions / rewards / imposters / done rows,
// all in flattened order, so every access of the launch is a full-line coalesced stream.  k_replay_push (one thread per float,
// a 64-bit division each) is kept as the SUSNET_REPLAY_PUSH=v1 comparison path: 0.346 ms for 1 Mi cfg4 transitions.
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#include "../../include/susnet_b200.h"

extern "C" int sus_internal_fail(int code, const char* msg);
extern "C" void sus_internal_count_launch(void);

namespace {

__global__ void __launch_bounds__(256) k_replay_push(const __grid_constant__ SusReplayPush p) {
  const int64_t TS = (int64_t)p.T * p.S;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= p.N * TS) return;
  const int64_t e = gid / TS;
  const int j = (int)(gid - e * TS);
  const int t = j / p.S, k = j - t * p.S;
  const int64_t slot = ((p.idx_dev ? *p.idx_dev : p.idx) + e) % p.M;
  const bool finished = p.done[e] || p.truncated[e];
  const float cur = p.seq_in[e * TS + j];
  // np.roll(sequence, -1, axis=0); last row <- the new state (replay_memory.py:121-126)
  const float nxt = t < p.T - 1 ? p.seq_in[e * TS + j + p.S] : p.next_flat[e * p.S + k];
  p.states[slot * TS + j] = cur;
  p.next_states[slot * TS + j] = nxt;
  // the next step starts from the rolled sequence, or from T copies of the reset state (train.py:440-445)
  p.seq_out[e * TS + j] = finished ? p.cur_flat[e * p.S + k] : nxt;
  if (j < p.A) {
    long long a;
    if (p.actions_dtype == SUS_I32) a = static_cast<const int32_t*>(p.actions)[e * p.A + j];
    else if (p.actions_dtype == SUS_I64) a = static_cast<const long long*>(p.actions)[e * p.A + j];
    else a = static_cast<const uint8_t*>(p.actions)[e * p.A + j];
    p.r_actions[slot * p.A + j] = a;
    p.r_rewards[slot * p.A + j] = p.rewards[e * p.A + j];
  }
  if (j < p.n_imposters) p.r_imposters[slot * p.n_imposters + j] = p.imposters[e * p.n_imposters + j];
  if (j == 0) p.r_dones[slot] = p.done[e];  // `done` only: a truncated transition still bootstraps (train.py:396)
}

__device__ __forceinline__ bool aligned16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }

// x / d for x < total: a 32-bit division where the whole stream is shorter than 2^32 elements (the usual case)
__device__ __forceinline__ int64_t div_small(int64_t x, int64_t d, bool narrow) {
  return narrow ? (int64_t)((uint32_t)x / (uint32_t)d) : x / d;
}

template <bool T1>  // T == 1: next_flat / cur_flat are contiguous in x as well
__global__ void __launch_bounds__(256) k_replay_push_v(const __grid_constant__ SusReplayPush p, unsigned vec_blocks) {
  const int64_t TS = (int64_t)p.T * p.S;
  const int64_t total = p.N * TS;
  const int64_t idx = p.idx_dev ? *p.idx_dev : p.idx;
  if (blockIdx.x < vec_blocks) {
    // ---------------------------------------------------------------- the three float streams, four values per thread
    const int64_t x0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
    if (x0 >= total) return;
    const int n = total - x0 < 4 ? (int)(total - x0) : 4;
    const bool full = n == 4;
    const bool io_vec = full && aligned16(p.seq_in) && aligned16(p.seq_out);
    const bool flat_vec = T1 && full && aligned16(p.next_flat) && aligned16(p.cur_flat);
    const int64_t e0 = div_small(x0, TS, total <= 0xffffffffll);
    const int j0 = (int)(x0 - e0 * TS);
    float cur[4] = {0.f, 0.f, 0.f, 0.f}, nxt[4] = {0.f, 0.f, 0.f, 0.f}, nw[4];
    if (io_vec) {
      const float4 t = *reinterpret_cast<const float4*>(p.seq_in + x0);
      cur[0] = t.x; cur[1] = t.y; cur[2] = t.z; cur[3] = t.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) if (i < n) cur[i] = p.seq_in[x0 + i];
    }
    // np.roll(sequence, -1, axis=0); last row <- the new state (replay_memory.py:121-126)
    if (T1) {
      if (flat_vec) {
        const float4 t = *reinterpret_cast<const float4*>(p.next_flat + x0);
        nxt[0] = t.x; nxt[1] = t.y; nxt[2] = t.z; nxt[3] = t.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) if (i < n) nxt[i] = p.next_flat[x0 + i];
      }
    }
    bool fin[4] = {false, false, false, false};
    bool any_fin = false;
    {
      int64_t e = e0;
      int j = j0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (i >= n) break;
        fin[i] = p.done[e] || p.truncated[e];
        any_fin |= fin[i];
        if (!T1) nxt[i] = j < TS - p.S ? p.seq_in[x0 + i + p.S] : p.next_flat[e * p.S + (j - (int)(TS - p.S))];
        if (++j == TS) { j = 0; ++e; }
      }
    }
    // the next step starts from the rolled sequence, or from T copies of the reset state (train.py:440-445)
#pragma unroll
    for (int i = 0; i < 4; ++i) nw[i] = nxt[i];
    if (any_fin) {
      if (flat_vec) {
        const float4 t = *reinterpret_cast<const float4*>(p.cur_flat + x0);
        const float c[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) if (fin[i]) nw[i] = c[i];
      } else {
        int64_t e = e0;
        int j = j0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (i >= n) break;
          if (fin[i]) nw[i] = p.cur_flat[e * p.S + j % p.S];
          if (++j == TS) { j = 0; ++e; }
        }
      }
    }
    if (io_vec) {
      *reinterpret_cast<float4*>(p.seq_out + x0) = make_float4(nw[0], nw[1], nw[2], nw[3]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) if (i < n) p.seq_out[x0 + i] = nw[i];
    }
    // ring rows: slot (idx + e) mod M, i.e. x + idx * TS up to the wrap point and x - wrap_x behind it
    const int64_t wrap_x = (p.M - idx) * TS;
    const bool ring_vec = full && aligned16(p.states) && aligned16(p.next_states) && ((idx * TS) & 3) == 0 && ((p.M * TS) & 3) == 0;
    if (ring_vec && (x0 + 4 <= wrap_x || x0 >= wrap_x)) {
      const int64_t d = x0 >= wrap_x ? x0 - wrap_x : x0 + idx * TS;
      *reinterpret_cast<float4*>(p.states + d) = make_float4(cur[0], cur[1], cur[2], cur[3]);
      *reinterpret_cast<float4*>(p.next_states + d) = make_float4(nxt[0], nxt[1], nxt[2], nxt[3]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (i >= n) break;
        const int64_t x = x0 + i;
        const int64_t d = x >= wrap_x ? x - wrap_x : x + idx * TS;
        p.states[d] = cur[i];
        p.next_states[d] = nxt[i];
      }
    }
  } else {
    // ---------------------------------------------------------------- per-agent scalars, one thread per (env, agent)
    const int64_t y = (int64_t)(blockIdx.x - vec_blocks) * 256 + threadIdx.x;
    const int64_t NA = p.N * p.A;
    if (y >= NA) return;
    const bool narrow = NA <= 0xffffffffll;
    {
      const int64_t e = div_small(y, p.A, narrow);
      const int64_t d = y + (idx + e >= p.M ? idx - p.M : idx) * p.A;  // slot * A + agent
      long long a;
      if (p.actions_dtype == SUS_I32) a = static_cast<const int32_t*>(p.actions)[y];
      else if (p.actions_dtype == SUS_I64) a = static_cast<const long long*>(p.actions)[y];
      else a = static_cast<const uint8_t*>(p.actions)[y];
      p.r_actions[d] = a;
      p.r_rewards[d] = p.rewards[y];
    }
    if (y < p.N * p.n_imposters) {
      const int64_t e = div_small(y, p.n_imposters, narrow);
      p.r_imposters[y + (idx + e >= p.M ? idx - p.M : idx) * p.n_imposters] = p.imposters[y];
    }
    if (y < p.N) p.r_dones[idx + y >= p.M ? idx + y - p.M : idx + y] = p.done[y];  // `done` only (train.py:396)
  }
}

}  // namespace

extern "C" int sus_replay_push(const SusReplayPush* a, int device, void* stream) {
  if (!a) return sus_internal_fail(SUS_ERR_INVALID_ARGUMENT, "args is NULL");
  if (a->N < 0 || a->M <= 0 || a->T <= 0 || a->S <= 0 || a->A <= 0 || a->n_imposters <= 0 || a->idx < 0 || a->idx >= a->M)
    return sus_internal_fail(SUS_ERR_INVALID_ARGUMENT, "replay push: bad sizes");
  if (a->N > a->M) return sus_internal_fail(SUS_ERR_INVALID_ARGUMENT, "replay push: more transitions than ring slots");
  if (a->T * a->S < a->A || a->T * a->S < a->n_imposters)
    return sus_internal_fail(SUS_ERR_INVALID_ARGUMENT, "replay push: sequence block smaller than the action row");
  if (a->N == 0) return SUS_OK;
  // T == 1: every thread reads and writes its own elements of the sequence only, so it may be advanced in place
  if (!a->seq_in || !a->seq_out || (a->seq_in == a->seq_out && a->T != 1) || !a->next_flat || !a->cur_flat || !a->actions || !a->rewards ||
      !a->done || !a->truncated || !a->imposters || !a->states || !a->r_actions || !a->r_rewards || !a->next_states ||
      !a->r_dones || !a->r_imposters)
    return sus_internal_fail(SUS_ERR_INVALID_ARGUMENT, "replay push: NULL or aliased buffer");
  int prev = -1;
  cudaGetDevice(&prev);
  if (prev != device) cudaSetDevice(device);
  const int64_t total = a->N * a->T * a->S;
  const char* ver = getenv("SUSNET_REPLAY_PUSH");  // read per call: tests and tools compare the two kernels inside one process
  const bool use_v1 = ver && ver[0] == 'v' && ver[1] == '1';
  if (use_v1) {
    k_replay_push<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*a);
  } else {
    const unsigned vec_blocks = (unsigned)(((total + 3) / 4 + 255) / 256);
    const unsigned row_blocks = (unsigned)((a->N * a->A + 255) / 256);
    if (a->T == 1) k_replay_push_v<true><<<vec_blocks + row_blocks, 256, 0, (cudaStream_t)stream>>>(*a, vec_blocks);
    else k_replay_push_v<false><<<vec_blocks + row_blocks, 256, 0, (cudaStream_t)stream>>>(*a, vec_blocks);
  }
  sus_internal_count_launch();
  const cudaError_t err = cudaGetLastError();
  if (prev != device && prev >= 0) cudaSetDevice(prev);
  if (err != cudaSuccess) return sus_internal_fail(SUS_ERR_CUDA, cudaGetErrorString(err));
  return SUS_OK;
}
