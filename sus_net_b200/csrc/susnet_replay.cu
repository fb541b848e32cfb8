// susnet_replay.cu -- row (f1): the reference's replay layout (src/replay_memory.py:33-44) filled on the GPU.
//
// One thread per (env, element of the T x S sequence block): reads the env's running sequence once and writes the
// `states` row, the rolled `next_states` row and the sequence the next step starts from, all coalesced along the
// flattened (t, s) index.  The per-transition scalars (actions -> int64, rewards, done, imposters) are written by
// the first threads of each env's block.
#include <cstdint>
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

}  // namespace

extern "C" int sus_replay_push(const SusReplayPush* a, int device, void* stream) {
  if (!a) return sus_internal_fail(SUS_ERR_INVALID_ARGUMENT, "args is NULL");
  if (a->N < 0 || a->M <= 0 || a->T <= 0 || a->S <= 0 || a->A <= 0 || a->n_imposters <= 0 || a->idx < 0 || a->idx >= a->M)
    return sus_internal_fail(SUS_ERR_INVALID_ARGUMENT, "replay push: bad sizes");
  if (a->N > a->M) return sus_internal_fail(SUS_ERR_INVALID_ARGUMENT, "replay push: more transitions than ring slots");
  if (a->T * a->S < a->A || a->T * a->S < a->n_imposters)
    return sus_internal_fail(SUS_ERR_INVALID_ARGUMENT, "replay push: sequence block smaller than the action row");
  if (a->N == 0) return SUS_OK;
  if (!a->seq_in || !a->seq_out || a->seq_in == a->seq_out || !a->next_flat || !a->cur_flat || !a->actions || !a->rewards ||
      !a->done || !a->truncated || !a->imposters || !a->states || !a->r_actions || !a->r_rewards || !a->next_states ||
      !a->r_dones || !a->r_imposters)
    return sus_internal_fail(SUS_ERR_INVALID_ARGUMENT, "replay push: NULL or aliased buffer");
  int prev = -1;
  cudaGetDevice(&prev);
  if (prev != device) cudaSetDevice(device);
  const int64_t total = a->N * a->T * a->S;
  k_replay_push<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*a);
  sus_internal_count_launch();
  const cudaError_t err = cudaGetLastError();
  if (prev != device && prev >= 0) cudaSetDevice(prev);
  if (err != cudaSuccess) return sus_internal_fail(SUS_ERR_CUDA, cudaGetErrorString(err));
  return SUS_OK;
}
