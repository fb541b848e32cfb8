// susnet_policy.cu -- row (f2): the acting part of train() (src/train.py:349-381) for every env at once, and the
// T-deep feature sequences the acting / training networks read (train.py:318-322,388-389,440-445).
//
//   k_select_actions   epsilon-greedy over Q-values the caller's networks produced: per agent view, an alive imposter
//                      explores with probability eps (uniform over its role list) or takes argmax Q_imposter, an alive
//                      crew member likewise with Q_crew, a dead agent keeps action 0.  One thread per env; the explore /
//                      random-action draws are Philox words keyed (seed, global env id, act epoch, P_POLICY), so acting is
//                      reproducible, independent of sharding and replayable inside a CUDA graph (device-resident ticks).
//   k_seq_roll         np.roll(sequence, -1, axis=0) with the newest item in the last row, or T copies of the newest item
//                      where the episode just ended -- for ENCODED features, so that a T > 1 loop encodes only the newest
//                      time step (the fused step kernel's output) instead of re-encoding all T states every iteration.
#include <cstdint>
#include <cuda_runtime.h>

#include "susnet_device.cuh"

extern "C" int sus_internal_fail(int code, const char* msg);
extern "C" void sus_internal_count_launch(void);
extern "C" int sus_internal_policy_params(sus_env_t env, susnet::DevConfig* c, susnet::StateArrays* st, uint64_t* tick,
                                          uint64_t** tick_dev, unsigned int** tick_ctr, int64_t* N, int* device);

using namespace susnet;

namespace {

constexpr uint32_t P_POLICY = 5;  // oracle/rng_spec.py: slot 2i = explore word of agent i, slot 2i + 1 = its random action

struct PolicyParams {
  DevConfig c;
  StateArrays st;
  const float* q_imp;
  const float* q_crew;
  const float* eps_dev;
  float eps_value;
  int32_t imp_per_view, actions_dtype;
  void* actions;
  uint64_t tick;
  uint64_t* tick_dev;
  unsigned int* tick_ctr;
  int64_t N;
};

__device__ __forceinline__ uint32_t role_actions(const DevConfig& c, uint32_t is_imp) {
  if (c.variant == SUS_VARIANT_TRAINING_GROUND) return 5u + is_imp;
  const uint32_t base = 6u + is_imp;
  return c.variant == SUS_VARIANT_TAGGING ? base + (uint32_t)c.A - 1u : base;
}

// torch.argmax: index of the FIRST maximal value (train.py:368-370,379-381)
__device__ __forceinline__ uint32_t argmax_row(const float* __restrict__ q, uint32_t n) {
  uint32_t best = 0;
  float bv = q[0];
  for (uint32_t j = 1; j < n; ++j) {
    const float v = q[j];
    if (v > bv) { bv = v; best = j; }
  }
  return best;
}

__global__ void __launch_bounds__(256) k_select_actions(const __grid_constant__ PolicyParams p) {
  __shared__ uint64_t tick_s;
  if (threadIdx.x == 0) tick_s = fetch_launch_tick(p.tick, p.tick_dev, p.tick_ctr, 1);
  __syncthreads();
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= p.N) return;
  const DevConfig& c = p.c;
  const int A = c.A;
  const uint32_t aux = p.st.aux[e].x;
  const uint32_t alive = aux & 0xff, imp = (aux >> 8) & 0xff;
  const double eps = (double)(p.eps_dev ? *p.eps_dev : p.eps_value);
  const uint32_t nia = role_actions(c, 1u), nca = role_actions(c, 0u);
  WordStream ws;
  ws.init(c, nullptr, (uint32_t)e, tick_s, P_POLICY);
  for (int i = 0; i < A; ++i) {
    uint32_t a = 0;
    if ((alive >> i) & 1u) {  // train.py:361,373: dead agents keep agent_actions[i] = 0
      const uint32_t is_imp = (imp >> i) & 1u;
      const uint32_t n = is_imp ? nia : nca;
      const float* q = nullptr;
      if (is_imp) {
        if (p.q_imp) q = p.imp_per_view ? p.q_imp + ((int64_t)i * p.N + e) * nia : p.q_imp + e * nia;
      } else if (p.q_crew) {
        q = p.q_crew + ((int64_t)i * p.N + e) * nca;
      }
      // np.random.random() <= eps (train.py:363,374); a team without a network acts uniformly (RandomEquiprobable)
      const bool explore = q == nullptr || (double)ws.word(2 * i) * 2.3283064365386963e-10 <= eps;
      a = explore ? bounded(ws.word(2 * i + 1), n) : argmax_row(q, n);
    }
    if (p.actions_dtype == SUS_U8) static_cast<uint8_t*>(p.actions)[e * A + i] = (uint8_t)a;
    else static_cast<int32_t*>(p.actions)[e * A + i] = (int32_t)a;
  }
}

struct RollParams {
  const float* in;
  float* out;
  const float* newest;
  const uint8_t* done;
  const uint8_t* trunc;
  int64_t rows, n_envs;
  int32_t T, R;
};

// one thread per 4 floats of a (row, t) item where R % 4 == 0 and the pointers are 16-byte aligned (VEC), else per float
template <bool VEC>
__global__ void __launch_bounds__(256) k_seq_roll(const __grid_constant__ RollParams p) {
  const int64_t W = VEC ? p.R / 4 : p.R;  // elements per (row, t) item
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= p.rows * p.T * W) return;
  const int64_t row = gid / (p.T * W);
  const int64_t rem = gid - row * p.T * W;
  const int t = (int)(rem / W);
  const int64_t k = rem - (int64_t)t * W;
  const int64_t env = row % p.n_envs;
  const bool finished = p.done[env] || p.trunc[env];
  const bool take_new = finished || t == p.T - 1;
  if (VEC) {
    const float4* src = take_new ? reinterpret_cast<const float4*>(p.newest) + row * W + k
                                 : reinterpret_cast<const float4*>(p.in) + (row * p.T + t + 1) * W + k;
    reinterpret_cast<float4*>(p.out)[(row * p.T + t) * W + k] = *src;
  } else {
    p.out[(row * p.T + t) * W + k] = take_new ? p.newest[row * W + k] : p.in[(row * p.T + t + 1) * W + k];
  }
}

}  // namespace

extern "C" int sus_env_select_actions(sus_env_t env, const SusPolicyIO* io, void* stream) {
  if (!env || !io) return sus_internal_fail(SUS_ERR_INVALID_ARGUMENT, "NULL argument");
  if (io->actions_dtype != SUS_I32 && io->actions_dtype != SUS_U8)
    return sus_internal_fail(SUS_ERR_INVALID_ARGUMENT, "select_actions: actions_dtype must be SUS_I32 or SUS_U8");
  PolicyParams p;
  int device = 0;
  if (int rc = sus_internal_policy_params(env, &p.c, &p.st, &p.tick, &p.tick_dev, &p.tick_ctr, &p.N, &device)) return rc;
  if (io->q_imposter && !io->imposter_per_view && p.c.nI != 1)
    return sus_internal_fail(SUS_ERR_INVALID_ARGUMENT, "select_actions: the [N][n_actions] imposter layout needs n_imposters == 1");
  if (p.N > 0 && !io->actions) return sus_internal_fail(SUS_ERR_INVALID_ARGUMENT, "select_actions: actions is NULL");
  p.q_imp = io->q_imposter; p.q_crew = io->q_crew; p.eps_dev = io->eps; p.eps_value = io->eps_value;
  p.imp_per_view = io->imposter_per_view; p.actions_dtype = io->actions_dtype; p.actions = io->actions;
  int prev = -1;
  cudaGetDevice(&prev);
  if (prev != device) cudaSetDevice(device);
  cudaError_t err = cudaSuccess;
  if (p.N > 0) {
    k_select_actions<<<(unsigned)((p.N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p);
    sus_internal_count_launch();
    err = cudaGetLastError();
  }
  if (prev != device && prev >= 0) cudaSetDevice(prev);
  if (err != cudaSuccess) return sus_internal_fail(SUS_ERR_CUDA, cudaGetErrorString(err));
  return SUS_OK;
}

extern "C" int sus_seq_roll(const float* seq_in, float* seq_out, const float* newest, const uint8_t* done,
                            const uint8_t* truncated, int64_t rows, int64_t n_envs, int32_t T, int32_t R, int device,
                            void* stream) {
  if (rows < 0 || n_envs <= 0 || T <= 0 || R <= 0 || rows % n_envs != 0)
    return sus_internal_fail(SUS_ERR_INVALID_ARGUMENT, "seq_roll: bad sizes (rows must be a multiple of n_envs)");
  if (rows == 0) return SUS_OK;
  if (!seq_in || !seq_out || seq_in == seq_out || !newest || !done || !truncated)
    return sus_internal_fail(SUS_ERR_INVALID_ARGUMENT, "seq_roll: NULL or aliased buffer");
  RollParams p = {seq_in, seq_out, newest, done, truncated, rows, n_envs, T, R};
  int prev = -1;
  cudaGetDevice(&prev);
  if (prev != device) cudaSetDevice(device);
  const bool vec = R % 4 == 0 && ((reinterpret_cast<uintptr_t>(seq_in) | reinterpret_cast<uintptr_t>(seq_out) |
                                   reinterpret_cast<uintptr_t>(newest)) & 15u) == 0;
  const int64_t total = rows * T * (vec ? R / 4 : R);
  if (vec) k_seq_roll<true><<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p);
  else k_seq_roll<false><<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p);
  sus_internal_count_launch();
  const cudaError_t err = cudaGetLastError();
  if (prev != device && prev >= 0) cudaSetDevice(prev);
  if (err != cudaSuccess) return sus_internal_fail(SUS_ERR_CUDA, cudaGetErrorString(err));
  return SUS_OK;
}
