// susnet_api.cu -- kernels and C ABI (include/susnet_b200.h) of the B200-native Sus-Net simulator.
//
// Kernels (all sm_100a; one thread owns one env / item; the wall grid is staged in shared memory):
//   K0    k_reset              FourRoomEnv.reset for a masked subset of envs
//   K1    k_step<V, false>     the step: action decode, ordered per-agent loop (move with wall collision, kill, fix,
//                              sabotage, tag), vote tally, win / reward / done / truncation, episode-stat flush,
//                              auto-reset.  One thread per env, 256-thread CTAs (step-only launches).
//   K1+K2 k_step_ws<V>         the step FUSED with the observation encode of the state the next action is taken
//                              from, warp-specialised: compute warps + one TMA emitter warp per persistent CTA
//                              (Global / Perspective; the headline kernel)
//         k_step_flat<V>       same fusion for Flat encodes: one thread per env at 32 warps / SM, rows staged as one byte
//                              per value in shared memory and expanded with lane-contiguous 128-bit stores
//         k_step_tma<V, true>  same fusion, every warp stages and bulk-stores its own tiles (SUSNET_PATH=tma)
//         k_step<V, true>      same fusion with direct register stores (fallback, SUSNET_PATH=direct)
//   K2    k_encode_ws / k_encode_flat / k_encode_tma / k_encode_env / k_encode_rows<T>
//                              featurizers on live env state / on (B*T, S) flattened rows (= featurizer.fit)
//   K3    k_sample_actions     role-aware uniform random actions
//         k_rollout<V>         n random-policy steps per launch with the env state in registers
//         device-resident launch ticks (sus_env_device_ticks, fetch_launch_tick in susnet_device.cuh) make the
//         launches above replayable inside CUDA graphs
//   plus small export / import kernels for the reference's flatten order, k_replay_push (susnet_replay.cu) and the
//   L2-compressible allocator the feature tensors live in (susnet_alloc.cu).
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "susnet_device.cuh"
#include "susnet_encode.cuh"
#include "susnet_tile.cuh"
#include "susnet_ws.cuh"

using namespace susnet;

namespace {

constexpr int kThreads = 256;
constexpr int kTmaMaxWarps = 16;  // per-warp TMA path (small staging blocks, e.g. Flat rows): up to 512 threads per CTA
constexpr int kWsMaxWarps = 12;   // warp-specialised path: up to 11 compute warps + the emitter
constexpr unsigned kFull = 0xffffffffu;

thread_local std::string g_last_error;
std::atomic<long long> g_launches{0};

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

#define SUS_CUDA(expr)                                                                                  \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess) return fail(SUS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    else if (prev == dev) prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// ------------------------------------------------------------------------------------------ params
struct StepParams {
  DevConfig c;
  DevEncode enc;
  StateArrays st;
  const void* actions;
  void* rewards;
  uint8_t* done;
  uint8_t* trunc;
  int32_t* actions_out;
  float* next_flat;
  long long* metrics;
  int16_t* imposters;
  void* spatial;  // float planes, or uint8_t planes with SUS_ENCODE_PLANES_U8
  float* non_spatial;
  const uint32_t* inj_step;
  const uint32_t* inj_reset;
  const uint32_t* inj_act;
  unsigned long long* stats;
  uint32_t* err;
  double* ret;       // [A][N] running return G of every agent (train.py:324,386), or nullptr when not tracked
  double* ret_sums;  // [2] sums over finished episodes of mean imposter / mean crew return (train.py:421-424)
  double gamma;
  uint64_t tick;
  uint64_t* tick_dev;        // device-resident tick (sus_env_device_ticks) or nullptr: use `tick`
  unsigned int* tick_ctr;    // CTAs of this launch that have read it (stage_tables_and_tick)
  int64_t N;
  int32_t actions_dtype, rewards_dtype;
  // compact host protocol (SusCompactLayout): packed result records instead of rewards / done / trunc
  uint8_t* packed_out;
  int32_t action_bits, action_bytes, reward_bits, result_bytes;
};

struct ResetParams {
  DevConfig c;
  StateArrays st;
  const uint8_t* mask;
  const uint32_t* inj_reset;
  uint64_t tick;
  uint64_t* tick_dev;
  unsigned int* tick_ctr;
  int64_t N;
};

struct EncodeParams {
  DevConfig c;
  DevEncode enc;
  StateArrays st;      // k_encode_env
  const void* rows;    // k_encode_rows
  void* spatial;
  float* non_spatial;
  int64_t n_items;
};

struct ExportParams {
  DevConfig c;
  StateArrays st;
  void* out;
  int64_t N;
};

struct ImportParams {
  DevConfig c;
  StateArrays st;
  const long long* flat;
  const uint8_t* imp;
  const int32_t* t;
  int64_t N;
};

struct ActParams {
  DevConfig c;
  StateArrays st;
  int32_t* out;
  const uint32_t* inj_act;
  uint64_t tick;
  uint64_t* tick_dev;
  unsigned int* tick_ctr;
  int64_t N;
};

// ------------------------------------------------------------------------------------------ kernels
// N == 0 / zero-step calls still consume ticks: advance the device-resident counter without a kernel body to do it
__global__ void k_advance_tick(uint64_t* tick, uint64_t n) { *tick += n; }

__device__ __forceinline__ uint32_t role_actions_rt(const DevConfig& c, uint32_t is_imp) {
  if (c.variant == SUS_VARIANT_TRAINING_GROUND) return 5u + is_imp;
  const uint32_t base = 6u + is_imp;
  return c.variant == SUS_VARIANT_TAGGING ? base + (uint32_t)c.A - 1u : base;
}

__global__ void __launch_bounds__(kThreads) k_reset(const __grid_constant__ ResetParams p) {
  __shared__ GridTables tb;
  stage_tables_and_tick(p.c, tb, p.tick, p.tick_dev, p.tick_ctr);
  const int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (e >= p.N) return;
  if (p.mask && !p.mask[e]) return;
  EnvState s;
  WordStream ws;
  ws.init(p.c, p.inj_reset ? p.inj_reset + e * (p.c.nI + p.c.A + p.c.J) : nullptr, (uint32_t)e, tb.tick, P_RESET);
  reset_env(p.c, tb, s, ws);
  store_state(p.st, e, s, true);
}

// the per-env "reward row" of a launch: float rewards [A] (f32 / f64) or one packed result record (compact host protocol)
__device__ __forceinline__ int reward_row_bytes(const StepParams& p) {
  return p.packed_out ? p.result_bytes : p.c.A * (p.rewards_dtype == SUS_F64 ? 8 : 4);
}
__device__ __forceinline__ uint8_t* reward_rows(const StepParams& p) {
  return p.packed_out ? p.packed_out : static_cast<uint8_t*>(p.rewards);
}

// ---- everything a step reads for one env, as raw words: loaded one group ahead by the TMA kernel so the DRAM
// latency of the state records and the action rows hides behind the previous group's work
struct StepInput {
  uint64_t pos, jobpos;
  uint4 aux, met;
  uint32_t raw[SUS_MAX_AGENTS];  // action words exactly as loaded (no arithmetic here: consumers would stall on the loads)
  uint32_t oob64;                // int64 actions only: a high word was non-zero
};

__device__ __forceinline__ void load_input(const StepParams& p, int64_t e, bool have, StepInput& in) {
  in.pos = in.jobpos = 0; in.oob64 = 0;
  in.aux = in.met = make_uint4(0, 0, 0, 0);
#pragma unroll
  for (int i = 0; i < SUS_MAX_AGENTS; ++i) in.raw[i] = 0;
  if (!have) return;
  const int A = p.c.A;
  in.pos = p.st.pos[e]; in.jobpos = p.st.jobpos[e]; in.aux = p.st.aux[e]; in.met = p.st.met[e];
  if (p.actions == nullptr) return;
  if (p.actions_dtype == SUS_I32) {
    const uint32_t* a = static_cast<const uint32_t*>(p.actions) + e * A;
#pragma unroll
    for (int i = 0; i < SUS_MAX_AGENTS; ++i)
      if (i < A) in.raw[i] = a[i];
  } else if (p.actions_dtype == SUS_U8) {
    const uint8_t* a = static_cast<const uint8_t*>(p.actions) + e * A;
#pragma unroll
    for (int i = 0; i < SUS_MAX_AGENTS; ++i)
      if (i < A) in.raw[i] = a[i];
  } else if (p.actions_dtype == SUS_PACKED) {  // the record's bytes, still packed (<= 4 bytes: 8 agents x 4 bits)
    const uint8_t* a = static_cast<const uint8_t*>(p.actions) + e * p.action_bytes;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (i < p.action_bytes) in.raw[i] = a[i];
  } else {
    const unsigned long long* a = static_cast<const unsigned long long*>(p.actions) + e * A;
#pragma unroll
    for (int i = 0; i < SUS_MAX_AGENTS; ++i)
      if (i < A) { const unsigned long long v = a[i]; in.raw[i] = (uint32_t)v; in.oob64 |= (uint32_t)(v >> 32); }
  }
}

// one byte per agent + "some action was negative or >= 256"
__device__ __forceinline__ uint64_t pack_actions(const StepParams& p, const StepInput& in, int A, bool& in_range) {
  uint64_t acts = 0;
  if (p.actions_dtype == SUS_PACKED) {
    const uint32_t rec = in.raw[0] | (in.raw[1] << 8) | (in.raw[2] << 16) | (in.raw[3] << 24);
    const uint32_t m = (1u << p.action_bits) - 1u;
#pragma unroll
    for (int i = 0; i < SUS_MAX_AGENTS; ++i)
      if (i < A) acts |= (uint64_t)((rec >> (i * p.action_bits)) & m) << (8 * i);
    in_range = true;
    return acts;
  }
  uint32_t oob = in.oob64;
#pragma unroll
  for (int i = 0; i < SUS_MAX_AGENTS; ++i)
    if (i < A) { oob |= in.raw[i] & ~0xffu; acts |= (uint64_t)(in.raw[i] & 0xffu) << (8 * i); }
  in_range = oob == 0;
  return acts;
}

__device__ __forceinline__ void unpack_state(const StepInput& in, EnvState& s) {
  s.pos = in.pos; s.jobpos = in.jobpos;
  s.alive = in.aux.x & 0xff; s.imp = (in.aux.x >> 8) & 0xff; s.jobdone = (in.aux.x >> 16) & 0xff; s.used = in.aux.x >> 24;
  s.nsteps = in.aux.y; s.timer = in.aux.z; s.tagcnt = in.aux.w;
  s.completed = in.met.x; s.sabotaged = in.met.y; s.misc = in.met.z;
}

// ---- the step of one env, shared by the direct-store kernel (k_step) and the TMA-staged kernel (k_step_tma).
// rew_row / nf_row point at THIS env's reward / next_flat row: in global memory (direct) or in the warp's
// shared-memory staging block (TMA path).
template <int VARIANT, int TA = 0, int TJ = 0>
__device__ __forceinline__ void step_one(const StepParams& p, const GridTables& tb, int64_t e, bool have,
                                         const StepInput& in, void* rew_row, float* nf_row, EnvState& s, StepResult& r,
                                         bool& stepped, bool& finished) {
  const DevConfig& c = p.c;
  const int A = TA ? TA : c.A;
  stepped = false;
  finished = false;
  if (!have) return;
  unpack_state(in, s);
  // ---- actions: role-list indices, one byte per agent
  bool ok = true;
  uint64_t acts = pack_actions(p, in, A, ok);
  if (p.actions == nullptr) {  // fused random policy == env.step(env.sample_actions()), base.py:326-330
    WordStream wa;
    wa.init(c, p.inj_act ? p.inj_act + e * A : nullptr, (uint32_t)e, tb.tick, P_ACT_FUSED);
    for (int i = 0; i < A; ++i)
      acts |= (uint64_t)bounded(wa.word(i), n_role_actions<VARIANT>(c, (s.imp >> i) & 1u)) << (8 * i);
  } else {
    for (int i = 0; i < A; ++i)
      if (get_byte(acts, i) >= n_role_actions<VARIANT>(c, (s.imp >> i) & 1u)) ok = false;
    if (!ok) acts = 0;
  }
  if (p.actions_out)
    for (int i = 0; i < A; ++i) p.actions_out[e * A + i] = (int32_t)get_byte(acts, i);
  r = StepResult{};
  if (!ok) {
    // reference: IndexError (base.py:381).  Here the env is left untouched, the error is counted (sus_env_check_actions)
    // and the step's outputs are DEFINED: NaN rewards (the all-ones reward code), done = truncated = 0, next_flat = the
    // unchanged state
    atomicAdd(p.err, 1u);
  } else {
    WordStream ws;
    ws.init(c, p.inj_step ? p.inj_step + e * (2 * A - 1) : nullptr, (uint32_t)e, tb.tick, P_STEP);
    step_env<VARIANT, TA, TJ>(c, tb, s, acts, ws, r);
    stepped = true;
    finished = r.done || r.trunc;
  }
  r.ret_imp = r.ret_crew = 0.0;
  if (p.packed_out && rew_row) {  // compact host protocol: reward codes + done / truncated bits in one record
    const int rb = p.reward_bits;
    uint64_t rec = 0;
    for (int i = 0; i < A; ++i)
      rec |= (uint64_t)(ok ? agent_reward_code(c, s, r, i) : (1u << rb) - 1u) << (i * rb);
    rec |= (uint64_t)(r.done ? 1u : 0u) << (A * rb);
    rec |= (uint64_t)(r.trunc ? 1u : 0u) << (A * rb + 1);
    uint8_t* o = static_cast<uint8_t*>(rew_row);
    for (int b = 0; b < p.result_bytes; ++b) o[b] = (uint8_t)(rec >> (8 * b));
  }
  const bool float_rewards = rew_row && !p.packed_out;
  if (float_rewards || (p.ret && ok)) {
    double g_imp = 0.0, g_crew = 0.0;
    for (int i = 0; i < A; ++i) {
      const double v = ok ? agent_reward<VARIANT>(c, s, r, i) : __longlong_as_double(0x7ff8000000000000ll);
      if (float_rewards) {
        if (p.rewards_dtype == SUS_F64) static_cast<double*>(rew_row)[i] = v;
        else static_cast<float*>(rew_row)[i] = (float)v;
      }
      if (p.ret && ok) {  // G = reward + gamma * G (train.py:386); reset to 0 when the episode ends (train.py:436)
        const double g = __dadd_rn(v, __dmul_rn(p.gamma, p.ret[(int64_t)i * p.N + e]));  // two roundings like numpy, no FMA
        if ((s.imp >> i) & 1u) g_imp += g; else g_crew += g;
        p.ret[(int64_t)i * p.N + e] = finished ? 0.0 : g;
      }
    }
    if (p.ret && finished) {  // G[imposter_mask].mean(), G[~imposter_mask].mean() (train.py:421-422)
      r.ret_imp = g_imp / (double)c.nI;
      r.ret_crew = g_crew / (double)(A - c.nI);
    }
  }
  if (p.done) p.done[e] = r.done;
  if (p.trunc) p.trunc[e] = r.trunc;
  if (nf_row) write_flat<float>(c, s, nf_row);
  if (p.imposters) {  // ascending ids = env.imposter_idxs of this episode (before any auto-reset)
    int k = 0;
    for (int i = 0; i < A; ++i)
      if ((s.imp >> i) & 1u) p.imposters[e * c.nI + k++] = (int16_t)i;
  }
  if (p.metrics) {
    long long* m = p.metrics + e * SUS_N_METRICS;
    m[SUS_M_TOTAL_TIME_STEPS] = s.nsteps; m[SUS_M_IMP_KILLED_CREW] = s.misc & 0xff;
    m[SUS_M_COMPLETED_JOBS] = s.completed; m[SUS_M_SABOTAGED_JOBS] = s.sabotaged;
    m[SUS_M_IMP_VOTED_OUT] = (s.misc >> 8) & 0xff; m[SUS_M_CREW_VOTED_OUT] = (s.misc >> 16) & 0xff;
    m[SUS_M_CREW_WON] = (s.misc >> 24) & 1u; m[SUS_M_IMPOSTER_WON] = (s.misc >> 25) & 1u;
  }
}

// ---- episode statistics, auto-reset and state write-back; all 32 lanes of the warp must call this together
__device__ __forceinline__ void finish_one(const StepParams& p, const GridTables& tb, int64_t e, int lane, EnvState& s,
                                           const StepResult& r, bool stepped, bool finished) {
  const DevConfig& c = p.c;
  if (__any_sync(kFull, finished)) {  // warp reduction, one atomic per statistic per warp
    const uint32_t f = finished ? 1u : 0u;
    uint32_t v[SUS_N_STATS];
    v[SUS_S_EPISODES] = f; v[SUS_S_CREW_WON] = f * ((s.misc >> 24) & 1u); v[SUS_S_IMPOSTER_WON] = f * ((s.misc >> 25) & 1u);
    v[SUS_S_IMP_KILLED_CREW] = f * (s.misc & 0xff); v[SUS_S_COMPLETED_JOBS] = f * s.completed;
    v[SUS_S_SABOTAGED_JOBS] = f * s.sabotaged; v[SUS_S_IMP_VOTED_OUT] = f * ((s.misc >> 8) & 0xff);
    v[SUS_S_CREW_VOTED_OUT] = f * ((s.misc >> 16) & 0xff); v[SUS_S_TOTAL_TIME_STEPS] = f * s.nsteps;
    v[SUS_S_TRUNCATED] = (finished && r.trunc) ? 1u : 0u;
#pragma unroll
    for (int k = 0; k < SUS_N_STATS; ++k) {
      // 32 lanes x < 2^26 each cannot overflow 32 bits for any sane episode length
      const uint32_t sum = __reduce_add_sync(kFull, v[k]);
      if (lane == 0 && sum) atomicAdd(p.stats + k, (unsigned long long)sum);
    }
    if (p.ret) {
      double a = r.ret_imp, b = r.ret_crew;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(kFull, a, o); b += __shfl_xor_sync(kFull, b, o); }
      if (lane == 0) { atomicAdd(p.ret_sums, a); atomicAdd(p.ret_sums + 1, b); }
    }
  }
  if (stepped) {
    const bool restart = finished && c.auto_reset;  // SURVEY.md A.7; train.py:419-445 does this on the host
    if (restart) {
      WordStream wr;
      wr.init(c, p.inj_reset ? p.inj_reset + e * (c.nI + c.A + c.J) : nullptr, (uint32_t)e, tb.tick, P_AUTORESET);
      reset_env(c, tb, s, wr);
    }
    store_state(p.st, e, s, restart);  // one copy of the store sequence for both outcomes (the job cells only after a reset)
  }
}

// K1 (+K2), direct-store path: one thread per env, outputs written straight from registers.  Used for ragged or
// unaligned outputs and as the fallback when the staging tiles do not fit in shared memory.
template <int VARIANT, bool ENCODE, int TA = 0, int TJ = 0>
__global__ void __launch_bounds__(kThreads) k_step(const __grid_constant__ StepParams p) {
  __shared__ GridTables tb;
  stage_tables_and_tick(p.c, tb, p.tick, p.tick_dev, p.tick_ctr);
  const int lane = threadIdx.x & 31;
  const int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const int64_t e0 = e - lane;
  const bool have = e < p.N;
  bool stepped, finished;
  EnvState s = {};
  StepResult r = {};
  StepInput in;
  load_input(p, e, have, in);
  step_one<VARIANT, TA, TJ>(p, tb, e, have, in, reward_rows(p) ? reward_rows(p) + e * reward_row_bytes(p) : nullptr,
                    p.next_flat ? p.next_flat + e * p.c.S : nullptr, s, r, stepped, finished);
  finish_one(p, tb, e, lane, s, r, stepped, finished);
  if (ENCODE) {
    const int64_t rem = p.N - e0;
    warp_encode(p.c, p.enc, tb, obs_of(s), e0, rem < 32 ? (int)(rem < 0 ? 0 : rem) : 32, have, p.N, p.spatial, p.non_spatial);
  }
}

// Per-warp staging block of k_step_flat (bytes; offsets are multiples of 16)
struct FlatStage {
  int32_t row_bytes;  // 32 rows x F bytes, rounded up to 16
  int32_t rew_off, nf_off, per_warp;
};

// Copy `n_words` 32-bit words from a warp's staging block to global memory with lane-contiguous stores
// (128-bit when both sides are 16-byte aligned).  All 32 lanes.
__device__ __forceinline__ void warp_copy_words(uint32_t* __restrict__ g, const uint32_t* __restrict__ s, int n_words, int lane) {
  int done = 0;
  if (((uint32_t)reinterpret_cast<uintptr_t>(g) & 15u) == 0) {
    const int n4 = n_words >> 2;
    for (int i = lane; i < n4; i += 32) reinterpret_cast<uint4*>(g)[i] = reinterpret_cast<const uint4*>(s)[i];
    done = n4 << 2;
  }
  for (int i = done + lane; i < n_words; i += 32) g[i] = s[i];
}

// Copy `n_bytes` from a warp's staging block to global memory: words where both the destination and the size allow,
// single bytes over a ragged tail / odd destination (packed result records of a ragged or unaligned batch).  All 32 lanes.
__device__ __forceinline__ void warp_copy_bytes(uint8_t* __restrict__ g, const uint8_t* __restrict__ s, int n_bytes, int lane) {
  if (((uint32_t)reinterpret_cast<uintptr_t>(g) & 3u) == 0) {
    warp_copy_words(reinterpret_cast<uint32_t*>(g), reinterpret_cast<const uint32_t*>(s), n_bytes >> 2, lane);
    for (int i = (n_bytes & ~3) + lane; i < n_bytes; i += 32) g[i] = s[i];
  } else {
    for (int i = lane; i < n_bytes; i += 32) g[i] = s[i];
  }
}

// Expand a warp's staged byte rows to `n` floats at `out`: word i of the block holds floats [4i, 4i + 4).  Lane-contiguous
// 128-bit stores when `out` is 16-byte aligned, scalar stores otherwise and over a ragged tail.  All 32 lanes, after the
// rows are complete (__syncwarp / any warp collective).
__device__ __forceinline__ void warp_expand_byte_rows(float* __restrict__ out, const uint32_t* __restrict__ w, int n, int lane) {
  int done = 0;
  if (((uint32_t)reinterpret_cast<uintptr_t>(out) & 15u) == 0) {
    const int n4 = n >> 2;
    // words loaded before the first is expanded: the LDS latency is paid once per batch.  (A variant whose full rounds run
    // without the per-word bounds test -- 11 instead of 20 instructions per float4 -- measured SLOWER, 0.0847 against 0.0820 ms
    // at 1 Mi envs: the expansion phase is bound by the store stream, not by issue slots.)
    constexpr int kBatch = 8;
    for (int i0 = lane; i0 < n4; i0 += 32 * kBatch) {
      uint32_t x[kBatch];
#pragma unroll
      for (int b = 0; b < kBatch; ++b) x[b] = i0 + 32 * b < n4 ? w[i0 + 32 * b] : 0u;
#pragma unroll
      for (int b = 0; b < kBatch; ++b)
        if (i0 + 32 * b < n4)
          reinterpret_cast<float4*>(out)[i0 + 32 * b] = make_float4(byte_row_value(x[b], 0), byte_row_value(x[b], 1),
                                                                    byte_row_value(x[b], 2), byte_row_value(x[b], 3));
    }
    done = n4 << 2;
  }
  for (int i = done + lane; i < n; i += 32) out[i] = byte_row_value(w[i >> 2], i & 3);
}

// K1+K2 for FLAT encodes (the reference's training recipes: one-hot positions + crew flags, component.py): one thread
// per env at high occupancy.  A flat row is F mostly-zero small integers, so each warp stages its 32 rows as ONE BYTE
// per value (prefilled with the byte of 0; the one-hot segments only set their ones), then all lanes expand the
// 32 x F bytes to floats with lane-contiguous 128-bit stores.  32 F bytes of shared memory per warp instead of the
// 128 F of the TMA staging rows is what lifts the occupancy from 16 to 32 warps per SM; the step is latency-bound
// (measured: time ~ 0.06 ms + 1.0 ms / warps per SM at 1 Mi envs), so occupancy is what pays.  Rewards and the replay
// row are staged as words and leave the same way.
constexpr int kFlatMinCtas = 4;
template <int VARIANT, int TA = 0, int TJ = 0>
__global__ void __launch_bounds__(kThreads, kFlatMinCtas) k_step_flat(const __grid_constant__ StepParams p,
                                                                       const __grid_constant__ FlatStage L) {
  extern __shared__ __align__(128) uint8_t dyn_smem[];
  __shared__ GridTables tb;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int A = p.c.A, S = p.c.S, F = p.enc.ns_floats;
  const int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x, e0 = e - lane;
  const bool have = e < p.N;
  StepInput in;
  load_input(p, e, have, in);  // in flight while the tables and the row prefill are set up
  stage_tables_and_tick(p.c, tb, p.tick, p.tick_dev, p.tick_ctr);
  if (e0 >= p.N) return;  // whole warp
  const int cnt = p.N - e0 < 32 ? (int)(p.N - e0) : 32;
  uint8_t* blk = dyn_smem + (size_t)warp * L.per_warp;
  {
    const uint32_t z = kByteRowBias * 0x01010101u;
    for (int i = lane; i < (L.row_bytes >> 4); i += 32) reinterpret_cast<uint4*>(blk)[i] = make_uint4(z, z, z, z);
  }
  const int rew_row = reward_row_bytes(p);
  uint8_t* rew = blk + L.rew_off;
  float* nf = reinterpret_cast<float*>(blk + L.nf_off);
  bool stepped, finished;
  EnvState s = {};
  StepResult r = {};
  step_one<VARIANT, TA, TJ>(p, tb, e, have, in, reward_rows(p) ? rew + lane * rew_row : nullptr, p.next_flat ? nf + lane * S : nullptr,
                    s, r, stepped, finished);
  finish_one(p, tb, e, lane, s, r, stepped, finished);
  __syncwarp();  // the prefill is complete before any lane sets bytes in its row
  if (have) flat_row<ByteRow, TA>(p.c, p.enc, tb, obs_of(s), blk + lane * F);
  __syncwarp();  // the staged rows are complete before the reads below
  if (reward_rows(p)) warp_copy_bytes(reward_rows(p) + e0 * rew_row, rew, cnt * rew_row, lane);
  if (p.next_flat) warp_copy_words(reinterpret_cast<uint32_t*>(p.next_flat + e0 * S), reinterpret_cast<const uint32_t*>(nf), cnt * S, lane);
  warp_expand_byte_rows(p.non_spatial + e0 * F, reinterpret_cast<const uint32_t*>(blk), cnt * F, lane);
}

// K1+K2 for FLAT encodes, warp-specialised (the Flat counterpart of k_step_ws): the step is latency-bound, so what pays is
// many resident compute warps with few instructions each.  Here up to 28 compute warps per persistent CTA only run the step and
// leave, per group of 32 envs, a ROW RECORD in shared memory -- the K non-zero (offset, value) entries of each env's row, 32 x K x
// 2 bytes instead of the 32 x F bytes / 128 F bytes the staged / TMA paths hold per warp -- and a few emitter warps turn records
// into rows: each owns ONE persistently-zero tile of 32 rows x F floats, sets the entries of a group (lane = row), hands the
// tile to the TMA engine with one cp.async.bulk, and clears the same entries once the engine has read the tile.  The compute
// warps never touch the F floats of a row, the emitters never compute a step; rewards / the replay row / packed result records
// are staged and bulk-stored by the compute warps themselves.
struct FlatWsLayout {
  int32_t compute_warps, emitter_warps;
  int32_t K;                              // entries per env in a record (multiple of 8)
  int32_t tile_bytes, rec_bytes;          // one emitter tile (32 x F floats); one record (32 x K x 2 bytes)
  int32_t dense_bytes, rew_off, nf_off;   // per compute warp: rewards / packed records + next_flat staging
  int32_t recs_off, dense_off, bars_off, total_bytes;  // inside dynamic shared memory (tiles first)
};

constexpr int kFlatWsMaxWarps = 32;
template <int VARIANT, int TA = 0, int TJ = 0>
__global__ void __launch_bounds__(kFlatWsMaxWarps * 32, 1) k_step_flat_ws(const __grid_constant__ StepParams p,
                                                                          const __grid_constant__ FlatWsLayout L) {
  extern __shared__ __align__(128) uint8_t dyn_smem[];
  __shared__ GridTables tb;
  stage_tables_and_tick(p.c, tb, p.tick, p.tick_dev, p.tick_ctr);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int CW = L.compute_warps, EW = L.emitter_warps, F = p.enc.ns_floats, K = L.K, S = p.c.S;
  uint64_t* bars = reinterpret_cast<uint64_t*>(dyn_smem + L.bars_off);  // full[w*2+s], then empty[w*2+s]
  {
    float4* t = reinterpret_cast<float4*>(dyn_smem);
    for (int i = threadIdx.x; i < (EW * L.tile_bytes) >> 4; i += blockDim.x) t[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x == 0)
      for (int i = 0; i < 4 * CW; ++i) mbar_init(&bars[i], 1);
  }
  __syncthreads();
  const int64_t n_groups = (p.N + 31) >> 5;
  const int64_t g_stride = (int64_t)gridDim.x * CW;
  const int rew_row = reward_row_bytes(p);
  if (warp < CW) {
    // ------------------------------------------------------------------ compute warp
    uint8_t* dense = dyn_smem + L.dense_off + (size_t)warp * L.dense_bytes;
    uint8_t* rew = dense + L.rew_off;
    float* nf = reinterpret_cast<float*>(dense + L.nf_off);
    const bool dense_out = reward_rows(p) != nullptr || p.next_flat != nullptr;
    int64_t g = (int64_t)blockIdx.x * CW + warp;
    for (int it = 0; g < n_groups; ++it, g += g_stride) {
      const int64_t e0 = g << 5, e = e0 + lane;
      const bool have = e < p.N;
      const int cnt = p.N - e0 < 32 ? (int)(p.N - e0) : 32;
      StepInput in;
      load_input(p, e, have, in);
      const int sl = it & 1;
      uint16_t* rec = reinterpret_cast<uint16_t*>(dyn_smem + L.recs_off + (size_t)(warp * 2 + sl) * L.rec_bytes);
      if (it >= 2) mbar_wait(&bars[2 * CW + warp * 2 + sl], (uint32_t)(((it >> 1) - 1) & 1));  // the emitter has read this record slot
      if (it >= 1 && dense_out) {  // my dense stores of the previous group have been read (single staging block)
        if (lane == 0) bulk_wait_read_all();
        __syncwarp();
      }
      bool stepped, finished;
      EnvState s = {};
      StepResult r = {};
      step_one<VARIANT, TA, TJ>(p, tb, e, have, in, reward_rows(p) ? rew + lane * rew_row : nullptr,
                                p.next_flat ? nf + lane * S : nullptr, s, r, stepped, finished);
      finish_one(p, tb, e, lane, s, r, stepped, finished);
      {
        uint4* mine = reinterpret_cast<uint4*>(rec + lane * K);
        for (int i = 0; i < (K >> 3); ++i) mine[i] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
        if (have) flat_row_record(p.c, p.enc, tb, obs_of(s), rec + lane * K, K);
      }
      if (dense_out) fence_proxy_async_smem();
      __syncwarp();
      if (dense_out) {
        bool any = false;
        if (reward_rows(p)) any |= drain(reward_rows(p) + e0 * rew_row, rew, (uint32_t)(cnt * rew_row), lane);
        if (p.next_flat) any |= drain(p.next_flat + e0 * S, nf, (uint32_t)(cnt * S * 4), lane);
        if (lane == 0 && any) bulk_commit();
      }
      if (lane == 0) mbar_arrive(&bars[warp * 2 + sl]);  // the record is ready for the emitter
    }
    if (lane == 0) bulk_wait_all();
  } else if (warp < CW + EW) {
    // ------------------------------------------------------------------ emitter warp: serves compute warps em, em + EW, ...
    const int em = warp - CW;
    float* tile = reinterpret_cast<float*>(dyn_smem + (size_t)em * L.tile_bytes);
    constexpr int kMaxK = 16;
    uint32_t prev[kMaxK];  // float indices this lane set in the tile (0xffffffff = none)
#pragma unroll
    for (int j = 0; j < kMaxK; ++j) prev[j] = 0xffffffffu;
    bool inflight = false;
    const int64_t base_g = (int64_t)blockIdx.x * CW;
    for (int it = 0; base_g + (int64_t)it * g_stride < n_groups; ++it) {
      const int sl = it & 1;
      for (int w = em; w < CW; w += EW) {
        const int64_t g = base_g + w + (int64_t)it * g_stride;
        if (g >= n_groups) break;
        mbar_wait(&bars[w * 2 + sl], (uint32_t)((it >> 1) & 1));
        const uint16_t* rec = reinterpret_cast<const uint16_t*>(dyn_smem + L.recs_off + (size_t)(w * 2 + sl) * L.rec_bytes) + lane * K;
        uint32_t ent[kMaxK];
#pragma unroll
        for (int q = 0; q < kMaxK / 8; ++q) {  // all loads first
          uint4 v = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
          if (q * 8 < K) v = reinterpret_cast<const uint4*>(rec)[q];
          ent[8 * q + 0] = v.x & 0xffffu; ent[8 * q + 1] = v.x >> 16; ent[8 * q + 2] = v.y & 0xffffu; ent[8 * q + 3] = v.y >> 16;
          ent[8 * q + 4] = v.z & 0xffffu; ent[8 * q + 5] = v.z >> 16; ent[8 * q + 6] = v.w & 0xffffu; ent[8 * q + 7] = v.w >> 16;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[2 * CW + w * 2 + sl]);  // the record slot may be refilled
        if (inflight) {
          if (lane == 0) bulk_wait_read_all();  // the engine has read my tile
          __syncwarp();
        }
#pragma unroll
        for (int j = 0; j < kMaxK; ++j)
          if (prev[j] != 0xffffffffu) tile[prev[j]] = 0.0f;
#pragma unroll
        for (int j = 0; j < kMaxK; ++j) {
          if (ent[j] != kRecordNone) {
            const uint32_t idx = (uint32_t)(lane * F) + (ent[j] & 0x3ffu);
            tile[idx] = record_value(ent[j]);
            prev[j] = idx;
          } else {
            prev[j] = 0xffffffffu;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        const int64_t e0 = g << 5;
        const int cnt = p.N - e0 < 32 ? (int)(p.N - e0) : 32;
        if (drain(p.non_spatial + e0 * F, tile, (uint32_t)(cnt * F * 4), lane)) {
          if (lane == 0) bulk_commit();
          inflight = true;
        } else {
          inflight = false;
        }
      }
    }
    if (lane == 0) bulk_wait_all();
  }
}

// Random-policy rollout: every env advances `n_steps` steps inside ONE launch with its state in registers
// (== n_steps calls of step(None): same ticks, same draws, same auto-resets, same episode statistics); only the final
// state and, optionally, the per-agent reward sums are written.  This is ReplayBuffer.populate's / a random-policy
// benchmark's inner loop (replay_memory.py:103-143) with no launch per step.
struct RolloutParams {
  DevConfig c;
  StateArrays st;
  unsigned long long* stats;
  double* reward_sums;  // [N][A] or nullptr: sum over the rollout of each agent index's rewards
  uint64_t tick0;
  uint64_t* tick_dev;
  unsigned int* tick_ctr;
  int64_t N;
  int32_t n_steps;
};

template <int VARIANT, int TA = 0, int TJ = 0>
__global__ void __launch_bounds__(kThreads) k_rollout(const __grid_constant__ RolloutParams p) {
  __shared__ GridTables tb;
  stage_tables_and_tick(p.c, tb, p.tick0, p.tick_dev, p.tick_ctr, (uint64_t)p.n_steps);
  const DevConfig& c = p.c;
  const int A = TA ? TA : c.A, lane = threadIdx.x & 31;
  const int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const bool have = e < p.N;
  EnvState s = {};
  if (have) load_state(p.st, e, s);
  double sums[SUS_MAX_AGENTS];
#pragma unroll
  for (int i = 0; i < SUS_MAX_AGENTS; ++i) sums[i] = 0.0;
  uint32_t acc[SUS_N_STATS];
#pragma unroll
  for (int k = 0; k < SUS_N_STATS; ++k) acc[k] = 0;
  const uint64_t tick0 = tb.tick;
  for (int t = 0; t < p.n_steps; ++t) {
    const uint64_t tick = tick0 + (uint64_t)t;
    if (have) {
      uint64_t acts = 0;
      WordStream wa;
      wa.init(c, nullptr, (uint32_t)e, tick, P_ACT_FUSED);
      for (int i = 0; i < A; ++i)
        acts |= (uint64_t)bounded(wa.word(i), n_role_actions<VARIANT>(c, (s.imp >> i) & 1u)) << (8 * i);
      WordStream ws;
      ws.init(c, nullptr, (uint32_t)e, tick, P_STEP);
      StepResult r = {};
      step_env<VARIANT, TA, TJ>(c, tb, s, acts, ws, r);
      if (p.reward_sums) {
#pragma unroll
        for (int i = 0; i < SUS_MAX_AGENTS; ++i)
          if (i < A) sums[i] += agent_reward<VARIANT>(c, s, r, i);
      }
      if (r.done || r.trunc) {  // per-thread accumulation; flushed once at the end of the rollout
        acc[SUS_S_EPISODES] += 1u; acc[SUS_S_CREW_WON] += (s.misc >> 24) & 1u; acc[SUS_S_IMPOSTER_WON] += (s.misc >> 25) & 1u;
        acc[SUS_S_IMP_KILLED_CREW] += s.misc & 0xff; acc[SUS_S_COMPLETED_JOBS] += s.completed;
        acc[SUS_S_SABOTAGED_JOBS] += s.sabotaged; acc[SUS_S_IMP_VOTED_OUT] += (s.misc >> 8) & 0xff;
        acc[SUS_S_CREW_VOTED_OUT] += (s.misc >> 16) & 0xff; acc[SUS_S_TOTAL_TIME_STEPS] += s.nsteps;
        acc[SUS_S_TRUNCATED] += r.trunc ? 1u : 0u;
        if (c.auto_reset) {
          WordStream wr;
          wr.init(c, nullptr, (uint32_t)e, tick, P_AUTORESET);
          reset_env(c, tb, s, wr);
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < SUS_N_STATS; ++k) {
    const unsigned long long sum = (unsigned long long)__reduce_add_sync(kFull, acc[k] & 0xffffu) +
                                   ((unsigned long long)__reduce_add_sync(kFull, acc[k] >> 16) << 16);
    if (lane == 0 && sum) atomicAdd(p.stats + k, sum);
  }
  if (have) {
    store_state(p.st, e, s, true);
    if (p.reward_sums)
      for (int i = 0; i < A; ++i) {
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < SUS_MAX_AGENTS; ++q) v = q == i ? sums[q] : v;
        p.reward_sums[e * A + i] = v;
      }
  }
}

// K1 (+K2), TMA path: persistent CTAs (one per SM), each warp walks groups of 32 envs; rewards, the replay-layout
// state row and the feature tensors are staged in shared memory and leave the SM as TMA bulk stores.
template <int VARIANT, bool ENCODE>
__global__ void __launch_bounds__(kTmaMaxWarps * 32, 1) k_step_tma(const __grid_constant__ StepParams p,
                                                          const __grid_constant__ TileLayout L) {
  extern __shared__ __align__(128) uint8_t dyn_smem[];
  __shared__ GridTables tb;
  stage_tables_and_tick(p.c, tb, p.tick, p.tick_dev, p.tick_ctr);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int A = p.c.A;
  WarpEmitter em;
  em.init(dyn_smem + (size_t)warp * L.per_warp, &L, lane);
  if (ENCODE && p.enc.sp_floats > 0) em.zero_spatial();
  const int64_t n_groups = (p.N + 31) >> 5;
  const int rew_row = reward_row_bytes(p);
  const int64_t g_stride = (int64_t)gridDim.x * L.warps;
  StepInput in, in_next;
  {
    const int64_t g = (int64_t)blockIdx.x * L.warps + warp;
    load_input(p, (g << 5) + lane, g < n_groups && (g << 5) + lane < p.N, in);
  }
  for (int64_t g = (int64_t)blockIdx.x * L.warps + warp; g < n_groups; g += g_stride) {
    const int64_t e0 = g << 5, e = e0 + lane;
    const bool have = e < p.N;
    {  // prefetch the next group's state records and action rows
      const int64_t gn = g + g_stride, en = (gn << 5) + lane;
      load_input(p, en, gn < n_groups && en < p.N, in_next);
    }
    const int64_t rem = p.N - e0;
    const int cnt = rem < 32 ? (int)rem : 32;
    bool stepped, finished;
    EnvState s = {};
    StepResult r = {};
    em.acquire_dense();  // the previous group's dense bulk stores have finished reading the staging rows
    step_one<VARIANT>(p, tb, e, have, in, reward_rows(p) ? em.rew() + lane * rew_row : nullptr,
                      p.next_flat ? em.nf() + lane * p.c.S : nullptr, s, r, stepped, finished);
    finish_one(p, tb, e, lane, s, r, stepped, finished);
    bool dense_any = false;
    if (reward_rows(p) || p.next_flat) {
      fence_proxy_async_smem();
      __syncwarp();
      if (reward_rows(p)) dense_any |= drain(reward_rows(p) + e0 * rew_row, em.rew(), (uint32_t)(cnt * rew_row), lane);
      if (p.next_flat) dense_any |= drain(p.next_flat + e0 * p.c.S, em.nf(), (uint32_t)(cnt * p.c.S * 4), lane);
      if (!ENCODE && dense_any) {
        if (lane == 0) bulk_commit();
        em.committed_dense();
      }
    }
    if (ENCODE)
      warp_encode_tma(p.c, p.enc, tb, em, obs_of(s), e0, cnt, have, p.N, p.spatial, p.non_spatial, true, dense_any);
    in = in_next;
  }
  em.finish();
}

// One plane sub-tile (8 envs) of the emitter warp: retire the store that last used this tile, clear the ones it
// carried, set the new ones from the group's plane records with all 32 lanes, hand the tile to the TMA engine.
// T: the plane element type -- float (the reference's tensors) or uint8_t (opt-in compact planes, SUS_ENCODE_PLANES_U8).
template <int TE, typename T>  // envs per plane tile: 8 or 16
struct EmitterTile {
  static constexpr int NE = TE / 2;  // plane entries per lane: TE envs x 16 entries over 32 lanes
  T* tile;
  uint32_t prev[NE];  // element indices this lane set in the tile (0xffffffff = none)
  bool inflight;
};

template <int TE, bool PERSPECTIVE, typename T>
__device__ __forceinline__ void emit_sub(EmitterTile<TE, T>& t, EmitterTile<TE, T>& other, int& last_commit, int my_id,
                                         const uint16_t* po, int g0, int cnt, int k, int A, int R, T* gdst, int lane) {
  constexpr int NE = EmitterTile<TE, T>::NE;
  if (t.inflight) {
    if (lane == 0) {
      if (last_commit != my_id) bulk_wait_read_all_but_newest();  // the newest group reads the other tile
      else bulk_wait_read_all();
    }
    if (last_commit == my_id) other.inflight = false;
    t.inflight = false;
  }
  __syncwarp();
#pragma unroll
  for (int r = 0; r < NE; ++r)
    if (t.prev[r] != 0xffffffffu) t.tile[t.prev[r]] = (T)0;
  const int gc = cnt - g0 < TE ? cnt - g0 : TE;
  const int el = lane & (TE - 1);
  const uint16_t* rec = po + (g0 + el) * 16 + lane / TE;
  uint32_t off[NE];
#pragma unroll
  for (int r = 0; r < NE; ++r) off[r] = el < gc ? (uint32_t)rec[(32 / TE) * r] : 0x3ffu;  // all loads first
#pragma unroll
  for (int r = 0; r < NE; ++r) {
    if (off[r] != 0x3ffu) {
      if (PERSPECTIVE) {
        const int q = lane / TE + (32 / TE) * r;
        if (k > 0 && q < A) off[r] += (uint32_t)((persp_channel_of_agent(k, q) - q) * 81);  // view k's channel order
      }
      const uint32_t idx = (uint32_t)(el * R) + off[r];
      t.tile[idx] = (T)1;
      t.prev[r] = idx;
    } else {
      t.prev[r] = 0xffffffffu;
    }
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (drain(gdst, t.tile, (uint32_t)(gc * R * (int)sizeof(T)), lane)) {
    if (lane == 0) bulk_commit();
    t.inflight = true;
    last_commit = my_id;
  }
}

// the emitter warp's loop over the groups its CTA's compute warps produce
template <int TE, bool PERSPECTIVE, typename T>
__device__ __forceinline__ void emitter_loop(int A, int R, int64_t N, T* __restrict__ spatial, const WsLayout& L,
                                             uint8_t* dyn_smem, uint64_t* bars, int64_t n_groups, int64_t g_stride, int lane) {
  constexpr int NE = EmitterTile<TE, T>::NE;
  const int CW = L.compute_warps;
  EmitterTile<TE, T> t0, t1;
  t0.tile = reinterpret_cast<T*>(dyn_smem);
  t1.tile = reinterpret_cast<T*>(dyn_smem + L.tile_bytes);
  t0.inflight = t1.inflight = false;
#pragma unroll
  for (int r = 0; r < NE; ++r) t0.prev[r] = t1.prev[r] = 0xffffffffu;
  int last_commit = -1;
  const int sp_views = PERSPECTIVE ? A : 1;
  const int64_t base_g = (int64_t)blockIdx.x * CW;
  for (int it = 0; base_g + (int64_t)it * g_stride < n_groups; ++it) {
    const int sl = it & 1;
    for (int w = 0; w < CW; ++w) {
      const int64_t g = base_g + w + (int64_t)it * g_stride;
      if (g >= n_groups) break;
      const uint8_t* slot = dyn_smem + L.slots_off + (size_t)(w * 2 + sl) * L.slot_bytes;
      mbar_wait(&bars[w * 2 + sl], (uint32_t)((it >> 1) & 1));
      const int64_t e0 = g << 5;
      const int64_t rem = N - e0;
      const int cnt = rem < 32 ? (int)rem : 32;
      const uint16_t* po = reinterpret_cast<const uint16_t*>(slot + L.po_off);
      for (int k = 0; k < sp_views; ++k) {
        T* gbase = spatial + ((int64_t)k * N + e0) * R;
        for (int g0 = 0; g0 < cnt; g0 += 2 * TE) {
          emit_sub<TE, PERSPECTIVE, T>(t0, t1, last_commit, 0, po, g0, cnt, k, A, R, gbase + (int64_t)g0 * R, lane);
          if (g0 + TE < cnt)
            emit_sub<TE, PERSPECTIVE, T>(t1, t0, last_commit, 1, po, g0 + TE, cnt, k, A, R, gbase + (int64_t)(g0 + TE) * R, lane);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[2 * CW + w * 2 + sl]);  // the slot may be refilled
    }
  }
  if (lane == 0) bulk_wait_all();
}

__device__ __forceinline__ void emitter_dispatch(bool persp, bool planes_u8, int A, int R, int64_t N, void* spatial,
                                                 const WsLayout& L, uint8_t* dyn_smem, uint64_t* bars, int64_t n_groups,
                                                 int64_t g_stride, int lane) {
  if (planes_u8) {  // 16-env tiles only: 16 x (A+2) x 81 bytes is a multiple of 16, 8 x is not
    uint8_t* sp = static_cast<uint8_t*>(spatial);
    if (persp) emitter_loop<16, true, uint8_t>(A, R, N, sp, L, dyn_smem, bars, n_groups, g_stride, lane);
    else emitter_loop<16, false, uint8_t>(A, R, N, sp, L, dyn_smem, bars, n_groups, g_stride, lane);
    return;
  }
  float* sp = static_cast<float*>(spatial);
  if (L.tile_envs == 16) {
    if (persp) emitter_loop<16, true, float>(A, R, N, sp, L, dyn_smem, bars, n_groups, g_stride, lane);
    else emitter_loop<16, false, float>(A, R, N, sp, L, dyn_smem, bars, n_groups, g_stride, lane);
  } else {
    if (persp) emitter_loop<8, true, float>(A, R, N, sp, L, dyn_smem, bars, n_groups, g_stride, lane);
    else emitter_loop<8, false, float>(A, R, N, sp, L, dyn_smem, bars, n_groups, g_stride, lane);
  }
}

// K1+K2, warp-specialised TMA path (susnet_ws.cuh): CW compute warps + one emitter warp per persistent CTA.
template <int VARIANT>
__global__ void __launch_bounds__(kWsMaxWarps * 32, 1) k_step_ws(const __grid_constant__ StepParams p,
                                                                  const __grid_constant__ WsLayout L) {
  extern __shared__ __align__(128) uint8_t dyn_smem[];
  __shared__ GridTables tb;
  stage_tables_and_tick(p.c, tb, p.tick, p.tick_dev, p.tick_ctr);
  const DevConfig& c = p.c;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int CW = L.compute_warps, A = c.A, R = p.enc.sp_floats, F = p.enc.ns_floats;
  uint64_t* bars = reinterpret_cast<uint64_t*>(dyn_smem + L.bars_off);  // full[w*2+s], then empty[w*2+s]
  {
    float4* t = reinterpret_cast<float4*>(dyn_smem);
    for (int i = threadIdx.x; i < (2 * L.tile_bytes) >> 4; i += blockDim.x) t[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x == 0)
      for (int i = 0; i < 4 * CW; ++i) mbar_init(&bars[i], 1);
  }
  __syncthreads();
  const int64_t n_groups = (p.N + 31) >> 5;
  const int64_t g_stride = (int64_t)gridDim.x * CW;
  const int rew_row = reward_row_bytes(p);
  if (warp < CW) {
    // ------------------------------------------------------------------ compute warp
    // (issuing the first group's loads before the set-up above changed nothing at 65 536 envs: 38.9 us either way)
    int64_t g = (int64_t)blockIdx.x * CW + warp;
    StepInput in, in_next;
    load_input(p, (g << 5) + lane, g < n_groups && (g << 5) + lane < p.N, in);
    for (int it = 0; g < n_groups; ++it, g += g_stride) {
      const int64_t e0 = g << 5, e = e0 + lane;
      const bool have = e < p.N;
      const int64_t rem = p.N - e0;
      const int cnt = rem < 32 ? (int)rem : 32;
      {
        const int64_t gn = g + g_stride, en = (gn << 5) + lane;
        load_input(p, en, gn < n_groups && en < p.N, in_next);
      }
      const int sl = it & 1;
      uint8_t* slot = dyn_smem + L.slots_off + (size_t)(warp * 2 + sl) * L.slot_bytes;
      if (it >= 2) {
        mbar_wait(&bars[2 * CW + warp * 2 + sl], (uint32_t)(((it >> 1) - 1) & 1));  // emitter is done with the slot
        if (lane == 0) bulk_wait_read_all_but_newest();  // my dense stores of two groups ago have been read
        __syncwarp();
      }
      bool stepped, finished;
      EnvState s = {};
      StepResult r = {};
      uint8_t* rew = slot + L.rew_off;
      float* nf = reinterpret_cast<float*>(slot + L.nf_off);
      step_one<VARIANT>(p, tb, e, have, in, reward_rows(p) ? rew + lane * rew_row : nullptr,
                        p.next_flat ? nf + lane * c.S : nullptr, s, r, stepped, finished);
      finish_one(p, tb, e, lane, s, r, stepped, finished);
      const ObsState o = obs_of(s);
      float* ns = reinterpret_cast<float*>(slot + L.ns_off);
      if (have) {
        if (p.enc.kind == SUS_ENCODE_GLOBAL) global_ns_rows(c, o, ns + lane * F, 32 * F);
        else
          for (int k = 0; k < A; ++k) persp_ns_row(c, o, k, ns + (k * 32 + lane) * F);
      }
      write_plane_record(c, o, have, reinterpret_cast<uint16_t*>(slot + L.po_off) + lane * 16);
      fence_proxy_async_smem();
      __syncwarp();
      if (reward_rows(p)) drain(reward_rows(p) + e0 * rew_row, rew, (uint32_t)(cnt * rew_row), lane);
      if (p.next_flat) drain(p.next_flat + e0 * c.S, nf, (uint32_t)(cnt * c.S * 4), lane);
      for (int k = 0; k < A; ++k)
        drain(p.non_spatial + ((int64_t)k * p.N + e0) * F, ns + k * 32 * F, (uint32_t)(cnt * F * 4), lane);
      __syncwarp();
      if (lane == 0) {
        bulk_commit();                     // one (possibly empty) DENSE group per group of envs
        mbar_arrive(&bars[warp * 2 + sl]);  // plane records are ready for the emitter
      }
      in = in_next;
    }
    if (lane == 0) bulk_wait_all();
  } else if (warp == CW) {
    // ------------------------------------------------------------------ emitter warp
    emitter_dispatch(p.enc.kind == SUS_ENCODE_PERSPECTIVE, p.enc.planes_u8 != 0, A, R, p.N, p.spatial, L, dyn_smem, bars, n_groups, g_stride, lane);
  }
}

__global__ void __launch_bounds__(kThreads) k_sample_actions(const __grid_constant__ ActParams p) {
  // each warp's 32 x A action words are contiguous in [N][A]: stage them in shared memory and write them with
  // lane-contiguous 4-byte stores (A strided stores per lane cost A x the store sectors)
  __shared__ int32_t stage[kThreads / 32][32 * SUS_MAX_AGENTS];
  __shared__ uint64_t tick;
  if (threadIdx.x == 0) tick = fetch_launch_tick(p.tick, p.tick_dev, p.tick_ctr, 1);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, A = p.c.A;
  const int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x, e0 = e - lane;
  if (e < p.N) {
    const uint32_t imp = (p.st.aux[e].x >> 8) & 0xff;
    uint32_t w[SUS_MAX_AGENTS];  // word i of the (env, act epoch, P_ACT) stream
    if (p.inj_act) {
#pragma unroll
      for (int i = 0; i < SUS_MAX_AGENTS; ++i) w[i] = i < A ? p.inj_act[e * A + i] : 0u;
    } else {  // == WordStream::word(0..7), with the two blocks' rounds interleaved
      const uint32_t env_id = p.c.env_id_base + (uint32_t)e;
      uint4 a = make_uint4(env_id, (uint32_t)tick, (uint32_t)(tick >> 32), P_ACT), b = a;
      b.w |= 1u << 8;
      if (A > 4) philox4x32_10_x2(a, b, p.c.seed_lo, p.c.seed_hi);
      else a = philox4x32_10(a, p.c.seed_lo, p.c.seed_hi);
      w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    }
#pragma unroll
    for (int i = 0; i < SUS_MAX_AGENTS; ++i)  // base.py:326-330 (R6): dead agents are sampled too
      if (i < A) stage[warp][lane * A + i] = (int32_t)bounded(w[i], role_actions_rt(p.c, (imp >> i) & 1u));
  }
  __syncwarp();
  const int64_t rem = p.N - e0;
  const int n = (int)(rem < 32 ? (rem < 0 ? 0 : rem) : 32) * A;
  for (int f = lane; f < n; f += 32) p.out[e0 * A + f] = stage[warp][f];
}

__global__ void __launch_bounds__(kThreads) k_encode_env(const __grid_constant__ EncodeParams p) {
  __shared__ GridTables tb;
  stage_tables(p.c, tb);
  const int lane = threadIdx.x & 31;
  const int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x, e0 = e - lane;
  const bool have = e < p.n_items;
  EnvState s = {};
  if (have) load_state(p.st, e, s);
  const int64_t rem = p.n_items - e0;
  warp_encode(p.c, p.enc, tb, obs_of(s), e0, rem < 32 ? (int)(rem < 0 ? 0 : rem) : 32, have, p.n_items, p.spatial,
              p.non_spatial);
}

// unflatten one row (gymnasium.spaces.unflatten casts to int64: truncation) into the packed observation
template <typename T>
__device__ __forceinline__ ObsState parse_row(const DevConfig& c, const T* __restrict__ row) {
  ObsState o = {};
  const int A = c.A, J = c.J;
  int k = 0;
  for (int i = 0; i < A; ++i) {
    const long long x = (long long)row[k], y = (long long)row[k + 1];
    k += 2;
    const uint32_t code = (x >= 0 && x <= 8 && y >= 0 && y <= 8) ? (uint32_t)((x << 4) | y) : 0xffu;
    o.pos |= (uint64_t)code << (8 * i);
  }
  for (int i = 0; i < A; ++i) o.alive |= ((long long)row[k++] != 0 ? 1u : 0u) << i;
  if (J > 0 || c.variant == SUS_VARIANT_TAGGING) {
    for (int j = 0; j < J; ++j) {
      const long long x = (long long)row[k], y = (long long)row[k + 1];
      k += 2;
      const uint32_t code = (x >= 0 && x <= 8 && y >= 0 && y <= 8) ? (uint32_t)((x << 4) | y) : 0xffu;
      o.jobpos |= (uint64_t)code << (8 * j);
    }
    for (int j = 0; j < J; ++j) o.jobdone |= ((long long)row[k++] != 0 ? 1u : 0u) << j;
  }
  if (c.variant == SUS_VARIANT_TAGGING) {
    for (int i = 0; i < A; ++i) o.used |= ((long long)row[k++] != 0 ? 1u : 0u) << i;
    for (int i = 0; i < A; ++i) {
      long long t = (long long)row[k++];
      t = t < 0 ? 0 : (t > 15 ? 15 : t);
      o.tagcnt |= (uint32_t)t << (4 * i);
    }
  }
  return o;
}

template <typename T>
__global__ void __launch_bounds__(kThreads) k_encode_rows(const __grid_constant__ EncodeParams p) {
  __shared__ GridTables tb;
  stage_tables(p.c, tb);
  const int lane = threadIdx.x & 31;
  const int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x, e0 = e - lane;
  const bool have = e < p.n_items;
  ObsState o = {};
  if (have) o = parse_row<T>(p.c, static_cast<const T*>(p.rows) + e * p.c.S);
  const int64_t rem = p.n_items - e0;
  warp_encode(p.c, p.enc, tb, o, e0, rem < 32 ? (int)(rem < 0 ? 0 : rem) : 32, have, p.n_items, p.spatial, p.non_spatial);
}

// K2 for FLAT encodes, byte-staged like k_step_flat: live env state (FROM_ROWS = false) or (B*T, S) rows.
template <typename T, bool FROM_ROWS>
__global__ void __launch_bounds__(kThreads, kFlatMinCtas) k_encode_flat(const __grid_constant__ EncodeParams p,
                                                                         const __grid_constant__ FlatStage L) {
  extern __shared__ __align__(128) uint8_t dyn_smem[];
  __shared__ GridTables tb;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, F = p.enc.ns_floats;
  const int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x, e0 = e - lane;
  const bool have = e < p.n_items;
  ObsState o = {};
  if (have) {
    if (FROM_ROWS) {
      o = parse_row<T>(p.c, static_cast<const T*>(p.rows) + e * p.c.S);
    } else {
      EnvState s;
      load_state(p.st, e, s);
      o = obs_of(s);
    }
  }
  stage_tables(p.c, tb);
  if (e0 >= p.n_items) return;  // whole warp
  const int cnt = p.n_items - e0 < 32 ? (int)(p.n_items - e0) : 32;
  uint8_t* blk = dyn_smem + (size_t)warp * L.per_warp;
  const uint32_t z = kByteRowBias * 0x01010101u;
  for (int i = lane; i < (L.row_bytes >> 4); i += 32) reinterpret_cast<uint4*>(blk)[i] = make_uint4(z, z, z, z);
  __syncwarp();
  if (have) flat_row<ByteRow>(p.c, p.enc, tb, o, blk + lane * F);
  __syncwarp();
  warp_expand_byte_rows(p.non_spatial + e0 * F, reinterpret_cast<const uint32_t*>(blk), cnt * F, lane);
}

// K2, TMA path (see k_step_tma): persistent CTAs, features staged in shared memory, bulk stores.
template <typename T, bool FROM_ROWS>
__global__ void __launch_bounds__(kTmaMaxWarps * 32, 1) k_encode_tma(const __grid_constant__ EncodeParams p,
                                                            const __grid_constant__ TileLayout L) {
  extern __shared__ __align__(128) uint8_t dyn_smem[];
  __shared__ GridTables tb;
  stage_tables(p.c, tb);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  WarpEmitter em;
  em.init(dyn_smem + (size_t)warp * L.per_warp, &L, lane);
  if (p.enc.sp_floats > 0) em.zero_spatial();
  const int64_t n_groups = (p.n_items + 31) >> 5;
  for (int64_t g = (int64_t)blockIdx.x * L.warps + warp; g < n_groups; g += (int64_t)gridDim.x * L.warps) {
    const int64_t e0 = g << 5, e = e0 + lane;
    const bool have = e < p.n_items;
    const int64_t rem = p.n_items - e0;
    ObsState o = {};
    if (have) {
      if (FROM_ROWS) {
        o = parse_row<T>(p.c, static_cast<const T*>(p.rows) + e * p.c.S);
      } else {
        EnvState s;
        load_state(p.st, e, s);
        o = obs_of(s);
      }
    }
    warp_encode_tma(p.c, p.enc, tb, em, o, e0, rem < 32 ? (int)rem : 32, have, p.n_items, p.spatial, p.non_spatial, false,
                    false);
  }
  em.finish();
}

// K2, warp-specialised (see k_step_ws): compute warps unflatten rows / load env state and leave plane records + dense
// rows; the emitter warp writes the plane tiles.
template <typename T, bool FROM_ROWS>
__global__ void __launch_bounds__(kWsMaxWarps * 32, 1) k_encode_ws(const __grid_constant__ EncodeParams p,
                                                                    const __grid_constant__ WsLayout L) {
  extern __shared__ __align__(128) uint8_t dyn_smem[];
  __shared__ GridTables tb;
  stage_tables(p.c, tb);
  const DevConfig& c = p.c;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int CW = L.compute_warps, A = c.A, R = p.enc.sp_floats, F = p.enc.ns_floats;
  uint64_t* bars = reinterpret_cast<uint64_t*>(dyn_smem + L.bars_off);
  {
    float4* t = reinterpret_cast<float4*>(dyn_smem);
    for (int i = threadIdx.x; i < (2 * L.tile_bytes) >> 4; i += blockDim.x) t[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x == 0)
      for (int i = 0; i < 4 * CW; ++i) mbar_init(&bars[i], 1);
  }
  __syncthreads();
  const int64_t n_groups = (p.n_items + 31) >> 5;
  const int64_t g_stride = (int64_t)gridDim.x * CW;
  if (warp < CW) {
    int it = 0;
    for (int64_t g = (int64_t)blockIdx.x * CW + warp; g < n_groups; g += g_stride, ++it) {
      const int64_t e0 = g << 5, e = e0 + lane;
      const bool have = e < p.n_items;
      const int64_t rem = p.n_items - e0;
      const int cnt = rem < 32 ? (int)rem : 32;
      ObsState o = {};
      if (have) {
        if (FROM_ROWS) {
          o = parse_row<T>(c, static_cast<const T*>(p.rows) + e * c.S);
        } else {
          EnvState s;
          load_state(p.st, e, s);
          o = obs_of(s);
        }
      }
      const int sl = it & 1;
      uint8_t* slot = dyn_smem + L.slots_off + (size_t)(warp * 2 + sl) * L.slot_bytes;
      if (it >= 2) {
        mbar_wait(&bars[2 * CW + warp * 2 + sl], (uint32_t)(((it >> 1) - 1) & 1));
        if (lane == 0) bulk_wait_read_all_but_newest();
        __syncwarp();
      }
      float* ns = reinterpret_cast<float*>(slot + L.ns_off);
      if (have) {
        if (p.enc.kind == SUS_ENCODE_GLOBAL) global_ns_rows(c, o, ns + lane * F, 32 * F);
        else
          for (int k = 0; k < A; ++k) persp_ns_row(c, o, k, ns + (k * 32 + lane) * F);
      }
      write_plane_record(c, o, have, reinterpret_cast<uint16_t*>(slot + L.po_off) + lane * 16);
      fence_proxy_async_smem();
      __syncwarp();
      for (int k = 0; k < A; ++k)
        drain(p.non_spatial + ((int64_t)k * p.n_items + e0) * F, ns + k * 32 * F, (uint32_t)(cnt * F * 4), lane);
      __syncwarp();
      if (lane == 0) {
        bulk_commit();
        mbar_arrive(&bars[warp * 2 + sl]);
      }
    }
    if (lane == 0) bulk_wait_all();
  } else if (warp == CW) {
    emitter_dispatch(p.enc.kind == SUS_ENCODE_PERSPECTIVE, p.enc.planes_u8 != 0, A, R, p.n_items, p.spatial, L, dyn_smem, bars, n_groups, g_stride, lane);
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) k_export_flat(const __grid_constant__ ExportParams p) {
  const int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (e >= p.N) return;
  EnvState s;
  load_state(p.st, e, s);
  write_flat<T>(p.c, s, static_cast<T*>(p.out) + e * p.c.S);
}

__global__ void __launch_bounds__(kThreads) k_export_imp(const __grid_constant__ ExportParams p) {
  const int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (e >= p.N) return;
  const uint32_t imp = (p.st.aux[e].x >> 8) & 0xff;
  for (int i = 0; i < p.c.A; ++i) static_cast<uint8_t*>(p.out)[e * p.c.A + i] = (imp >> i) & 1u;
}

__global__ void __launch_bounds__(kThreads) k_export_metrics(const __grid_constant__ ExportParams p) {
  const int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (e >= p.N) return;
  EnvState s;
  load_state(p.st, e, s);
  long long* m = static_cast<long long*>(p.out) + e * SUS_N_METRICS;
  m[SUS_M_TOTAL_TIME_STEPS] = s.nsteps; m[SUS_M_IMP_KILLED_CREW] = s.misc & 0xff;
  m[SUS_M_COMPLETED_JOBS] = s.completed; m[SUS_M_SABOTAGED_JOBS] = s.sabotaged;
  m[SUS_M_IMP_VOTED_OUT] = (s.misc >> 8) & 0xff; m[SUS_M_CREW_VOTED_OUT] = (s.misc >> 16) & 0xff;
  m[SUS_M_CREW_WON] = (s.misc >> 24) & 1u; m[SUS_M_IMPOSTER_WON] = (s.misc >> 25) & 1u;
}

__global__ void __launch_bounds__(kThreads) k_import_flat(const __grid_constant__ ImportParams p) {
  const int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (e >= p.N) return;
  const DevConfig& c = p.c;
  const ObsState o = parse_row<long long>(c, p.flat + e * c.S);
  EnvState s = {};
  s.pos = o.pos; s.jobpos = o.jobpos; s.alive = o.alive; s.jobdone = o.jobdone; s.used = o.used; s.tagcnt = o.tagcnt;
  if (c.variant == SUS_VARIANT_TAGGING) s.timer = (uint32_t)((long long)c.tag_interval - p.flat[e * c.S + c.S - 1]);
  for (int i = 0; i < c.A; ++i) s.imp |= (p.imp[e * c.A + i] ? 1u : 0u) << i;
  s.nsteps = p.t ? (uint32_t)p.t[e] : 0u;
  store_state(p.st, e, s, true);
}

// ------------------------------------------------------------------------------------------ host side
int flat_size(const SusConfig& c) {
  const int A = c.n_imposters + c.n_crew, J = c.n_jobs;
  int S = 3 * A + ((J > 0 || c.variant == SUS_VARIANT_TAGGING) ? 3 * J : 0);  // base.py:211-228
  if (c.variant == SUS_VARIANT_TAGGING) S += 2 * A + 1;                        // tagging.py:42-60
  return S;
}

int role_actions_host(const SusConfig& c, int is_imposter) {
  const int A = c.n_imposters + c.n_crew;
  if (c.variant == SUS_VARIANT_TRAINING_GROUND) return is_imposter ? 6 : 5;
  const int base = is_imposter ? 7 : 6;
  return c.variant == SUS_VARIANT_TAGGING ? base + A - 1 : base;
}

// record geometry of the compact host protocol (SusCompactLayout in the header)
void compact_layout(const SusConfig& c, SusCompactLayout& L) {
  const int A = c.n_imposters + c.n_crew;
  const int n_max = role_actions_host(c, 1);  // the imposter list is the longer one
  int ab = 1;
  while ((1 << ab) < n_max) ++ab;
  L.action_bits = ab;
  L.action_bytes = (A * ab + 7) / 8;
  L.n_codes = (int)n_live_codes(c.variant) + 1;
  int rb = 1;
  while ((1 << rb) - 1 < L.n_codes) ++rb;  // the all-ones code is reserved for "actions rejected"
  L.reward_bits = rb;
  L.result_bytes = (A * rb + 2 + 7) / 8;
  L.invalid_code = (1 << rb) - 1;
}

int validate_config(const SusConfig& c) {
  // _validate_init_args: base.py:243-249, pred_prey.py:75-76
  if (c.variant < 0 || c.variant > 2) return fail(SUS_ERR_INVALID_ARGUMENT, "unknown variant");
  if (c.variant == SUS_VARIANT_TRAINING_GROUND) {
    if (c.n_crew <= 0) return fail(SUS_ERR_INVALID_ARGUMENT, "Must have at least one crew member.");
    if (c.n_imposters != 1) return fail(SUS_ERR_INVALID_ARGUMENT, "ImposterTrainingGround has exactly one imposter");
  } else {
    if (c.n_imposters <= 0) return fail(SUS_ERR_INVALID_ARGUMENT, "Must have at least one imposter.");
    if (c.n_crew <= 0) return fail(SUS_ERR_INVALID_ARGUMENT, "Must have at least one crew member.");
    if (c.n_imposters >= c.n_crew) return fail(SUS_ERR_INVALID_ARGUMENT, "Must be more crew members than imposters.");
  }
  if (c.n_jobs < 0) return fail(SUS_ERR_INVALID_ARGUMENT, "Must non-negative jobs.");
  if (c.n_imposters + c.n_crew > SUS_MAX_AGENTS) return fail(SUS_ERR_UNSUPPORTED, "more than 8 agents is not supported");
  if (c.n_jobs > SUS_MAX_JOBS) return fail(SUS_ERR_UNSUPPORTED, "more than 8 jobs is not supported");
  if (c.variant == SUS_VARIANT_TAGGING && c.n_jobs == 0)
    return fail(SUS_ERR_UNSUPPORTED, "the tagging env with n_jobs == 0 has no consistent state layout in the reference");
  if (c.max_time_steps < 1) return fail(SUS_ERR_INVALID_ARGUMENT, "max_time_steps must be >= 1");
  if (c.variant == SUS_VARIANT_TAGGING && c.tag_reset_interval < 1)
    return fail(SUS_ERR_INVALID_ARGUMENT, "tag_reset_interval must be >= 1");
  if (c.num_envs < 0 || (uint64_t)c.env_id_base + (uint64_t)c.num_envs > 0x100000000ull)
    return fail(SUS_ERR_INVALID_ARGUMENT, "env_id_base + num_envs must fit in 32 bits");
  return SUS_OK;
}

void make_dev_config(const SusConfig& c, DevConfig& d) {
  std::memset(&d, 0, sizeof(d));
  d.variant = c.variant; d.nI = c.n_imposters; d.A = c.n_imposters + c.n_crew; d.J = c.n_jobs;
  d.S = flat_size(c);
  d.order_random = c.is_action_order_random; d.shuffle_imp = c.shuffle_imposter_index; d.auto_reset = c.auto_reset;
  d.max_time_steps = (uint32_t)c.max_time_steps; d.tag_interval = (uint32_t)c.tag_reset_interval;
  d.seed_lo = (uint32_t)c.seed; d.seed_hi = (uint32_t)(c.seed >> 32); d.env_id_base = c.env_id_base;
  d.r_kill = c.kill_reward; d.r_fix = c.complete_job_reward; d.r_sab = c.sabotage_reward; d.r_tsr = c.time_step_reward;
  d.r_end = c.game_end_reward; d.r_dead = c.dead_penalty; d.r_vote = c.vote_reward;
  // geometry: base.py:171-199
  static const int walls[13][2] = {{0, 4}, {2, 4}, {3, 4}, {4, 4}, {5, 4}, {6, 4}, {8, 4},
                                   {4, 0}, {4, 2}, {4, 3}, {4, 5}, {4, 6}, {4, 8}};
  bool grid[9][9];
  for (auto& row : grid) for (bool& g : row) g = true;
  if (c.include_walls) for (auto& w : walls) grid[w[0]][w[1]] = false;
  int V = 0;
  for (int x = 0; x < 9; ++x)
    for (int y = 0; y < 9; ++y)
      if (grid[x][y]) {
        const uint32_t code = (uint32_t)(x << 4 | y);
        d.valid_bits[code >> 5] |= 1u << (code & 31u);
        d.cell_code[V++] = (uint8_t)code;
      }
  d.V = V;
}

int component_size(const SusConfig& c, int comp) {
  const int A = c.n_imposters + c.n_crew;
  switch (comp) {
    case SUS_FC_ONEHOT_POS: return A * 18;
    case SUS_FC_COORDS: return 2 * A;
    case SUS_FC_ALIVE_CREW: return A - 1;
    case SUS_FC_CLOSEST_CREW: return c.n_imposters == 1 ? c.n_crew : -1;  // indexes l1_crew[agent_idx - 1]
    case SUS_FC_L1_CREW: return c.n_imposters == 1 ? c.n_crew : -1;
    case SUS_FC_DIST_TO_IMPOSTER: return 2 * (A - 1);
    case SUS_FC_WALLS: return 9;
    case SUS_FC_ROOMS: return 8;
    case SUS_FC_SCENT: return 4;
    case SUS_FC_STATE_ALIVE: return A;
    case SUS_FC_STATE_JOB_STATUS: return c.n_jobs;
    case SUS_FC_STATE_USED_TAGS: return c.variant == SUS_VARIANT_TAGGING ? A : -1;
    case SUS_FC_STATE_TAG_COUNTS: return c.variant == SUS_VARIANT_TAGGING ? A : -1;
  }
  return -1;
}

int make_dev_encode(const SusConfig& c, const SusEncodeSpec* spec, DevEncode& d, SusEncodeShape* shape) {
  std::memset(&d, 0, sizeof(d));
  const int A = c.n_imposters + c.n_crew, J = c.n_jobs;
  SusEncodeShape sh = {0, 0, 0, 0};
  if (!spec || spec->kind == SUS_ENCODE_NONE) {
    d.kind = SUS_ENCODE_NONE;
  } else if (spec->kind == SUS_ENCODE_GLOBAL || spec->kind == SUS_ENCODE_PERSPECTIVE) {
    if (J == 0)  // the reference raises IndexError: the state tuple has no job fields
      return fail(SUS_ERR_INVALID_ARGUMENT, "Global/Perspective featurizers need n_jobs > 0");
    const int tags = c.variant == SUS_VARIANT_TAGGING ? A : 0;
    d.kind = spec->kind;
    d.planes_u8 = (spec->flags & SUS_ENCODE_PLANES_U8) ? 1 : 0;
    sh.spatial_floats = (A + 2) * 81;
    sh.non_spatial_views = A;
    if (spec->kind == SUS_ENCODE_GLOBAL) { sh.spatial_views = 1; sh.non_spatial_floats = A + tags + J + A; }
    else { sh.spatial_views = A; sh.non_spatial_floats = A + tags + J; }
  } else if (spec->kind == SUS_ENCODE_FLAT) {
    if (spec->n_components < 1 || spec->n_components > SUS_MAX_FLAT_COMPONENTS)
      return fail(SUS_ERR_INVALID_ARGUMENT, "flat encode needs 1..16 components");
    d.kind = SUS_ENCODE_FLAT;
    d.n_components = spec->n_components;
    int F = 0;
    for (int q = 0; q < spec->n_components; ++q) {
      const int n = component_size(c, spec->components[q]);
      if (n < 0) return fail(SUS_ERR_INVALID_ARGUMENT, "flat component " + std::to_string(spec->components[q]) + " is not defined for this env");
      d.components[q] = spec->components[q];
      F += n;
    }
    sh.non_spatial_floats = F;
    sh.non_spatial_views = 1;
  } else {
    return fail(SUS_ERR_INVALID_ARGUMENT, "unknown encode kind");
  }
  d.sp_floats = sh.spatial_floats;
  d.ns_floats = sh.non_spatial_floats;
  if (shape) *shape = sh;
  return SUS_OK;
}

inline int32_t align128(int64_t b) { return (int32_t)((b + 127) & ~127ll); }

struct DeviceInfo {
  int sms = 0;
  int max_dyn_smem = 0;
};

int device_info(int device, DeviceInfo& d) {
  static DeviceInfo cache[64];
  static bool have[64] = {};
  if (device >= 0 && device < 64 && have[device]) { d = cache[device]; return SUS_OK; }
  SUS_CUDA(cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, device));
  SUS_CUDA(cudaDeviceGetAttribute(&d.max_dyn_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  if (device >= 0 && device < 64) { cache[device] = d; have[device] = true; }
  return SUS_OK;
}

// false only for SUSNET_PATH=direct (register / LSU stores everywhere).  Which staged path runs is decided by want_ws() /
// want_staged_flat() below: unset -> warp-specialised emitter for Global / Perspective, byte-staged rows for Flat.
bool want_tma() {
  const char* v = std::getenv("SUSNET_PATH");
  return !(v && std::strcmp(v, "direct") == 0);
}

// Per-warp staging layout; returns false if not even two warps fit (then the direct path is used).
bool make_layout(const DevConfig& c, const DevEncode& enc, int rew_elem, bool want_nf, int max_dyn_smem, TileLayout& L) {
  if (enc.planes_u8) return false;  // byte planes: warp-specialised or direct path only
  const int views = enc.kind == SUS_ENCODE_NONE ? 0 : (enc.kind == SUS_ENCODE_FLAT ? 1 : c.A);
  TileLayout best = {};
  // tuning overrides (tools/kernel_sweep.py): SUSNET_TILE_G in {4, 8, 16}, SUSNET_TILE_WARPS in 2..8
  const char* env_g = std::getenv("SUSNET_TILE_G");
  const char* env_w = std::getenv("SUSNET_TILE_WARPS");
  const int force_g = env_g ? std::atoi(env_g) : 0, force_w = env_w ? std::atoi(env_w) : 0;
  for (int G : {8, 4, 16}) {
    if (force_g ? G != force_g : G == 16) continue;
    TileLayout t = {};
    t.G = G;
    int off = 0;
    t.sp_off = off; t.sp_bytes = align128((int64_t)G * enc.sp_floats * 4); off += t.sp_bytes;
    t.ns_off = off; t.ns_bytes = align128((int64_t)views * 32 * enc.ns_floats * 4); off += t.ns_bytes;
    t.rew_off = off; t.rew_bytes = align128((int64_t)32 * c.A * rew_elem); off += t.rew_bytes;
    t.nf_off = off; t.nf_bytes = want_nf ? align128((int64_t)32 * c.S * 4) : 0; off += t.nf_bytes;
    t.per_warp = off > 0 ? off : 128;
    const int budget = max_dyn_smem - 2048;  // static tables + slack
    t.warps = budget / t.per_warp;
    const char* env_mw = std::getenv("SUSNET_TILE_MAXWARPS");
    const int max_warps = env_mw ? std::atoi(env_mw) : kTmaMaxWarps;
    if (t.warps > max_warps) t.warps = max_warps;
    if (t.warps > kTmaMaxWarps) t.warps = kTmaMaxWarps;
    if (force_w > 0 && force_w < t.warps) t.warps = force_w;
    if (t.warps > best.warps) best = t;
    if (!force_g && best.warps >= 8) break;  // G = 8 tiles with >= 8 warps measured best; otherwise try G = 4
  }
  if (best.warps < 2) return false;
  L = best;
  return true;
}

// Staged-bytes layout of k_step_flat; false if a component is float-valued or kFlatMinCtas CTAs do not fit on an SM.
bool make_flat_stage(const DevConfig& c, const DevEncode& enc, int rew_elem, bool want_nf, int max_dyn_smem, FlatStage& L) {
  if (enc.kind != SUS_ENCODE_FLAT) return false;
  for (int q = 0; q < enc.n_components; ++q)
    if (enc.components[q] == SUS_FC_SCENT) return false;
  L.row_bytes = (32 * enc.ns_floats + 15) & ~15;
  L.rew_off = L.row_bytes;
  L.nf_off = L.rew_off + ((32 * c.A * rew_elem + 15) & ~15);
  L.per_warp = L.nf_off + (want_nf ? (32 * c.S * 4 + 15) & ~15 : 0);
  const size_t per_cta = (size_t)L.per_warp * (kThreads / 32) + 1024;  // + static tables and the driver's reserve
  return per_cta <= (size_t)max_dyn_smem && per_cta * kFlatMinCtas <= 227u * 1024u;
}

// Upper bound of the non-zero values of a flat row (the record size of k_step_flat_ws); -1 if a component is float-valued.
int flat_max_nonzeros(const DevConfig& c, const DevEncode& enc) {
  const int A = c.A, J = c.J;
  int k = 0;
  for (int q = 0; q < enc.n_components; ++q) {
    switch (enc.components[q]) {
      case SUS_FC_ONEHOT_POS: k += 2 * A; break;
      case SUS_FC_COORDS: k += 2 * A; break;
      case SUS_FC_ALIVE_CREW: k += A - 1; break;
      case SUS_FC_CLOSEST_CREW: k += 1; break;
      case SUS_FC_L1_CREW: k += A - 1; break;
      case SUS_FC_DIST_TO_IMPOSTER: k += 2 * (A - 1); break;
      case SUS_FC_WALLS: k += 9; break;
      case SUS_FC_ROOMS: k += 1 + (A - 1 < 4 ? A - 1 : 4); break;
      case SUS_FC_STATE_ALIVE: k += A; break;
      case SUS_FC_STATE_JOB_STATUS: k += J; break;
      case SUS_FC_STATE_USED_TAGS: k += A; break;
      case SUS_FC_STATE_TAG_COUNTS: k += A; break;
      default: return -1;  // SUS_FC_SCENT: float-valued
    }
  }
  return k;
}

// Layout of k_step_flat_ws; false if it does not apply (float-valued component, too many non-zeros per row, rows longer than
// the 10-bit offsets) or does not fit.  Measured at 1 Mi envs, cfg4-alt Flat-98 (tools/ab_flat.py): see DESIGN.md.
bool make_flat_ws_layout(const DevConfig& c, const DevEncode& enc, int rew_row_bytes, bool want_nf, int max_dyn_smem,
                         FlatWsLayout& L) {
  if (enc.kind != SUS_ENCODE_FLAT || enc.ns_floats >= 1023) return false;
  const int nz = flat_max_nonzeros(c, enc);
  if (nz < 0 || nz > 16) return false;
  FlatWsLayout t = {};
  t.K = (nz + 7) & ~7;
  if (t.K == 0) t.K = 8;
  t.tile_bytes = align128((int64_t)32 * enc.ns_floats * 4);
  t.rec_bytes = 32 * t.K * 2;
  t.rew_off = 0;
  t.nf_off = align128((int64_t)32 * rew_row_bytes);
  t.dense_bytes = t.nf_off + (want_nf ? align128((int64_t)32 * c.S * 4) : 0);
  const char* env_e = std::getenv("SUSNET_FLATWS_EMITTERS");
  const char* env_w = std::getenv("SUSNET_FLATWS_WARPS");
  t.emitter_warps = env_e && std::atoi(env_e) > 0 ? std::atoi(env_e) : 4;
  const int budget = max_dyn_smem - 2048 - t.emitter_warps * t.tile_bytes - 512;
  int cw = budget / (2 * t.rec_bytes + t.dense_bytes);
  if (cw > kFlatWsMaxWarps - t.emitter_warps) cw = kFlatWsMaxWarps - t.emitter_warps;
  if (env_w && std::atoi(env_w) > 0 && std::atoi(env_w) < cw) cw = std::atoi(env_w);
  if (cw < 2 * t.emitter_warps) return false;
  t.compute_warps = cw;
  t.recs_off = t.emitter_warps * t.tile_bytes;
  t.dense_off = t.recs_off + 2 * cw * t.rec_bytes;
  t.bars_off = t.dense_off + cw * t.dense_bytes;
  t.total_bytes = t.bars_off + 4 * cw * 8;
  L = t;
  return true;
}

// Warp-specialised layout (susnet_ws.cuh); returns false if it does not apply or does not fit.
// `fast_sink`: the plane tensor lives in L2-compressible memory (sus_alloc_compressible), where the tile stores drain
// ~18 % faster.  Measured at 1 Mi envs (tools/ab_flat.py under SUSNET_WS_TILE / SUSNET_WS_WARPS): Global into a fast sink
// is compute-bound and best with 8-env tiles + 7 compute warps (8 warps = 2 per scheduler: 0.381 ms against 0.403 with 6
// and 0.412 with 8); everything else is emitter-bound and best with 6 (Global into cudaMalloc memory 0.465 against
// 0.504 with 7, Perspective 0.399 against 0.414).
// The encode-only kernels pass step = false (their row-parsing warps are cheap).
bool make_ws_layout(const DevConfig& c, const DevEncode& enc, int rew_elem, bool want_nf, int max_dyn_smem, bool step,
                    bool fast_sink, WsLayout& L) {
  if (enc.sp_floats <= 0 || (enc.kind != SUS_ENCODE_GLOBAL && enc.kind != SUS_ENCODE_PERSPECTIVE)) return false;
  WsLayout t = {};
  const bool compute_bound = step && fast_sink && enc.kind == SUS_ENCODE_GLOBAL;
  const char* env_te = std::getenv("SUSNET_WS_TILE");
  const int elem = enc.planes_u8 ? 1 : 4;
  t.tile_envs = enc.planes_u8 ? 16 : (env_te ? (std::atoi(env_te) == 8 ? 8 : 16) : (compute_bound ? 8 : 16));
  t.tile_bytes = align128((int64_t)t.tile_envs * enc.sp_floats * elem);
  if (!enc.planes_u8 && t.tile_envs == 16 && max_dyn_smem - 2048 - 2 * t.tile_bytes - 512 < 4 * (32 * 32 + align128((int64_t)c.A * 32 * enc.ns_floats * 4) + align128((int64_t)32 * c.A * rew_elem))) {
    t.tile_envs = 8;  // big configs: fall back to 8-env tiles so that at least two compute warps fit
    t.tile_bytes = align128((int64_t)8 * enc.sp_floats * 4);
  }
  int off = 0;
  t.po_off = off; off += 32 * 32;
  t.ns_off = off; off += align128((int64_t)c.A * 32 * enc.ns_floats * 4);
  t.rew_off = off; off += align128((int64_t)32 * c.A * rew_elem);
  t.nf_off = off; off += want_nf ? align128((int64_t)32 * c.S * 4) : 0;
  t.slot_bytes = off;
  const int budget = max_dyn_smem - 2048 - 2 * t.tile_bytes - 512;
  int cw = budget / (2 * t.slot_bytes);
  if (cw > kWsMaxWarps - 1) cw = kWsMaxWarps - 1;
  const char* env_w = std::getenv("SUSNET_WS_WARPS");
  // encode-only launches: any count >= 5 is the same for Global (0.393-0.398 ms at 1 Mi rows); Perspective is best
  // with 5 (0.419 ms at 256 Ki rows against 0.427 with 7-8 and 0.459 with 11)
  // byte planes (4x fewer plane bytes): the step arithmetic binds, so every compute warp that fits is used
  const int want_cw = env_w && std::atoi(env_w) > 0 ? std::atoi(env_w)
                      : (enc.planes_u8 ? cw : (!step ? (enc.kind == SUS_ENCODE_PERSPECTIVE ? 5 : cw) : (compute_bound ? 7 : 6)));
  if (want_cw < cw) cw = want_cw;
  if (cw < 2) return false;
  t.compute_warps = cw;
  t.slots_off = 2 * t.tile_bytes;
  t.bars_off = t.slots_off + 2 * cw * t.slot_bytes;
  t.total_bytes = t.bars_off + 4 * cw * 8;
  L = t;
  return true;
}

// SUSNET_PATH forces an output path (the GPU tests run every case on all of them): "direct" = register/LSU stores,
// "tma" = every warp stages and bulk-stores its own tiles, "ws" = warp-specialised emitter where planes are written
// ("tma" elsewhere), "staged" = byte-staged rows for Flat encodes ("tma" elsewhere).  Unset: "ws" for Global /
// Perspective, "staged" for Flat.
bool want_staged_flat() {
  const char* v = std::getenv("SUSNET_PATH");
  return !v || std::strcmp(v, "staged") == 0;
}

bool want_ws() {
  const char* v = std::getenv("SUSNET_PATH");
  return !v || std::strcmp(v, "ws") == 0;
}

// Flat encodes: the warp-specialised record kernel (k_step_flat_ws) only when asked for.  Measured at 1 Mi envs, cfg4-alt
// Flat-98 (profiles/r02_flat98_ws_sweep.jsonl): 0.0881 ms with 4 emitter + 28 compute warps against 0.0881 ms for the staged
// bytes kernel (and 0.092-0.135 ms with fewer warps / emitters) -- halving the compute warps' instructions did not shorten a
// group's latency, which is what binds both kernels, so the simpler non-persistent staged kernel stays the default.
bool want_flat_ws() {
  const char* v = std::getenv("SUSNET_PATH");
  return v && std::strcmp(v, "ws") == 0;
}

unsigned persistent_grid(int64_t n_items, const TileLayout& L, int sms) {
  const int64_t groups = (n_items + 31) / 32;
  const int64_t ctas = (groups + L.warps - 1) / L.warps;
  return (unsigned)(ctas < sms ? ctas : sms);
}

unsigned ws_grid(int64_t n_items, const WsLayout& W, int sms) {
  const int64_t groups = (n_items + 31) / 32;
  const int64_t ctas = (groups + W.compute_warps - 1) / W.compute_warps;
  return (unsigned)(ctas < sms ? ctas : sms);
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is per kernel and per device: remember what was granted so the
// attribute is only set when it has to grow (keyed by the kernel's address; instantiations share a pointer TYPE).
template <typename K>
int allow_big_smem(K kernel, size_t bytes) {
  struct Entry { const void* fn; int dev; size_t bytes; };
  static std::mutex mu;
  static std::vector<Entry> granted;
  int dev = 0;
  SUS_CUDA(cudaGetDevice(&dev));
  const void* key = reinterpret_cast<const void*>(kernel);
  std::lock_guard<std::mutex> lock(mu);
  for (Entry& en : granted)
    if (en.fn == key && en.dev == dev) {
      if (bytes <= en.bytes) return SUS_OK;
      SUS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
      en.bytes = bytes;
      return SUS_OK;
    }
  SUS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  granted.push_back({key, dev, bytes});
  return SUS_OK;
}

inline unsigned grid_for(int64_t n) { return (unsigned)((n + kThreads - 1) / kThreads); }

int after_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(SUS_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  return SUS_OK;
}

}  // namespace

struct SusEnv {
  SusConfig cfg;
  DevConfig dc;
  int device;
  int64_t N;
  StateArrays st;
  unsigned long long* stats;
  uint32_t* err;
  double* ret;       // [A][N] + 2 trailing sums, allocated by sus_env_track_returns
  double gamma;
  uint64_t step_tick, reset_epoch, act_epoch;
  uint64_t* dev_ticks;  // [3] = {step_tick, reset_epoch, act_epoch} in device memory once sus_env_device_ticks enabled them,
                        // followed by three 32-bit CTA counters (see fetch_launch_tick)
  const uint32_t *inj_step, *inj_reset, *inj_act;
};

enum { TICK_STEP = 0, TICK_RESET = 1, TICK_ACT = 2 };

namespace {
inline unsigned int* tick_counter(SusEnv* e, int which) {
  return e->dev_ticks ? reinterpret_cast<unsigned int*>(e->dev_ticks + 3) + which : nullptr;
}

// device-resident ticks: a call that launches no kernel (empty batch, zero steps) still consumes its ticks
int advance_tick(SusEnv* e, int which, uint64_t n, cudaStream_t st) {
  if (!e->dev_ticks || n == 0) return SUS_OK;
  k_advance_tick<<<1, 1, 0, st>>>(e->dev_ticks + which, n);
  return after_launch("k_advance_tick");
}
}  // namespace

extern "C" int sus_internal_is_compressible(const void* p);  // susnet_alloc.cu

extern "C" {

// shared with the other translation units of the library (not part of the public ABI)
int sus_internal_fail(int code, const char* msg) { return fail(code, msg); }
void sus_internal_count_launch(void) { g_launches.fetch_add(1, std::memory_order_relaxed); }

// susnet_policy.cu: everything k_select_actions needs from a handle; consumes one act epoch (like sus_env_sample_actions)
int sus_internal_policy_params(sus_env_t e, DevConfig* c, StateArrays* st, uint64_t* tick, uint64_t** tick_dev,
                               unsigned int** tick_ctr, int64_t* N, int* device) {
  *c = e->dc; *st = e->st; *N = e->N; *device = e->device;
  *tick = e->act_epoch++;
  *tick_dev = e->dev_ticks ? e->dev_ticks + TICK_ACT : nullptr;
  *tick_ctr = tick_counter(e, TICK_ACT);
  if (e->N == 0) {
    DeviceGuard g(e->device);
    return advance_tick(e, TICK_ACT, 1, nullptr);
  }
  return SUS_OK;
}

int sus_abi_version(void) { return SUS_ABI_VERSION; }
const char* sus_last_error(void) { return g_last_error.c_str(); }
int64_t sus_launch_count(void) { return g_launches.load(); }

int sus_flat_state_size(const SusConfig* cfg) {
  if (!cfg) return fail(SUS_ERR_INVALID_ARGUMENT, "cfg is NULL");
  if (int rc = validate_config(*cfg)) return rc;
  return flat_size(*cfg);
}

int sus_n_role_actions(const SusConfig* cfg, int is_imposter) {
  if (!cfg) return fail(SUS_ERR_INVALID_ARGUMENT, "cfg is NULL");
  if (int rc = validate_config(*cfg)) return rc;
  return role_actions_host(*cfg, is_imposter);
}

int sus_compact_layout(const SusConfig* cfg, SusCompactLayout* out) {
  if (!cfg || !out) return fail(SUS_ERR_INVALID_ARGUMENT, "NULL argument");
  if (int rc = validate_config(*cfg)) return rc;
  compact_layout(*cfg, *out);
  return SUS_OK;
}

int sus_reward_lut(const SusConfig* cfg, double* out) {
  if (!cfg || !out) return fail(SUS_ERR_INVALID_ARGUMENT, "NULL argument");
  if (int rc = validate_config(*cfg)) return rc;
  SusCompactLayout L;
  compact_layout(*cfg, L);
  DevConfig d;
  make_dev_config(*cfg, d);
  const int stride = L.invalid_code + 1;
  for (int i = 0; i < d.A; ++i)
    for (int code = 0; code < stride; ++code)
      out[i * stride + code] = code < L.n_codes ? reward_of_code(d, i, (uint32_t)code) : std::nan("");
  return SUS_OK;
}

int sus_encode_shape(const SusConfig* cfg, const SusEncodeSpec* spec, SusEncodeShape* out) {
  if (!cfg || !spec || !out) return fail(SUS_ERR_INVALID_ARGUMENT, "NULL argument");
  if (int rc = validate_config(*cfg)) return rc;
  DevEncode d;
  return make_dev_encode(*cfg, spec, d, out);
}

int sus_env_create(const SusConfig* cfg, int device, sus_env_t* out) {
  if (!cfg || !out) return fail(SUS_ERR_INVALID_ARGUMENT, "NULL argument");
  if (int rc = validate_config(*cfg)) return rc;
  DeviceGuard g(device);
  if (!g.ok) return fail(SUS_ERR_CUDA, "cannot select CUDA device " + std::to_string(device));
  SusEnv* e = new SusEnv();
  e->cfg = *cfg;
  make_dev_config(*cfg, e->dc);
  e->device = device;
  e->N = cfg->num_envs;
  const size_t n = (size_t)(e->N > 0 ? e->N : 1);
  cudaError_t err = cudaSuccess;
  if (err == cudaSuccess) err = cudaMalloc(&e->st.pos, n * sizeof(uint64_t));
  if (err == cudaSuccess) err = cudaMalloc(&e->st.jobpos, n * sizeof(uint64_t));
  if (err == cudaSuccess) err = cudaMalloc(&e->st.aux, n * sizeof(uint4));
  if (err == cudaSuccess) err = cudaMalloc(&e->st.met, n * sizeof(uint4));
  if (err == cudaSuccess) err = cudaMalloc(&e->stats, SUS_N_STATS * sizeof(unsigned long long));
  if (err == cudaSuccess) err = cudaMalloc(&e->err, sizeof(uint32_t));
  if (err == cudaSuccess) err = cudaMemset(e->st.pos, 0, n * sizeof(uint64_t));
  if (err == cudaSuccess) err = cudaMemset(e->st.jobpos, 0, n * sizeof(uint64_t));
  if (err == cudaSuccess) err = cudaMemset(e->st.aux, 0, n * sizeof(uint4));
  if (err == cudaSuccess) err = cudaMemset(e->st.met, 0, n * sizeof(uint4));
  if (err == cudaSuccess) err = cudaMemset(e->stats, 0, SUS_N_STATS * sizeof(unsigned long long));
  if (err == cudaSuccess) err = cudaMemset(e->err, 0, sizeof(uint32_t));
  if (err == cudaSuccess) err = cudaDeviceSynchronize();
  if (err != cudaSuccess) {
    const std::string msg = std::string("allocating env state: ") + cudaGetErrorString(err);
    sus_env_destroy(e);
    return fail(SUS_ERR_CUDA, msg);
  }
  *out = e;
  return SUS_OK;
}

int sus_env_destroy(sus_env_t e) {
  if (!e) return SUS_OK;
  DeviceGuard g(e->device);
  cudaDeviceSynchronize();
  cudaFree(e->st.pos); cudaFree(e->st.jobpos); cudaFree(e->st.aux); cudaFree(e->st.met);
  cudaFree(e->stats); cudaFree(e->err); cudaFree(e->ret); cudaFree(e->dev_ticks);
  delete e;
  return SUS_OK;
}

int sus_env_reset(sus_env_t e, const uint8_t* mask, void* stream) {
  if (!e) return fail(SUS_ERR_INVALID_ARGUMENT, "env is NULL");
  DeviceGuard g(e->device);
  ResetParams p;
  p.c = e->dc; p.st = e->st; p.mask = mask; p.inj_reset = e->inj_reset; p.tick = e->reset_epoch++; p.N = e->N;
  p.tick_dev = e->dev_ticks ? e->dev_ticks + TICK_RESET : nullptr; p.tick_ctr = tick_counter(e, TICK_RESET);
  e->inj_reset = nullptr;
  if (e->N == 0) return advance_tick(e, TICK_RESET, 1, (cudaStream_t)stream);
  k_reset<<<grid_for(e->N), kThreads, 0, (cudaStream_t)stream>>>(p);
  return after_launch("k_reset");
}

static int step_launch(sus_env_t e, const SusStepIO* io, void* stream);

int sus_env_step(sus_env_t e, const SusStepIO* io, void* stream) {
  if (!e || !io) return fail(SUS_ERR_INVALID_ARGUMENT, "NULL argument");
  if (int rc = step_launch(e, io, stream)) return rc;
  if (e->N > 0) return SUS_OK;  // (the kernel advanced a device-resident tick itself)
  DeviceGuard g(e->device);
  return advance_tick(e, TICK_STEP, 1, (cudaStream_t)stream);
}

static int step_launch(sus_env_t e, const SusStepIO* io, void* stream) {
  if (io->actions && io->actions_dtype != SUS_U8 && io->actions_dtype != SUS_I32 && io->actions_dtype != SUS_I64 &&
      io->actions_dtype != SUS_PACKED)
    return fail(SUS_ERR_INVALID_ARGUMENT, "actions_dtype must be SUS_U8, SUS_I32, SUS_I64 or SUS_PACKED");
  if (io->packed_out && (io->rewards || io->done || io->truncated))
    return fail(SUS_ERR_INVALID_ARGUMENT, "packed_out replaces rewards / done / truncated: pass those as NULL");
  if (io->rewards && io->rewards_dtype != SUS_F32 && io->rewards_dtype != SUS_F64)
    return fail(SUS_ERR_INVALID_ARGUMENT, "rewards_dtype must be SUS_F32 or SUS_F64");
  DeviceGuard g(e->device);
  StepParams p;
  std::memset(&p, 0, sizeof(p));
  p.c = e->dc;
  if (int rc = make_dev_encode(e->cfg, io->encode, p.enc, nullptr)) return rc;
  const bool enc = p.enc.kind != SUS_ENCODE_NONE;
  if (enc && e->N > 0 && (!io->non_spatial || (p.enc.sp_floats > 0 && !io->spatial)))
    return fail(SUS_ERR_INVALID_ARGUMENT, "fused encode requested without output tensors");
  p.st = e->st;
  p.actions = io->actions; p.actions_dtype = io->actions_dtype; p.rewards = io->rewards; p.rewards_dtype = io->rewards_dtype;
  p.done = io->done; p.trunc = io->truncated; p.actions_out = io->actions_out; p.next_flat = io->next_flat;
  p.metrics = reinterpret_cast<long long*>(io->metrics); p.imposters = io->imposters; p.spatial = io->spatial; p.non_spatial = io->non_spatial;
  p.inj_step = e->inj_step; p.inj_reset = e->inj_reset; p.inj_act = e->inj_act;
  p.stats = e->stats; p.err = e->err; p.tick = e->step_tick++; p.N = e->N;
  p.tick_dev = e->dev_ticks ? e->dev_ticks + TICK_STEP : nullptr; p.tick_ctr = tick_counter(e, TICK_STEP);
  p.ret = e->ret; p.ret_sums = e->ret ? e->ret + (size_t)e->N * p.c.A : nullptr; p.gamma = e->gamma;
  {
    SusCompactLayout cl;
    compact_layout(e->cfg, cl);
    p.packed_out = io->packed_out;
    p.action_bits = cl.action_bits; p.action_bytes = cl.action_bytes; p.reward_bits = cl.reward_bits; p.result_bytes = cl.result_bytes;
  }
  e->inj_step = e->inj_reset = e->inj_act = nullptr;
  if (e->N == 0) return SUS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  DeviceInfo di;
  if (int rc = device_info(e->device, di)) return rc;
  TileLayout L;
  // step-only launches move < 100 B per env and are latency/issue bound: the one-thread-per-env kernel with its
  // higher occupancy wins there (measured 14.6e9 vs 7.3e9 env-steps/s); the TMA path pays off once features are written
  WsLayout W;
  if ((want_ws() || (p.enc.planes_u8 && want_tma())) && enc &&
      make_ws_layout(p.c, p.enc, io->rewards_dtype == SUS_F64 ? 8 : 4, io->next_flat != nullptr,
                                         di.max_dyn_smem, true, sus_internal_is_compressible(io->spatial) != 0, W)) {
    const int64_t groups = (e->N + 31) / 32;
    const int64_t ctas = (groups + W.compute_warps - 1) / W.compute_warps;
    const unsigned gr = (unsigned)(ctas < di.sms ? ctas : di.sms);
    const unsigned threads = (unsigned)(W.compute_warps + 1) * 32;
    switch (e->cfg.variant) {
      case SUS_VARIANT_BASE:
        if (int rc = allow_big_smem(k_step_ws<SUS_VARIANT_BASE>, W.total_bytes)) return rc;
        k_step_ws<SUS_VARIANT_BASE><<<gr, threads, W.total_bytes, st>>>(p, W);
        break;
      case SUS_VARIANT_TAGGING:
        if (int rc = allow_big_smem(k_step_ws<SUS_VARIANT_TAGGING>, W.total_bytes)) return rc;
        k_step_ws<SUS_VARIANT_TAGGING><<<gr, threads, W.total_bytes, st>>>(p, W);
        break;
      default:
        if (int rc = allow_big_smem(k_step_ws<SUS_VARIANT_TRAINING_GROUND>, W.total_bytes)) return rc;
        k_step_ws<SUS_VARIANT_TRAINING_GROUND><<<gr, threads, W.total_bytes, st>>>(p, W);
        break;
    }
    return after_launch("k_step_ws");
  }
  FlatWsLayout FW;
  if (want_flat_ws() && enc && make_flat_ws_layout(p.c, p.enc, io->packed_out ? p.result_bytes : p.c.A * (io->rewards_dtype == SUS_F64 ? 8 : 4),
                                              io->next_flat != nullptr, di.max_dyn_smem, FW)) {
    const int64_t groups = (e->N + 31) / 32;
    const int64_t ctas = (groups + FW.compute_warps - 1) / FW.compute_warps;
    const unsigned gr = (unsigned)(ctas < di.sms ? ctas : di.sms);
    const unsigned threads = (unsigned)(FW.compute_warps + FW.emitter_warps) * 32;
    switch (e->cfg.variant) {
      case SUS_VARIANT_BASE:
        if (int rc = allow_big_smem(k_step_flat_ws<SUS_VARIANT_BASE>, FW.total_bytes)) return rc;
        k_step_flat_ws<SUS_VARIANT_BASE><<<gr, threads, FW.total_bytes, st>>>(p, FW);
        break;
      case SUS_VARIANT_TAGGING:
        if (int rc = allow_big_smem(k_step_flat_ws<SUS_VARIANT_TAGGING>, FW.total_bytes)) return rc;
        k_step_flat_ws<SUS_VARIANT_TAGGING><<<gr, threads, FW.total_bytes, st>>>(p, FW);
        break;
      default:
        if (p.c.A == 5 && p.c.J == 0) {  // the reference's training shape: compile-time agent / job counts
          if (int rc = allow_big_smem(k_step_flat_ws<SUS_VARIANT_TRAINING_GROUND, 5, 0>, FW.total_bytes)) return rc;
          k_step_flat_ws<SUS_VARIANT_TRAINING_GROUND, 5, 0><<<gr, threads, FW.total_bytes, st>>>(p, FW);
        } else {
          if (int rc = allow_big_smem(k_step_flat_ws<SUS_VARIANT_TRAINING_GROUND>, FW.total_bytes)) return rc;
          k_step_flat_ws<SUS_VARIANT_TRAINING_GROUND><<<gr, threads, FW.total_bytes, st>>>(p, FW);
        }
        break;
    }
    return after_launch("k_step_flat_ws");
  }
  FlatStage FS;
  if (want_staged_flat() && enc &&
      make_flat_stage(p.c, p.enc, io->rewards_dtype == SUS_F64 ? 8 : 4, io->next_flat != nullptr, di.max_dyn_smem, FS)) {
    const size_t smem = (size_t)FS.per_warp * (kThreads / 32);
    const unsigned gr = grid_for(e->N);
    switch (e->cfg.variant) {
      case SUS_VARIANT_BASE:
        if (int rc = allow_big_smem(k_step_flat<SUS_VARIANT_BASE>, smem)) return rc;
        k_step_flat<SUS_VARIANT_BASE><<<gr, kThreads, smem, st>>>(p, FS);
        break;
      case SUS_VARIANT_TAGGING:
        if (int rc = allow_big_smem(k_step_flat<SUS_VARIANT_TAGGING>, smem)) return rc;
        k_step_flat<SUS_VARIANT_TAGGING><<<gr, kThreads, smem, st>>>(p, FS);
        break;
      default:
        if (p.c.A == 5 && p.c.J == 0) {  // the reference's training shape: compile-time agent / job counts
          if (int rc = allow_big_smem(k_step_flat<SUS_VARIANT_TRAINING_GROUND, 5, 0>, smem)) return rc;
          k_step_flat<SUS_VARIANT_TRAINING_GROUND, 5, 0><<<gr, kThreads, smem, st>>>(p, FS);
        } else {
          if (int rc = allow_big_smem(k_step_flat<SUS_VARIANT_TRAINING_GROUND>, smem)) return rc;
          k_step_flat<SUS_VARIANT_TRAINING_GROUND><<<gr, kThreads, smem, st>>>(p, FS);
        }
        break;
    }
    return after_launch("k_step_flat");
  }
  if (want_tma() && enc &&
      make_layout(p.c, p.enc, io->rewards_dtype == SUS_F64 ? 8 : 4, io->next_flat != nullptr, di.max_dyn_smem, L)) {
    const size_t smem = (size_t)L.per_warp * L.warps;
    const unsigned gr = persistent_grid(e->N, L, di.sms);
#define SUS_LAUNCH_STEP_TMA(V)                                                        \
  if (enc) {                                                                          \
    if (int rc = allow_big_smem(k_step_tma<V, true>, smem)) return rc;                \
    k_step_tma<V, true><<<gr, L.warps * 32, smem, st>>>(p, L);                        \
  } else {                                                                            \
    if (int rc = allow_big_smem(k_step_tma<V, false>, smem)) return rc;               \
    k_step_tma<V, false><<<gr, L.warps * 32, smem, st>>>(p, L);                       \
  }
    switch (e->cfg.variant) {
      case SUS_VARIANT_BASE: SUS_LAUNCH_STEP_TMA(SUS_VARIANT_BASE); break;
      case SUS_VARIANT_TAGGING: SUS_LAUNCH_STEP_TMA(SUS_VARIANT_TAGGING); break;
      default: SUS_LAUNCH_STEP_TMA(SUS_VARIANT_TRAINING_GROUND); break;
    }
#undef SUS_LAUNCH_STEP_TMA
    return after_launch("k_step_tma");
  }
  const unsigned gr = grid_for(e->N);
  const int dA = p.c.A, dJ = p.c.J;
  // step-only launches of the common shapes use instantiations with compile-time agent / job counts
#define SUS_LAUNCH_STEP_ONLY(V)                                                                   \
  if (dA == 5 && dJ == 5) k_step<V, false, 5, 5><<<gr, kThreads, 0, st>>>(p);                     \
  else if (dA == 3 && dJ == 5) k_step<V, false, 3, 5><<<gr, kThreads, 0, st>>>(p);                \
  else k_step<V, false><<<gr, kThreads, 0, st>>>(p)
#define SUS_LAUNCH_STEP(V)                                      \
  if (enc) k_step<V, true><<<gr, kThreads, 0, st>>>(p);         \
  else { SUS_LAUNCH_STEP_ONLY(V); }
  switch (e->cfg.variant) {
    case SUS_VARIANT_BASE: SUS_LAUNCH_STEP(SUS_VARIANT_BASE); break;
    case SUS_VARIANT_TAGGING: SUS_LAUNCH_STEP(SUS_VARIANT_TAGGING); break;
    default:
      if (enc) k_step<SUS_VARIANT_TRAINING_GROUND, true><<<gr, kThreads, 0, st>>>(p);
      else if (dA == 2 && dJ == 0) k_step<SUS_VARIANT_TRAINING_GROUND, false, 2, 0><<<gr, kThreads, 0, st>>>(p);
      else if (dA == 5 && dJ == 0) k_step<SUS_VARIANT_TRAINING_GROUND, false, 5, 0><<<gr, kThreads, 0, st>>>(p);
      else k_step<SUS_VARIANT_TRAINING_GROUND, false><<<gr, kThreads, 0, st>>>(p);
      break;
  }
#undef SUS_LAUNCH_STEP
#undef SUS_LAUNCH_STEP_ONLY
  return after_launch("k_step");
}

int sus_env_rollout(sus_env_t e, int32_t n_steps, double* reward_sums, void* stream) {
  if (!e) return fail(SUS_ERR_INVALID_ARGUMENT, "env is NULL");
  if (n_steps < 0) return fail(SUS_ERR_INVALID_ARGUMENT, "n_steps < 0");
  if (e->inj_step || e->inj_reset || e->inj_act) return fail(SUS_ERR_INVALID_ARGUMENT, "rollout does not take injected words");
  if (e->ret) return fail(SUS_ERR_UNSUPPORTED, "rollout does not maintain the tracked returns; use sus_env_step");
  DeviceGuard g(e->device);
  RolloutParams p;
  p.c = e->dc; p.st = e->st; p.stats = e->stats; p.reward_sums = reward_sums; p.tick0 = e->step_tick; p.N = e->N;
  p.tick_dev = e->dev_ticks ? e->dev_ticks + TICK_STEP : nullptr; p.tick_ctr = tick_counter(e, TICK_STEP);
  p.n_steps = n_steps;
  e->step_tick += (uint64_t)n_steps;
  if (e->N == 0 || n_steps == 0) return advance_tick(e, TICK_STEP, (uint64_t)n_steps, (cudaStream_t)stream);
  const unsigned gr = grid_for(e->N);
  cudaStream_t st = (cudaStream_t)stream;
  const int dA = p.c.A, dJ = p.c.J;
#define SUS_LAUNCH_ROLLOUT(V)                                                             \
  if (dA == 5 && dJ == 5) k_rollout<V, 5, 5><<<gr, kThreads, 0, st>>>(p);                 \
  else if (dA == 3 && dJ == 5) k_rollout<V, 3, 5><<<gr, kThreads, 0, st>>>(p);            \
  else k_rollout<V><<<gr, kThreads, 0, st>>>(p)
  switch (e->cfg.variant) {
    case SUS_VARIANT_BASE: SUS_LAUNCH_ROLLOUT(SUS_VARIANT_BASE); break;
    case SUS_VARIANT_TAGGING: SUS_LAUNCH_ROLLOUT(SUS_VARIANT_TAGGING); break;
    default:
      if (dA == 2 && dJ == 0) k_rollout<SUS_VARIANT_TRAINING_GROUND, 2, 0><<<gr, kThreads, 0, st>>>(p);
      else if (dA == 5 && dJ == 0) k_rollout<SUS_VARIANT_TRAINING_GROUND, 5, 0><<<gr, kThreads, 0, st>>>(p);
      else k_rollout<SUS_VARIANT_TRAINING_GROUND><<<gr, kThreads, 0, st>>>(p);
      break;
  }
#undef SUS_LAUNCH_ROLLOUT
  return after_launch("k_rollout");
}

int sus_env_check_actions(sus_env_t e, void* stream) {
  if (!e) return fail(SUS_ERR_INVALID_ARGUMENT, "env is NULL");
  DeviceGuard g(e->device);
  uint32_t n = 0;
  SUS_CUDA(cudaMemcpyAsync(&n, e->err, sizeof(n), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  SUS_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  if (n) {
    SUS_CUDA(cudaMemsetAsync(e->err, 0, sizeof(uint32_t), (cudaStream_t)stream));
    return fail(SUS_ERR_INVALID_ACTION, std::to_string(n) + " env step(s) received an action index outside the agent's role list");
  }
  return SUS_OK;
}

int sus_env_sample_actions(sus_env_t e, int32_t* out, void* stream) {
  if (!e) return fail(SUS_ERR_INVALID_ARGUMENT, "env is NULL");
  DeviceGuard g(e->device);
  if (e->N == 0) { e->act_epoch++; e->inj_act = nullptr; return advance_tick(e, TICK_ACT, 1, (cudaStream_t)stream); }
  if (!out) return fail(SUS_ERR_INVALID_ARGUMENT, "NULL argument");
  ActParams p;
  p.c = e->dc; p.st = e->st; p.out = out; p.inj_act = e->inj_act; p.tick = e->act_epoch++; p.N = e->N;
  p.tick_dev = e->dev_ticks ? e->dev_ticks + TICK_ACT : nullptr; p.tick_ctr = tick_counter(e, TICK_ACT);
  e->inj_act = nullptr;
  k_sample_actions<<<grid_for(e->N), kThreads, 0, (cudaStream_t)stream>>>(p);
  return after_launch("k_sample_actions");
}

int sus_env_export_flat(sus_env_t e, int32_t dtype, void* out, void* stream) {
  if (!e) return fail(SUS_ERR_INVALID_ARGUMENT, "env is NULL");
  if (e->N == 0) return SUS_OK;
  if (!out) return fail(SUS_ERR_INVALID_ARGUMENT, "NULL argument");
  DeviceGuard g(e->device);
  ExportParams p;
  p.c = e->dc; p.st = e->st; p.out = out; p.N = e->N;
  if (e->N == 0) return SUS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == SUS_F32) k_export_flat<float><<<grid_for(e->N), kThreads, 0, st>>>(p);
  else if (dtype == SUS_F64) k_export_flat<double><<<grid_for(e->N), kThreads, 0, st>>>(p);
  else if (dtype == SUS_I64) k_export_flat<long long><<<grid_for(e->N), kThreads, 0, st>>>(p);
  else return fail(SUS_ERR_INVALID_ARGUMENT, "export dtype must be SUS_F32, SUS_F64 or SUS_I64");
  return after_launch("k_export_flat");
}

int sus_env_import_flat(sus_env_t e, const int64_t* flat, const uint8_t* imposter_mask, const int32_t* t, void* stream) {
  if (!e) return fail(SUS_ERR_INVALID_ARGUMENT, "env is NULL");
  if (e->N == 0) return SUS_OK;
  if (!flat || !imposter_mask) return fail(SUS_ERR_INVALID_ARGUMENT, "NULL argument");
  DeviceGuard g(e->device);
  ImportParams p;
  p.c = e->dc; p.st = e->st; p.flat = reinterpret_cast<const long long*>(flat); p.imp = imposter_mask; p.t = t; p.N = e->N;
  if (e->N == 0) return SUS_OK;
  k_import_flat<<<grid_for(e->N), kThreads, 0, (cudaStream_t)stream>>>(p);
  return after_launch("k_import_flat");
}

int sus_env_export_imposter_mask(sus_env_t e, uint8_t* out, void* stream) {
  if (!e) return fail(SUS_ERR_INVALID_ARGUMENT, "env is NULL");
  if (e->N == 0) return SUS_OK;
  if (!out) return fail(SUS_ERR_INVALID_ARGUMENT, "NULL argument");
  DeviceGuard g(e->device);
  ExportParams p;
  p.c = e->dc; p.st = e->st; p.out = out; p.N = e->N;
  if (e->N == 0) return SUS_OK;
  k_export_imp<<<grid_for(e->N), kThreads, 0, (cudaStream_t)stream>>>(p);
  return after_launch("k_export_imp");
}

int sus_env_export_metrics(sus_env_t e, int64_t* out, void* stream) {
  if (!e) return fail(SUS_ERR_INVALID_ARGUMENT, "env is NULL");
  if (e->N == 0) return SUS_OK;
  if (!out) return fail(SUS_ERR_INVALID_ARGUMENT, "NULL argument");
  DeviceGuard g(e->device);
  ExportParams p;
  p.c = e->dc; p.st = e->st; p.out = out; p.N = e->N;
  if (e->N == 0) return SUS_OK;
  k_export_metrics<<<grid_for(e->N), kThreads, 0, (cudaStream_t)stream>>>(p);
  return after_launch("k_export_metrics");
}

int sus_env_encode(sus_env_t e, const SusEncodeSpec* spec, void* spatial, float* non_spatial, void* stream) {
  if (!e || !spec) return fail(SUS_ERR_INVALID_ARGUMENT, "NULL argument");
  DeviceGuard g(e->device);
  EncodeParams p;
  std::memset(&p, 0, sizeof(p));
  p.c = e->dc;
  if (int rc = make_dev_encode(e->cfg, spec, p.enc, nullptr)) return rc;
  if (p.enc.kind == SUS_ENCODE_NONE) return fail(SUS_ERR_INVALID_ARGUMENT, "encode kind is NONE");
  if (e->N == 0) return SUS_OK;
  if (!non_spatial || (p.enc.sp_floats > 0 && !spatial)) return fail(SUS_ERR_INVALID_ARGUMENT, "missing output tensor");
  p.st = e->st; p.spatial = spatial; p.non_spatial = non_spatial; p.n_items = e->N;
  if (e->N == 0) return SUS_OK;
  DeviceInfo di;
  if (int rc = device_info(e->device, di)) return rc;
  WsLayout W;
  if ((want_ws() || (p.enc.planes_u8 && want_tma())) && make_ws_layout(p.c, p.enc, 0, false, di.max_dyn_smem, false, false, W)) {
    if (int rc = allow_big_smem(k_encode_ws<float, false>, W.total_bytes)) return rc;
    k_encode_ws<float, false><<<ws_grid(e->N, W, di.sms), (W.compute_warps + 1) * 32, W.total_bytes, (cudaStream_t)stream>>>(p, W);
    return after_launch("k_encode_ws");
  }
  FlatStage FS;
  if (want_staged_flat() && make_flat_stage(p.c, p.enc, 0, false, di.max_dyn_smem, FS)) {
    const size_t smem = (size_t)FS.per_warp * (kThreads / 32);
    if (int rc = allow_big_smem(k_encode_flat<float, false>, smem)) return rc;
    k_encode_flat<float, false><<<grid_for(e->N), kThreads, smem, (cudaStream_t)stream>>>(p, FS);
    return after_launch("k_encode_flat");
  }
  TileLayout L;
  if (want_tma() && make_layout(p.c, p.enc, 0, false, di.max_dyn_smem, L)) {
    const size_t smem = (size_t)L.per_warp * L.warps;
    if (int rc = allow_big_smem(k_encode_tma<float, false>, smem)) return rc;
    k_encode_tma<float, false><<<persistent_grid(e->N, L, di.sms), L.warps * 32, smem, (cudaStream_t)stream>>>(p, L);
    return after_launch("k_encode_tma");
  }
  k_encode_env<<<grid_for(e->N), kThreads, 0, (cudaStream_t)stream>>>(p);
  return after_launch("k_encode_env");
}

int sus_encode_from_flat(const SusConfig* cfg, const SusEncodeSpec* spec, const void* states, int32_t dtype,
                         int64_t n_items, void* spatial, float* non_spatial, int device, void* stream) {
  if (!cfg || !spec) return fail(SUS_ERR_INVALID_ARGUMENT, "NULL argument");
  if (int rc = validate_config(*cfg)) return rc;
  if (n_items > 0 && !states) return fail(SUS_ERR_INVALID_ARGUMENT, "states is NULL");
  if (n_items < 0) return fail(SUS_ERR_INVALID_ARGUMENT, "n_items < 0");
  DeviceGuard g(device);
  if (!g.ok) return fail(SUS_ERR_CUDA, "cannot select CUDA device " + std::to_string(device));
  EncodeParams p;
  std::memset(&p, 0, sizeof(p));
  make_dev_config(*cfg, p.c);
  if (int rc = make_dev_encode(*cfg, spec, p.enc, nullptr)) return rc;
  if (p.enc.kind == SUS_ENCODE_NONE) return fail(SUS_ERR_INVALID_ARGUMENT, "encode kind is NONE");
  if (n_items == 0) return SUS_OK;
  if (!non_spatial || (p.enc.sp_floats > 0 && !spatial)) return fail(SUS_ERR_INVALID_ARGUMENT, "missing output tensor");
  p.rows = states; p.spatial = spatial; p.non_spatial = non_spatial; p.n_items = n_items;
  if (n_items == 0) return SUS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype != SUS_F32 && dtype != SUS_F64 && dtype != SUS_I64)
    return fail(SUS_ERR_INVALID_ARGUMENT, "states dtype must be SUS_F32, SUS_F64 or SUS_I64");
  DeviceInfo di;
  if (int rc = device_info(device, di)) return rc;
  WsLayout W;
  if ((want_ws() || (p.enc.planes_u8 && want_tma())) && make_ws_layout(p.c, p.enc, 0, false, di.max_dyn_smem, false, false, W)) {
    const unsigned gr = ws_grid(n_items, W, di.sms), threads = (unsigned)(W.compute_warps + 1) * 32;
    if (dtype == SUS_F32) {
      if (int rc = allow_big_smem(k_encode_ws<float, true>, W.total_bytes)) return rc;
      k_encode_ws<float, true><<<gr, threads, W.total_bytes, st>>>(p, W);
    } else if (dtype == SUS_F64) {
      if (int rc = allow_big_smem(k_encode_ws<double, true>, W.total_bytes)) return rc;
      k_encode_ws<double, true><<<gr, threads, W.total_bytes, st>>>(p, W);
    } else {
      if (int rc = allow_big_smem(k_encode_ws<long long, true>, W.total_bytes)) return rc;
      k_encode_ws<long long, true><<<gr, threads, W.total_bytes, st>>>(p, W);
    }
    return after_launch("k_encode_ws");
  }
  FlatStage FS;
  if (want_staged_flat() && make_flat_stage(p.c, p.enc, 0, false, di.max_dyn_smem, FS)) {
    const size_t smem = (size_t)FS.per_warp * (kThreads / 32);
    const unsigned gr = grid_for(n_items);
    if (dtype == SUS_F32) {
      if (int rc = allow_big_smem(k_encode_flat<float, true>, smem)) return rc;
      k_encode_flat<float, true><<<gr, kThreads, smem, st>>>(p, FS);
    } else if (dtype == SUS_F64) {
      if (int rc = allow_big_smem(k_encode_flat<double, true>, smem)) return rc;
      k_encode_flat<double, true><<<gr, kThreads, smem, st>>>(p, FS);
    } else {
      if (int rc = allow_big_smem(k_encode_flat<long long, true>, smem)) return rc;
      k_encode_flat<long long, true><<<gr, kThreads, smem, st>>>(p, FS);
    }
    return after_launch("k_encode_flat");
  }
  TileLayout L;
  if (want_tma() && make_layout(p.c, p.enc, 0, false, di.max_dyn_smem, L)) {
    const size_t smem = (size_t)L.per_warp * L.warps;
    const unsigned gr = persistent_grid(n_items, L, di.sms);
    if (dtype == SUS_F32) {
      if (int rc = allow_big_smem(k_encode_tma<float, true>, smem)) return rc;
      k_encode_tma<float, true><<<gr, L.warps * 32, smem, st>>>(p, L);
    } else if (dtype == SUS_F64) {
      if (int rc = allow_big_smem(k_encode_tma<double, true>, smem)) return rc;
      k_encode_tma<double, true><<<gr, L.warps * 32, smem, st>>>(p, L);
    } else {
      if (int rc = allow_big_smem(k_encode_tma<long long, true>, smem)) return rc;
      k_encode_tma<long long, true><<<gr, L.warps * 32, smem, st>>>(p, L);
    }
    return after_launch("k_encode_tma");
  }
  if (dtype == SUS_F32) k_encode_rows<float><<<grid_for(n_items), kThreads, 0, st>>>(p);
  else if (dtype == SUS_F64) k_encode_rows<double><<<grid_for(n_items), kThreads, 0, st>>>(p);
  else k_encode_rows<long long><<<grid_for(n_items), kThreads, 0, st>>>(p);
  return after_launch("k_encode_rows");
}

int sus_env_stats(sus_env_t e, int64_t* out, void* stream) {
  if (!e || !out) return fail(SUS_ERR_INVALID_ARGUMENT, "NULL argument");
  DeviceGuard g(e->device);
  SUS_CUDA(cudaMemcpyAsync(out, e->stats, SUS_N_STATS * sizeof(int64_t), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return SUS_OK;
}

int sus_env_clear_stats(sus_env_t e, void* stream) {
  if (!e) return fail(SUS_ERR_INVALID_ARGUMENT, "env is NULL");
  DeviceGuard g(e->device);
  SUS_CUDA(cudaMemsetAsync(e->stats, 0, SUS_N_STATS * sizeof(int64_t), (cudaStream_t)stream));
  return SUS_OK;
}

int sus_env_track_returns(sus_env_t e, double gamma, void* stream) {
  if (!e) return fail(SUS_ERR_INVALID_ARGUMENT, "env is NULL");
  DeviceGuard g(e->device);
  const size_t n = ((size_t)e->N * e->dc.A + 2) * sizeof(double);
  if (!e->ret) SUS_CUDA(cudaMalloc(&e->ret, n));
  SUS_CUDA(cudaMemsetAsync(e->ret, 0, n, (cudaStream_t)stream));
  e->gamma = gamma;
  return SUS_OK;
}

int sus_env_return_sums(sus_env_t e, double* out, void* stream) {
  if (!e || !out) return fail(SUS_ERR_INVALID_ARGUMENT, "NULL argument");
  if (!e->ret) return fail(SUS_ERR_INVALID_ARGUMENT, "returns are not tracked: call sus_env_track_returns first");
  DeviceGuard g(e->device);
  SUS_CUDA(cudaMemcpyAsync(out, e->ret + (size_t)e->N * e->dc.A, 2 * sizeof(double), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return SUS_OK;
}

int sus_env_device_ticks(sus_env_t e, int32_t enable, void* stream) {
  if (!e) return fail(SUS_ERR_INVALID_ARGUMENT, "env is NULL");
  DeviceGuard g(e->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (enable && !e->dev_ticks) {
    uint64_t* d = nullptr;
    SUS_CUDA(cudaMalloc(&d, 5 * sizeof(uint64_t)));  // three ticks + three 32-bit CTA counters
    const uint64_t h[5] = {e->step_tick, e->reset_epoch, e->act_epoch, 0, 0};
    SUS_CUDA(cudaMemcpyAsync(d, h, sizeof(h), cudaMemcpyHostToDevice, st));  // pageable source: staged before the call returns
    SUS_CUDA(cudaStreamSynchronize(st));
    e->dev_ticks = d;
  } else if (!enable && e->dev_ticks) {
    uint64_t h[3];
    SUS_CUDA(cudaDeviceSynchronize());
    SUS_CUDA(cudaMemcpy(h, e->dev_ticks, sizeof(h), cudaMemcpyDeviceToHost));
    e->step_tick = h[0]; e->reset_epoch = h[1]; e->act_epoch = h[2];
    SUS_CUDA(cudaFree(e->dev_ticks));
    e->dev_ticks = nullptr;
  }
  return SUS_OK;
}

int sus_env_get_ticks(sus_env_t e, uint64_t* step_tick, uint64_t* reset_epoch, uint64_t* act_epoch) {
  if (!e) return fail(SUS_ERR_INVALID_ARGUMENT, "env is NULL");
  if (e->dev_ticks) {  // the device holds the truth (graph replays advance it without the host): synchronises
    DeviceGuard g(e->device);
    uint64_t h[3];
    SUS_CUDA(cudaDeviceSynchronize());
    SUS_CUDA(cudaMemcpy(h, e->dev_ticks, sizeof(h), cudaMemcpyDeviceToHost));
    e->step_tick = h[0]; e->reset_epoch = h[1]; e->act_epoch = h[2];
  }
  if (step_tick) *step_tick = e->step_tick;
  if (reset_epoch) *reset_epoch = e->reset_epoch;
  if (act_epoch) *act_epoch = e->act_epoch;
  return SUS_OK;
}

int sus_env_set_ticks(sus_env_t e, uint64_t step_tick, uint64_t reset_epoch, uint64_t act_epoch) {
  if (!e) return fail(SUS_ERR_INVALID_ARGUMENT, "env is NULL");
  e->step_tick = step_tick; e->reset_epoch = reset_epoch; e->act_epoch = act_epoch;
  if (e->dev_ticks) {
    DeviceGuard g(e->device);
    const uint64_t h[3] = {step_tick, reset_epoch, act_epoch};
    SUS_CUDA(cudaDeviceSynchronize());
    SUS_CUDA(cudaMemcpy(e->dev_ticks, h, sizeof(h), cudaMemcpyHostToDevice));
  }
  return SUS_OK;
}

int sus_env_state_arrays(sus_env_t e, void** ptrs, int32_t* bytes_per_env) {
  if (!e || !ptrs || !bytes_per_env) return fail(SUS_ERR_INVALID_ARGUMENT, "NULL argument");
  ptrs[0] = e->st.pos; ptrs[1] = e->st.jobpos; ptrs[2] = e->st.aux; ptrs[3] = e->st.met;
  bytes_per_env[0] = 8; bytes_per_env[1] = 8; bytes_per_env[2] = 16; bytes_per_env[3] = 16;
  return SUS_OK;
}

int sus_env_aux_arrays(sus_env_t e, void** ptrs, int64_t* bytes) {
  if (!e || !ptrs || !bytes) return fail(SUS_ERR_INVALID_ARGUMENT, "NULL argument");
  ptrs[0] = e->stats; bytes[0] = SUS_N_STATS * (int64_t)sizeof(unsigned long long);
  ptrs[1] = e->err; bytes[1] = sizeof(uint32_t);
  ptrs[2] = e->ret; bytes[2] = e->ret ? (int64_t)(((size_t)e->N * e->dc.A + 2) * sizeof(double)) : 0;
  return SUS_OK;
}

int sus_env_debug_inject_words(sus_env_t e, const uint32_t* step_words, const uint32_t* reset_words,
                               const uint32_t* act_words) {
  if (!e) return fail(SUS_ERR_INVALID_ARGUMENT, "env is NULL");
  if (step_words) e->inj_step = step_words;
  if (reset_words) e->inj_reset = reset_words;
  if (act_words) e->inj_act = act_words;
  return SUS_OK;
}

}  // extern "C"
