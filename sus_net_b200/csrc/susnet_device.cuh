// susnet_device.cuh -- device-side game logic of the B200-native Sus-Net simulator.
//
// One thread owns one environment.  The whole mutable game state of an env lives in a handful of
// packed 32/64-bit registers (positions as one byte per agent, liveness/role/job flags as bitmasks),
// so the sequential per-agent loop of the reference (src/environment/base.py:377-382) indexes agents
// with shifts instead of local-memory arrays.  Global memory holds the same words as a structure of
// arrays: one u64 (positions), one u64 (job positions), and two 128-bit records per env, each read
// and written with a single fully coalesced vector access per warp.
//
// Semantics follow the reference line by line; anchors are given at each block.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/susnet_b200.h"

namespace susnet {

// ---------------------------------------------------------------------------------------------
// configuration as the kernels see it (passed by value as a __grid_constant__ kernel parameter)
// ---------------------------------------------------------------------------------------------
struct DevConfig {
  int32_t variant, A, J, nI, V, S;
  int32_t order_random, shuffle_imp, auto_reset;
  uint32_t max_time_steps, tag_interval;
  uint32_t seed_lo, seed_hi, env_id_base;
  double r_kill, r_fix, r_sab, r_tsr, r_end, r_dead, r_vote;
  uint32_t valid_bits[8];  // bit c = 1 iff position code c = (x<<4|y) is inside the grid and not a wall
  uint8_t cell_code[84];   // valid-cell index (row-major argwhere(grid), base.py:199) -> position code
};

// shared-memory copy of the wall grid / spawn table ("the wall grid is staged in shared memory")
struct GridTables {
  uint32_t valid_bits[8];
  uint8_t cell_code[84];
  uint64_t tick;  // the launch's Philox tick (kernels that draw; see stage_tables_and_tick)
};

__device__ __forceinline__ void stage_tables(const DevConfig& c, GridTables& t) {
  for (int i = threadIdx.x; i < 8; i += blockDim.x) t.valid_bits[i] = c.valid_bits[i];
  for (int i = threadIdx.x; i < 84; i += blockDim.x) t.cell_code[i] = c.cell_code[i];
  __syncthreads();
}

// The tick that keys a launch's draws: a kernel parameter, or -- device-resident ticks (sus_env_device_ticks), which is
// what makes a launch replayable inside a CUDA graph -- a counter in device memory.  Thread 0 of every CTA reads it once
// and counts the CTA in; the CTA that completes the count (every CTA of the launch has read the tick by then) advances
// the counter by `n_ticks` for the next launch and re-arms the count.  No extra launch, nothing at kernel exit.
__device__ __forceinline__ uint64_t fetch_launch_tick(uint64_t host_tick, uint64_t* tick_dev, unsigned int* tick_ctr,
                                                      uint64_t n_ticks) {
  if (!tick_dev) return host_tick;
  const uint64_t tick = *reinterpret_cast<volatile uint64_t*>(tick_dev);
  __threadfence();
  if (atomicAdd(tick_ctr, 1u) == gridDim.x - 1) {
    *reinterpret_cast<volatile uint64_t*>(tick_dev) = tick + n_ticks;
    *reinterpret_cast<volatile unsigned int*>(tick_ctr) = 0u;
    __threadfence();
  }
  return tick;
}
__device__ __forceinline__ void stage_tables_and_tick(const DevConfig& c, GridTables& t, uint64_t host_tick,
                                                      uint64_t* tick_dev, unsigned int* tick_ctr, uint64_t n_ticks = 1) {
  for (int i = threadIdx.x; i < 8; i += blockDim.x) t.valid_bits[i] = c.valid_bits[i];
  for (int i = threadIdx.x; i < 84; i += blockDim.x) t.cell_code[i] = c.cell_code[i];
  if (threadIdx.x == 0) t.tick = fetch_launch_tick(host_tick, tick_dev, tick_ctr, n_ticks);
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// per-env state in registers and its HBM layout
// ---------------------------------------------------------------------------------------------
struct EnvState {
  uint64_t pos;     // byte i = (x << 4) | y of agent i            (base.py: agent_positions)
  uint64_t jobpos;  // byte j = (x << 4) | y of job j              (base.py: job_positions)
  uint32_t alive;   // bit i                                       (alive_agents)
  uint32_t imp;     // bit i: agent i is an imposter               (imposter_mask)
  uint32_t jobdone; // bit j                                       (completed_jobs)
  uint32_t used;    // bit i: tag used in this vote window         (tagging.py: used_tag_actions)
  uint32_t tagcnt;  // 4 bits per agent                            (tagging.py: tag_counts)
  uint32_t timer;   //                                             (tagging.py: tag_reset_timer)
  uint32_t nsteps;  // step() calls since reset; t = min(nsteps, max_time_steps-1)  (base.py:392-395)
  uint32_t completed, sabotaged;  // SusMetrics.COMPLETED_JOBS / SABOTAGED_JOBS of the episode
  uint32_t misc;    // kills | imp_voted << 8 | crew_voted << 16 | crew_won << 24 | imposter_won << 25
};

struct StateArrays {
  uint64_t* pos;
  uint64_t* jobpos;
  uint4* aux;  // {alive | imp<<8 | jobdone<<16 | used<<24, nsteps, timer, tagcnt}
  uint4* met;  // {completed, sabotaged, misc, 0}
};

__device__ __forceinline__ void load_state(const StateArrays& a, int64_t e, EnvState& s) {
  s.pos = a.pos[e];
  s.jobpos = a.jobpos[e];
  const uint4 x = a.aux[e];
  const uint4 m = a.met[e];
  s.alive = x.x & 0xff; s.imp = (x.x >> 8) & 0xff; s.jobdone = (x.x >> 16) & 0xff; s.used = x.x >> 24;
  s.nsteps = x.y; s.timer = x.z; s.tagcnt = x.w;
  s.completed = m.x; s.sabotaged = m.y; s.misc = m.z;
}

__device__ __forceinline__ void store_state(const StateArrays& a, int64_t e, const EnvState& s, bool jobs_too) {
  a.pos[e] = s.pos;
  if (jobs_too) a.jobpos[e] = s.jobpos;
  a.aux[e] = make_uint4(s.alive | (s.imp << 8) | (s.jobdone << 16) | (s.used << 24), s.nsteps, s.timer, s.tagcnt);
  a.met[e] = make_uint4(s.completed, s.sabotaged, s.misc, 0u);
}

__device__ __forceinline__ uint32_t get_byte(uint64_t w, int i) { return (uint32_t)(w >> (8 * i)) & 0xffu; }
__device__ __forceinline__ uint64_t set_byte(uint64_t w, int i, uint32_t b) {
  const int sh = 8 * i;
  return (w & ~(0xffull << sh)) | ((uint64_t)b << sh);
}
__device__ __forceinline__ uint32_t code_x(uint32_t c) { return c >> 4; }
__device__ __forceinline__ uint32_t code_y(uint32_t c) { return c & 15u; }
__device__ __forceinline__ uint32_t code_cell(uint32_t c) { return c - 7u * (c >> 4); }  // x*9 + y

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 and the draw spec (oracle/rng_spec.py is the written spec)
// ---------------------------------------------------------------------------------------------
enum : uint32_t { P_STEP = 0, P_AUTORESET = 1, P_ACT = 2, P_RESET = 3, P_ACT_FUSED = 4 };

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ k0, lo1, hi0 ^ ctr.w ^ k1, lo0);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return ctr;
}

// two independent blocks in one loop: the two 10-round dependency chains interleave (k_sample_actions with A > 4)
__device__ __forceinline__ void philox4x32_10_x2(uint4& a, uint4& b, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t ah0 = __umulhi(0xD2511F53u, a.x), al0 = 0xD2511F53u * a.x, ah1 = __umulhi(0xCD9E8D57u, a.z), al1 = 0xCD9E8D57u * a.z;
    const uint32_t bh0 = __umulhi(0xD2511F53u, b.x), bl0 = 0xD2511F53u * b.x, bh1 = __umulhi(0xCD9E8D57u, b.z), bl1 = 0xCD9E8D57u * b.z;
    a = make_uint4(ah1 ^ a.y ^ k0, al1, ah0 ^ a.w ^ k1, al0);
    b = make_uint4(bh1 ^ b.y ^ k0, bl1, bh0 ^ b.w ^ k1, bl0);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

__device__ __forceinline__ uint32_t bounded(uint32_t u, uint32_t k) { return __umulhi(u, k); }

// One logical word stream (env, tick, purpose): either injected raw words (parity mode) or Philox,
// generated four words at a time and cached.
struct WordStream {
  const uint32_t* inj;  // this env's injected words or nullptr
  uint32_t env_id, tick_lo, tick_hi, purpose, k0, k1;
  int cached_block;
  uint4 cache;
  __device__ __forceinline__ void init(const DevConfig& c, const uint32_t* inj_row, uint32_t env, uint64_t tick,
                                       uint32_t purp) {
    inj = inj_row; env_id = c.env_id_base + env; tick_lo = (uint32_t)tick; tick_hi = (uint32_t)(tick >> 32);
    purpose = purp; k0 = c.seed_lo; k1 = c.seed_hi; cached_block = -1; cache = make_uint4(0, 0, 0, 0);
  }
  __device__ __forceinline__ uint32_t word(int slot) {
    if (inj) return inj[slot];
    const int b = slot >> 2;
    if (b != cached_block) {
      cache = philox4x32_10(make_uint4(env_id, tick_lo, tick_hi, purpose | ((uint32_t)b << 8)), k0, k1);
      cached_block = b;
    }
    const int l = slot & 3;
    return l == 0 ? cache.x : l == 1 ? cache.y : l == 2 ? cache.z : cache.w;
  }
};

// r-th smallest id not yet chosen; chosen ids are a bitmask over < 128 ids (rng_spec.pick_distinct:
// every chosen id <= the running r, visited ascending, shifts r up by one)
__device__ __forceinline__ int pick_unchosen(int r, uint64_t lo, uint64_t hi) {
  while (lo) { const int c = __ffsll((long long)lo) - 1; if (r < c) return r; ++r; lo &= lo - 1; }
  while (hi) { const int c = 63 + __ffsll((long long)hi); if (r < c) return r; ++r; hi &= hi - 1; }
  return r;
}

// ---------------------------------------------------------------------------------------------
// reset: base.py:251-324, tagging.py:62-66
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void reset_env(const DevConfig& c, const GridTables& tb, EnvState& s, WordStream& ws) {
  const int A = c.A, J = c.J, nI = c.nI;
  s.imp = 0;
  if (c.shuffle_imp) {  // base.py:273-276 (R1), ascending-sorted distinct subset
    for (int m = 0; m < nI; ++m) {
      const int r = (int)bounded(ws.word(m), (uint32_t)(A - m));
      s.imp |= 1u << pick_unchosen(r, (uint64_t)s.imp, 0ull);
    }
  } else {
    s.imp = (1u << nI) - 1u;  // base.py:278
  }
  s.pos = 0;
  for (int i = 0; i < A; ++i) {  // base.py:288-291 (R2): iid uniform over valid cells, with replacement
    const uint32_t cell = bounded(ws.word(nI + i), (uint32_t)c.V);
    s.pos |= (uint64_t)tb.cell_code[cell] << (8 * i);
  }
  s.jobpos = 0;
  uint64_t lo = 0, hi = 0;
  for (int j = 0; j < J; ++j) {  // base.py:295-299 (R3): distinct cells
    const int r = (int)bounded(ws.word(nI + A + j), (uint32_t)(c.V - j));
    const int cell = pick_unchosen(r, lo, hi);
    if (cell < 64) lo |= 1ull << cell; else hi |= 1ull << (cell - 64);
    s.jobpos |= (uint64_t)tb.cell_code[cell] << (8 * j);
  }
  s.alive = (1u << A) - 1u;  // base.py:301
  s.jobdone = 0;             // base.py:302
  s.used = 0; s.tagcnt = 0; s.timer = 0;  // tagging.py:64-66
  s.nsteps = 0;                           // base.py:315
  s.completed = 0; s.sabotaged = 0; s.misc = 0;  // base.py:270
}

// ---------------------------------------------------------------------------------------------
// action decoding (role lists: base.py:82-99, pred_prey.py:4-19, tagging.py:69-75)
// ---------------------------------------------------------------------------------------------
template <int VARIANT>
__device__ __forceinline__ uint32_t n_role_actions(const DevConfig& c, uint32_t is_imp) {
  if (VARIANT == SUS_VARIANT_TRAINING_GROUND) return 5u + is_imp;
  const uint32_t base = 6u + is_imp;
  return VARIANT == SUS_VARIANT_TAGGING ? base + (uint32_t)c.A - 1u : base;
}

// ---------------------------------------------------------------------------------------------
// step: base.py:332-407 (+ tagging.py:120-235, pred_prey.py:78-99)
// ---------------------------------------------------------------------------------------------
struct StepResult {
  uint32_t kill_m, fix_m, sab_m;  // last event ASSIGNED to each agent in this step (base.py:514-515,523,532)
  double team_reward;
  uint32_t team_code;  // win (0 none, 1 crew, 2 imposters) + 3 * vote (0 none, 1 crew member ejected, 2 imposter ejected)
  double ret_imp, ret_crew;  // mean return of the imposters / the crew if the episode ended at this step (else 0)
  bool done, trunc;
};

__device__ __forceinline__ void assign_event(StepResult& r, int which, uint32_t bit) {
  r.kill_m &= ~bit; r.fix_m &= ~bit; r.sab_m &= ~bit;
  if (which == 0) r.kill_m |= bit; else if (which == 1) r.fix_m |= bit; else r.sab_m |= bit;
}

// TA / TJ: compile-time agent / job counts of a specialised instantiation (TA == 0: read them from the config).  Only
// the step-only and rollout kernels are specialised: their time is the step arithmetic, and constant trip counts let
// the compiler unroll the agent / job loops (measured +18 % / +26 %); the fused encode kernels are store-bound.
template <int VARIANT, int TA = 0, int TJ = 0>
__device__ __forceinline__ void step_env(const DevConfig& c, const GridTables& tb, EnvState& s, uint64_t acts,
                                         WordStream& ws, StepResult& out) {
  const int A = TA ? TA : c.A, J = TA ? TJ : c.J;
  out.kill_m = out.fix_m = out.sab_m = 0;
  out.team_reward = 0.0;
  out.team_code = 0;

  // action order: identity or Fisher-Yates (base.py:372-374, R4); one nibble per slot
  uint32_t order = 0x76543210u;
  if (c.order_random) {
    for (int k = A - 1; k > 0; --k) {
      const uint32_t j = bounded(ws.word(A - 1 - k), (uint32_t)(k + 1));
      const uint32_t vk = (order >> (4 * k)) & 15u, vj = (order >> (4 * j)) & 15u;
      order = (order & ~(15u << (4 * k))) | (vj << (4 * k));
      order = (order & ~(15u << (4 * j))) | (vk << (4 * j));
    }
  }

  int kill_events = 0;
  for (int k = 0; k < A; ++k) {  // base.py:377-382
    const int i = (int)((order >> (4 * k)) & 15u);
    const uint32_t bit = 1u << i;
    const uint32_t a = get_byte(acts, i);
    const uint32_t is_imp = (s.imp >> i) & 1u;
    if (VARIANT == SUS_VARIANT_TAGGING) {
      const uint32_t nbase = 6u + is_imp;
      if (a >= nbase) {  // tag action: tagging.py:103-110 -- the tagger's own liveness is NOT checked
        const uint32_t q = a - nbase;
        const uint32_t target = q < (uint32_t)i ? q : q + 1u;
        if (!(s.used & bit) && ((s.alive >> target) & 1u)) {
          s.tagcnt += 1u << (4 * target);
          s.used |= bit;
        }
        continue;
      }
    }
    if (!(s.alive & bit)) continue;  // base.py:477
    const uint32_t me = get_byte(s.pos, i);
    if (a <= 4u) {  // move: base.py:484-487, move(): base.py:69-79; validity: base.py:548-551
      const uint32_t nb = (me + (uint32_t)((0x10F0FF0100ull >> (8 * a)) & 0xffull)) & 0xffu;
      if ((tb.valid_bits[nb >> 5] >> (nb & 31u)) & 1u) s.pos = set_byte(s.pos, i, nb);
    } else if (VARIANT == SUS_VARIANT_TRAINING_GROUND || (is_imp && a == 6u)) {  // KILL: base.py:490-515
      uint32_t m = 0;
      for (int j = 0; j < A; ++j) m |= (get_byte(s.pos, j) == me ? 1u : 0u) << j;
      m &= s.alive & ~s.imp;  // alive crew on the killer's cell, ascending (base.py:535-542)
      const int n = __popc(m);
      if (n > 0) {
        const uint32_t pick = n == 1 ? 0u : bounded(ws.word(A - 1 + kill_events), (uint32_t)n);  // base.py:497 (R5)
        ++kill_events;
        const int victim = (int)__fns(m, 0, (int)pick + 1);
        s.alive &= ~(1u << victim);
        s.misc += 1u;  // IMP_KILLED_CREW
        assign_event(out, 0, 1u << victim);
        assign_event(out, 0, bit);
      }
    } else {  // FIX (crew, index 5) or SABOTAGE (imposter, index 5): base.py:518-533
      int job = -1;
      for (int j = J - 1; j >= 0; --j)
        if (get_byte(s.jobpos, j) == me) job = j;  // first job at the cell (base.py:544-546)
      if (job >= 0) {
        const uint32_t jb = 1u << job;
        if (!is_imp && !(s.jobdone & jb)) {
          s.jobdone |= jb; s.completed += 1u; assign_event(out, 1, bit);
        } else if (is_imp && (s.jobdone & jb)) {
          s.jobdone &= ~jb; s.sabotaged += 1u; assign_event(out, 2, bit);
        }
      }
    }
  }

  if (VARIANT == SUS_VARIANT_TAGGING) {  // tagging.py:180-207
    uint32_t keep = 0;
    for (int i = 0; i < A; ++i) keep |= ((s.alive >> i) & 1u) ? (15u << (4 * i)) : 0u;
    s.tagcnt &= keep;  // tag_counts *= alive_agents
    s.timer += 1u;
    if (s.timer >= c.tag_interval) {
      uint32_t best = 0, best_cnt = s.tagcnt & 15u;
      for (int i = 1; i < A; ++i) {  // np.argmax: first maximum
        const uint32_t cnt = (s.tagcnt >> (4 * i)) & 15u;
        if (cnt > best_cnt) { best_cnt = cnt; best = (uint32_t)i; }
      }
      const uint32_t quorum = ((uint32_t)__popc(s.alive) + 1u) >> 1;  // counted BEFORE the eject
      if (best_cnt >= quorum) {
        s.alive &= ~(1u << best);
        const bool was_imp = (s.imp >> best) & 1u;
        out.team_reward += c.r_vote * (was_imp ? -1.0 : 1.0);  // tagging.py:196
        out.team_code = was_imp ? 6u : 3u;
        s.misc += was_imp ? (1u << 8) : (1u << 16);            // IMP_VOTED_OUT / CREW_VOTED_OUT
      }
      s.tagcnt = 0; s.used = 0; s.timer = 0;  // tagging.py:237-240
    }
  }

  // win condition
  const int alive_imp = __popc(s.alive & s.imp), alive_crew = __popc(s.alive & ~s.imp);
  const int n_done = __popc(s.jobdone);
  bool done = false;
  double win = 0.0;
  if (VARIANT == SUS_VARIANT_TRAINING_GROUND) {  // pred_prey.py:78-99
    if (J != 0 && n_done == J) { done = true; win = c.r_end; s.misc |= 1u << 24; out.team_code += 1u; }
    else if (alive_crew == 0) { done = true; win = -1.0 * c.r_end; s.misc |= 1u << 25; out.team_code += 2u; }
  } else {  // base.py:409-460: crew win is tested first; true at once when J == 0
    if (alive_imp == 0 || n_done == J) { done = true; win = c.r_end; s.misc |= 1u << 24; out.team_code += 1u; }
    else if (alive_crew <= alive_imp) { done = true; win = -1.0 * c.r_end; s.misc |= 1u << 25; out.team_code += 2u; }
  }
  out.team_reward += win;
  out.done = done;
  out.trunc = s.nsteps >= c.max_time_steps - 1u;  // t == max_time_steps - 1 (base.py:392-395)
  if (s.nsteps != 0xffffffffu) s.nsteps += 1u;
}

// reward of agent i after _merge_rewards (base.py:553-563) and the zero replacement (base.py:389-390);
// computed in double like numpy so arbitrary float reward constants round identically.
template <int VARIANT>
__device__ __forceinline__ double agent_reward(const DevConfig& c, const EnvState& s, const StepResult& r, int i) {
  const uint32_t bit = 1u << i;
  double v = VARIANT == SUS_VARIANT_TAGGING ? 1.0 * c.r_tsr : 0.0;  // tagging.py:162 / base.py:369
  if (r.kill_m & bit) v = c.r_kill;
  else if (r.fix_m & bit) v = c.r_fix;
  else if (r.sab_m & bit) v = -1.0 * c.r_sab;
  v += r.team_reward;
  if (i < c.nI) v *= -1.0;             // by INDEX, not by role (base.py:559)
  if (!(s.alive & bit)) v = c.r_dead;  // base.py:562
  if (VARIANT != SUS_VARIANT_TAGGING && v == 0.0) v = c.r_tsr;
  return v;
}

// ---- compact host protocol (include/susnet_b200.h, SusCompactLayout): the same reward as a small code.
// code = dead ? n_live : event + 4 * team_code; n_live = 12 (no vote phase) or 36 (tagging).
__host__ __device__ __forceinline__ uint32_t n_live_codes(int variant) { return variant == SUS_VARIANT_TAGGING ? 36u : 12u; }

__device__ __forceinline__ uint32_t agent_reward_code(const DevConfig& c, const EnvState& s, const StepResult& r, int i) {
  const uint32_t bit = 1u << i;
  if (!(s.alive & bit)) return n_live_codes(c.variant);
  const uint32_t ev = (r.kill_m & bit) ? 1u : (r.fix_m & bit) ? 2u : (r.sab_m & bit) ? 3u : 0u;
  return ev + 4u * r.team_code;
}

// The float64 reward a code stands for, by the SAME operation sequence as step_env() + agent_reward() (so the host's
// decode table is bit-exact for arbitrary reward constants; multiplications by +-1 are exact, so FMA contraction of
// `team += r_vote * s` cannot change a bit).
inline double reward_of_code(const DevConfig& c, int i, uint32_t code) {
  const bool tagging = c.variant == SUS_VARIANT_TAGGING;
  const uint32_t n_live = n_live_codes(c.variant);
  double v;
  if (code == n_live) {
    v = c.r_dead;
  } else {
    const uint32_t ev = code & 3u, team = code >> 2, win = team % 3u, vote = team / 3u;
    v = tagging ? 1.0 * c.r_tsr : 0.0;
    if (ev == 1u) v = c.r_kill; else if (ev == 2u) v = c.r_fix; else if (ev == 3u) v = -1.0 * c.r_sab;
    double t = 0.0;
    if (vote) t += c.r_vote * (vote == 2u ? -1.0 : 1.0);
    t += win == 1u ? c.r_end : (win == 2u ? -1.0 * c.r_end : 0.0);
    v += t;
    if (i < c.nI) v *= -1.0;
  }
  if (!tagging && v == 0.0) v = c.r_tsr;
  return v;
}

// ---------------------------------------------------------------------------------------------
// flatten order (base.py:211-241, tagging.py:42-60,221-230)
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void write_flat(const DevConfig& c, const EnvState& s, T* __restrict__ o) {
  const int A = c.A, J = c.J;
  int k = 0;
  for (int i = 0; i < A; ++i) { const uint32_t b = get_byte(s.pos, i); o[k++] = (T)code_x(b); o[k++] = (T)code_y(b); }
  for (int i = 0; i < A; ++i) o[k++] = (T)((s.alive >> i) & 1u);
  if (J > 0 || c.variant == SUS_VARIANT_TAGGING) {
    for (int j = 0; j < J; ++j) { const uint32_t b = get_byte(s.jobpos, j); o[k++] = (T)code_x(b); o[k++] = (T)code_y(b); }
    for (int j = 0; j < J; ++j) o[k++] = (T)((s.jobdone >> j) & 1u);
  }
  if (c.variant == SUS_VARIANT_TAGGING) {
    for (int i = 0; i < A; ++i) o[k++] = (T)((s.used >> i) & 1u);
    for (int i = 0; i < A; ++i) o[k++] = (T)((s.tagcnt >> (4 * i)) & 15u);
    o[k++] = (T)((int64_t)c.tag_interval - (int64_t)s.timer);
  }
}

// The observable part of a state, as the featurizers see it after unflatten.  Rows handed to
// sus_encode_from_flat must be valid flattened states (coordinates 0..8, flags 0/1, tag counts 0..15);
// a coordinate outside the grid becomes position code 0xff, which every featurizer skips (memory safe).
struct ObsState {
  uint64_t pos, jobpos;
  uint32_t alive, jobdone, used, tagcnt;
};

__device__ __forceinline__ ObsState obs_of(const EnvState& s) {
  ObsState o;
  o.pos = s.pos; o.jobpos = s.jobpos; o.alive = s.alive; o.jobdone = s.jobdone; o.used = s.used; o.tagcnt = s.tagcnt;
  return o;
}

}  // namespace susnet
