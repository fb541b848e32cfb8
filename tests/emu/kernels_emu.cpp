// Whole kernels of susnet_api.cu on the host with real barriers and warp collectives (simt.h): k_reset, k_sample_actions, the
// direct-store fused kernel k_step<V, ENCODE> (warp_encode: planes zero-filled by all lanes, ones scattered per lane) and the
// byte-staged Flat kernel k_step_flat<V[, A, J]> (rows prefilled and built in shared memory, expanded by all lanes).
// KERNEL_SOURCE is the kernel text of susnet_api.cu (everything between the parameter structs and the host side) plus the host
// functions that build DevConfig / DevEncode / FlatStage, cut out by the test; the TMA headers are copies whose PTX is removed
// (those kernels compile but are not run here: bulk copies and mbarriers have no host counterpart).
#define EMU_SIMT
#include <cuda_runtime.h>

#include <string>

static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }

#include "susnet_device.cuh"
#include "susnet_encode.cuh"
#include "susnet_tile.cuh"
#include "susnet_ws.cuh"

using namespace susnet;

namespace {
alignas(128) uint8_t dyn_smem[256 * 1024];
int fail(int code, const std::string&) { return code; }
#include KERNEL_SOURCE

unsigned blocks_for(int64_t n) { return (unsigned)((n + kThreads - 1) / kThreads); }
}  // namespace

extern "C" int emu_k_reset(const SusConfig* cfg, uint64_t epoch, uint64_t* pos, uint64_t* jobpos, uint4* aux, uint4* met) {
  ResetParams p;
  std::memset(&p, 0, sizeof(p));
  make_dev_config(*cfg, p.c);
  p.st = StateArrays{pos, jobpos, aux, met};
  p.tick = epoch; p.N = cfg->num_envs;
  simt::launch(blocks_for(p.N), kThreads, [&] { k_reset(p); });
  return 0;
}

extern "C" int emu_k_sample_actions(const SusConfig* cfg, uint64_t act_epoch, uint4* aux, int32_t* out) {
  ActParams p;
  std::memset(&p, 0, sizeof(p));
  make_dev_config(*cfg, p.c);
  p.st.aux = aux; p.out = out; p.tick = act_epoch; p.N = cfg->num_envs;
  simt::launch(blocks_for(p.N), kThreads, [&] { k_sample_actions(p); });
  return 0;
}

// path 0: k_step<V, ENCODE> (direct stores); path 1: k_step_flat<V> (byte-staged rows); path 2: k_step_flat<V, A, J> where the
// library instantiates the compile-time shape (ImposterTrainingGround, 5 agents, no jobs).  Returns -1 if the path does not apply.
extern "C" int emu_k_step(const SusConfig* cfg, int path, uint64_t tick, uint64_t* pos, uint64_t* jobpos, uint4* aux, uint4* met,
                          const int32_t* actions, void* rewards, int rewards_dtype, uint8_t* done, uint8_t* trunc, float* next_flat,
                          const SusEncodeSpec* spec, float* spatial, float* non_spatial, unsigned long long* stats, uint32_t* err) {
  StepParams p;
  std::memset(&p, 0, sizeof(p));
  make_dev_config(*cfg, p.c);
  if (int rc = make_dev_encode(*cfg, spec, p.enc, nullptr)) return rc;
  p.st = StateArrays{pos, jobpos, aux, met};
  p.actions = actions; p.actions_dtype = SUS_I32;
  p.rewards = rewards; p.rewards_dtype = rewards_dtype;
  p.done = done; p.trunc = trunc; p.next_flat = next_flat; p.spatial = spatial; p.non_spatial = non_spatial;
  p.stats = stats; p.err = err; p.tick = tick; p.N = cfg->num_envs;
  const unsigned grid = blocks_for(p.N);
  const bool enc = p.enc.kind != SUS_ENCODE_NONE;
  if (path == 0) {
    switch (cfg->variant) {
      case SUS_VARIANT_BASE:
        if (enc) simt::launch(grid, kThreads, [&] { k_step<SUS_VARIANT_BASE, true>(p); });
        else simt::launch(grid, kThreads, [&] { k_step<SUS_VARIANT_BASE, false>(p); });
        break;
      case SUS_VARIANT_TAGGING:
        if (enc) simt::launch(grid, kThreads, [&] { k_step<SUS_VARIANT_TAGGING, true>(p); });
        else simt::launch(grid, kThreads, [&] { k_step<SUS_VARIANT_TAGGING, false>(p); });
        break;
      default:
        if (enc) simt::launch(grid, kThreads, [&] { k_step<SUS_VARIANT_TRAINING_GROUND, true>(p); });
        else simt::launch(grid, kThreads, [&] { k_step<SUS_VARIANT_TRAINING_GROUND, false>(p); });
        break;
    }
    return 0;
  }
  FlatStage FS;
  if (!make_flat_stage(p.c, p.enc, rewards_dtype == SUS_F64 ? 8 : 4, next_flat != nullptr, 227 * 1024, FS)) return -1;
  if ((size_t)FS.per_warp * (kThreads / 32) > sizeof(dyn_smem)) return -1;
  if (path == 2) {
    if (!(cfg->variant == SUS_VARIANT_TRAINING_GROUND && p.c.A == 5 && p.c.J == 0)) return -1;
    simt::launch(grid, kThreads, [&] { k_step_flat<SUS_VARIANT_TRAINING_GROUND, 5, 0>(p, FS); });
    return 0;
  }
  switch (cfg->variant) {
    case SUS_VARIANT_BASE: simt::launch(grid, kThreads, [&] { k_step_flat<SUS_VARIANT_BASE>(p, FS); }); break;
    case SUS_VARIANT_TAGGING: simt::launch(grid, kThreads, [&] { k_step_flat<SUS_VARIANT_TAGGING>(p, FS); }); break;
    default: simt::launch(grid, kThreads, [&] { k_step_flat<SUS_VARIANT_TRAINING_GROUND>(p, FS); }); break;
  }
  return 0;
}

// n random-policy steps in one launch (env.rollout): generic instantiation, or (shape != 0) the compile-time agent / job counts
// the library picks for the common shapes.  Returns -1 if the shape does not apply.
extern "C" int emu_k_rollout(const SusConfig* cfg, int shape, uint64_t tick0, int32_t n_steps, uint64_t* pos, uint64_t* jobpos,
                             uint4* aux, uint4* met, unsigned long long* stats, double* reward_sums) {
  RolloutParams p;
  std::memset(&p, 0, sizeof(p));
  make_dev_config(*cfg, p.c);
  p.st = StateArrays{pos, jobpos, aux, met};
  p.stats = stats; p.reward_sums = reward_sums; p.tick0 = tick0; p.N = cfg->num_envs; p.n_steps = n_steps;
  const unsigned grid = blocks_for(p.N);
  const int A = p.c.A, J = p.c.J;
  if (shape) {
    if (cfg->variant == SUS_VARIANT_TRAINING_GROUND && A == 5 && J == 0) simt::launch(grid, kThreads, [&] { k_rollout<SUS_VARIANT_TRAINING_GROUND, 5, 0>(p); });
    else if (cfg->variant == SUS_VARIANT_TRAINING_GROUND && A == 2 && J == 0) simt::launch(grid, kThreads, [&] { k_rollout<SUS_VARIANT_TRAINING_GROUND, 2, 0>(p); });
    else if (cfg->variant == SUS_VARIANT_BASE && A == 5 && J == 5) simt::launch(grid, kThreads, [&] { k_rollout<SUS_VARIANT_BASE, 5, 5>(p); });
    else if (cfg->variant == SUS_VARIANT_TAGGING && A == 3 && J == 5) simt::launch(grid, kThreads, [&] { k_rollout<SUS_VARIANT_TAGGING, 3, 5>(p); });
    else return -1;
    return 0;
  }
  switch (cfg->variant) {
    case SUS_VARIANT_BASE: simt::launch(grid, kThreads, [&] { k_rollout<SUS_VARIANT_BASE>(p); }); break;
    case SUS_VARIANT_TAGGING: simt::launch(grid, kThreads, [&] { k_rollout<SUS_VARIANT_TAGGING>(p); }); break;
    default: simt::launch(grid, kThreads, [&] { k_rollout<SUS_VARIANT_TRAINING_GROUND>(p); }); break;
  }
  return 0;
}

// SequenceStateFeaturizer.fit on (n, S) replay rows of dtype f32 / f64 / i64: k_encode_rows<T> (direct stores, every encode kind)
// or, for integer-valued Flat rows, the byte-staged k_encode_flat<T, true>.  Returns -1 if the staged kernel does not apply.
template <typename T>
static int encode_rows_t(const SusConfig* cfg, const SusEncodeSpec* spec, int staged, const void* rows, int64_t n, float* spatial,
                         float* non_spatial) {
  EncodeParams p;
  std::memset(&p, 0, sizeof(p));
  make_dev_config(*cfg, p.c);
  if (int rc = make_dev_encode(*cfg, spec, p.enc, nullptr)) return rc;
  p.rows = rows; p.spatial = spatial; p.non_spatial = non_spatial; p.n_items = n;
  if (!staged) {
    simt::launch(blocks_for(n), kThreads, [&] { k_encode_rows<T>(p); });
    return 0;
  }
  FlatStage FS;
  if (!make_flat_stage(p.c, p.enc, 4, false, 227 * 1024, FS)) return -1;
  simt::launch(blocks_for(n), kThreads, [&] { k_encode_flat<T, true>(p, FS); });
  return 0;
}

extern "C" int emu_k_encode_rows(const SusConfig* cfg, const SusEncodeSpec* spec, int staged, int dtype, const void* rows, int64_t n,
                                 float* spatial, float* non_spatial) {
  if (dtype == SUS_F32) return encode_rows_t<float>(cfg, spec, staged, rows, n, spatial, non_spatial);
  if (dtype == SUS_F64) return encode_rows_t<double>(cfg, spec, staged, rows, n, spatial, non_spatial);
  return encode_rows_t<long long>(cfg, spec, staged, rows, n, spatial, non_spatial);
}
