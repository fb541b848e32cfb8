// The library's DEVICE code for reset / step / flatten / the per-env feature rows (susnet_device.cuh, susnet_encode.cuh and the
// step glue of susnet_api.cu: load_input, step_one, finish_one) compiled for the host and run one env at a time, so that the CPU
// suite can hold the very functions the kernels run against the oracle.  KERNEL_SOURCE is the text cut out of susnet_api.cu by
// the test: StepParams, the step glue, and the host functions that build DevConfig / DevEncode from a SusConfig.
#include <cuda_runtime.h>

#include <string>

#include "susnet_device.cuh"
#include "susnet_encode.cuh"

using namespace susnet;

namespace {
constexpr int kThreads = 256;
constexpr unsigned kFull = 0xffffffffu;
int fail(int code, const std::string&) { return code; }
#include KERNEL_SOURCE

void tables_of(const DevConfig& c, uint64_t tick, GridTables& tb) {
  for (int i = 0; i < 8; ++i) tb.valid_bits[i] = c.valid_bits[i];
  for (int i = 0; i < 84; ++i) tb.cell_code[i] = c.cell_code[i];
  tb.tick = tick;
}
}  // namespace

extern "C" int emu_reset(const SusConfig* cfg, uint64_t epoch, uint64_t* pos, uint64_t* jobpos, uint4* aux, uint4* met) {
  DevConfig c;
  make_dev_config(*cfg, c);
  GridTables tb;
  tables_of(c, epoch, tb);
  StateArrays st{pos, jobpos, aux, met};
  for (int64_t e = 0; e < cfg->num_envs; ++e) {  // the body of k_reset
    EnvState s;
    WordStream ws;
    ws.init(c, nullptr, (uint32_t)e, tb.tick, P_RESET);
    reset_env(c, tb, s, ws);
    store_state(st, e, s, true);
  }
  return 0;
}

// one step launch (step_tick) of k_step<VARIANT, false>: actions [N][A] int32 or NULL (fused random policy)
template <int V>
static void step_all(const StepParams& p, const GridTables& tb) {
  const int rew_row = reward_row_bytes(p);
  for (int64_t e = 0; e < p.N; ++e) {
    StepInput in;
    load_input(p, e, true, in);
    bool stepped, finished;
    EnvState s = {};
    StepResult r = {};
    step_one<V>(p, tb, e, true, in, reward_rows(p) + e * rew_row, p.next_flat ? p.next_flat + e * p.c.S : nullptr, s, r, stepped, finished);
    finish_one(p, tb, e, 0, s, r, stepped, finished);
  }
}

extern "C" int emu_step(const SusConfig* cfg, uint64_t tick, uint64_t* pos, uint64_t* jobpos, uint4* aux, uint4* met,
                        const int32_t* actions, double* rewards, uint8_t* done, uint8_t* trunc, float* next_flat,
                        int32_t* actions_out, unsigned long long* stats, uint32_t* err) {
  StepParams p;
  std::memset(&p, 0, sizeof(p));
  make_dev_config(*cfg, p.c);
  make_dev_encode(*cfg, nullptr, p.enc, nullptr);
  p.st = StateArrays{pos, jobpos, aux, met};
  p.actions = actions; p.actions_dtype = SUS_I32;
  p.rewards = rewards; p.rewards_dtype = SUS_F64;
  p.done = done; p.trunc = trunc; p.next_flat = next_flat; p.actions_out = actions_out;
  p.stats = stats; p.err = err; p.tick = tick; p.N = cfg->num_envs;
  GridTables tb;
  tables_of(p.c, tick, tb);
  switch (cfg->variant) {
    case SUS_VARIANT_BASE: step_all<SUS_VARIANT_BASE>(p, tb); break;
    case SUS_VARIANT_TAGGING: step_all<SUS_VARIANT_TAGGING>(p, tb); break;
    default: step_all<SUS_VARIANT_TRAINING_GROUND>(p, tb); break;
  }
  return 0;
}

// the same launch through the compact host protocol: bit-packed action records in, reward codes + done / truncated bits out
extern "C" int emu_step_compact(const SusConfig* cfg, uint64_t tick, uint64_t* pos, uint64_t* jobpos, uint4* aux, uint4* met,
                                const uint8_t* packed_actions, uint8_t* packed_out, unsigned long long* stats, uint32_t* err) {
  StepParams p;
  std::memset(&p, 0, sizeof(p));
  make_dev_config(*cfg, p.c);
  make_dev_encode(*cfg, nullptr, p.enc, nullptr);
  SusCompactLayout cl;
  compact_layout(*cfg, cl);
  p.st = StateArrays{pos, jobpos, aux, met};
  p.actions = packed_actions; p.actions_dtype = SUS_PACKED;
  p.packed_out = packed_out;
  p.action_bits = cl.action_bits; p.action_bytes = cl.action_bytes; p.reward_bits = cl.reward_bits; p.result_bytes = cl.result_bytes;
  p.stats = stats; p.err = err; p.tick = tick; p.N = cfg->num_envs;
  GridTables tb;
  tables_of(p.c, tick, tb);
  switch (cfg->variant) {
    case SUS_VARIANT_BASE: step_all<SUS_VARIANT_BASE>(p, tb); break;
    case SUS_VARIANT_TAGGING: step_all<SUS_VARIANT_TAGGING>(p, tb); break;
    default: step_all<SUS_VARIANT_TRAINING_GROUND>(p, tb); break;
  }
  return 0;
}

extern "C" int emu_export_flat(const SusConfig* cfg, uint64_t* pos, uint64_t* jobpos, uint4* aux, uint4* met, long long* out) {
  DevConfig c;
  make_dev_config(*cfg, c);
  StateArrays st{pos, jobpos, aux, met};
  for (int64_t e = 0; e < cfg->num_envs; ++e) {
    EnvState s;
    load_state(st, e, s);
    write_flat<long long>(c, s, out + e * c.S);
  }
  return 0;
}

// the per-env feature rows: Flat rows (float rows, byte-staged rows expanded like the kernel does, with and without the
// compile-time agent count), Global / Perspective non-spatial rows, and the plane offsets as dense planes
static void encode_item(const DevConfig& c, const DevEncode& enc, const GridTables& tb, const ObsState& o, int mode, int64_t e,
                        int64_t N, float* spatial, float* non_spatial) {
  const int A = c.A, F = enc.ns_floats, R = enc.sp_floats;
  {
    if (enc.kind == SUS_ENCODE_FLAT) {
      float* row = non_spatial + e * F;
      if (mode == 0) {
        flat_row<FloatRow>(c, enc, tb, o, row);
      } else {  // byte-staged row (k_step_flat): prefill with the byte of 0, set values, expand
        uint8_t bytes[1024];
        std::memset(bytes, (int)kByteRowBias, sizeof(bytes));
        if (mode == 1) flat_row<ByteRow>(c, enc, tb, o, bytes);
        else if (A == 5) flat_row<ByteRow, 5>(c, enc, tb, o, bytes);
        else if (A == 2) flat_row<ByteRow, 2>(c, enc, tb, o, bytes);
        else flat_row<ByteRow>(c, enc, tb, o, bytes);
        for (int i = 0; i < F; ++i) {
          uint32_t w;
          std::memcpy(&w, bytes + (i & ~3), 4);
          row[i] = byte_row_value(w, i & 3);
        }
      }
    } else {
      const int views = enc.kind == SUS_ENCODE_GLOBAL ? 1 : A;
      for (int k = 0; k < A; ++k) {
        float* row = non_spatial + ((int64_t)k * N + e) * F;
        if (enc.kind == SUS_ENCODE_GLOBAL) global_ns_row(c, o, k, row);
        else persp_ns_row(c, o, k, row);
      }
      for (int k = 0; k < views; ++k) {
        float* plane = spatial + ((int64_t)k * N + e) * R;
        for (int i = 0; i < R; ++i) plane[i] = 0.0f;
        const PlaneOffsets po = enc.kind == SUS_ENCODE_GLOBAL ? plane_offsets(c, o, [](int i) { return i; })
                                                              : plane_offsets(c, o, [k](int i) { return persp_channel_of_agent(k, i); });
        put_planes(po, plane, 1.0f);
      }
    }
  }
}

extern "C" int emu_encode(const SusConfig* cfg, const SusEncodeSpec* spec, int mode, uint64_t* pos, uint64_t* jobpos, uint4* aux,
                          uint4* met, float* spatial, float* non_spatial) {
  DevConfig c;
  make_dev_config(*cfg, c);
  DevEncode enc;
  if (int rc = make_dev_encode(*cfg, spec, enc, nullptr)) return rc;
  GridTables tb;
  tables_of(c, 0, tb);
  StateArrays st{pos, jobpos, aux, met};
  for (int64_t e = 0; e < cfg->num_envs; ++e) {
    EnvState s;
    load_state(st, e, s);
    encode_item(c, enc, tb, obs_of(s), mode, e, cfg->num_envs, spatial, non_spatial);
  }
  return 0;
}

// the same from flattened state rows [n][S] (int64), the way SequenceStateFeaturizer.fit feeds replay batches (parse_row)
extern "C" int emu_encode_rows(const SusConfig* cfg, const SusEncodeSpec* spec, int mode, const long long* rows, int64_t n,
                               float* spatial, float* non_spatial) {
  DevConfig c;
  make_dev_config(*cfg, c);
  DevEncode enc;
  if (int rc = make_dev_encode(*cfg, spec, enc, nullptr)) return rc;
  GridTables tb;
  tables_of(c, 0, tb);
  for (int64_t e = 0; e < n; ++e) encode_item(c, enc, tb, parse_row<long long>(c, rows + e * c.S), mode, e, n, spatial, non_spatial);
  return 0;
}
