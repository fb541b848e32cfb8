/*
 * susnet_b200.h -- C ABI of the B200-native batched Sus-Net simulator + observation encoder.
 *
 * The reference (jhrudden/Sus-Net) is pure Python and has no FFI of its own; the drop-in boundary is the
 * duck-typed surface its callers use.  Each entry point below names the reference
 * interface it replaces (file:line under the reference's src/).  The Python host side
 * (sus_net_b200/env.py, featurizers.py) binds these with ctypes and mirrors the reference's class and
 * method names; INTEGRATION.md shows the stub a Sus-Net maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer argument is a DEVICE pointer on the env's device
 *     unless the comment says "host";
 *   - every call returns 0 on success or a negative SUS_ERR_* code; sus_last_error() returns the message
 *     of the calling thread's last failure;
 *   - all work is enqueued on the caller's cudaStream_t (passed as void*); no call synchronises the
 *     device except sus_env_check_actions() and sus_env_create()/destroy();
 *   - calls on one handle are not thread-safe (neither is the reference: it uses numpy's global RNG).
 *
 * Batch layout: env e of a handle has the global id env_id_base + e, which keys its Philox stream
 * together with `seed` and the launch tick, so results do not depend on how envs are sharded over GPUs.
 */
#ifndef SUSNET_B200_H_
#define SUSNET_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SUS_ABI_VERSION 2

#define SUS_MAX_AGENTS 8
#define SUS_MAX_JOBS 8
#define SUS_N_METRICS 8  /* per-episode counters, order of SusMetricIndex */
#define SUS_N_STATS 10   /* finished-episode accumulators, order of SusStatIndex */
#define SUS_MAX_FLAT_COMPONENTS 16

enum SusError {
  SUS_OK = 0,
  SUS_ERR_INVALID_ARGUMENT = -1, /* reference: AssertionError from _validate_init_args / step asserts */
  SUS_ERR_UNSUPPORTED = -2,      /* valid for the reference, outside this build's limits (A, J > 8 ...) */
  SUS_ERR_CUDA = -3,
  SUS_ERR_INVALID_ACTION = -4    /* reference: IndexError from agent_action_map[i][a] (base.py:381) */
};

/* src/environment/{base.py:102, tagging.py:9, pred_prey.py:20} */
enum SusVariant { SUS_VARIANT_BASE = 0, SUS_VARIANT_TAGGING = 1, SUS_VARIANT_TRAINING_GROUND = 2 };

enum SusDtype { SUS_U8 = 0, SUS_I32 = 1, SUS_I64 = 2, SUS_F32 = 3, SUS_F64 = 4,
                SUS_PACKED = 5 /* actions only: bit-packed records of the compact host protocol, see SusCompactLayout */ };

/* src/metrics.py:7-32 -- the counters the env touches (base.py:366,508,522,531; tagging.py:199,201;
 * base.py:433,444; pred_prey.py:90,96). */
enum SusMetricIndex {
  SUS_M_TOTAL_TIME_STEPS = 0, SUS_M_IMP_KILLED_CREW, SUS_M_COMPLETED_JOBS, SUS_M_SABOTAGED_JOBS,
  SUS_M_IMP_VOTED_OUT, SUS_M_CREW_VOTED_OUT, SUS_M_CREW_WON, SUS_M_IMPOSTER_WON
};

/* Sums over FINISHED episodes (what EpisodicMetricHandler.step(info) receives at train.py:427). */
enum SusStatIndex {
  SUS_S_EPISODES = 0, SUS_S_CREW_WON, SUS_S_IMPOSTER_WON, SUS_S_IMP_KILLED_CREW, SUS_S_COMPLETED_JOBS,
  SUS_S_SABOTAGED_JOBS, SUS_S_IMP_VOTED_OUT, SUS_S_CREW_VOTED_OUT, SUS_S_TOTAL_TIME_STEPS, SUS_S_TRUNCATED
};

/* Constructor arguments of FourRoomEnv (base.py:103-120), FourRoomEnvWithTagging (tagging.py:10-12) and
 * ImposterTrainingGround (pred_prey.py:26-66; the host normalises its fixed arguments: n_imposters = 1,
 * dead_penalty = 0, is_action_order_random = 0), plus the batch/sharding fields the reference has no
 * analogue for. */
typedef struct SusConfig {
  int32_t variant;                /* SusVariant */
  int32_t n_imposters;
  int32_t n_crew;
  int32_t n_jobs;
  int32_t include_walls;          /* base.py:171-193 */
  int32_t is_action_order_random; /* base.py:372-374 */
  int32_t shuffle_imposter_index; /* base.py:273-278 */
  int32_t max_time_steps;         /* base.py:118,392-395 */
  int32_t tag_reset_interval;     /* tagging.py:11,184 */
  int32_t auto_reset;             /* 1: a step that ends an episode resets the env in the same launch
                                     (train.py:419-445 does this on the host) */
  double kill_reward;             /* base.py:110 */
  double complete_job_reward;
  double sabotage_reward;
  double time_step_reward;
  double game_end_reward;
  double dead_penalty;
  double vote_reward;             /* tagging.py:11 */
  int64_t num_envs;
  uint64_t seed;                  /* Philox key */
  uint32_t env_id_base;           /* global id of env 0 (sharding across GPUs) */
  uint32_t reserved;
} SusConfig;

/* Observation encodings of src/features/model_ready.py. */
enum SusEncodeKind {
  SUS_ENCODE_NONE = 0,
  SUS_ENCODE_GLOBAL = 1,      /* GlobalFeaturizer      model_ready.py:219-306 */
  SUS_ENCODE_PERSPECTIVE = 2, /* PerspectiveFeaturizer model_ready.py:82-216  */
  SUS_ENCODE_FLAT = 3         /* FlatFeaturizer over a CompositeFeaturizer, model_ready.py:309-367 */
};

/* Per-state featurizers of src/features/component.py usable inside SUS_ENCODE_FLAT. */
enum SusFlatComponent {
  SUS_FC_ONEHOT_POS = 0,       /* OneHotAgentPositionFeaturizer        component.py:221-247 */
  SUS_FC_COORDS = 1,           /* CoordinateAgentPositionsFeaturizer   component.py:384-403 */
  SUS_FC_ALIVE_CREW = 2,       /* AliveCrewFeaturizer                  component.py:406-425 */
  SUS_FC_CLOSEST_CREW = 3,     /* ClosestAliveCrewFeaturizer           component.py:455-482 */
  SUS_FC_L1_CREW = 4,          /* L1CrewFeaturizer                     component.py:428-452 */
  SUS_FC_DIST_TO_IMPOSTER = 5, /* DistanceToImposterFeaturizer         component.py:250-278 */
  SUS_FC_WALLS = 6,            /* WallsFeaturizer                      component.py:281-300 */
  SUS_FC_ROOMS = 7,            /* ImposterVSCrewRoomLocaionFeaturizer  component.py:303-334 */
  SUS_FC_SCENT = 8,            /* ImposterScentFeaturizer              component.py:339-380 */
  SUS_FC_STATE_ALIVE = 9,      /* StateFieldFeaturizer(ALIVE_AGENTS)   component.py:200-218 */
  SUS_FC_STATE_JOB_STATUS = 10,
  SUS_FC_STATE_USED_TAGS = 11,
  SUS_FC_STATE_TAG_COUNTS = 12,
  SUS_FC_COUNT = 13
};

/* SusEncodeSpec.flags.  SUS_ENCODE_PLANES_U8 (opt-in, no reference analogue): the spatial tensor of GLOBAL / PERSPECTIVE
 * holds ONE BYTE per cell (0 / 1) instead of one float32 -- same shape, same [channel][x][y] order, 4x fewer bytes -- for a
 * consumer that casts in its first layer.  The non-spatial tensor stays float32. */
#define SUS_ENCODE_PLANES_U8 1

typedef struct SusEncodeSpec {
  int32_t kind;         /* SusEncodeKind */
  int32_t n_components; /* SUS_ENCODE_FLAT only */
  int32_t components[SUS_MAX_FLAT_COMPONENTS];
  int32_t flags;        /* 0, or SUS_ENCODE_PLANES_U8 */
} SusEncodeSpec;

/* Output shapes of an encode spec for a config (host call, no GPU work):
 *   spatial_floats     per item per view: (A+2)*81 for GLOBAL/PERSPECTIVE, 0 for FLAT
 *   non_spatial_floats per item per view: F
 *   spatial_views      1 (GLOBAL: all views share the planes), A (PERSPECTIVE), 0 (FLAT)
 *   non_spatial_views  A (GLOBAL, PERSPECTIVE), 1 (FLAT: every view is identical)
 * Replaces SequenceStateFeaturizer.featurized_shape (model_ready.py:115-123,249-253,318-323). */
typedef struct SusEncodeShape {
  int32_t spatial_floats, non_spatial_floats, spatial_views, non_spatial_views;
} SusEncodeShape;

/* Inputs and outputs of one step launch.  NULL output pointers are skipped. */
typedef struct SusStepIO {
  const void *actions;     /* [N][A] role-list indices (base.py:381); NULL = fused random policy, i.e.
                              env.step(env.sample_actions()) with the SUS_P_ACT_FUSED draws */
  int32_t actions_dtype;   /* SUS_U8 / SUS_I32 / SUS_I64 */
  int32_t rewards_dtype;   /* SUS_F32 (replay layout, replay_memory.py:38) or SUS_F64 (numpy step result) */
  void *rewards;           /* [N][A] */
  uint8_t *done;           /* [N] base.py:404 */
  uint8_t *truncated;      /* [N] base.py:405 */
  int32_t *actions_out;    /* [N][A] the actions applied (useful with the random policy) */
  float *next_flat;        /* [N][S] post-step, PRE-reset state in flatten order = the row the reference
                              stores in next_states[:, -1] (train.py:388-399, replay_memory.py:120-126) */
  int64_t *metrics;        /* [N][SUS_N_METRICS] pre-reset episode counters = the step's `info` dict */
  int16_t *imposters;      /* [N][n_imposters] ascending agent ids of the imposters of the episode the step belonged
                              to (pre-reset env.imposter_idxs: the `imposters` replay column, replay_memory.py:42-44) */
  const SusEncodeSpec *encode; /* host pointer or NULL: fused encode of the state the NEXT action is
                              taken from (post auto-reset), written to spatial / non_spatial below */
  void *spatial;           /* [views_s][N][spatial_floats] float32 (uint8 with SUS_ENCODE_PLANES_U8) */
  float *non_spatial;      /* [views_n][N][non_spatial_floats] */
  uint8_t *packed_out;     /* [N][result_bytes] compact result records (SusCompactLayout) INSTEAD of rewards / done /
                              truncated, which must then be NULL: what a host consumer pulls over PCIe per step */
} SusStepIO;

/* Compact host protocol (no reference analogue: the reference's step() hands numpy arrays to a caller in the same
 * process, train.py:383-399; here every byte of a step's actions and results crosses PCIe).  A step's reward of an
 * agent is one of a handful of values -- the event it was last assigned (none / kill / fix / sabotage, base.py:511-532),
 * plus the team reward (win +-game_end_reward, base.py:428-446; vote +-vote_reward, tagging.py:196), negated for agent
 * indices < n_imposters (base.py:559), or dead_penalty (base.py:562), with zeros replaced by time_step_reward
 * (base.py:389-390) -- so it travels as a small CODE and the host decodes it through the float64 table of
 * sus_reward_lut(), which is computed with the same float64 operation sequence as the kernel: decoding is bit-exact
 * for arbitrary reward constants.
 *   packed action record, action_bytes per env, little-endian bit order: agent i's role-list index in bits
 *     [i*action_bits, (i+1)*action_bits)                                   (SusStepIO.actions with SUS_PACKED)
 *   packed result record, result_bytes per env: agent i's reward code in bits [i*reward_bits, (i+1)*reward_bits),
 *     done in bit A*reward_bits, truncated in bit A*reward_bits + 1          (SusStepIO.packed_out)
 *   code = dead ? n_live_codes : event + 4 * (win + 3 * vote); event 0 none, 1 kill, 2 fix, 3 sabotage; win 0 none,
 *     1 crew, 2 imposters; vote 0 none, 1 crew member ejected, 2 imposter ejected.  An env whose actions were rejected
 *     (SUS_ERR_INVALID_ACTION) reports the all-ones code for every agent (decodes to NaN) with done = truncated = 0. */
typedef struct SusCompactLayout {
  int32_t action_bits, action_bytes;
  int32_t reward_bits, result_bytes;
  int32_t n_codes;      /* valid codes are 0 .. n_codes-1; n_codes-1 = dead */
  int32_t invalid_code; /* (1 << reward_bits) - 1 */
} SusCompactLayout;

typedef struct SusEnv *sus_env_t;

int sus_abi_version(void);
const char *sus_last_error(void); /* host string, valid until the thread's next failing call */

/* S = env.flattened_state_size (base.py:230-232); n_actions: env.n_imposter_actions / n_crew_actions
 * (base.py:203-204, tagging.py:35-36, pred_prey.py:69-73).  Host calls. */
int sus_flat_state_size(const SusConfig *cfg);
int sus_n_role_actions(const SusConfig *cfg, int is_imposter);
int sus_encode_shape(const SusConfig *cfg, const SusEncodeSpec *spec, SusEncodeShape *out /*host*/);
/* Record geometry of the compact host protocol for a config.  Host call. */
int sus_compact_layout(const SusConfig *cfg, SusCompactLayout *out /*host*/);
/* Decode table of the reward codes: out[i * (invalid_code + 1) + code] = float64 reward of agent index i for `code`
 * (NaN for codes >= n_codes); A * (invalid_code + 1) doubles.  Host call, no GPU work. */
int sus_reward_lut(const SusConfig *cfg, double *out /*host*/);

/* HOST helpers of the compact protocol (plain CPU code over host pointers, `threads` <= 0: pick; no GPU work).
 * sus_host_pack_actions: [N][A] role-list indices (SUS_U8 / SUS_I32 / SUS_I64) -> [N][action_bytes] records; an index that does
 * not fit its field becomes the field's maximum, which the kernel then rejects like any index outside the role list.
 * sus_host_decode_results: [N][result_bytes] records -> what step() returns per env (base.py:397-407): rewards [N][A] float32 or
 * float64 through the table of sus_reward_lut (NaN for a rejected env), done [N], truncated [N]; NULL outputs are skipped. */
int sus_host_pack_actions(const SusConfig *cfg, const void *actions /*host*/, int32_t dtype, int64_t n_envs,
                          uint8_t *packed /*host*/, int32_t threads);
int sus_host_decode_results(const SusConfig *cfg, const uint8_t *records /*host*/, int64_t n_envs, void *rewards /*host*/,
                            int32_t rewards_dtype, uint8_t *done /*host*/, uint8_t *truncated /*host*/, int32_t threads);

/* FourRoomEnv.__init__ & co. (base.py:103-228): validates like _validate_init_args (base.py:243-249,
 * pred_prey.py:75-76), builds the wall grid, allocates the structure-of-arrays state for num_envs envs on
 * `device`.  The envs are NOT reset. */
int sus_env_create(const SusConfig *cfg, int device, sus_env_t *out /*host*/);
int sus_env_destroy(sus_env_t env);

/* FourRoomEnv.reset (base.py:251-324; tagging.py:62-101) for every env whose mask byte is non-zero
 * (all envs if mask is NULL).  Kernel K0. */
int sus_env_reset(sus_env_t env, const uint8_t *mask_or_null, void *stream);

/* FourRoomEnv.step (base.py:332-407), FourRoomEnvWithTagging.step (tagging.py:120-235),
 * ImposterTrainingGround.check_win_condition (pred_prey.py:78-99), with auto-reset and the optional fused
 * observation encode.  Kernel K1 (+K2 fused).  Invalid role-list indices do not abort the launch: the env's
 * step is skipped and a device-side error counter is raised (see sus_env_check_actions). */
int sus_env_step(sus_env_t env, const SusStepIO *io /*host*/, void *stream);

/* n_steps random-policy steps of every env inside ONE launch (state stays in registers): identical to n_steps calls
 * of sus_env_step with actions == NULL -- same ticks, draws, auto-resets and episode statistics -- but only the final
 * state is written.  reward_sums: optional [N][A] float64, sum of each agent index's rewards over the rollout.
 * The inner loop of ReplayBuffer.populate / random-policy evaluation (replay_memory.py:103-143) without a launch
 * per step. */
int sus_env_rollout(sus_env_t env, int32_t n_steps, double *reward_sums, void *stream);

/* Synchronises `stream` and returns SUS_ERR_INVALID_ACTION if any step since the last check saw an action
 * index outside its agent's role list (reference: IndexError, base.py:381), else 0. */
int sus_env_check_actions(sus_env_t env, void *stream);

/* FourRoomEnv.sample_actions (base.py:326-330): uniform over each agent's role-specific list, dead agents
 * included.  Kernel K3.  out: [N][A] int32. */
int sus_env_sample_actions(sus_env_t env, int32_t *out, void *stream);

/* env.flatten_state(current state) (base.py:234-235) for all envs; dtype SUS_F32 / SUS_F64 / SUS_I64. */
int sus_env_export_flat(sus_env_t env, int32_t dtype, void *out /*[N][S]*/, void *stream);
/* Load states from flatten-order rows (inverse of the above) plus role masks and time steps. */
int sus_env_import_flat(sus_env_t env, const int64_t *flat /*[N][S]*/, const uint8_t *imposter_mask /*[N][A]*/,
                        const int32_t *t_or_null /*[N]*/, void *stream);
/* env.imposter_mask (base.py:280-281) as [N][A] bytes; env.imposter_idxs is its ascending index list. */
int sus_env_export_imposter_mask(sus_env_t env, uint8_t *out, void *stream);
/* env.metrics.get_metrics() (metrics.py:60-61) for all envs: [N][SUS_N_METRICS] int64. */
int sus_env_export_metrics(sus_env_t env, int64_t *out, void *stream);

/* SequenceStateFeaturizer.fit + generate_featurized_states on the envs' CURRENT states, T = 1.  Kernel K2. */
int sus_env_encode(sus_env_t env, const SusEncodeSpec *spec /*host*/, void *spatial, float *non_spatial,
                   void *stream);
/* SequenceStateFeaturizer.fit on a (B, T, S) batch of flattened states (train.py:70-74,346-348):
 * n_items = B*T rows of S values, dtype SUS_F32 / SUS_F64 / SUS_I64 (floats are truncated like
 * gymnasium.spaces.unflatten does).  Output item order = row order, so (B,T,...) views are free. */
int sus_encode_from_flat(const SusConfig *cfg /*host*/, const SusEncodeSpec *spec /*host*/, const void *states,
                         int32_t dtype, int64_t n_items, void *spatial, float *non_spatial, int device,
                         void *stream);

/* Finished-episode accumulators (SusStatIndex) of this handle since creation (or the last clear), int64[10].
 * These are what the multi-GPU driver all-reduces with NCCL at the end of a run. */
int sus_env_stats(sus_env_t env, int64_t *out, void *stream);
int sus_env_clear_stats(sus_env_t env, void *stream);

/* Per-agent running returns exactly as train() keeps them: G = reward + gamma * G after every step (train.py:386),
 * G[imposter_mask].mean() and G[~imposter_mask].mean() summed over finished episodes (train.py:421-424), G = 0 at
 * every episode start (train.py:324,436).  Together with sus_env_stats these are the 12 entries of the episode-stat
 * vector that the multi-GPU driver all-reduces.  Tracking costs 16 * A bytes of state traffic per env-step. */
int sus_env_track_returns(sus_env_t env, double gamma, void *stream);
/* out: float64[2] = {sum of imposter returns, sum of crew returns} over the episodes counted in SUS_S_EPISODES. */
int sus_env_return_sums(sus_env_t env, double *out, void *stream);

/* Device-resident launch ticks.  By default the tick of a launch (the Philox counter word that makes step t differ from
 * step t+1) is a host counter passed as a kernel parameter, so a captured CUDA graph would replay the SAME draws.  With
 * enable != 0 the three counters live in device memory: every launch reads its tick there and the last of its thread
 * blocks to have read it advances it for the next launch (no extra launch), which makes reset / sample_actions / step /
 * rollout launches capturable in a CUDA graph and replayable any number of times with the same results as individual
 * calls (no reference analogue: the reference has no launches to amortise).  sus_env_get_ticks / sus_env_set_ticks then synchronise the device.
 * enable == 0 copies the counters back to the host and returns to kernel-parameter ticks.  Not during a capture. */
int sus_env_device_ticks(sus_env_t env, int32_t enable, void *stream);

/* Launch ticks of the three Philox streams (host values), for checkpoint/resume. */
int sus_env_get_ticks(sus_env_t env, uint64_t *step_tick, uint64_t *reset_epoch, uint64_t *act_epoch /*host*/);
int sus_env_set_ticks(sus_env_t env, uint64_t step_tick, uint64_t reset_epoch, uint64_t act_epoch);

/* Raw structure-of-arrays state (device pointers, bytes per env) for state_dict()/load_state_dict(). */
int sus_env_state_arrays(sus_env_t env, void **ptrs /*host [4]*/, int32_t *bytes_per_env /*host [4]*/);

/* Device pointers of the finished-episode accumulators (int64[SUS_N_STATS]), the pending invalid-action counter (uint32)
 * and -- once sus_env_track_returns was called, else NULL / 0 -- the [A][N] running returns followed by the two return sums
 * (float64), with their sizes in bytes: what a checkpoint has to carry besides sus_env_state_arrays and the ticks. */
int sus_env_aux_arrays(sus_env_t env, void **ptrs /*host [3]*/, int64_t *bytes /*host [3]*/);

/* Parity mode: raw 32-bit words that replace the Philox output of the NEXT launch that draws from the
 * stream (step_words [N][2A-1], reset_words [N][n_imp+A+J], act_words [N][A]); NULL leaves a stream alone. */
int sus_env_debug_inject_words(sus_env_t env, const uint32_t *step_words, const uint32_t *reset_words,
                               const uint32_t *act_words);

/* ---- row (f1): replay ring in the reference layout (src/replay_memory.py:33-44) -------------------------
 * One launch stores the N transitions of a batched step at ring slots (idx + e) mod M and advances every env's
 * T-deep state sequence exactly like ReplayBuffer.populate / train() do on the host: next_sequence =
 * np.roll(sequence, -1, axis=0) with the new state in the last row (replay_memory.py:120-126, train.py:388-389);
 * an env whose episode ended restarts from T copies of its reset state (replay_memory.py:107-112,
 * train.py:440-445). */
typedef struct SusReplayPush {
  int64_t N, M, idx;          /* transitions per launch, ring capacity, first slot */
  int32_t T, S, A, n_imposters;
  const float *seq_in;        /* [N][T][S] running state sequence of every env before the step */
  float *seq_out;             /* [N][T][S] the sequence the next step starts from (may be seq_in itself when T == 1,
                                 must not alias it otherwise) */
  const float *next_flat;     /* [N][S] post-step, pre-reset state (SusStepIO.next_flat) */
  const float *cur_flat;      /* [N][S] state the next action is taken from (sus_env_export_flat) */
  const void *actions;        /* [N][A] */
  int32_t actions_dtype;      /* SUS_U8 / SUS_I32 / SUS_I64 */
  int32_t reserved;
  const float *rewards;       /* [N][A] */
  const uint8_t *done;        /* [N] */
  const uint8_t *truncated;   /* [N] */
  const int16_t *imposters;   /* [N][n_imposters] (SusStepIO.imposters) */
  float *states;              /* [M][T][S]  ReplayBuffer.states      */
  int64_t *r_actions;         /* [M][A]     ReplayBuffer.actions     */
  float *r_rewards;           /* [M][A]     ReplayBuffer.rewards     */
  float *next_states;         /* [M][T][S]  ReplayBuffer.next_states */
  uint8_t *r_dones;           /* [M][1]     ReplayBuffer.dones       */
  int16_t *r_imposters;       /* [M][n_imp] ReplayBuffer.imposters   */
  const int64_t *idx_dev;     /* optional DEVICE copy of the first slot: used instead of `idx` when not NULL, so that a captured
                                 CUDA graph can be replayed while the caller advances the index on the device */
} SusReplayPush;

/* ReplayBuffer.add (replay_memory.py:50-72) for a whole batched step; the caller advances idx/size. */
int sus_replay_push(const SusReplayPush *args /*host*/, int device, void *stream);

/* ---- row (f2): the acting part of train() (src/train.py:349-381) for all envs in one launch ---------------------
 * Per agent view i of env e: an ALIVE imposter explores with probability eps -- `np.random.random() <= eps`, uniform over
 * its role list (train.py:363-366) -- or takes the FIRST argmax of its network's Q-values (train.py:367-370); an alive crew
 * member likewise with the crew network (train.py:373-381); a dead agent keeps action 0 (train.py:352).  The Q-values come
 * from the caller's networks (any framework; dense layers are outside this library):
 *   q_imposter  [N][n_imposter_actions], row e = Q-values of env e's single imposter evaluated on ITS agent view
 *               (imposter_per_view == 0, needs n_imposters == 1), or [A][N][n_imposter_actions] (imposter_per_view == 1);
 *               NULL = the imposters act uniformly at random (the reference's RandomEquiprobable model)
 *   q_crew      [A][N][n_crew_actions], view-major; NULL = the crew acts uniformly at random
 *   eps         DEVICE pointer to epsilon (so a captured CUDA graph sees the current value), or NULL to use eps_value
 * Draws: Philox words keyed (seed, global env id, act epoch, purpose 5); slot 2i = agent i's explore word
 * (explore iff word * 2^-32 <= eps), slot 2i + 1 = its random action (bounded).  Consumes one act epoch like
 * sus_env_sample_actions.  actions: [N][A] role-list indices, SUS_I32 or SUS_U8. */
typedef struct SusPolicyIO {
  const float *q_imposter;
  const float *q_crew;
  const float *eps;
  float eps_value;
  int32_t imposter_per_view;
  int32_t actions_dtype;
  int32_t reserved;
  void *actions;
} SusPolicyIO;
int sus_env_select_actions(sus_env_t env, const SusPolicyIO *io /*host*/, void *stream);

/* Q-network inference for the reference's MLP estimator (src/models/dqn.py:72-108: `make_mlp`, a Linear + activation stack on
 * the flattened non-spatial features; train.py:367-370,378-381 evaluate it per agent per step): out = MLP(x) for n_rows rows in
 * ONE launch, activations resident in shared memory, fp32 FFMA (same arithmetic as the float32 reference up to summation order).
 * weight[l] is [dims[l+1]][dims[l]] row-major (torch.nn.Linear.weight), bias[l] [dims[l+1]] or NULL, alpha[l] the single PReLU
 * slope after layer l (torch.nn.PReLU(), num_parameters = 1) -- DEVICE pointers to the live parameters, nothing is copied.
 * The activation follows every layer but the last.  SUS_ERR_UNSUPPORTED if two adjacent widths do not fit in shared memory. */
#define SUS_MLP_MAX_LAYERS 8
enum SusActivation { SUS_ACT_NONE = 0, SUS_ACT_RELU = 1, SUS_ACT_PRELU = 2 };
typedef struct SusMlpSpec {
  int32_t n_layers;
  int32_t activation;                       /* SusActivation */
  int32_t dims[SUS_MLP_MAX_LAYERS + 1];     /* dims[0] = input width ... dims[n_layers] = outputs */
  int32_t reserved;
  const float *weight[SUS_MLP_MAX_LAYERS];
  const float *bias[SUS_MLP_MAX_LAYERS];
  const float *alpha[SUS_MLP_MAX_LAYERS];
} SusMlpSpec;
int sus_mlp_forward(const SusMlpSpec *spec /*host*/, const float *x /*[n_rows][dims[0]]*/, int64_t n_rows,
                    float *out /*[n_rows][dims[n_layers]]*/, int device, void *stream);
/* The same with a caller-owned DEVICE workspace of sus_mlp_workspace_bytes(spec) bytes (16-byte aligned; 0 = invalid spec): a
 * small kernel in front of the forward repacks the live weights into it, chunk by chunk in the order the forward kernel stages
 * them, so that its staging becomes straight 128-bit copies (0.467 -> 0.43 ms at cfg5).  The workspace belongs to one network
 * and one stream at a time; workspace == NULL is sus_mlp_forward. */
int64_t sus_mlp_workspace_bytes(const SusMlpSpec *spec /*host*/);
int sus_mlp_forward_ws(const SusMlpSpec *spec /*host*/, const float *x, int64_t n_rows, float *out, void *workspace /*device*/,
                       int64_t workspace_bytes, int device, void *stream);

/* T-deep sequences of ENCODED features (train.py:318-322,388-389,440-445 keep them as raw states and re-encode all T
 * every iteration): seq_out[r][t] = newest[r] if t == T-1 or env (r mod n_envs)'s episode just ended (done | truncated),
 * else seq_in[r][t+1].  rows = views * n_envs items of R floats per time step (view-major tensors [views][N][T][R]);
 * `newest` [rows][R] is what the fused step wrote.  seq_in and seq_out must not alias. */
int sus_seq_roll(const float *seq_in, float *seq_out, const float *newest, const uint8_t *done, const uint8_t *truncated,
                 int64_t rows, int64_t n_envs, int32_t T, int32_t R, int device, void *stream);

/* L2-compressible device memory for the feature tensors (no reference analogue; the reference's tensors live in host
 * memory).  The planes and flat rows are almost all zeros; in an allocation made with cuMemCreate +
 * CU_MEM_ALLOCATION_COMP_GENERIC Blackwell's L2 keeps such lines compressed on their way to and from HBM (measured on
 * B200: the fused kernel's tile-store stream 6.4 -> 7.5 TB/s, reading the tiles back 6.9 -> 9.3 TB/s).  Any pointer this
 * library takes may point into such a block.  `bytes` is rounded up to the allocation granularity (2 MiB); `allocated`
 * (optional) receives the rounded size.  SUS_ERR_UNSUPPORTED if the device or driver has no generic compression.
 * sus_free_compressible synchronises the device. */
int sus_alloc_compressible(int device, uint64_t bytes, void **ptr /*host*/, uint64_t *allocated /*host, may be NULL*/);
int sus_free_compressible(void *ptr);

/* Number of kernels this library has launched in the calling process (bench.py's gpu_launches). */
int64_t sus_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SUSNET_B200_H_ */
