/*
 * susnet_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar CPU restatement of Sus-Net's grid environment (reset / step / sample_actions /
 * flatten) and of its observation featurizers, one environment at a time exactly like
 * the Python reference does it.  It exists so that the CUDA kernels in
 * sus_net_b200/csrc can be checked bit for bit on a box that has no copy of the
 * reference.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this file's shared object; the product path never does.
 *
 * Parity pinning: the reference has no golden vectors.  This file is
 * pinned against the reference ITSELF: tools/make_golden.py and tests/test_oracle_vs_reference.py
 * run the unmodified Python reference on injected draws (oracle/ref_harness.py) and
 * require identical states, rewards, dones, truncations, metrics and features; the
 * committed fixtures under tests/golden/ carry those reference outputs to the GPU box.
 *
 * Reference anchors (file:line under /root/reference/src):
 *   geometry            environment/base.py:171-199
 *   reset               environment/base.py:251-324, tagging.py:62-101
 *   sample_actions      environment/base.py:326-330
 *   step (base)         environment/base.py:332-407
 *   _agent_step         environment/base.py:462-533
 *   win (base)          environment/base.py:409-460
 *   win (training gr.)  environment/pred_prey.py:78-99
 *   _merge_rewards      environment/base.py:553-563
 *   step (tagging)      environment/tagging.py:103-118,120-241
 *   flatten order       environment/base.py:211-241, tagging.py:42-60
 *   featurizers         features/component.py:83-131,200-482, features/model_ready.py:82-370
 *
 * Random draws follow the susnet Philox spec written down in oracle/rng_spec.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAX_AGENTS 8
#define ORC_MAX_JOBS 8
#define ORC_N_METRICS 8
#define ORC_N_STATS 10

enum { V_BASE = 0, V_TAGGING = 1, V_TRAINING_GROUND = 2 };
enum { ACT_STAY = 0, ACT_UP, ACT_DOWN, ACT_LEFT, ACT_RIGHT, ACT_KILL, ACT_FIX, ACT_SABOTAGE, ACT_TAG0 = 100 };
enum { P_STEP = 0, P_AUTORESET = 1, P_ACT = 2, P_RESET = 3, P_ACT_FUSED = 4 };
enum { M_STEPS = 0, M_KILLS, M_COMPLETED, M_SABOTAGED, M_IMP_VOTED, M_CREW_VOTED, M_CREW_WON, M_IMP_WON };

typedef struct {
  int32_t variant;
  int32_t n_imposters, n_crew, n_jobs;
  int32_t include_walls, is_action_order_random, shuffle_imposter_index;
  int32_t max_time_steps, tag_reset_interval;
  double kill_reward, complete_job_reward, sabotage_reward, time_step_reward;
  double game_end_reward, dead_penalty, vote_reward;
} orc_config;

typedef struct {
  int64_t pos[ORC_MAX_AGENTS][2];
  uint8_t alive[ORC_MAX_AGENTS];
  int64_t jobpos[ORC_MAX_JOBS][2];
  uint8_t completed[ORC_MAX_JOBS];
  uint8_t imposter[ORC_MAX_AGENTS];
  int64_t tag_counts[ORC_MAX_AGENTS];
  uint8_t used_tag[ORC_MAX_AGENTS];
  int32_t tag_timer;
  int32_t t;
  int64_t metrics[ORC_N_METRICS];
  double G[ORC_MAX_AGENTS]; /* running returns as train() keeps them (train.py:324,386,436) */
} orc_env;

typedef struct {
  orc_config cfg;
  int32_t N, A, J, V, S;
  uint64_t seed;
  uint32_t env_id_base;
  int32_t auto_reset;
  uint64_t step_tick, reset_epoch, act_epoch;
  uint8_t grid[9][9];      /* 1 = free */
  int32_t valid_xy[81][2]; /* row-major argwhere(grid): base.py:199 */
  orc_env *envs;
  int64_t stats[ORC_N_STATS];
  int32_t track_returns;
  double gamma, ret_sums[2]; /* sums of G[imposter_mask].mean() / G[~imposter_mask].mean() at episode ends */
  const uint32_t *inj_step, *inj_reset, *inj_act; /* injected raw words for the next launch */
} orc_handle;

/* ------------------------------------------------------------------ Philox4x32-10 */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static uint32_t spec_word(const orc_handle *h, int env, uint64_t tick, int purpose, int slot) {
  uint32_t ctr[4] = {h->env_id_base + (uint32_t)env, (uint32_t)tick, (uint32_t)(tick >> 32),
                     (uint32_t)purpose | ((uint32_t)(slot >> 2) << 8)};
  uint32_t key[2] = {(uint32_t)h->seed, (uint32_t)(h->seed >> 32)};
  uint32_t out[4];
  philox4x32_10(ctr, key, out);
  return out[slot & 3];
}

static inline uint32_t bounded(uint32_t u, uint32_t k) { return (uint32_t)(((uint64_t)u * k) >> 32); }

static inline int n_step_slots(int A) { return 2 * A - 1; }
static inline int n_reset_slots(const orc_handle *h) { return h->cfg.n_imposters + h->A + h->J; }

static uint32_t step_word(const orc_handle *h, int env, uint64_t tick, int slot) {
  if (h->inj_step) return h->inj_step[(size_t)env * n_step_slots(h->A) + slot];
  return spec_word(h, env, tick, P_STEP, slot);
}
static uint32_t reset_word(const orc_handle *h, int env, uint64_t tick, int purpose, int slot) {
  if (h->inj_reset) return h->inj_reset[(size_t)env * n_reset_slots(h) + slot];
  return spec_word(h, env, tick, purpose, slot);
}
static uint32_t act_word(const orc_handle *h, int env, uint64_t tick, int purpose, int slot) {
  if (h->inj_act) return h->inj_act[(size_t)env * h->A + slot];
  return spec_word(h, env, tick, purpose, slot);
}

/* r-th smallest id not yet in chosen[] (kept ascending); appends it. rng_spec.pick_distinct */
static int pick_unchosen(int r, int *chosen_sorted, int n_chosen) {
  for (int i = 0; i < n_chosen; ++i)
    if (r >= chosen_sorted[i]) r++;
  int i = n_chosen;
  while (i > 0 && chosen_sorted[i - 1] > r) { chosen_sorted[i] = chosen_sorted[i - 1]; --i; }
  chosen_sorted[i] = r;
  return r;
}

/* ------------------------------------------------------------------ geometry */
static void build_geometry(orc_handle *h) {
  static const int walls[13][2] = {{0, 4}, {2, 4}, {3, 4}, {4, 4}, {5, 4}, {6, 4}, {8, 4},
                                   {4, 0}, {4, 2}, {4, 3}, {4, 5}, {4, 6}, {4, 8}}; /* base.py:172-188 */
  memset(h->grid, 1, sizeof(h->grid));
  if (h->cfg.include_walls)
    for (int w = 0; w < 13; ++w) h->grid[walls[w][0]][walls[w][1]] = 0; /* base.py:195-197 */
  h->V = 0;
  for (int i = 0; i < 9; ++i)
    for (int j = 0; j < 9; ++j)
      if (h->grid[i][j]) { h->valid_xy[h->V][0] = i; h->valid_xy[h->V][1] = j; h->V++; } /* base.py:199 */
}

static int flat_size(const orc_config *c) {
  int A = c->n_imposters + c->n_crew, J = c->n_jobs;
  int S = 3 * A + (J > 0 || c->variant == V_TAGGING ? 3 * J : 0); /* base.py:211-228 */
  if (c->variant == V_TAGGING) S += 2 * A + 1;                    /* tagging.py:42-60 */
  return S;
}

static int n_role_actions(const orc_config *c, int is_imposter) {
  int A = c->n_imposters + c->n_crew;
  if (c->variant == V_TRAINING_GROUND) return is_imposter ? 6 : 5; /* pred_prey.py:4-19 */
  int base = is_imposter ? 7 : 6;                                  /* base.py:82-99 */
  return c->variant == V_TAGGING ? base + A - 1 : base;            /* tagging.py:35-36,69-75 */
}

/* role-list index -> Action (or ACT_TAG0 + target); -1 if out of the role list (reference: IndexError) */
static int decode_action(const orc_config *c, int agent, int is_imposter, int idx) {
  if (idx < 0 || idx >= n_role_actions(c, is_imposter)) return -1;
  if (idx <= 4) return idx; /* STAY, UP, DOWN, LEFT, RIGHT share positions in every list */
  if (c->variant == V_TRAINING_GROUND) return ACT_KILL; /* imposter idx 5 */
  int nbase = is_imposter ? 7 : 6;
  if (idx < nbase) {
    if (!is_imposter) return ACT_FIX;          /* crew idx 5 */
    return idx == 5 ? ACT_SABOTAGE : ACT_KILL; /* imposter idx 5, 6 */
  }
  int k = idx - nbase; /* tagging.py:70-75: targets = all other agent ids ascending */
  int target = k < agent ? k : k + 1;
  return ACT_TAG0 + target;
}

/* ------------------------------------------------------------------ reset */
static void env_reset(orc_handle *h, int e, uint64_t tick, int purpose) {
  orc_env *s = &h->envs[e];
  const orc_config *c = &h->cfg;
  int A = h->A, J = h->J, nI = c->n_imposters;
  memset(s->metrics, 0, sizeof(s->metrics)); /* base.py:270 */
  memset(s->imposter, 0, sizeof(s->imposter));
  if (c->shuffle_imposter_index) { /* base.py:273-276 (R1): ascending-sorted distinct subset */
    int chosen[ORC_MAX_AGENTS];
    for (int m = 0; m < nI; ++m) {
      int r = (int)bounded(reset_word(h, e, tick, purpose, m), (uint32_t)(A - m));
      s->imposter[pick_unchosen(r, chosen, m)] = 1;
    }
  } else {
    for (int m = 0; m < nI; ++m) s->imposter[m] = 1; /* base.py:278 */
  }
  for (int i = 0; i < A; ++i) { /* base.py:288-291 (R2) */
    int cell = (int)bounded(reset_word(h, e, tick, purpose, nI + i), (uint32_t)h->V);
    s->pos[i][0] = h->valid_xy[cell][0];
    s->pos[i][1] = h->valid_xy[cell][1];
    s->alive[i] = 1; /* base.py:301 */
  }
  int chosen[ORC_MAX_JOBS];
  for (int j = 0; j < J; ++j) { /* base.py:295-299 (R3) */
    int r = (int)bounded(reset_word(h, e, tick, purpose, nI + A + j), (uint32_t)(h->V - j));
    int cell = pick_unchosen(r, chosen, j);
    s->jobpos[j][0] = h->valid_xy[cell][0];
    s->jobpos[j][1] = h->valid_xy[cell][1];
    s->completed[j] = 0; /* base.py:302 */
  }
  s->t = 0; /* base.py:315 */
  memset(s->tag_counts, 0, sizeof(s->tag_counts)); /* tagging.py:64-66 */
  memset(s->used_tag, 0, sizeof(s->used_tag));
  s->tag_timer = 0;
}

/* ------------------------------------------------------------------ flatten */
static void env_flatten(const orc_handle *h, const orc_env *s, int64_t *out) {
  const orc_config *c = &h->cfg;
  int A = h->A, J = h->J, k = 0;
  for (int i = 0; i < A; ++i) { out[k++] = s->pos[i][0]; out[k++] = s->pos[i][1]; }
  for (int i = 0; i < A; ++i) out[k++] = s->alive[i];
  if (J > 0 || c->variant == V_TAGGING) {
    for (int j = 0; j < J; ++j) { out[k++] = s->jobpos[j][0]; out[k++] = s->jobpos[j][1]; }
    for (int j = 0; j < J; ++j) out[k++] = s->completed[j];
  }
  if (c->variant == V_TAGGING) { /* tagging.py:221-230 */
    for (int i = 0; i < A; ++i) out[k++] = s->used_tag[i];
    for (int i = 0; i < A; ++i) out[k++] = s->tag_counts[i];
    out[k++] = c->tag_reset_interval - s->tag_timer;
  }
}

/* ------------------------------------------------------------------ step */
static int is_valid_position(const orc_handle *h, int64_t x, int64_t y) { /* base.py:548-551 */
  if (x < 0 || y < 0 || x >= 9 || y >= 9) return 0;
  return h->grid[y][x];
}

typedef struct { int kill_events; } step_ctx;

static void agent_step(orc_handle *h, int e, uint64_t tick, orc_env *s, double *rew, int agent, int action,
                       step_ctx *ctx) { /* base.py:462-533 */
  int A = h->A, J = h->J;
  const orc_config *c = &h->cfg;
  if (!s->alive[agent]) return; /* base.py:477 */
  int64_t x = s->pos[agent][0], y = s->pos[agent][1];
  if (action <= ACT_RIGHT) { /* base.py:484-487, move(): base.py:69-79 */
    int64_t nx = x, ny = y;
    if (action == ACT_UP) ny = y + 1;
    else if (action == ACT_DOWN) ny = y - 1;
    else if (action == ACT_LEFT) nx = x - 1;
    else if (action == ACT_RIGHT) nx = x + 1;
    if (is_valid_position(h, nx, ny)) { s->pos[agent][0] = nx; s->pos[agent][1] = ny; }
  } else if (action == ACT_KILL) { /* base.py:490-515 */
    int cand[ORC_MAX_AGENTS], k = 0;
    for (int i = 0; i < A; ++i) /* alive crew at the killer's cell, ascending: base.py:535-542 */
      if (s->alive[i] && !s->imposter[i] && s->pos[i][0] == x && s->pos[i][1] == y) cand[k++] = i;
    if (k > 0) {
      uint32_t w = step_word(h, e, tick, A - 1 + ctx->kill_events);
      ctx->kill_events++;
      int victim = cand[bounded(w, (uint32_t)k)]; /* base.py:497 (R5) */
      s->metrics[M_KILLS] += 1;
      s->alive[victim] = 0;
      rew[victim] = c->kill_reward; /* assignment, not accumulation: base.py:514-515 */
      rew[agent] = c->kill_reward;
    }
  } else if (action == ACT_FIX || action == ACT_SABOTAGE) { /* base.py:518-533 */
    int job = -1;
    for (int j = 0; j < J; ++j) /* first job at the cell: base.py:544-546 */
      if (s->jobpos[j][0] == x && s->jobpos[j][1] == y) { job = j; break; }
    if (job >= 0) {
      if (action == ACT_FIX && !s->completed[job]) {
        s->completed[job] = 1;
        s->metrics[M_COMPLETED] += 1;
        rew[agent] = c->complete_job_reward;
      } else if (action == ACT_SABOTAGE && s->completed[job]) {
        s->completed[job] = 0;
        s->metrics[M_SABOTAGED] += 1;
        rew[agent] = -1 * c->sabotage_reward;
      }
    }
  }
}

static void check_win(const orc_handle *h, orc_env *s, int *done, double *team_reward) {
  const orc_config *c = &h->cfg;
  int A = h->A, J = h->J, alive_imp = 0, alive_crew = 0, n_completed = 0;
  for (int i = 0; i < A; ++i) {
    if (s->alive[i] && s->imposter[i]) alive_imp++;
    if (s->alive[i] && !s->imposter[i]) alive_crew++;
  }
  for (int j = 0; j < J; ++j) n_completed += s->completed[j];
  *done = 0;
  *team_reward = 0;
  if (c->variant == V_TRAINING_GROUND) { /* pred_prey.py:78-99 */
    if (J != 0 && n_completed == J) { s->metrics[M_CREW_WON] = 1; *done = 1; *team_reward = c->game_end_reward; return; }
    if (alive_crew == 0) { s->metrics[M_IMP_WON] = 1; *done = 1; *team_reward = -1 * c->game_end_reward; return; }
    return;
  }
  if (alive_imp == 0 || n_completed == J) { /* base.py:428-435 (true at once when J == 0) */
    s->metrics[M_CREW_WON] = 1; *done = 1; *team_reward = c->game_end_reward;
  } else if (alive_crew <= alive_imp) { /* base.py:438-446 */
    s->metrics[M_IMP_WON] = 1; *done = 1; *team_reward = -1 * c->game_end_reward;
  }
}

static void merge_rewards(const orc_handle *h, const orc_env *s, double *rew, double team_reward) { /* base.py:553-563 */
  for (int i = 0; i < h->A; ++i) rew[i] += team_reward;
  for (int i = 0; i < h->cfg.n_imposters; ++i) rew[i] *= -1; /* by INDEX, not by mask: base.py:559 */
  for (int i = 0; i < h->A; ++i) if (!s->alive[i]) rew[i] = h->cfg.dead_penalty;
}

/* returns 0 ok, -1 invalid action */
static int env_step(orc_handle *h, int e, uint64_t tick, const int32_t *actions, double *rew, uint8_t *done_out,
                    uint8_t *trunc_out) {
  orc_env *s = &h->envs[e];
  const orc_config *c = &h->cfg;
  int A = h->A;
  int decoded[ORC_MAX_AGENTS];
  for (int i = 0; i < A; ++i) { /* validated up front (the reference raises mid-step: documented deviation) */
    decoded[i] = decode_action(c, i, s->imposter[i], actions[i]);
    if (decoded[i] < 0) return -1;
  }
  s->metrics[M_STEPS] += 1; /* base.py:366, tagging.py:152 */
  double team_reward = 0;
  for (int i = 0; i < A; ++i) rew[i] = (c->variant == V_TAGGING) ? 1.0 * c->time_step_reward : 0.0; /* tagging.py:162 / base.py:369 */
  int order[ORC_MAX_AGENTS];
  for (int i = 0; i < A; ++i) order[i] = i;
  if (c->is_action_order_random) /* base.py:372-374 (R4) */
    for (int k = A - 1; k > 0; --k) {
      int j = (int)bounded(step_word(h, e, tick, A - 1 - k), (uint32_t)(k + 1));
      int tmp = order[k]; order[k] = order[j]; order[j] = tmp;
    }
  step_ctx ctx = {0};
  for (int k = 0; k < A; ++k) {
    int agent = order[k], act = decoded[agent];
    if (act >= ACT_TAG0) { /* tagging.py:103-110: the tagger's own liveness is NOT checked */
      int target = act - ACT_TAG0;
      if (!s->used_tag[agent] && s->alive[target]) { s->tag_counts[target] += 1; s->used_tag[agent] = 1; }
    } else {
      agent_step(h, e, tick, s, rew, agent, act, &ctx);
    }
  }
  if (c->variant == V_TAGGING) { /* tagging.py:180-207 */
    for (int i = 0; i < A; ++i) s->tag_counts[i] *= s->alive[i];
    s->tag_timer += 1;
    if (s->tag_timer >= c->tag_reset_interval) {
      int best = 0, n_alive = 0;
      for (int i = 1; i < A; ++i) if (s->tag_counts[i] > s->tag_counts[best]) best = i; /* argmax: first max */
      for (int i = 0; i < A; ++i) n_alive += s->alive[i];
      int64_t quorum = (n_alive + 1) / 2; /* counted BEFORE the eject */
      if (s->tag_counts[best] >= quorum) {
        s->alive[best] = 0;
        team_reward += c->vote_reward * (s->imposter[best] ? -1 : 1);
        s->metrics[s->imposter[best] ? M_IMP_VOTED : M_CREW_VOTED] += 1;
      }
      memset(s->tag_counts, 0, sizeof(s->tag_counts));
      memset(s->used_tag, 0, sizeof(s->used_tag));
      s->tag_timer = 0;
    }
  }
  int done; double win_reward;
  check_win(h, s, &done, &win_reward);
  team_reward += win_reward;
  merge_rewards(h, s, rew, team_reward);
  if (c->variant != V_TAGGING) /* base.py:389-390; the tagging step has no such replacement */
    for (int i = 0; i < A; ++i) if (rew[i] == 0) rew[i] = c->time_step_reward;
  int trunc = 0;
  if (s->t == c->max_time_steps - 1) trunc = 1; else s->t += 1; /* base.py:392-395 */
  *done_out = (uint8_t)done;
  *trunc_out = (uint8_t)trunc;
  return 0;
}

/* ------------------------------------------------------------------ public API */
int orc_flat_size(const orc_config *c) { return flat_size(c); }
int orc_n_metrics(void) { return ORC_N_METRICS; }
int orc_n_stats(void) { return ORC_N_STATS; }
int orc_n_role_actions(const orc_config *c, int is_imposter) { return n_role_actions(c, is_imposter); }

int orc_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  (void)n; return 1;
#endif
}

static int check_config(const orc_config *c) {
  int A = c->n_imposters + c->n_crew;
  if (c->variant < 0 || c->variant > 2) return -1;
  if (c->n_imposters < 1 || c->n_crew < 1 || c->n_jobs < 0) return -1;
  if (A > ORC_MAX_AGENTS || c->n_jobs > ORC_MAX_JOBS) return -1;
  if (c->variant == V_TAGGING && c->n_jobs == 0) return -1;
  if (c->variant == V_TRAINING_GROUND && c->n_imposters != 1) return -1;
  if (c->max_time_steps < 1) return -1;
  return 0;
}

int orc_create(const orc_config *cfg, int num_envs, uint64_t seed, uint32_t env_id_base, int auto_reset,
               orc_handle **out) {
  if (check_config(cfg) || num_envs < 0) return -1;
  orc_handle *h = (orc_handle *)calloc(1, sizeof(orc_handle));
  h->cfg = *cfg;
  h->N = num_envs;
  h->A = cfg->n_imposters + cfg->n_crew;
  h->J = cfg->n_jobs;
  h->S = flat_size(cfg);
  h->seed = seed;
  h->env_id_base = env_id_base;
  h->auto_reset = auto_reset;
  build_geometry(h);
  h->envs = (orc_env *)calloc(num_envs > 0 ? num_envs : 1, sizeof(orc_env));
  *out = h;
  return 0;
}

void orc_destroy(orc_handle *h) {
  if (!h) return;
  free(h->envs);
  free(h);
}

int orc_inject_words(orc_handle *h, const uint32_t *step_words, const uint32_t *reset_words, const uint32_t *act_words) {
  h->inj_step = step_words; h->inj_reset = reset_words; h->inj_act = act_words;
  return 0;
}

int orc_reset(orc_handle *h, const uint8_t *mask) {
  uint64_t tick = h->reset_epoch++;
#pragma omp parallel for schedule(static)
  for (int e = 0; e < h->N; ++e)
    if (!mask || mask[e]) env_reset(h, e, tick, P_RESET);
  h->inj_reset = NULL;
  return 0;
}

int orc_sample_actions(orc_handle *h, int32_t *out) { /* base.py:326-330 (R6): dead agents are sampled too */
  uint64_t tick = h->act_epoch++;
#pragma omp parallel for schedule(static)
  for (int e = 0; e < h->N; ++e)
    for (int i = 0; i < h->A; ++i)
      out[(size_t)e * h->A + i] =
          (int32_t)bounded(act_word(h, e, tick, P_ACT, i), (uint32_t)n_role_actions(&h->cfg, h->envs[e].imposter[i]));
  h->inj_act = NULL;
  return 0;
}

/* actions == NULL: fused random policy (P_ACT_FUSED); the drawn actions go to actions_out if given.
 * next_flat / metrics_out report the post-step, PRE-reset state and episode counters. */
int orc_step(orc_handle *h, const int32_t *actions, int32_t *actions_out, double *rewards, uint8_t *done,
             uint8_t *trunc, int64_t *next_flat, int64_t *metrics_out) {
  uint64_t tick = h->step_tick++;
  int A = h->A, bad = 0;
  int64_t stats[ORC_N_STATS] = {0};
#pragma omp parallel
  {
    int64_t loc[ORC_N_STATS] = {0};
#pragma omp for schedule(static)
    for (int e = 0; e < h->N; ++e) {
      int32_t act[ORC_MAX_AGENTS];
      double rew[ORC_MAX_AGENTS];
      orc_env *s = &h->envs[e];
      for (int i = 0; i < A; ++i)
        act[i] = actions ? actions[(size_t)e * A + i]
                         : (int32_t)bounded(act_word(h, e, tick, P_ACT_FUSED, i),
                                            (uint32_t)n_role_actions(&h->cfg, s->imposter[i]));
      if (actions_out) for (int i = 0; i < A; ++i) actions_out[(size_t)e * A + i] = act[i];
      uint8_t d = 0, t = 0;
      if (env_step(h, e, tick, act, rew, &d, &t)) {
#pragma omp atomic write
        bad = 1;
        continue;
      }
      for (int i = 0; i < A; ++i) rewards[(size_t)e * A + i] = rew[i];
      done[e] = d; trunc[e] = t;
      if (next_flat) env_flatten(h, s, next_flat + (size_t)e * h->S);
      if (metrics_out) memcpy(metrics_out + (size_t)e * ORC_N_METRICS, s->metrics, sizeof(s->metrics));
      if (h->track_returns) { /* train.py:386: G = reward + gamma * G */
        double gi = 0, gc = 0;
        int ni = 0, nc = 0;
        for (int i = 0; i < A; ++i) {
          s->G[i] = rew[i] + h->gamma * s->G[i];
          if (s->imposter[i]) { gi += s->G[i]; ni++; } else { gc += s->G[i]; nc++; }
        }
        if (d || t) { /* train.py:421-424,436 */
#pragma omp critical
          { h->ret_sums[0] += gi / ni; h->ret_sums[1] += gc / nc; }
          for (int i = 0; i < A; ++i) s->G[i] = 0;
        }
      }
      if (d || t) {
        loc[0] += 1; loc[1] += s->metrics[M_CREW_WON]; loc[2] += s->metrics[M_IMP_WON];
        loc[3] += s->metrics[M_KILLS]; loc[4] += s->metrics[M_COMPLETED]; loc[5] += s->metrics[M_SABOTAGED];
        loc[6] += s->metrics[M_IMP_VOTED]; loc[7] += s->metrics[M_CREW_VOTED]; loc[8] += s->metrics[M_STEPS];
        loc[9] += t;
        if (h->auto_reset) env_reset(h, e, tick, P_AUTORESET); /* SURVEY.md A.7, train.py:419-445 */
      }
    }
#pragma omp critical
    for (int k = 0; k < ORC_N_STATS; ++k) stats[k] += loc[k];
  }
  for (int k = 0; k < ORC_N_STATS; ++k) h->stats[k] += stats[k];
  h->inj_step = NULL; h->inj_reset = NULL; h->inj_act = NULL;
  return bad ? -2 : 0;
}

int orc_export_flat(const orc_handle *h, int64_t *out) {
#pragma omp parallel for schedule(static)
  for (int e = 0; e < h->N; ++e) env_flatten(h, &h->envs[e], out + (size_t)e * h->S);
  return 0;
}

int orc_export_metrics(const orc_handle *h, int64_t *out) {
  for (int e = 0; e < h->N; ++e) memcpy(out + (size_t)e * ORC_N_METRICS, h->envs[e].metrics, sizeof(h->envs[e].metrics));
  return 0;
}

int orc_imposter_mask(const orc_handle *h, uint8_t *out) {
  for (int e = 0; e < h->N; ++e) memcpy(out + (size_t)e * h->A, h->envs[e].imposter, (size_t)h->A);
  return 0;
}

int orc_track_returns(orc_handle *h, double gamma) {
  h->track_returns = 1; h->gamma = gamma; h->ret_sums[0] = h->ret_sums[1] = 0;
  for (int e = 0; e < h->N; ++e) memset(h->envs[e].G, 0, sizeof(h->envs[e].G));
  return 0;
}
int orc_return_sums(const orc_handle *h, double *out) { out[0] = h->ret_sums[0]; out[1] = h->ret_sums[1]; return 0; }
int orc_export_returns(const orc_handle *h, double *out) { /* [N][A] */
  for (int e = 0; e < h->N; ++e) memcpy(out + (size_t)e * h->A, h->envs[e].G, sizeof(double) * (size_t)h->A);
  return 0;
}

int orc_stats(const orc_handle *h, int64_t *out) { memcpy(out, h->stats, sizeof(h->stats)); return 0; }

/* Load env state from flat rows (reference flatten order) + imposter mask; t/metrics as given. */
int orc_import_flat(orc_handle *h, const int64_t *flat, const uint8_t *imposter_mask, const int32_t *t) {
  int A = h->A, J = h->J;
  for (int e = 0; e < h->N; ++e) {
    orc_env *s = &h->envs[e];
    const int64_t *f = flat + (size_t)e * h->S;
    int k = 0;
    for (int i = 0; i < A; ++i) { s->pos[i][0] = f[k++]; s->pos[i][1] = f[k++]; }
    for (int i = 0; i < A; ++i) s->alive[i] = f[k++] != 0;
    if (J > 0 || h->cfg.variant == V_TAGGING) {
      for (int j = 0; j < J; ++j) { s->jobpos[j][0] = f[k++]; s->jobpos[j][1] = f[k++]; }
      for (int j = 0; j < J; ++j) s->completed[j] = f[k++] != 0;
    }
    if (h->cfg.variant == V_TAGGING) {
      for (int i = 0; i < A; ++i) s->used_tag[i] = f[k++] != 0;
      for (int i = 0; i < A; ++i) s->tag_counts[i] = f[k++];
      s->tag_timer = h->cfg.tag_reset_interval - (int32_t)f[k++];
    }
    memcpy(s->imposter, imposter_mask + (size_t)e * A, (size_t)A);
    s->t = t ? t[e] : 0;
    memset(s->metrics, 0, sizeof(s->metrics));
    s->metrics[M_STEPS] = s->t;
  }
  return 0;
}

/* ------------------------------------------------------------------ featurizers */
typedef struct {
  int64_t pos[ORC_MAX_AGENTS][2];
  int64_t alive[ORC_MAX_AGENTS];
  int64_t jobpos[ORC_MAX_JOBS][2];
  int64_t completed[ORC_MAX_JOBS];
  int64_t used[ORC_MAX_AGENTS];
  int64_t tags[ORC_MAX_AGENTS];
} flat_view;

/* unflatten: gymnasium.spaces.unflatten on the observation Tuple (base.py:237-241).  For the tagging env the
 * fields are read in TUPLE order (tagging.py:221-230), not through its inconsistent state_fields map
 * (tagging.py:15-28) -- documented deviation. */
static void view_from_flat(const orc_config *c, const int64_t *f, flat_view *v) {
  int A = c->n_imposters + c->n_crew, J = c->n_jobs, k = 0;
  memset(v, 0, sizeof(*v));
  for (int i = 0; i < A; ++i) { v->pos[i][0] = f[k++]; v->pos[i][1] = f[k++]; }
  for (int i = 0; i < A; ++i) v->alive[i] = f[k++];
  if (J > 0 || c->variant == V_TAGGING) {
    for (int j = 0; j < J; ++j) { v->jobpos[j][0] = f[k++]; v->jobpos[j][1] = f[k++]; }
    for (int j = 0; j < J; ++j) v->completed[j] = f[k++];
  }
  if (c->variant == V_TAGGING) {
    for (int i = 0; i < A; ++i) v->used[i] = f[k++];
    for (int i = 0; i < A; ++i) v->tags[i] = f[k++];
  }
}

static inline int in_grid(int64_t x, int64_t y) { return x >= 0 && y >= 0 && x < 9 && y < 9; }

/* component.py:83-131: planes [agent i][x][y] (alive only), then job planes [done?][x][y] */
static void spatial_planes(const orc_config *c, const flat_view *v, const int *chan_of_agent, float *out /*[A+2][81]*/) {
  int A = c->n_imposters + c->n_crew, J = c->n_jobs;
  memset(out, 0, sizeof(float) * (size_t)(A + 2) * 81);
  for (int i = 0; i < A; ++i)
    if (v->alive[i] && in_grid(v->pos[i][0], v->pos[i][1]))
      out[(size_t)chan_of_agent[i] * 81 + v->pos[i][0] * 9 + v->pos[i][1]] = 1.0f;
  for (int j = 0; j < J; ++j)
    if (in_grid(v->jobpos[j][0], v->jobpos[j][1]))
      out[(size_t)(A + (v->completed[j] ? 1 : 0)) * 81 + v->jobpos[j][0] * 9 + v->jobpos[j][1]] = 1.0f;
}

int orc_global_nonspatial_size(const orc_config *c) {
  int A = c->n_imposters + c->n_crew;
  return A + (c->variant == V_TAGGING ? A : 0) + c->n_jobs + A; /* model_ready.py:237-253 */
}
int orc_perspective_nonspatial_size(const orc_config *c) {
  int A = c->n_imposters + c->n_crew;
  return A + (c->variant == V_TAGGING ? A : 0) + c->n_jobs; /* model_ready.py:99-123 */
}

/* GlobalFeaturizer: spatial [n][A+2][9][9] shared by all views; non_spatial [A][n][F] (view k = one-hot k appended).
 * model_ready.py:227-306 */
int orc_encode_global(const orc_config *c, const int64_t *flat, int64_t n, float *spatial, float *nonspatial) {
  int A = c->n_imposters + c->n_crew, J = c->n_jobs, S = flat_size(c), F = orc_global_nonspatial_size(c);
  if (J == 0) return -1; /* the reference raises IndexError */
  int ident[ORC_MAX_AGENTS];
  for (int i = 0; i < A; ++i) ident[i] = i;
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < n; ++e) {
    flat_view v;
    view_from_flat(c, flat + e * S, &v);
    spatial_planes(c, &v, ident, spatial + e * (A + 2) * 81);
    for (int k = 0; k < A; ++k) {
      float *o = nonspatial + ((size_t)k * n + e) * F;
      int p = 0;
      for (int i = 0; i < A; ++i) o[p++] = (float)v.alive[i];
      if (c->variant == V_TAGGING) for (int i = 0; i < A; ++i) o[p++] = (float)v.tags[i];
      for (int j = 0; j < J; ++j) o[p++] = (float)v.completed[j];
      for (int i = 0; i < A; ++i) o[p++] = (i == k) ? 1.0f : 0.0f;
    }
  }
  return 0;
}

/* PerspectiveFeaturizer: spatial [A][n][A+2][9][9], view k has agent channels [k,0..k-1,k+1..A-1] then the job
 * planes; non_spatial [A][n][F] = per-agent fields permuted the same way (field-major) then job status.
 * model_ready.py:93-216 */
int orc_encode_perspective(const orc_config *c, const int64_t *flat, int64_t n, float *spatial, float *nonspatial) {
  int A = c->n_imposters + c->n_crew, J = c->n_jobs, S = flat_size(c), F = orc_perspective_nonspatial_size(c);
  if (J == 0) return -1;
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < n; ++e) {
    flat_view v;
    view_from_flat(c, flat + e * S, &v);
    for (int k = 0; k < A; ++k) {
      int order[ORC_MAX_AGENTS], chan_of_agent[ORC_MAX_AGENTS]; /* order[ch] = agent shown in channel ch */
      order[0] = k;
      for (int i = 0, ch = 1; i < A; ++i) if (i != k) order[ch++] = i;
      for (int ch = 0; ch < A; ++ch) chan_of_agent[order[ch]] = ch;
      spatial_planes(c, &v, chan_of_agent, spatial + ((size_t)k * n + e) * (A + 2) * 81);
      float *o = nonspatial + ((size_t)k * n + e) * F;
      int p = 0;
      for (int ch = 0; ch < A; ++ch) o[p++] = (float)v.alive[order[ch]];
      if (c->variant == V_TAGGING) for (int ch = 0; ch < A; ++ch) o[p++] = (float)v.tags[order[ch]];
      for (int j = 0; j < J; ++j) o[p++] = (float)v.completed[j];
    }
  }
  return 0;
}

enum {
  FC_ONEHOT_POS = 0, FC_COORDS, FC_ALIVE_CREW, FC_CLOSEST_CREW, FC_L1_CREW, FC_DIST_TO_IMPOSTER, FC_WALLS, FC_ROOMS,
  FC_SCENT, FC_STATE_ALIVE, FC_STATE_JOB_STATUS, FC_STATE_USED_TAGS, FC_STATE_TAG_COUNTS, FC_COUNT
};

static int comp_size(const orc_config *c, int comp) {
  int A = c->n_imposters + c->n_crew;
  switch (comp) {
    case FC_ONEHOT_POS: return A * 18;          /* component.py:243-247 */
    case FC_COORDS: return 2 * A;               /* component.py:401-403 */
    case FC_ALIVE_CREW: return A - 1;           /* component.py:423-425 */
    case FC_CLOSEST_CREW: return c->n_crew;     /* component.py:480-482 */
    case FC_L1_CREW: return c->n_crew;          /* component.py:450-452 */
    case FC_DIST_TO_IMPOSTER: return 2 * (A - 1); /* component.py:275-278 */
    case FC_WALLS: return 9;
    case FC_ROOMS: return 8;
    case FC_SCENT: return 4;
    case FC_STATE_ALIVE: return A;
    case FC_STATE_JOB_STATUS: return c->n_jobs;
    case FC_STATE_USED_TAGS: return c->variant == V_TAGGING ? A : -1;
    case FC_STATE_TAG_COUNTS: return c->variant == V_TAGGING ? A : -1;
  }
  return -1;
}

int orc_flat_feature_size(const orc_config *c, const int32_t *comps, int n_comps) {
  int F = 0;
  for (int i = 0; i < n_comps; ++i) {
    int s = comp_size(c, comps[i]);
    if (s < 0) return -1;
    if ((comps[i] == FC_CLOSEST_CREW || comps[i] == FC_L1_CREW) && c->n_imposters != 1) return -1;
    F += s;
  }
  return F;
}

static int room_of(int m, int64_t x, int64_t y) { /* component.py:8-17 */
  switch (m) {
    case 0: return x < 5 && y < 5;
    case 1: return x < 5 && y >= 5;
    case 2: return x >= 5 && y >= 5;
    default: return x >= 5 && y < 5;
  }
}

static int emit_component(const orc_config *c, const uint8_t grid[9][9], const flat_view *v, int comp, float *o) {
  int A = c->n_imposters + c->n_crew, J = c->n_jobs, p = 0;
  int64_t ix = v->pos[0][0], iy = v->pos[0][1]; /* "imposter is agent 0": component.py:262,289,355,440,467 */
  switch (comp) {
    case FC_ONEHOT_POS: /* component.py:226-240 */
      for (int i = 0; i < A; ++i)
        for (int q = 0; q < 18; ++q)
          o[p++] = (v->alive[i] && (q < 9 ? v->pos[i][0] == q : v->pos[i][1] == q - 9)) ? 1.0f : 0.0f;
      break;
    case FC_COORDS: /* component.py:389-399 */
      for (int i = 0; i < A; ++i) { o[p++] = (float)v->pos[i][0]; o[p++] = (float)v->pos[i][1]; }
      break;
    case FC_ALIVE_CREW: /* component.py:411-421 */
      for (int i = 1; i < A; ++i) o[p++] = v->alive[i] ? 1.0f : 0.0f;
      break;
    case FC_CLOSEST_CREW: { /* component.py:460-478 */
      float l1[ORC_MAX_AGENTS];
      int best = 0;
      for (int i = 1; i < A; ++i)
        l1[i - 1] = v->alive[i] ? (float)(llabs(ix - v->pos[i][0]) + llabs(iy - v->pos[i][1])) : 18.0f;
      for (int i = 1; i < A - 1; ++i) if (l1[i] < l1[best]) best = i;
      for (int i = 0; i < A - 1; ++i) o[p++] = (i == best) ? 1.0f : 0.0f;
    } break;
    case FC_L1_CREW: /* component.py:433-448 */
      for (int i = 1; i < A; ++i)
        o[p++] = v->alive[i] ? (float)(llabs(ix - v->pos[i][0]) + llabs(iy - v->pos[i][1])) : -1.0f;
      break;
    case FC_DIST_TO_IMPOSTER: { /* component.py:255-273: alive others compacted, trailing zeros */
      int n = 2 * (A - 1);
      for (int q = 0; q < n; ++q) o[q] = 0.0f;
      for (int i = 1; i < A; ++i)
        if (v->alive[i]) { o[p++] = (float)(ix - v->pos[i][0]); o[p++] = (float)(iy - v->pos[i][1]); }
      p = n;
    } break;
    case FC_WALLS: /* component.py:286-296: 3x3 patch of the zero-padded grid around agent 0 */
      for (int dx = -1; dx <= 1; ++dx)
        for (int dy = -1; dy <= 1; ++dy) {
          int64_t x = ix + dx, y = iy + dy;
          o[p++] = (in_grid(x, y) && grid[x][y]) ? 1.0f : 0.0f;
        }
      break;
    case FC_ROOMS: /* component.py:308-329 */
      for (int q = 0; q < 8; ++q) o[q] = 0.0f;
      for (int i = 0; i < A; ++i) {
        if (!v->alive[i]) continue;
        for (int m = 0; m < 4; ++m) o[(i == 0 ? 0 : 4) + m] += (float)room_of(m, v->pos[i][0], v->pos[i][1]);
      }
      p = 8;
      break;
    case FC_SCENT: { /* component.py:344-375: float32 accumulation of (9 - d) / 9 */
      float sc[4] = {0, 0, 0, 0};
      for (int i = 1; i < A; ++i) {
        if (!v->alive[i]) continue;
        double xs = (9.0 - (double)(v->pos[i][0] - ix)) / 9.0, ys = (9.0 - (double)(v->pos[i][1] - iy)) / 9.0;
        if (xs > 0) sc[0] += (float)xs; else sc[1] += (float)xs;
        if (ys > 0) sc[2] += (float)ys; else sc[3] += (float)ys;
      }
      for (int q = 0; q < 4; ++q) o[p++] = sc[q];
    } break;
    case FC_STATE_ALIVE: for (int i = 0; i < A; ++i) o[p++] = (float)v->alive[i]; break; /* component.py:210-214 */
    case FC_STATE_JOB_STATUS: for (int j = 0; j < J; ++j) o[p++] = (float)v->completed[j]; break;
    case FC_STATE_USED_TAGS: for (int i = 0; i < A; ++i) o[p++] = (float)v->used[i]; break;
    case FC_STATE_TAG_COUNTS: for (int i = 0; i < A; ++i) o[p++] = (float)v->tags[i]; break;
  }
  return p;
}

/* FlatFeaturizer over a CompositeFeaturizer of the listed components: out [n][F]. model_ready.py:309-367 */
int orc_encode_flat(const orc_config *c, const int32_t *comps, int n_comps, const int64_t *flat, int64_t n, float *out) {
  int F = orc_flat_feature_size(c, comps, n_comps), S = flat_size(c);
  if (F < 0) return -1;
  orc_handle g;
  memset(&g, 0, sizeof(g));
  g.cfg = *c;
  build_geometry(&g);
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < n; ++e) {
    flat_view v;
    view_from_flat(c, flat + e * S, &v);
    float *o = out + e * F;
    for (int q = 0; q < n_comps; ++q) o += emit_component(c, g.grid, &v, comps[q], o);
  }
  return 0;
}
