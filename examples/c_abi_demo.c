/*
 * c_abi_demo.c -- the library used from plain C through include/susnet_b200.h only (no Python, no torch):
 * create 65 536 FourRoomEnv(1 imposter, 4 crew, 5 jobs) envs, reset, run 200 fused step + Global-encode launches
 * with the random policy, read back the episode statistics and a checksum of the feature tensors.
 *
 *   gcc -O2 -Iinclude examples/c_abi_demo.c -o c_abi_demo -Lsus_net_b200 -lsusnet_b200 \
 *       -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/sus_net_b200 && ./c_abi_demo
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <cuda_runtime_api.h>

#include "susnet_b200.h"

#define CHECK(x) do { int rc_ = (x); if (rc_ < 0) { fprintf(stderr, "%s failed: %s\n", #x, sus_last_error()); return 1; } } while (0)
#define CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

int main(int argc, char **argv) {
  const int64_t N = argc > 1 ? atoll(argv[1]) : 65536;
  const int steps = argc > 2 ? atoi(argv[2]) : 200;
  SusConfig cfg = {0};
  cfg.variant = SUS_VARIANT_BASE; cfg.n_imposters = 1; cfg.n_crew = 4; cfg.n_jobs = 5; cfg.include_walls = 1;
  cfg.is_action_order_random = 1; cfg.shuffle_imposter_index = 1; cfg.max_time_steps = 1000; cfg.tag_reset_interval = 50;
  cfg.auto_reset = 1; cfg.kill_reward = -5; cfg.complete_job_reward = 3; cfg.sabotage_reward = 3; cfg.time_step_reward = 0;
  cfg.game_end_reward = 10; cfg.dead_penalty = -2; cfg.vote_reward = 3; cfg.num_envs = N; cfg.seed = 2026; cfg.env_id_base = 0;

  SusEncodeSpec spec = {0};
  spec.kind = SUS_ENCODE_GLOBAL;
  SusEncodeShape shape;
  CHECK(sus_encode_shape(&cfg, &spec, &shape));
  const int A = cfg.n_imposters + cfg.n_crew;
  printf("ABI %d, S = %d, spatial %d floats x %d view(s), non-spatial %d floats x %d views\n", sus_abi_version(),
         sus_flat_state_size(&cfg), shape.spatial_floats, shape.spatial_views, shape.non_spatial_floats, shape.non_spatial_views);

  sus_env_t env;
  CHECK(sus_env_create(&cfg, 0, &env));
  cudaStream_t stream;
  CUDA(cudaStreamCreate(&stream));
  float *rewards, *spatial, *non_spatial;
  uint8_t *done, *trunc;
  int64_t *stats;
  CUDA(cudaMalloc((void **)&rewards, (size_t)N * A * sizeof(float)));
  CUDA(cudaMalloc((void **)&done, (size_t)N));
  CUDA(cudaMalloc((void **)&trunc, (size_t)N));
  /* the planes are almost all zeros: put them in L2-compressible memory where the device has it */
  const uint64_t sp_bytes = (uint64_t)shape.spatial_views * N * shape.spatial_floats * sizeof(float);
  int compressible = sus_alloc_compressible(0, sp_bytes, (void **)&spatial, NULL) == SUS_OK;
  if (!compressible) CUDA(cudaMalloc((void **)&spatial, sp_bytes));
  printf("plane tensor in %s memory\n", compressible ? "L2-compressible" : "cudaMalloc");
  CUDA(cudaMalloc((void **)&non_spatial, (size_t)shape.non_spatial_views * N * shape.non_spatial_floats * sizeof(float)));
  CUDA(cudaMalloc((void **)&stats, SUS_N_STATS * sizeof(int64_t)));

  CHECK(sus_env_reset(env, NULL, stream));
  SusStepIO io = {0};
  io.actions = NULL; /* fused random policy */
  io.rewards = rewards; io.rewards_dtype = SUS_F32; io.done = done; io.truncated = trunc;
  io.encode = &spec; io.spatial = spatial; io.non_spatial = non_spatial;
  cudaEvent_t t0, t1;
  CUDA(cudaEventCreate(&t0)); CUDA(cudaEventCreate(&t1));
  CUDA(cudaEventRecord(t0, stream));
  for (int s = 0; s < steps; ++s) CHECK(sus_env_step(env, &io, stream));
  CUDA(cudaEventRecord(t1, stream));
  CHECK(sus_env_stats(env, stats, stream));
  CUDA(cudaStreamSynchronize(stream));
  float ms;
  CUDA(cudaEventElapsedTime(&ms, t0, t1));

  int64_t h_stats[SUS_N_STATS];
  CUDA(cudaMemcpy(h_stats, stats, sizeof(h_stats), cudaMemcpyDeviceToHost));
  const size_t n_sp = (size_t)N * shape.spatial_floats;
  float *h_sp = (float *)malloc(n_sp * sizeof(float));
  CUDA(cudaMemcpy(h_sp, spatial, n_sp * sizeof(float), cudaMemcpyDeviceToHost));
  double ones = 0;
  for (size_t i = 0; i < n_sp; ++i) ones += h_sp[i];
  printf("%d steps x %lld envs in %.3f ms = %.3e env-steps/s; %lld kernel launches\n", steps, (long long)N, ms,
         (double)N * steps / (ms * 1e-3), (long long)sus_launch_count());
  printf("episodes %lld, crew won %lld, imposters won %lld, kills %lld, total steps of finished episodes %lld\n",
         (long long)h_stats[SUS_S_EPISODES], (long long)h_stats[SUS_S_CREW_WON], (long long)h_stats[SUS_S_IMPOSTER_WON],
         (long long)h_stats[SUS_S_IMP_KILLED_CREW], (long long)h_stats[SUS_S_TOTAL_TIME_STEPS]);
  /* every env shows <= A alive-agent cells and exactly 5 job cells in its planes */
  printf("ones in the spatial planes: %.0f (per env %.3f, must be in [5, %d])\n", ones, ones / (double)N, A + 5);
  const int ok = h_stats[SUS_S_EPISODES] == h_stats[SUS_S_CREW_WON] + h_stats[SUS_S_IMPOSTER_WON] + h_stats[SUS_S_TRUNCATED] &&
                 ones / (double)N >= 5.0 && ones / (double)N <= A + 5.0;
  free(h_sp);
  if (compressible) CHECK(sus_free_compressible(spatial));
  CHECK(sus_env_destroy(env));
  printf(ok ? "OK\n" : "MISMATCH\n");
  return ok ? 0 : 2;
}
