// k_select_actions and k_seq_roll of susnet_policy.cu run on the host, one emulated thread at a time
// (tests/test_host_kernel_emulation.py).  KERNEL_SOURCE: the body of the file's anonymous namespace plus the host functions of
// susnet_api.cu that build a DevConfig from a SusConfig.
#include <cuda_runtime.h>

#include <string>

#include "susnet_device.cuh"

using namespace susnet;

namespace {
int fail(int code, const std::string&) { return code; }
#include KERNEL_SOURCE

template <typename K>
void run_grid(int64_t threads, K&& kernel) {
  const unsigned blocks = (unsigned)((threads + 255) / 256);
  gridDim = {blocks, 1, 1};
  blockDim = {256, 1, 1};
  for (unsigned b = 0; b < blocks; ++b)
    for (unsigned t = 0; t < 256; ++t) {
      blockIdx = {b, 0, 0};
      threadIdx = {t, 0, 0};
      kernel();
    }
}
}  // namespace

extern "C" int emu_select_actions(const SusConfig* cfg, uint64_t act_epoch, uint4* aux, const float* q_imp, const float* q_crew,
                                  const float* eps_dev, float eps_value, int imp_per_view, int actions_dtype, void* actions) {
  PolicyParams p;
  std::memset(&p, 0, sizeof(p));
  make_dev_config(*cfg, p.c);
  p.st.aux = aux;
  p.q_imp = q_imp; p.q_crew = q_crew; p.eps_dev = eps_dev; p.eps_value = eps_value;
  p.imp_per_view = imp_per_view; p.actions_dtype = actions_dtype; p.actions = actions;
  p.tick = act_epoch; p.N = cfg->num_envs;
  run_grid(p.N, [&] { k_select_actions(p); });
  return 0;
}

extern "C" int emu_seq_roll(const float* in, float* out, const float* newest, const uint8_t* done, const uint8_t* trunc, int64_t rows,
                            int64_t n_envs, int32_t T, int32_t R, int vec) {
  RollParams p = {in, out, newest, done, trunc, rows, n_envs, T, R};
  if (vec) run_grid(rows * T * (R / 4), [&] { k_seq_roll<true>(p); });
  else run_grid(rows * T * R, [&] { k_seq_roll<false>(p); });
  return 0;
}
