// Runs k_replay_push (one thread per float) and k_replay_push_v (vector streams + per-agent rows) of susnet_replay.cu on the
// host, one emulated thread at a time, so that the CPU suite can compare their results for arbitrary sizes, ring positions and
// buffer alignments.  KERNEL_SOURCE is the text of the file's anonymous namespace, cut out by the test.
#include <cuda_runtime.h>

#include "susnet_b200.h"
#include KERNEL_SOURCE

template <typename K>
static void run_grid(unsigned blocks, K&& kernel) {
  gridDim = {blocks, 1, 1};
  blockDim = {256, 1, 1};
  for (unsigned b = 0; b < blocks; ++b)
    for (unsigned t = 0; t < 256; ++t) {
      blockIdx = {b, 0, 0};
      threadIdx = {t, 0, 0};
      kernel();
    }
}

extern "C" void emu_replay_push_v1(const SusReplayPush* a) {
  const int64_t total = a->N * a->T * a->S;
  run_grid((unsigned)((total + 255) / 256), [&] { k_replay_push(*a); });
}

extern "C" void emu_replay_push_v2(const SusReplayPush* a) {
  const int64_t total = a->N * a->T * a->S;
  const unsigned vec_blocks = (unsigned)(((total + 3) / 4 + 255) / 256);
  const unsigned row_blocks = (unsigned)((a->N * a->A + 255) / 256);
  run_grid(vec_blocks + row_blocks, [&] {
    if (a->T == 1) k_replay_push_v<true>(*a, vec_blocks);
    else k_replay_push_v<false>(*a, vec_blocks);
  });
}
