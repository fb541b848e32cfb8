// Runs k_mlp_pack of susnet_mlp.cu on the host, one emulated thread at a time (tests/test_host_kernel_emulation.py): the packed
// weight image must be what the forward kernel's 128-bit chunk copies expect.  KERNEL_SOURCE is the text of the pieces the packer
// needs (constants, MlpParams, the layout helpers, the kernel), cut out of the file by the test.
#include <cuda_runtime.h>

#include "susnet_b200.h"
namespace {
#include KERNEL_SOURCE
}

extern "C" void emu_mlp_pack(const SusMlpSpec* spec, float* packed) {
  MlpParams p{};
  p.s = *spec;
  p.packed = packed;
  int64_t off = 0;
  for (int l = 0; l < SUS_MLP_MAX_LAYERS; ++l) {
    p.packed_off[l] = off;
    if (l < spec->n_layers) off += mlp_packed_floats(spec->dims[l], spec->dims[l + 1]);
  }
  const unsigned blocks = (unsigned)((off + 255) / 256);
  gridDim = {blocks, 1, 1};
  blockDim = {256, 1, 1};
  for (unsigned b = 0; b < blocks; ++b)
    for (unsigned t = 0; t < 256; ++t) {
      blockIdx = {b, 0, 0};
      threadIdx = {t, 0, 0};
      k_mlp_pack(p, off);
    }
}
