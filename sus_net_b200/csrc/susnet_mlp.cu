// susnet_mlp.cu -- row (f2): Q-network INFERENCE of the acting loop for the reference's MLP estimator (src/models/dqn.py:72-108,
// make_mlp :316-324: Linear + PReLU stack on the flattened non-spatial features; train.py:367-370 evaluates it once per agent
// per step) for ALL envs in one launch.
//
// Why a kernel: at cfg5 (131 072 envs per GPU, [98, 256, 128, 64, 16, 6]) the torch forward -- five cuBLAS SGEMMs with K as
// small as 98 plus separate bias / PReLU passes over 134 MB activations -- took 0.88 ms of a 1.04 ms loop iteration (20 TFLOP/s
// of fp32 FFMA).  Here one persistent CTA pushes a tile of 128 rows through EVERY layer with the activations resident in
// shared memory (k-major, [K][128], two regions that alternate between layers), so the only HBM traffic is the 392-byte input
// row and the 24-byte Q row; weights are read in their torch layout ([out][in], the live parameter tensors: no copies) in
// chunks of 16 k through a double-buffered shared-memory stage; each thread owns an 8 x CT register micro-tile (CT = 8 for
// 128-wide column blocks), i.e. 4 LDS.128 per 64 FFMA in the inner loop.  fp32 FFMA only: no tensor cores (the north star
// excludes them), same arithmetic as the reference's CPU float32 forward up to summation order.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "../../include/susnet_b200.h"

extern "C" int sus_internal_fail(int code, const char* msg);
extern "C" void sus_internal_count_launch(void);

namespace {

// Two tile geometries (template parameter ROWS = rows per CTA tile; 2 * ROWS threads = ROWS / 8 row groups x 16 column groups):
//   128 rows, 256 threads, ONE CTA per SM (213 KB of shared memory at cfg5);
//    64 rows, 128 threads, TWO CTAs per SM (2 x 112.5 KB): the same eight warps per SM, but while one CTA stages its input tile,
//    waits at a chunk barrier or runs the narrow tail layers the other one keeps the FFMA pipe busy (SUSNET_MLP_ROWS=64).
constexpr int kKc = 16;  // k per weight chunk

struct MlpParams {
  SusMlpSpec s;
  const float* x;
  float* out;
  int64_t n_rows;
  int32_t region_a_floats, region_b_floats;  // activations regions (per row-tile): dims at odd / even positions
};

// One layer on the CTA's row tile: out[m][r] = act(bias[m] + sum_k W[m][k] * in[k][r]) for m < M, or straight to global memory
// for the last layer.  CT = columns per thread (column block = 16 * CT); M is processed in blocks of 16 * CT columns.
template <int ROWS, int CT>
__device__ __forceinline__ void layer(const float* __restrict__ W, const float* __restrict__ bias, const float* __restrict__ alpha,
                                      int act, int K, int M, int in_off, int out_off,
                                      float* __restrict__ gout, int64_t row0, int64_t n_rows, int out_stride, int wst_off) {
  // offsets into the dynamic shared array, NOT generic pointers: with pointers picked from a runtime-indexed table the compiler
  // emitted generic LD.E.128 for the operand loads of the inner loop instead of LDS.128
  extern __shared__ __align__(128) float smem[];
  const float* in = smem + in_off;
  float* outs = smem + out_off;
  float* wst = smem + wst_off;
  const int tid = threadIdx.x;
  constexpr int kRows = ROWS, kThreads = 2 * ROWS, RG = ROWS / 8, HALF = ROWS / 2;
  const int rg = tid % RG, cg = tid / RG;  // row group, column group (CT columns)
  // a thread's 8 rows are r0 .. r0+3 and HALF+r0 .. HALF+r0+3 with r0 = 4 * rg: the row-group threads of a (half-)warp read
  // contiguous bytes per LDS.128 (rows 8 * rg .. would put four threads on every bank: measured 22 TFLOP/s)
  const int r0 = rg * 4;
  constexpr int CB = 16 * CT;  // columns per block
  const float a = (act == SUS_ACT_PRELU && alpha) ? alpha[0] : 0.0f;
  for (int m0 = 0; m0 < M; m0 += CB) {
    float acc[8][CT];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < CT; ++j) acc[i][j] = 0.0f;
    const int n_chunks = (K + kKc - 1) / kKc;
    constexpr int CBP = CB + 4;                  // padded row of a staged chunk: 2-way instead of 16-way bank conflicts
    constexpr int PER = (CB * kKc + kThreads - 1) / kThreads;  // staged weights per thread and chunk (8 / 4 / 1)
    float pre[PER];
    // chunk c of the block: w[kk][col] = W[m0 + col][c * kKc + kk] (zero outside M / K).  Consecutive threads walk k, so the
    // global reads are contiguous runs of a weight row; loaded into registers one chunk ahead, stored after the math.
    auto fetch = [&](int c) {
#pragma unroll
      for (int q = 0; q < PER; ++q) {
        const int idx = tid + q * kThreads;
        const int col = idx / kKc, kk = idx - col * kKc;
        const int m = m0 + col, k = c * kKc + kk;
        pre[q] = (idx < CB * kKc && m < M && k < K) ? W[(int64_t)m * K + k] : 0.0f;
      }
    };
    auto put = [&](int buf) {
      float* dst = wst + buf * (kKc * CBP);
#pragma unroll
      for (int q = 0; q < PER; ++q) {
        const int idx = tid + q * kThreads;
        const int col = idx / kKc, kk = idx - col * kKc;
        if (idx < CB * kKc) dst[kk * CBP + col] = pre[q];
      }
    };
    fetch(0);
    put(0);
    __syncthreads();
    for (int c = 0; c < n_chunks; ++c) {
      const int buf = c & 1;
      if (c + 1 < n_chunks) fetch(c + 1);
      const float* w = wst + buf * (kKc * CBP) + cg * CT;
      const float* xin = in + (int64_t)(c * kKc) * kRows + r0;
      const int kmax = K - c * kKc < kKc ? K - c * kKc : kKc;
      auto kstep = [&](int kk) {
        const float4 x0 = *reinterpret_cast<const float4*>(xin + kk * kRows);
        const float4 x1 = *reinterpret_cast<const float4*>(xin + kk * kRows + HALF);
        const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
        float wv[CT];
        if (CT >= 4) {
#pragma unroll
          for (int j = 0; j < CT; j += 4) {
            const float4 t = *reinterpret_cast<const float4*>(w + kk * CBP + j);
            wv[j] = t.x; wv[j + 1] = t.y; wv[j + 2] = t.z; wv[j + 3] = t.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < CT; ++j) wv[j] = w[kk * CBP + j];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < CT; ++j) acc[i][j] = fmaf(xv[i], wv[j], acc[i][j]);
      };
      if (kmax == kKc) {  // full chunk: straight-line code, the loads of the next k-steps overlap the FFMAs
#pragma unroll
        for (int kk = 0; kk < kKc; ++kk) kstep(kk);
      } else {
        for (int kk = 0; kk < kmax; ++kk) kstep(kk);
      }
      if (c + 1 < n_chunks) put(buf ^ 1);  // (the other buffer was last read before the barrier that ended chunk c - 1)
      __syncthreads();
    }
    // epilogue: bias + activation; to the next layer's k-major region, or (last layer) to global memory [row][out_stride]
#pragma unroll
    for (int j = 0; j < CT; ++j) {
      const int m = m0 + cg * CT + j;
      if (m >= M) continue;
      const float b = bias ? bias[m] : 0.0f;
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float t = acc[i][j] + b;
        if (act == SUS_ACT_RELU) t = t > 0.0f ? t : 0.0f;
        else if (act == SUS_ACT_PRELU) t = t > 0.0f ? t : a * t;
        v[i] = t;
      }
      if (gout) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int64_t row = row0 + r0 + (i & 3) + (i >> 2) * HALF;
          if (row < n_rows) gout[row * out_stride + m] = v[i];
        }
      } else {
        float* o = outs + (int64_t)m * kRows + r0;
        *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(o + HALF) = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
  }
  __syncthreads();
}

template <int ROWS>
__global__ void __launch_bounds__(2 * ROWS, ROWS == 64 ? 2 : 1) k_mlp_forward(const __grid_constant__ MlpParams p) {
  extern __shared__ __align__(128) float smem[];
  constexpr int kRows = ROWS, kThreads = 2 * ROWS;
  const int region_off[2] = {0, p.region_b_floats * kRows};  // [0]: even positions (input, h2, ...), [1]: odd
  const int wst_off = region_off[1] + p.region_a_floats * kRows;  // 2 x kKc x (128 + 4) floats
  float* region[2] = {smem + region_off[0], smem + region_off[1]};
  const SusMlpSpec& s = p.s;
  const int K0 = s.dims[0];
  const int64_t n_tiles = (p.n_rows + kRows - 1) / kRows;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * kRows;
    // input rows -> k-major region 0.  The tile's rows are one contiguous block of global memory: copy it with coalesced
    // 128-bit loads into region 1 (free until layer 0 writes its output there), then transpose shared -> shared (lane = row:
    // conflict-free stores, 2-way conflicts on the loads).  Strided 4-byte global loads took 15 % of the kernel, more when
    // the features live in L2-compressible memory.
    {
      const int64_t n_valid = p.n_rows - row0 < kRows ? p.n_rows - row0 : kRows;
      const int64_t n_floats = n_valid * K0;
      const float* src = p.x + row0 * K0;
      float* raw = region[1];
      if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
        const int64_t n4 = n_floats >> 2;
        for (int64_t i = threadIdx.x; i < n4; i += kThreads) reinterpret_cast<float4*>(raw)[i] = reinterpret_cast<const float4*>(src)[i];
        for (int64_t i = (n4 << 2) + threadIdx.x; i < n_floats; i += kThreads) raw[i] = src[i];
      } else {
        for (int64_t i = threadIdx.x; i < n_floats; i += kThreads) raw[i] = src[i];
      }
      __syncthreads();
      const int r = threadIdx.x % kRows, half = threadIdx.x / kRows;
      const bool ok = r < n_valid;
      for (int k = half; k < K0; k += 2) region[0][(int64_t)k * kRows + r] = ok ? raw[(int64_t)r * K0 + k] : 0.0f;
    }
    __syncthreads();
    for (int l = 0; l < s.n_layers; ++l) {
      const int K = s.dims[l], M = s.dims[l + 1];
      const bool last = l == s.n_layers - 1;
      const int in = (l & 1) ? region_off[1] : region_off[0];
      const int outs = (l & 1) ? region_off[0] : region_off[1];
      const int act = last ? SUS_ACT_NONE : s.activation;
      float* gout = last ? p.out : nullptr;
      if (M > 64) layer<ROWS, 8>(s.weight[l], s.bias[l], s.alpha[l], act, K, M, in, outs, gout, row0, p.n_rows, M, wst_off);
      else if (M > 16) layer<ROWS, 4>(s.weight[l], s.bias[l], s.alpha[l], act, K, M, in, outs, gout, row0, p.n_rows, M, wst_off);
      else layer<ROWS, 1>(s.weight[l], s.bias[l], s.alpha[l], act, K, M, in, outs, gout, row0, p.n_rows, M, wst_off);
    }
  }
}

}  // namespace

extern "C" int sus_mlp_forward(const SusMlpSpec* spec, const float* x, int64_t n_rows, float* out, int device, void* stream) {
  if (!spec || (n_rows > 0 && (!x || !out))) return sus_internal_fail(SUS_ERR_INVALID_ARGUMENT, "mlp_forward: NULL argument");
  if (spec->n_layers < 1 || spec->n_layers > SUS_MLP_MAX_LAYERS) return sus_internal_fail(SUS_ERR_INVALID_ARGUMENT, "mlp_forward: 1..8 layers");
  if (spec->activation < SUS_ACT_NONE || spec->activation > SUS_ACT_PRELU)
    return sus_internal_fail(SUS_ERR_INVALID_ARGUMENT, "mlp_forward: unknown activation");
  int a = 0, b = 0;  // widest activation at odd / even positions of the chain (the last layer's output goes to global memory)
  for (int l = 0; l <= spec->n_layers; ++l) {
    if (spec->dims[l] < 1) return sus_internal_fail(SUS_ERR_INVALID_ARGUMENT, "mlp_forward: layer width < 1");
    if (l < spec->n_layers && !spec->weight[l]) return sus_internal_fail(SUS_ERR_INVALID_ARGUMENT, "mlp_forward: NULL weight");
    if (l == spec->n_layers) break;
    if (l & 1) { if (spec->dims[l] > a) a = spec->dims[l]; } else { if (spec->dims[l] > b) b = spec->dims[l]; }
  }
  if (a < spec->dims[0]) a = spec->dims[0];  // region 1 also stages the raw (row-major) input tile before layer 0 runs
  int prev = -1;
  cudaGetDevice(&prev);
  if (prev != device) cudaSetDevice(device);
  int max_smem = 0, sms = 0;
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  auto smem_of = [&](int rows) { return ((size_t)(a + b) * rows + 2 * kKc * 132) * sizeof(float); };
  const char* rows_env = getenv("SUSNET_MLP_ROWS");  // read per call: tests and tools switch geometries inside one process
  // 128-row tiles unless asked for 64-row tiles, or unless only a 64-row tile of the two widest adjacent layers fits
  const int rows = ((rows_env && atoi(rows_env) == 64) || smem_of(128) > (size_t)max_smem) ? 64 : 128;
  const size_t smem = smem_of(rows);
  int rc = SUS_OK;
  if (smem > (size_t)max_smem) {
    rc = sus_internal_fail(SUS_ERR_UNSUPPORTED, "mlp_forward: the two widest adjacent layers do not fit in shared memory (run the module itself)");
  } else if (n_rows > 0) {
    static size_t granted[2][64] = {};
    const int v = rows == 64;
    if (device >= 0 && device < 64 && granted[v][device] < smem) {
      if (v) {
        cudaFuncSetAttribute(k_mlp_forward<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_mlp_forward<64>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      } else {
        cudaFuncSetAttribute(k_mlp_forward<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      }
      granted[v][device] = smem;
    }
    MlpParams p;
    p.s = *spec; p.x = x; p.out = out; p.n_rows = n_rows; p.region_a_floats = a; p.region_b_floats = b;
    const int64_t tiles = (n_rows + rows - 1) / rows;
    if (v) {
      // CTAs per SM: 2 when two tiles' activations fit on an SM (cfg5: 2 x 112.5 KB); queried once per (device, size)
      static size_t occ_smem[64] = {};
      static int occ_ctas[64] = {};
      int per_sm = 1;
      if (device >= 0 && device < 64 && occ_smem[device] == smem) {
        per_sm = occ_ctas[device];
      } else {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_mlp_forward<64>, 128, smem);
        if (per_sm < 1) per_sm = 1;
        if (device >= 0 && device < 64) { occ_smem[device] = smem; occ_ctas[device] = per_sm; }
      }
      static const bool verbose = getenv("SUSNET_MLP_VERBOSE") != nullptr;
      if (verbose) fprintf(stderr, "sus_mlp_forward: 64-row tiles, %zu bytes of shared memory, %d CTAs per SM\n", smem, per_sm);
      const int64_t grid = (int64_t)sms * per_sm;
      k_mlp_forward<64><<<(unsigned)(tiles < grid ? tiles : grid), 128, smem, (cudaStream_t)stream>>>(p);
    } else {
      k_mlp_forward<128><<<(unsigned)(tiles < sms ? tiles : sms), 256, smem, (cudaStream_t)stream>>>(p);
    }
    sus_internal_count_launch();
    const cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) rc = sus_internal_fail(SUS_ERR_CUDA, cudaGetErrorString(err));
  }
  if (prev != device && prev >= 0) cudaSetDevice(prev);
  return rc;
}
