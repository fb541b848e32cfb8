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
constexpr int kKc = 16;        // k per weight chunk (over all k-parts)
constexpr int kDefaultSplit = 1;  // k-parts per CTA unless SUSNET_MLP_SPLIT says otherwise

struct MlpParams {
  SusMlpSpec s;
  const float* x;
  float* out;
  int64_t n_rows;
  int32_t region_a_floats, region_b_floats;  // activations regions (per row-tile): dims at odd / even positions
  const float* packed;                        // repacked weights (k_mlp_pack) or NULL: stage from the torch layout
  int64_t packed_off[SUS_MLP_MAX_LAYERS];     // float offset of every layer inside `packed`
};

// columns per thread of a layer with M outputs (column block = 16 * CT): the same rule in the kernel, the packer and the host
__host__ __device__ inline int mlp_ct(int M) { return M > 64 ? 8 : (M > 16 ? 4 : 1); }
// floats of one layer in the packed image: [column block][chunk of kKc k's][kk][column], zero outside M / K
__host__ __device__ inline int64_t mlp_packed_floats(int K, int M) {
  const int CB = 16 * mlp_ct(M);
  return (int64_t)((M + CB - 1) / CB) * CB * ((K + 16 - 1) / 16) * 16;
}

// One layer on the CTA's row tile: out[m][r] = act(bias[m] + sum_k W[m][k] * in[k][r]) for m < M, or straight to global memory
// for the last layer.  CT = columns per thread (column block = 16 * CT); M is processed in blocks of 16 * CT columns.
// SPLIT = 2: the CTA has two PARTS of 2 * ROWS threads; both own the same (rows, columns) micro-tiles, each sums half of the k
// range, part 1 leaves its partial sums in the layer's output region and part 0 adds them in its epilogue.  Twice the warps per
// scheduler for the same registers per thread and the same operand loads per FFMA.
// PACKED: the layer's weights come from the repacked image (k_mlp_pack): a chunk is CB * kKc contiguous floats that are copied
// with 128-bit loads and stores -- 4 instructions per 4 weights instead of ~12 per weight for the transposing gather from the
// torch layout (address arithmetic, bounds tests), which was 10 % of the kernel's issue slots (ncu source counters).
template <int ROWS, int SPLIT, int CT, bool PACKED>
__device__ __forceinline__ void layer(const float* __restrict__ W, const float* __restrict__ Wp, const float* __restrict__ bias, const float* __restrict__ alpha,
                                      int act, int K, int M, int in_off, int out_off,
                                      float* __restrict__ gout, int64_t row0, int64_t n_rows, int out_stride, int wst_off) {
  // offsets into the dynamic shared array, NOT generic pointers: with pointers picked from a runtime-indexed table the compiler
  // emitted generic LD.E.128 for the operand loads of the inner loop instead of LDS.128
  extern __shared__ __align__(128) float smem[];
  const float* in = smem + in_off;
  float* outs = smem + out_off;
  constexpr int kRows = ROWS, PT = 2 * ROWS, RG = ROWS / 8, HALF = ROWS / 2;  // PT: threads per part
  constexpr int KC = kKc / SPLIT;                                              // k per chunk and part
  const int part = SPLIT > 1 ? (int)threadIdx.x / PT : 0;
  const int tid = SPLIT > 1 ? (int)threadIdx.x % PT : (int)threadIdx.x;        // thread within its part
  const int rg = tid % RG, cg = tid / RG;  // row group, column group (CT columns)
  // a thread's 8 rows are r0 .. r0+3 and HALF+r0 .. HALF+r0+3 with r0 = 4 * rg: the row-group threads of a (half-)warp read
  // contiguous bytes per LDS.128 (rows 8 * rg .. would put four threads on every bank: measured 22 TFLOP/s)
  const int r0 = rg * 4;
  constexpr int CB = 16 * CT;  // columns per block
  constexpr int CBP = CB + 4;  // padded row of a staged chunk: conflict-free transposing stores
  float* wst = smem + wst_off + part * (2 * KC * CBP);
  const float a = (act == SUS_ACT_PRELU && alpha) ? alpha[0] : 0.0f;
  // this part's k range; every part runs the same number of chunks (and barriers), trailing ones may be short or empty
  const int Kp = (K + SPLIT - 1) / SPLIT;
  const int k_lo = part * Kp;
  const int k_hi = k_lo + Kp < K ? k_lo + Kp : K;
  const int n_chunks = (Kp + KC - 1) / KC;
  for (int m0 = 0; m0 < M; m0 += CB) {
    float acc[8][CT];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < CT; ++j) acc[i][j] = 0.0f;
    constexpr int PER = PACKED ? (CB * KC / 4 + PT - 1) / PT : (CB * KC + PT - 1) / PT;  // staged float4s / floats per thread and chunk
    float4 pre4[PACKED ? PER : 1];
    float pre[PACKED ? 1 : PER];
    // chunk c of the block: w[kk][col] = W[m0 + col][k_lo + c * KC + kk] (zero outside M / the part's k range).
    // Unpacked: consecutive threads walk k, so the global reads are contiguous runs of a weight row.  Either way the chunk is
    // loaded into registers one chunk ahead and stored after the math.
    const float4* wp4 = PACKED ? reinterpret_cast<const float4*>(Wp + (int64_t)(m0 / CB) * ((K + kKc - 1) / kKc) * (kKc * CB)) : nullptr;
    auto fetch = [&](int c) {
      if constexpr (PACKED) {
#pragma unroll
        for (int q = 0; q < PER; ++q) {
          const int f = tid + q * PT;
          if (f < CB * KC / 4) pre4[q] = wp4[c * (CB * KC / 4) + f];
        }
      } else {
#pragma unroll
        for (int q = 0; q < PER; ++q) {
          const int idx = tid + q * PT;
          const int col = idx / KC, kk = idx - col * KC;
          const int m = m0 + col, k = k_lo + c * KC + kk;
          pre[q] = (idx < CB * KC && m < M && k < k_hi) ? W[(int64_t)m * K + k] : 0.0f;
        }
      }
    };
    auto put = [&](int buf) {
      float* dst = wst + buf * (KC * CBP);
      if constexpr (PACKED) {
#pragma unroll
        for (int q = 0; q < PER; ++q) {
          const int f = tid + q * PT;
          const int kk = f / (CB / 4), c4 = f - kk * (CB / 4);
          if (f < CB * KC / 4) *reinterpret_cast<float4*>(dst + kk * CBP + c4 * 4) = pre4[q];
        }
      } else {
#pragma unroll
        for (int q = 0; q < PER; ++q) {
          const int idx = tid + q * PT;
          const int col = idx / KC, kk = idx - col * KC;
          if (idx < CB * KC) dst[kk * CBP + col] = pre[q];
        }
      }
    };
    fetch(0);
    put(0);
    __syncthreads();
    for (int c = 0; c < n_chunks; ++c) {
      const int buf = c & 1;
      if (c + 1 < n_chunks) fetch(c + 1);
      const float* w = wst + buf * (KC * CBP) + cg * CT;
      const float* xin = in + (int64_t)(k_lo + c * KC) * kRows + r0;
      const int left = k_hi - (k_lo + c * KC);
      const int kmax = left < KC ? (left < 0 ? 0 : left) : KC;
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
      if (kmax == KC) {  // full chunk: straight-line code, the loads of the next k-steps overlap the FFMAs
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) kstep(kk);
      } else {
        for (int kk = 0; kk < kmax; ++kk) kstep(kk);
      }
      if (c + 1 < n_chunks) put(buf ^ 1);  // (the other buffer was last read before the barrier that ended chunk c - 1)
      __syncthreads();
    }
    if (SPLIT > 1) {  // part 1's partial sums travel through the (still unused) output positions of this column block
      if (part == 1) {
#pragma unroll
        for (int j = 0; j < CT; ++j) {
          const int m = m0 + cg * CT + j;
          if (m >= M) continue;
          float* o = outs + (int64_t)m * kRows + r0;
          *reinterpret_cast<float4*>(o) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
          *reinterpret_cast<float4*>(o + HALF) = make_float4(acc[4][j], acc[5][j], acc[6][j], acc[7][j]);
        }
      }
      __syncthreads();
      if (part == 0) {
#pragma unroll
        for (int j = 0; j < CT; ++j) {
          const int m = m0 + cg * CT + j;
          if (m >= M) continue;
          const float* o = outs + (int64_t)m * kRows + r0;
          const float4 p0 = *reinterpret_cast<const float4*>(o);
          const float4 p1 = *reinterpret_cast<const float4*>(o + HALF);
          acc[0][j] += p0.x; acc[1][j] += p0.y; acc[2][j] += p0.z; acc[3][j] += p0.w;
          acc[4][j] += p1.x; acc[5][j] += p1.y; acc[6][j] += p1.z; acc[7][j] += p1.w;
        }
      }
    }
    // epilogue: bias + activation; to the next layer's k-major region, or (last layer) to global memory [row][out_stride]
    if (part == 0) {
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
  }
  __syncthreads();
}

template <int ROWS, int SPLIT, bool PACKED>
__global__ void __launch_bounds__(2 * ROWS * SPLIT, ROWS == 64 ? 2 : 1) k_mlp_forward(const __grid_constant__ MlpParams p) {
  extern __shared__ __align__(128) float smem[];
  constexpr int kRows = ROWS, kThreads = 2 * ROWS * SPLIT;
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
      const int r = threadIdx.x % kRows, first = threadIdx.x / kRows;
      const bool ok = r < n_valid;
      for (int k = first; k < K0; k += kThreads / kRows) region[0][(int64_t)k * kRows + r] = ok ? raw[(int64_t)r * K0 + k] : 0.0f;
    }
    __syncthreads();
    for (int l = 0; l < s.n_layers; ++l) {
      const int K = s.dims[l], M = s.dims[l + 1];
      const bool last = l == s.n_layers - 1;
      const int in = (l & 1) ? region_off[1] : region_off[0];
      const int outs = (l & 1) ? region_off[0] : region_off[1];
      const int act = last ? SUS_ACT_NONE : s.activation;
      float* gout = last ? p.out : nullptr;
      const float* wp = PACKED ? p.packed + p.packed_off[l] : nullptr;
      if (M > 64) layer<ROWS, SPLIT, 8, PACKED>(s.weight[l], wp, s.bias[l], s.alpha[l], act, K, M, in, outs, gout, row0, p.n_rows, M, wst_off);
      else if (M > 16) layer<ROWS, SPLIT, 4, PACKED>(s.weight[l], wp, s.bias[l], s.alpha[l], act, K, M, in, outs, gout, row0, p.n_rows, M, wst_off);
      else layer<ROWS, SPLIT, 1, PACKED>(s.weight[l], wp, s.bias[l], s.alpha[l], act, K, M, in, outs, gout, row0, p.n_rows, M, wst_off);
    }
  }
}

// Weights of every layer -> the packed image (see mlp_packed_floats): one thread per packed float.  The live parameters change
// with every optimizer step, so this runs in front of every forward that was given a workspace (67 k floats at cfg5: ~3 us).
__global__ void __launch_bounds__(256) k_mlp_pack(const __grid_constant__ MlpParams p, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  int l = 0;
  while (l + 1 < p.s.n_layers && i >= p.packed_off[l + 1]) ++l;
  const int K = p.s.dims[l], M = p.s.dims[l + 1];
  const int CB = 16 * mlp_ct(M);
  const int n_chunks = (K + kKc - 1) / kKc;
  int64_t r = i - p.packed_off[l];
  const int col = (int)(r % CB); r /= CB;
  const int kk = (int)(r % kKc); r /= kKc;
  const int c = (int)(r % n_chunks);
  const int mb = (int)(r / n_chunks);
  const int m = mb * CB + col, k = c * kKc + kk;
  const_cast<float*>(p.packed)[i] = (m < M && k < K) ? p.s.weight[l][(int64_t)m * K + k] : 0.0f;
}

// one geometry: grants the dynamic shared memory once per device and size, sizes the persistent grid, launches
template <int ROWS, int SPLIT, bool PACKED>
cudaError_t launch_geometry(const MlpParams& p, size_t smem, int device, int sms, cudaStream_t stream, bool verbose) {
  constexpr int kThreads = 2 * ROWS * SPLIT;
  static size_t granted[64] = {};
  static int ctas_per_sm[64] = {};
  const int d = device >= 0 && device < 64 ? device : 0;
  if (granted[d] != smem) {  // (re)query: the CTAs per SM depend on the size
    cudaFuncSetAttribute(k_mlp_forward<ROWS, SPLIT, PACKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ROWS == 64)
      cudaFuncSetAttribute(k_mlp_forward<ROWS, SPLIT, PACKED>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int per_sm = 1;  // 64-row tiles: 2 when two tiles' activations fit on an SM (cfg5: 2 x 112.5 KB)
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_mlp_forward<ROWS, SPLIT, PACKED>, kThreads, smem);
    ctas_per_sm[d] = per_sm < 1 ? 1 : per_sm;
    granted[d] = smem;
    if (verbose)
      fprintf(stderr, "sus_mlp_forward: %d-row tiles, %d threads, %zu bytes of shared memory, %d CTAs per SM, %s weights\n", ROWS, kThreads,
              smem, ctas_per_sm[d], PACKED ? "packed" : "torch-layout");
  }
  const int64_t tiles = (p.n_rows + ROWS - 1) / ROWS;
  const int64_t grid = (int64_t)sms * ctas_per_sm[d];
  k_mlp_forward<ROWS, SPLIT, PACKED><<<(unsigned)(tiles < grid ? tiles : grid), kThreads, smem, stream>>>(p);
  return cudaGetLastError();
}

// validates the chain; a / b = widest activation at odd / even positions.  The last layer's output goes to global memory, but
// with two k-parts its partial sums pass through the region it would occupy, so it is counted as well.
int check_spec(const SusMlpSpec* spec, int& a, int& b) {
  if (!spec) return sus_internal_fail(SUS_ERR_INVALID_ARGUMENT, "mlp_forward: NULL argument");
  if (spec->n_layers < 1 || spec->n_layers > SUS_MLP_MAX_LAYERS) return sus_internal_fail(SUS_ERR_INVALID_ARGUMENT, "mlp_forward: 1..8 layers");
  if (spec->activation < SUS_ACT_NONE || spec->activation > SUS_ACT_PRELU)
    return sus_internal_fail(SUS_ERR_INVALID_ARGUMENT, "mlp_forward: unknown activation");
  a = b = 0;
  for (int l = 0; l <= spec->n_layers; ++l) {
    if (spec->dims[l] < 1) return sus_internal_fail(SUS_ERR_INVALID_ARGUMENT, "mlp_forward: layer width < 1");
    if (l < spec->n_layers && !spec->weight[l]) return sus_internal_fail(SUS_ERR_INVALID_ARGUMENT, "mlp_forward: NULL weight");
    if (l & 1) { if (spec->dims[l] > a) a = spec->dims[l]; } else { if (spec->dims[l] > b) b = spec->dims[l]; }
  }
  if (a < spec->dims[0]) a = spec->dims[0];  // region 1 also stages the raw (row-major) input tile before layer 0 runs
  return SUS_OK;
}

}  // namespace

extern "C" int64_t sus_mlp_workspace_bytes(const SusMlpSpec* spec) {
  int a, b;
  if (check_spec(spec, a, b) != SUS_OK) return 0;
  int64_t floats = 0;
  for (int l = 0; l < spec->n_layers; ++l) floats += mlp_packed_floats(spec->dims[l], spec->dims[l + 1]);
  return floats * (int64_t)sizeof(float);
}

extern "C" int sus_mlp_forward_ws(const SusMlpSpec* spec, const float* x, int64_t n_rows, float* out, void* workspace,
                                  int64_t workspace_bytes, int device, void* stream) {
  if (n_rows > 0 && (!x || !out)) return sus_internal_fail(SUS_ERR_INVALID_ARGUMENT, "mlp_forward: NULL argument");
  int a, b;
  if (int rc = check_spec(spec, a, b)) return rc;
  if (workspace && (workspace_bytes < sus_mlp_workspace_bytes(spec) || (reinterpret_cast<uintptr_t>(workspace) & 15u)))
    return sus_internal_fail(SUS_ERR_INVALID_ARGUMENT, "mlp_forward: workspace smaller than sus_mlp_workspace_bytes() or not 16-byte aligned");
  int prev = -1;
  cudaGetDevice(&prev);
  if (prev != device) cudaSetDevice(device);
  int max_smem = 0, sms = 0;
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  auto smem_of = [&](int rows) { return ((size_t)(a + b) * rows + 2 * kKc * 132) * sizeof(float); };
  // geometry knobs, read per call (tests and tools switch inside one process): SUSNET_MLP_ROWS = 64 | 128 rows per tile,
  // SUSNET_MLP_SPLIT = 1 | 2 k-parts, SUSNET_MLP_PACKED = 0 ignores the workspace.  128-row tiles unless only a 64-row tile of
  // the two widest adjacent layers fits.
  const char* rows_env = getenv("SUSNET_MLP_ROWS");
  const char* split_env = getenv("SUSNET_MLP_SPLIT");
  const char* packed_env = getenv("SUSNET_MLP_PACKED");
  const int rows = ((rows_env && atoi(rows_env) == 64) || smem_of(128) > (size_t)max_smem) ? 64 : 128;
  const int split = split_env ? (atoi(split_env) == 2 ? 2 : 1) : kDefaultSplit;
  const bool packed = workspace && split == 1 && !(packed_env && atoi(packed_env) == 0);  // (the packed chunks follow one k-part)
  const size_t smem = smem_of(rows);
  int rc = SUS_OK;
  if (smem > (size_t)max_smem) {
    rc = sus_internal_fail(SUS_ERR_UNSUPPORTED, "mlp_forward: the two widest adjacent layers do not fit in shared memory (run the module itself)");
  } else if (n_rows > 0) {
    MlpParams p;
    p.s = *spec; p.x = x; p.out = out; p.n_rows = n_rows; p.region_a_floats = a; p.region_b_floats = b;
    p.packed = packed ? static_cast<const float*>(workspace) : nullptr;
    int64_t off = 0;
    for (int l = 0; l < SUS_MLP_MAX_LAYERS; ++l) {
      p.packed_off[l] = off;
      if (l < spec->n_layers) off += mlp_packed_floats(spec->dims[l], spec->dims[l + 1]);
    }
    static const bool verbose = getenv("SUSNET_MLP_VERBOSE") != nullptr;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t err = cudaSuccess;
    if (packed) {
      k_mlp_pack<<<(unsigned)((off + 255) / 256), 256, 0, st>>>(p, off);
      sus_internal_count_launch();
      err = cudaGetLastError();
    }
    if (err == cudaSuccess) {
      if (rows == 128) {
        if (packed) err = launch_geometry<128, 1, true>(p, smem, device, sms, st, verbose);
        else err = split == 2 ? launch_geometry<128, 2, false>(p, smem, device, sms, st, verbose) : launch_geometry<128, 1, false>(p, smem, device, sms, st, verbose);
      } else {
        if (packed) err = launch_geometry<64, 1, true>(p, smem, device, sms, st, verbose);
        else err = split == 2 ? launch_geometry<64, 2, false>(p, smem, device, sms, st, verbose) : launch_geometry<64, 1, false>(p, smem, device, sms, st, verbose);
      }
      sus_internal_count_launch();
    }
    if (err != cudaSuccess) rc = sus_internal_fail(SUS_ERR_CUDA, cudaGetErrorString(err));
  }
  if (prev != device && prev >= 0) cudaSetDevice(prev);
  return rc;
}

extern "C" int sus_mlp_forward(const SusMlpSpec* spec, const float* x, int64_t n_rows, float* out, int device, void* stream) {
  return sus_mlp_forward_ws(spec, x, n_rows, out, nullptr, 0, device, stream);
}
