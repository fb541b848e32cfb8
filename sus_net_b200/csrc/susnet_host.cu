// susnet_host.cu -- HOST side of the compact host protocol (include/susnet_b200.h, SusCompactLayout): plain CPU code, no GPU work.
// The kernels move bit-packed action records in and reward-code records out; a host consumer that wants the reference's
// arrays back (rewards (N, A) float, done / truncated flags: what env.step() returns, base.py:397-407) decodes the records here
// through the float64 table of sus_reward_lut -- several threads over N, memory-bound (1 Mi envs: ~1 ms instead of 55 ms of numpy).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/susnet_b200.h"

extern "C" int sus_internal_fail(int code, const char* msg);

namespace {

template <typename F>
void parallel_for(int64_t n, int threads, F f) {
  if (threads <= 0) threads = (int)std::min<unsigned>(std::thread::hardware_concurrency(), 16u);
  threads = (int)std::max<int64_t>(1, std::min<int64_t>(threads, n / 65536 + 1));
  if (threads == 1) { f(0, n); return; }
  std::vector<std::thread> pool;
  const int64_t chunk = (n + threads - 1) / threads;
  for (int t = 0; t < threads; ++t) {
    const int64_t lo = t * chunk, hi = std::min<int64_t>(n, lo + chunk);
    if (lo < hi) pool.emplace_back([=] { f(lo, hi); });
  }
  for (auto& th : pool) th.join();
}

}  // namespace

extern "C" int sus_host_pack_actions(const SusConfig* cfg, const void* actions, int32_t dtype, int64_t n_envs, uint8_t* packed,
                                     int32_t threads) {
  if (!cfg || (n_envs > 0 && (!actions || !packed))) return sus_internal_fail(SUS_ERR_INVALID_ARGUMENT, "pack_actions: NULL argument");
  if (dtype != SUS_U8 && dtype != SUS_I32 && dtype != SUS_I64)
    return sus_internal_fail(SUS_ERR_INVALID_ARGUMENT, "pack_actions: dtype must be SUS_U8, SUS_I32 or SUS_I64");
  SusCompactLayout L;
  if (int rc = sus_compact_layout(cfg, &L)) return rc;
  const int A = cfg->n_imposters + cfg->n_crew, ab = L.action_bits, nb = L.action_bytes;
  const uint64_t mask = (1ull << ab) - 1ull;
  parallel_for(n_envs, threads, [=](int64_t lo, int64_t hi) {
    for (int64_t e = lo; e < hi; ++e) {
      uint64_t rec = 0;
      for (int i = 0; i < A; ++i) {
        long long a;
        if (dtype == SUS_U8) a = static_cast<const uint8_t*>(actions)[e * A + i];
        else if (dtype == SUS_I32) a = static_cast<const int32_t*>(actions)[e * A + i];
        else a = static_cast<const long long*>(actions)[e * A + i];
        // an index that does not fit the field becomes the field's maximum, which no role list reaches: the kernel rejects it
        const uint64_t v = (a < 0 || (uint64_t)a > mask) ? mask : (uint64_t)a;
        rec |= v << (i * ab);
      }
      for (int b = 0; b < nb; ++b) packed[e * nb + b] = (uint8_t)(rec >> (8 * b));
    }
  });
  return SUS_OK;
}

extern "C" int sus_host_decode_results(const SusConfig* cfg, const uint8_t* records, int64_t n_envs, void* rewards,
                                       int32_t rewards_dtype, uint8_t* done, uint8_t* truncated, int32_t threads) {
  if (!cfg || (n_envs > 0 && !records)) return sus_internal_fail(SUS_ERR_INVALID_ARGUMENT, "decode_results: NULL argument");
  if (rewards && rewards_dtype != SUS_F32 && rewards_dtype != SUS_F64)
    return sus_internal_fail(SUS_ERR_INVALID_ARGUMENT, "decode_results: rewards_dtype must be SUS_F32 or SUS_F64");
  SusCompactLayout L;
  if (int rc = sus_compact_layout(cfg, &L)) return rc;
  const int A = cfg->n_imposters + cfg->n_crew, rb = L.reward_bits, nb = L.result_bytes, stride = L.invalid_code + 1;
  std::vector<double> lut((size_t)A * stride);
  if (int rc = sus_reward_lut(cfg, lut.data())) return rc;
  std::vector<float> lut32(lut.begin(), lut.end());
  const double* lt = lut.data();
  const float* lt32 = lut32.data();
  const uint64_t mask = (1ull << rb) - 1ull;
  parallel_for(n_envs, threads, [=](int64_t lo, int64_t hi) {
    for (int64_t e = lo; e < hi; ++e) {
      uint64_t rec = 0;
      for (int b = 0; b < nb; ++b) rec |= (uint64_t)records[e * nb + b] << (8 * b);
      if (rewards) {
        if (rewards_dtype == SUS_F64) {
          double* r = static_cast<double*>(rewards) + e * A;
          for (int i = 0; i < A; ++i) r[i] = lt[i * stride + ((rec >> (i * rb)) & mask)];
        } else {
          float* r = static_cast<float*>(rewards) + e * A;
          for (int i = 0; i < A; ++i) r[i] = lt32[i * stride + ((rec >> (i * rb)) & mask)];
        }
      }
      if (done) done[e] = (uint8_t)((rec >> (A * rb)) & 1ull);
      if (truncated) truncated[e] = (uint8_t)((rec >> (A * rb + 1)) & 1ull);
    }
  });
  return SUS_OK;
}
