// Per-show archive statistics on sm_100a.
// Replaces computeArchiveShowStats (reference public/app.js:3898-3953).
//
// ONE kernel.  A CTA owns 256 consecutive shows; their entries are one contiguous
// row range of the entry columns, walked in chunks of kChunk rows:
//   phase A  entry-parallel.  Each thread classifies rows of the chunk: status / launched /
//            primaryIssue are read as aligned 32-bit words (fetch_words) and compared word-wise after
//            SIMD-in-register case folding — no byte loops, no early exits — and delaySec is staged.
//            Result: one code byte + one double per row in SHARED memory (never written to HBM).
//   phase B  show-parallel.  Thread t walks the rows of show t that fall in the chunk, in entry
//            order: packed 8-bit counters, the left-to-right delaySec sum (bit-exact with
//            Array.prototype.reduce), Math.max, and the first-occurrence order of the issues.
// Per-thread state lives in registers across chunks, so shows of any length (ragged, > kChunk,
// empty) are handled by the same code.  HBM traffic = inputs once + the plane-major table once.
#include "pie_device.cuh"
#include "pie_kernels.h"

namespace pie {

unsigned long long g_launches = 0;

constexpr int kShowsPerCta = 256;
constexpr int kChunk = 2048;

// code byte: bits 0-1 status (0 other, 1 completed, 2 no-launch, 3 abort), bit 2 launched == 'yes',
// bits 3-6 issue (0 none, k+1 = PRIMARY_ISSUES[k]), bit 7 Number.isFinite(delaySec)

#define PIE_W6(s) {lit_word(s, 0), lit_word(s, 1), lit_word(s, 2), lit_word(s, 3), lit_word(s, 4), lit_word(s, 5)}
__constant__ uint32_t c_issue_words[PIE_N_ISSUES + 1][6] = {
    {0, 0, 0, 0, 0, 0},
    PIE_W6("Tracking lost"), PIE_W6("Failed to launch"), PIE_W6("Command delay"), PIE_W6("RF link"),
    PIE_W6("Battery"), PIE_W6("Motor or prop"), PIE_W6("Sensor or IMU"), PIE_W6("Software or show control"),
    PIE_W6("Operator input"), PIE_W6("Other")};

// String(entry?.status || '').toLowerCase() -> 1 completed / 2 no-launch / 3 abort / 0   (:3907-3914)
__device__ __forceinline__ uint32_t status_code(const uint8_t* __restrict__ p, int n) {
  if (n != 9 && n != 5) return 0;
  uint32_t x[3];
  fetch_words<3>(p, n, x);
  const uint32_t a = lower4(x[0]), b = lower4(x[1]), c = lower4(x[2]);
  const bool completed = (a == lit_word("completed", 0)) & (b == lit_word("completed", 1)) & (c == lit_word("completed", 2));
  const bool no_launch = (a == lit_word("no-launch", 0)) & (b == lit_word("no-launch", 1)) & (c == lit_word("no-launch", 2));
  const bool abort_ = (a == lit_word("abort", 0)) & (b == lit_word("abort", 1)) & (c == 0);
  return (n == 9) ? (completed ? 1u : (no_launch ? 2u : 0u)) : (abort_ ? 3u : 0u);
}

// PRIMARY_ISSUES.includes(issue) ? issue : 'Other' on the TRIMMED, non-empty string -> 1..10 (:3923)
__device__ __forceinline__ uint32_t issue_code(const uint8_t* __restrict__ p, int n, const uint32_t (*tbl)[6]) {
  if (n != 7 && n != 13 && n != 14 && n != 16 && n != 24) return 10;
  uint32_t x[6];
  fetch_words<6>(p, n, x);
  uint32_t cand;
  if (n == 7) cand = (x[0] == lit_word("RF link", 0)) ? 4u : 5u;
  else if (n == 13) cand = (x[0] == lit_word("Tracking lost", 0)) ? 1u : (x[0] == lit_word("Command delay", 0)) ? 3u
                         : (x[0] == lit_word("Motor or prop", 0)) ? 6u : 7u;
  else cand = (n == 14) ? 9u : (n == 16) ? 2u : 8u;
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 6; ++k) ok &= (x[k] == tbl[cand][k]);
  return ok ? cand : 10u;
}

__device__ __forceinline__ uint32_t classify_entry(const pie_archive_view& v, int64_t i, const uint32_t (*tbl)[6],
                                                   double* delay_out) {
  uint32_t code;
  {
    const int b = v.status.offsets[i], e = v.status.offsets[i + 1];
    code = status_code(v.status.data + b, e - b);
  }
  {  // String(entry?.launched || '').toLowerCase() === 'yes'   (:3915)
    const int b = v.launched.offsets[i], e = v.launched.offsets[i + 1];
    if (e - b == 3) {
      uint32_t x[1];
      fetch_words<1>(v.launched.data + b, 3, x);
      if (lower4(x[0]) == lit_word("yes", 0)) code |= 4u;
    }
  }
  {  // typeof primaryIssue === 'string' ? primaryIssue.trim() : ''   (:3921)
    int b = v.primary_issue.offsets[i], e = v.primary_issue.offsets[i + 1];
    if (e > b) {
      const uint8_t* __restrict__ s = v.primary_issue.data;
      const uint8_t first = s[b], last = s[e - 1];
      if (first <= 0x20 || first >= 0x80 || last <= 0x20 || last >= 0x80) {  // rare: may need trimming
        while (b < e) {
          const int l = js_ws_len_at(s, b, e);
          if (!l) break;
          b += l;
        }
        while (e > b) {
          const int l = js_ws_len_before(s, b, e);
          if (!l) break;
          e -= l;
        }
      }
      if (e > b) code |= issue_code(s + b, e - b, tbl) << 3;
    }
  }
  const double d = v.delay_sec[i];
  if (v.delay_valid[i] && is_finite_f64(d)) code |= 0x80u;  // Number.isFinite(entry?.delaySec)   (:3918)
  *delay_out = d;
  return code;
}

struct ShowAcc {
  // packed 8-bit counters (flushed into the wide ones before any byte can overflow)
  uint32_t p0;  // completed | no-launch << 8 | abort << 16 | launched << 24
  uint32_t p1;  // delay count | issue1 << 8 | issue2 << 16 | issue3 << 24
  uint32_t p2;  // issue4 .. issue7
  uint32_t p3;  // issue8 .. issue10
  uint32_t since;
  int32_t wide[15];  // completed, no-launch, abort, launched, delay count, issue1..10
  uint32_t seen, nd;
  unsigned long long order;
  double sum, mx;
  bool any_delay;

  __device__ __forceinline__ void init() {
    p0 = p1 = p2 = p3 = since = 0;
#pragma unroll
    for (int k = 0; k < 15; ++k) wide[k] = 0;
    seen = nd = 0;
    order = 0;
    sum = 0.0;
    mx = 0.0;
    any_delay = false;
  }
  __device__ __forceinline__ void flush() {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      wide[j] += (p0 >> (8 * j)) & 0xFF;
      wide[4 + j] += (p1 >> (8 * j)) & 0xFF;
      wide[8 + j] += (p2 >> (8 * j)) & 0xFF;
      if (j < 3) wide[12 + j] += (p3 >> (8 * j)) & 0xFF;
    }
    p0 = p1 = p2 = p3 = since = 0;
  }
  __device__ __forceinline__ void add(uint32_t c, double d) {
    const uint32_t st = c & 3u, iss = (c >> 3) & 15u;
    p0 += ((1u << (8 * st)) >> 8) + ((c & 4u) << 22);
    const uint32_t bit = 1u << (8 * (iss & 3u));
    const uint32_t word = iss >> 2;
    p1 += (c >> 7) + ((word == 0 && iss != 0) ? bit : 0u);
    p2 += (word == 1) ? bit : 0u;
    p3 += (word == 2) ? bit : 0u;
    const uint32_t ibit = 1u << iss;
    if (iss != 0 && !(seen & ibit)) {  // first time this issue appears: next nibble of the order code
      order |= (unsigned long long)iss << (4 * nd);
      nd += 1;
      seen |= ibit;
    }
    if (c & 0x80u) {
      sum = sum + d;  // left to right, initial 0   (:3928)
      mx = any_delay ? js_max(mx, d) : d;
      any_delay = true;
    }
    if (++since == 255) flush();
  }
};

__global__ void __launch_bounds__(kShowsPerCta) show_stats_kernel(pie_archive_view v, int32_t* __restrict__ si,
                                                                 double* __restrict__ sf, int64_t stride) {
  __shared__ uint8_t s_code[kChunk];
  __shared__ double s_delay[kChunk];
  __shared__ uint32_t s_tbl[PIE_N_ISSUES + 1][6];  // row stride 6 words: rows 1..10 start in distinct banks
  const int tid = threadIdx.x;
  if (tid < (PIE_N_ISSUES + 1) * 6) (&s_tbl[0][0])[tid] = (&c_issue_words[0][0])[tid];

  const int64_t s0 = (int64_t)blockIdx.x * kShowsPerCta;
  const int64_t s1 = (s0 + kShowsPerCta < v.n_shows) ? s0 + kShowsPerCta : v.n_shows;
  const int64_t s = s0 + tid;
  const bool have_show = s < s1;
  const int e0 = have_show ? v.entry_offsets[s] : 0;
  const int e1 = have_show ? v.entry_offsets[s + 1] : 0;
  const int tile_begin = v.entry_offsets[s0], tile_end = v.entry_offsets[s1];

  ShowAcc acc;
  acc.init();
  __syncthreads();

  for (int c0 = tile_begin; c0 < tile_end; c0 += kChunk) {
    const int c1 = (tile_end - c0 > kChunk) ? c0 + kChunk : tile_end;
    // phase A: classify the rows of this chunk
    for (int i = c0 + tid; i < c1; i += kShowsPerCta) {
      double d;
      const uint32_t code = classify_entry(v, i, s_tbl, &d);
      s_code[i - c0] = (uint8_t)code;
      s_delay[i - c0] = d;
    }
    __syncthreads();
    // phase B: this thread's show, rows that fall in the chunk, in entry order
    const int lo = e0 > c0 ? e0 : c0, hi = e1 < c1 ? e1 : c1;
    for (int e = lo; e < hi; ++e) acc.add(s_code[e - c0], s_delay[e - c0]);
    __syncthreads();
  }
  if (!have_show) return;
  acc.flush();

  const int total = e1 - e0;
  const int delay_n = acc.wide[4];
  const double nan = quiet_nan();
  si[PIE_SI_TOTAL * stride + s] = total;
  si[PIE_SI_COMPLETED * stride + s] = acc.wide[0];
  si[PIE_SI_NO_LAUNCH * stride + s] = acc.wide[1];
  si[PIE_SI_ABORT * stride + s] = acc.wide[2];
  si[PIE_SI_LAUNCHED * stride + s] = acc.wide[3];
  si[PIE_SI_DELAY_COUNT * stride + s] = delay_n;
  si[PIE_SI_ISSUE_ORDER_LO * stride + s] = (int32_t)(uint32_t)(acc.order & 0xFFFFFFFFull);
  si[PIE_SI_ISSUE_ORDER_HI * stride + s] = (int32_t)(uint32_t)(acc.order >> 32);
  sf[PIE_SF_AVG_DELAY * stride + s] = delay_n ? acc.sum / (double)delay_n : nan;  // :3929
  sf[PIE_SF_MAX_DELAY * stride + s] = delay_n ? acc.mx : nan;                     // :3930
  // (count / totalEntries) * 100  (:3931-3937).  13 quotients share one denominator: one IEEE
  // reciprocal + the exact correction step instead of 13 divisions (bit-identical; §pie_device.cuh).
  const double dt = (double)total;
  const bool fast = total > 0 && total <= kFastDivMax;
  const double y = fast ? 1.0 / dt : 0.0;
  auto rate = [&](int count) -> double {
    if (total == 0) return nan;
    const double q = fast ? div_by_shared_reciprocal((double)count, dt, y) : (double)count / dt;
    return q * 100.0;
  };
  sf[PIE_SF_COMPLETION_RATE * stride + s] = rate(acc.wide[0]);
  sf[PIE_SF_LAUNCH_RATE * stride + s] = rate(acc.wide[3]);
  sf[PIE_SF_ABORT_RATE * stride + s] = rate(acc.wide[2]);
#pragma unroll
  for (int k = 0; k < PIE_N_ISSUES; ++k) {
    si[(PIE_SI_ISSUE_COUNT0 + k) * stride + s] = acc.wide[5 + k];
    sf[(PIE_SF_ISSUE_RATE0 + k) * stride + s] = rate(acc.wide[5 + k]);
  }
}

cudaError_t launch_show_stats(const pie_archive_view& v, int32_t* si, double* sf, int64_t stride, int sm_count,
                              cudaStream_t stream) {
  (void)sm_count;
  if (v.n_shows > 0) {
    const unsigned grid = (unsigned)((v.n_shows + kShowsPerCta - 1) / kShowsPerCta);
    show_stats_kernel<<<grid, kShowsPerCta, 0, stream>>>(v, si, sf, stride);
    g_launches += 1;
  }
  return cudaGetLastError();
}

// ---- self test: shared-reciprocal quotient vs IEEE division --------------------------------
__global__ void selftest_fast_div_kernel(int max_b, unsigned long long* mismatches) {
  const int b = blockIdx.x + 1;
  if (b > max_b) return;
  const double db = (double)b, y = 1.0 / db;
  unsigned long long bad = 0;
  for (int a = threadIdx.x; a <= b; a += blockDim.x) {
    const double q = div_by_shared_reciprocal((double)a, db, y), ref = (double)a / db;
    bad += (__double_as_longlong(q) != __double_as_longlong(ref));
    bad += (__double_as_longlong(q * 100.0) != __double_as_longlong(ref * 100.0));
  }
  if (bad) atomicAdd(mismatches, bad);
}

cudaError_t launch_selftest_fast_div(int max_b, unsigned long long* d_mismatches, cudaStream_t stream) {
  selftest_fast_div_kernel<<<max_b, 128, 0, stream>>>(max_b, d_mismatches);
  g_launches += 1;
  return cudaGetLastError();
}

}  // namespace pie
