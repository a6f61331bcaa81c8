// Per-show archive statistics on sm_100a.
// Replaces computeArchiveShowStats (reference public/app.js:3898-3953).
//
// ONE kernel.  A CTA owns 256 consecutive shows; their entries are one contiguous
// row range of the entry columns, walked in chunks of kChunk rows:
//   phase A  entry-parallel.  status / launched are read as aligned 32-bit words and compared with
//            one LOP3 per word ((x ^ literal) & mask; the mask makes letters case-insensitive) —
//            no byte loops, no early exits; delaySec is staged.  Rows with a non-empty
//            primaryIssue are pushed on a shared-memory queue.
//   phase A' the queue is drained with all lanes busy (only ~1 row in 4 has an issue, so doing
//            this inline leaves ~3/4 of every warp idle): trim, exact word-wise match against
//            PRIMARY_ISSUES, else 'Other'.
//            Result of A/A': one code byte + one double per row in SHARED memory (never HBM).
//   phase B  show-parallel.  Thread t walks the rows of show t that fall in the chunk, in entry
//            order: packed 8-bit counters fed from a 128-entry LUT, the left-to-right delaySec sum
//            (bit-exact with Array.prototype.reduce), Math.max, first-occurrence order of issues.
// Per-thread state lives in registers across chunks, so shows of any length (ragged, > kChunk,
// empty) are handled by the same code.  HBM traffic = inputs once + the plane-major table once.
#include "pie_device.cuh"
#include "pie_kernels.h"

namespace pie {

std::atomic<unsigned long long> g_launches{0};
int g_sm_count = 0;

// tunables (overridable with -D for scripts/sweep_stats.py; defaults are the measured best)
#ifndef PIE_STATS_THREADS
#define PIE_STATS_THREADS 256
#endif
#ifndef PIE_STATS_CHUNK
#define PIE_STATS_CHUNK 4096
#endif
#ifndef PIE_STATS_ROWS
#define PIE_STATS_ROWS 2
#endif
#ifndef PIE_STATS_MIN_BLOCKS
#define PIE_STATS_MIN_BLOCKS 4
#endif
constexpr int kShowsPerCta = PIE_STATS_THREADS;  // shows (= threads) per CTA
constexpr int kChunk = PIE_STATS_CHUNK;          // rows staged in shared memory per iteration
constexpr int kRows = PIE_STATS_ROWS;            // rows per thread in flight in phase A

// code byte: bits 0-1 status (0 other, 1 completed, 2 no-launch, 3 abort), bit 2 launched == 'yes',
// bits 3-6 issue (0 none, k+1 = PRIMARY_ISSUES[k]), bit 7 Number.isFinite(delaySec)

#define PIE_W6(s) {lit_word(s, 0), lit_word(s, 1), lit_word(s, 2), lit_word(s, 3), lit_word(s, 4), lit_word(s, 5)}
#define PIE_M6(s) {lit_mask(s, 0, false), lit_mask(s, 1, false), lit_mask(s, 2, false), lit_mask(s, 3, false), \
                   lit_mask(s, 4, false), lit_mask(s, 5, false)}
// [0] = literal words, [1] = byte masks (exact compare) of PRIMARY_ISSUES[k-1]; row 0 never matches
__constant__ uint32_t c_issue_words[2][PIE_N_ISSUES + 1][6] = {
    {{0, 0, 0, 0, 0, 0},
     PIE_W6("Tracking lost"), PIE_W6("Failed to launch"), PIE_W6("Command delay"), PIE_W6("RF link"),
     PIE_W6("Battery"), PIE_W6("Motor or prop"), PIE_W6("Sensor or IMU"), PIE_W6("Software or show control"),
     PIE_W6("Operator input"), PIE_W6("Other")},
    {{0, 0, 0, 0, 0, 0},
     PIE_M6("Tracking lost"), PIE_M6("Failed to launch"), PIE_M6("Command delay"), PIE_M6("RF link"),
     PIE_M6("Battery"), PIE_M6("Motor or prop"), PIE_M6("Sensor or IMU"), PIE_M6("Software or show control"),
     PIE_M6("Operator input"), PIE_M6("Other")}};

struct StatsSmem {
  double delay[kChunk];
  uint4 lut[256];                          // packed counter increments per code byte
  uint32_t tbl[2][PIE_N_ISSUES + 1][6];    // row stride 6 words: rows 1..10 start in distinct banks
  uint16_t queue[kChunk];                  // rows of the chunk that have a primaryIssue
  uint8_t code[kChunk];
  uint32_t queue_n;
};

// String(entry?.status || '').toLowerCase() -> 1 completed / 2 no-launch / 3 abort / 0   (:3907-3914)
// Branch-free: the loads are predicated on the length, so that several rows' loads can be in
// flight before any of them is consumed.
__device__ __forceinline__ void status_fetch(const uint8_t* __restrict__ p, int n, uint32_t (&x)[3]) {
  const bool ok = (n == 9) | (n == 5);
  fetch_words_raw<3>(p, ok ? n : 0, x);  // n = 0 -> `last` < 0 -> no load is issued
}
__device__ __forceinline__ uint32_t status_match(const uint32_t (&x)[3], int n) {
  const uint32_t c9 = words_equal_ci(x, "completed") ? 1u : (words_equal_ci(x, "no-launch") ? 2u : 0u);
  const uint32_t c5 = words_equal_ci(x, "abort") ? 3u : 0u;
  return n == 9 ? c9 : (n == 5 ? c5 : 0u);
}

// PRIMARY_ISSUES.includes(issue) ? issue : 'Other' on the TRIMMED, non-empty string -> 1..10 (:3923)
__device__ __forceinline__ uint32_t issue_code(const uint8_t* __restrict__ p, int n, const StatsSmem& sm) {
  if (n != 7 && n != 13 && n != 14 && n != 16 && n != 24) return 10;
  uint32_t x[6];
  fetch_words_raw<6>(p, n, x);
  uint32_t cand;
  if (n == 7) cand = (x[0] == lit_word("RF link", 0)) ? 4u : 5u;
  else if (n == 13) cand = (x[0] == lit_word("Tracking lost", 0)) ? 1u : (x[0] == lit_word("Command delay", 0)) ? 3u
                         : (x[0] == lit_word("Motor or prop", 0)) ? 6u : 7u;
  else cand = (n == 14) ? 9u : (n == 16) ? 2u : 8u;
  uint32_t diff = 0;  // the candidate has exactly n bytes, so its masks cover exactly the string
#pragma unroll
  for (int k = 0; k < 6; ++k) diff |= (x[k] ^ sm.tbl[0][cand][k]) & sm.tbl[1][cand][k];
  return diff == 0 ? cand : 10u;
}

struct ShowAcc {
  // packed 8-bit counters (flushed into the wide ones before any byte can overflow)
  //   p.x completed | no-launch << 8 | abort << 16 | launched << 24
  //   p.y delay count | issue1 << 8 | issue2 << 16 | issue3 << 24
  //   p.z issue4 .. issue7          p.w issue8 .. issue10
  uint4 p;
  uint32_t pending;  // rows added since the last flush (<= 255)
  int32_t wide[15];  // completed, no-launch, abort, launched, delay count, issue1..10
  uint32_t seen, nd;
  unsigned long long order;
  double sum;
  long long max_key;  // Math.max as an integer max over order-preserving keys (ordered_key)

  __device__ __forceinline__ void init() {
    p = make_uint4(0, 0, 0, 0);
    pending = 0;
#pragma unroll
    for (int k = 0; k < 15; ++k) wide[k] = 0;
    seen = nd = 0;
    order = 0;
    sum = 0.0;
    max_key = kKeyLowest;
  }
  __device__ __forceinline__ void flush() {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      wide[j] += (p.x >> (8 * j)) & 0xFF;
      wide[4 + j] += (p.y >> (8 * j)) & 0xFF;
      wide[8 + j] += (p.z >> (8 * j)) & 0xFF;
      if (j < 3) wide[12 + j] += (p.w >> (8 * j)) & 0xFF;
    }
    p = make_uint4(0, 0, 0, 0);
    pending = 0;
  }
  __device__ __forceinline__ void add(uint32_t c, double d, const StatsSmem& sm) {
    const uint4 inc = sm.lut[c];
    p.x += inc.x;
    p.y += inc.y;
    p.z += inc.z;
    p.w += inc.w;
    const uint32_t iss = (c >> 3) & 15u;
    const uint32_t ibit = 1u << iss;
    if (iss != 0 && !(seen & ibit)) {  // first time this issue appears: next nibble of the order code
      order |= (unsigned long long)iss << (4 * nd);
      nd += 1;
      seen |= ibit;
    }
    // left to right, initial 0 (:3928).  Phase A stored -0.0 for rows whose delaySec does not
    // count; x + (-0.0) == x bit-for-bit for every x this sum can hold (it starts at +0 and can
    // never become -0), so the add is unconditional.
    sum = sum + d;
    const long long k = (c & 0x80u) ? ordered_key(d) : kKeyLowest;
    max_key = k > max_key ? k : max_key;
  }
};

__device__ __forceinline__ uint4 lut_entry(uint32_t c) {
  const uint32_t st = c & 3u, iss = (c >> 3) & 15u;
  const uint32_t bit = 1u << (8 * (iss & 3u)), word = iss >> 2;
  uint4 r;
  r.x = ((1u << (8 * st)) >> 8) + ((c & 4u) << 22);
  r.y = ((word == 0 && iss != 0) ? bit : 0u) + (c >> 7);  // byte 0: delaySec counted
  r.z = (word == 1) ? bit : 0u;
  r.w = (word == 2) ? bit : 0u;
  return r;
}

__global__ void __launch_bounds__(kShowsPerCta, PIE_STATS_MIN_BLOCKS)
    show_stats_kernel(pie_archive_view v, int32_t* __restrict__ si, double* __restrict__ sf, int64_t stride) {
  extern __shared__ __align__(16) uint8_t smem_raw[];  // dynamic: StatsSmem exceeds the 48 KB static limit
  StatsSmem& sm = *reinterpret_cast<StatsSmem*>(smem_raw);
  const int tid = threadIdx.x;
  for (int c = tid; c < 256; c += kShowsPerCta) sm.lut[c] = lut_entry(c);
  for (int k = tid; k < 2 * (PIE_N_ISSUES + 1) * 6; k += kShowsPerCta)
    (&sm.tbl[0][0][0])[k] = (&c_issue_words[0][0][0])[k];
  if (tid == 0) sm.queue_n = 0;

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
    // ---- phase A: status, launched, delaySec of every row of the chunk
    // kRows rows per thread at a time: all offset loads, then all string-word loads, then the
    // compares — 2 load waits per kRows rows instead of 4 per row.
    for (int i0 = c0 + tid; i0 < c1; i0 += kRows * kShowsPerCta) {
      int sb[kRows], sn[kRows], lb[kRows], ln[kRows], in[kRows];
      double d[kRows];
      uint32_t dv[kRows];
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        const int i = i0 + r * kShowsPerCta;
        const int j = i < c1 ? i : c0;  // clamp: loads stay in bounds, results are discarded
        sb[r] = v.status.offsets[j];
        sn[r] = v.status.offsets[j + 1] - sb[r];
        lb[r] = v.launched.offsets[j];
        ln[r] = v.launched.offsets[j + 1] - lb[r];
        in[r] = v.primary_issue.offsets[j + 1] - v.primary_issue.offsets[j];
        d[r] = v.delay_sec[j];
        dv[r] = v.delay_valid[j];
      }
      uint32_t xs[kRows][3], xl[kRows][1];
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        status_fetch(v.status.data + sb[r], sn[r], xs[r]);
        fetch_words_raw<1>(v.launched.data + lb[r], ln[r] == 3 ? 3 : 0, xl[r]);
      }
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        const int i = i0 + r * kShowsPerCta;
        if (i < c1) {
          uint32_t code = status_match(xs[r], sn[r]);
          // String(entry?.launched || '').toLowerCase() === 'yes'   (:3915)
          if (ln[r] == 3 && words_equal_ci(xl[r], "yes")) code |= 4u;
          if (dv[r] && is_finite_f64(d[r])) code |= 0x80u;  // Number.isFinite(entry?.delaySec) (:3918)
          if (in[r] > 0) sm.queue[atomicAdd(&sm.queue_n, 1u)] = (uint16_t)(i - c0);
          sm.code[i - c0] = (uint8_t)code;
          sm.delay[i - c0] = (code & 0x80u) ? d[r] : -0.0;  // -0.0: exact no-op in the ordered sum
        }
      }
    }
    __syncthreads();
    // ---- phase A': rows with a primaryIssue, densely packed
    const int qn = (int)sm.queue_n;
    for (int q = tid; q < qn; q += kShowsPerCta) {
      const int r = sm.queue[q];
      int b = v.primary_issue.offsets[c0 + r], e = v.primary_issue.offsets[c0 + r + 1];
      const uint8_t* __restrict__ str = v.primary_issue.data;
      const uint8_t first = str[b], last = str[e - 1];
      if (first <= 0x20 || first >= 0x80 || last <= 0x20 || last >= 0x80) {  // rare: primaryIssue.trim() (:3921)
        while (b < e) {
          const int l = js_ws_len_at(str, b, e);
          if (!l) break;
          b += l;
        }
        while (e > b) {
          const int l = js_ws_len_before(str, b, e);
          if (!l) break;
          e -= l;
        }
      }
      if (e > b) sm.code[r] |= (uint8_t)(issue_code(str + b, e - b, sm) << 3);
    }
    __syncthreads();
    // ---- phase B: this thread's show, rows that fall in the chunk, in entry order
    const int lo = e0 > c0 ? e0 : c0, hi = e1 < c1 ? e1 : c1;
    for (int blk = lo; blk < hi; blk += 255) {
      const int end = (hi - blk > 255) ? blk + 255 : hi;
      if (acc.pending + (uint32_t)(end - blk) > 255u) acc.flush();
      acc.pending += (uint32_t)(end - blk);
      for (int e = blk; e < end; ++e) acc.add(sm.code[e - c0], sm.delay[e - c0], sm);
    }
    if (tid == 0) sm.queue_n = 0;
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
  sf[PIE_SF_MAX_DELAY * stride + s] = delay_n ? from_ordered_key(acc.max_key) : nan;  // :3930
  // (count / totalEntries) * 100  (:3931-3937).  13 quotients share one denominator: one IEEE
  // reciprocal + the exact correction step instead of 13 divisions (bit-identical; pie_device.cuh).
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
    static int configured_device = -1;  // the attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (configured_device != dev) {
      cudaError_t e = cudaFuncSetAttribute(show_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)sizeof(StatsSmem));
      if (e != cudaSuccess) return e;
      configured_device = dev;
    }
    const unsigned grid = (unsigned)((v.n_shows + kShowsPerCta - 1) / kShowsPerCta);
    show_stats_kernel<<<grid, kShowsPerCta, sizeof(StatsSmem), stream>>>(v, si, sf, stride);
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
