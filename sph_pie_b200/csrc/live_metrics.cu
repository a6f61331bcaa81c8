// Live show metrics on sm_100a.
// Replaces computeMetrics(show) (reference public/app.js:5024-5047) for every show of a batch: the counts
// of planned / completed / no-launch / abort entries, Math.round(completed / plannedYes * 100), the average
// delay as (sum / n).toFixed(2), and the three most frequent primary issues of the entries that did not
// complete.
//
// A CTA per 128 shows, whose entries are one contiguous row range of the entry columns: the rows are classified
// row-parallel (coalesced loads, a byte + a double per row into shared memory), then a thread per show walks its
// rows in entry order from shared memory (the left-to-right float sum needs the order).  The issue
// ranking needs no per-show table of distinct strings: an entry that is the FIRST carrier of its issue
// counts the later carriers and enters a three-slot ranking; "first" and "later" are decided by comparing
// strings — O(k^2) in the k entries of the show that carry an issue, k <= 21 by the reference's own rules
// (one entry per operator, sqlProvider.js:434-457).
#include "pie_device.cuh"
#include "pie_kernels.h"
#include "pie_numfmt.cuh"

namespace pie {

// Number::toString tables for the |x| >= 1e21 branch of toFixed (this translation unit's own copy: the
// library is built without relocatable device code)
static __device__ const uint64_t d_pow5_inv[PIE_RYU_POW5_INV_SPLIT_N][2] = PIE_RYU_POW5_INV_SPLIT_INIT;
static __device__ const uint64_t d_pow5[PIE_RYU_POW5_SPLIT_N][2] = PIE_RYU_POW5_SPLIT_INIT;

// Number.prototype.toFixed(2) (ECMA-262 21.1.3.3): n = the integer closest to |x| * 100 computed on the EXACT
// binary value (ties: the larger n), printed as n/100 with two decimals; "-" iff x < 0 (so (-0.001).toFixed(2)
// is "-0.00"); NaN, and |x| >= 1e21 (incl. +-Infinity) through Number::toString.
__device__ int js_to_fixed2(double x, char* out) {
  if (x != x) {
    out[0] = 'N'; out[1] = 'a'; out[2] = 'N';
    return 3;
  }
  const double ax = fabs(x);
  if (ax >= 1e21) {
    const RyuTables t{d_pow5_inv, d_pow5};
    return js_number_to_string(x, out, t);
  }
  int o = 0;
  if (x < 0) out[o++] = '-';
  const uint64_t bits = (uint64_t)__double_as_longlong(ax);
  const uint32_t expo = (uint32_t)(bits >> 52);
  uint64_t m = bits & ((1ull << 52) - 1);
  int e;  // ax = m * 2^e
  if (expo == 0) {
    e = -1074;
  } else {
    m |= 1ull << 52;
    e = (int)expo - 1075;
  }
  // integer part as up to three 9-digit chunks (most significant first), two fraction digits
  uint32_t chunk[3] = {0, 0, 0};
  uint32_t frac;
  if (e >= 0) {  // an integer below 1e21 < 2^70: m << e in three 32-bit limbs, divided by 10^9 twice
    uint32_t limb[3];
    const unsigned __int128 big = (unsigned __int128)m << e;
    limb[0] = (uint32_t)big;
    limb[1] = (uint32_t)(big >> 32);
    limb[2] = (uint32_t)(big >> 64);
    for (int c = 2; c >= 0; --c) {
      uint64_t r = 0;
      for (int i = 2; i >= 0; --i) {
        const uint64_t cur = (r << 32) | limb[i];
        limb[i] = (uint32_t)(cur / 1000000000ull);
        r = cur % 1000000000ull;
      }
      chunk[c] = (uint32_t)r;
    }
    frac = 0;
  } else {
    const int k = -e;
    uint64_t n = 0;  // round-half-up(m * 100 / 2^k); m * 100 < 2^60
    if (k <= 62) n = (m * 100ull + (1ull << (k - 1))) >> k;
    frac = (uint32_t)(n % 100ull);
    const uint64_t ip = n / 100ull;  // < 2^53
    chunk[2] = (uint32_t)(ip % 1000000000ull);
    chunk[1] = (uint32_t)(ip / 1000000000ull);  // < 10^7
  }
  bool started = false;
  for (int c = 0; c < 3; ++c) {
    uint32_t v = chunk[c];
    char d[9];
    for (int i = 8; i >= 0; --i) {
      d[i] = (char)('0' + v % 10u);
      v /= 10u;
    }
    for (int i = 0; i < 9; ++i) {
      if (!started && d[i] == '0' && !(c == 2 && i == 8)) continue;
      started = true;
      out[o++] = d[i];
    }
  }
  out[o++] = '.';
  out[o++] = (char)('0' + frac / 10u);
  out[o++] = (char)('0' + frac % 10u);
  return o;
}

// array index key (ECMA-262 6.1.7): the canonical decimal string of an integer in 0 .. 2^32 - 2.
// OrdinaryOwnPropertyKeys lists such keys first, ascending, before the string keys in creation order.
__device__ __forceinline__ bool array_index_key(const uint8_t* __restrict__ s, int n, uint32_t* value) {
  if (n < 1 || n > 10 || (n > 1 && s[0] == '0')) return false;
  uint64_t v = 0;
  for (int i = 0; i < n; ++i) {
    if (s[i] < '0' || s[i] > '9') return false;
    v = v * 10 + (uint64_t)(s[i] - '0');
  }
  if (v > 4294967294ull) return false;
  *value = (uint32_t)v;
  return true;
}

// a[0..n) == b[0..n), four bytes at a time (aligned words, funnel-shifted to the strings' alignment)
__device__ __forceinline__ bool words_equal(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int n) {
  for (int i = 0; i < n; i += 4) {
    const int m = n - i < 4 ? n - i : 4;
    uint32_t xa[1], xb[1];
    fetch_words<1>(a + i, m, xa);
    fetch_words<1>(b + i, m, xb);
    if (xa[0] != xb[0]) return false;
  }
  return true;
}

struct Ranked {  // a distinct issue: where it was first seen, how often, and its place in Object.entries order
  int first;
  int count;
  unsigned long long order;  // array-index keys: their value; other keys: 2^32 + creation position
};
// does a come before b after Object.entries(...).sort((a, b) => b[1] - a[1])?  (the sort is stable)
__device__ __forceinline__ bool ranks_before(const Ranked& a, const Ranked& b) {
  return a.count != b.count ? a.count > b.count : a.order < b.order;
}

constexpr int kCmShows = 128;   // shows per CTA: a thread per show in the show-parallel phases
constexpr int kCmChunk = 2048;  // rows of the entry columns staged in shared memory per round

// what pass 1 needs of a row, in one byte
constexpr uint32_t kCmPlanned = 1, kCmCompleted = 2, kCmNoLaunch = 4, kCmAbort = 8, kCmDelay = 16, kCmCarries = 32;

__global__ void __launch_bounds__(kCmShows) compute_metrics_kernel(pie_archive_view v, int32_t* __restrict__ out,
                                                                   uint8_t* __restrict__ text, int64_t stride) {
  __shared__ double sm_delay[kCmChunk];
  __shared__ uint8_t sm_code[kCmChunk];
  __shared__ uint8_t sm_show[kCmChunk];    // which of the CTA's shows a row belongs to
  __shared__ uint16_t sm_queue[kCmChunk];  // the rows that carry an issue, in any order
  __shared__ uint16_t sm_rank[kCmChunk];   // a first carrier: how many rows of its show carry its issue; else 0
  __shared__ uint16_t sm_print[kCmChunk];  // a carrier's issue in 16 bits (length, first byte): unequal prints, unequal strings
  __shared__ int sm_queued;
  const int64_t s_first = (int64_t)blockIdx.x * kCmShows;
  const int64_t s_last = (s_first + kCmShows < v.n_shows ? s_first + kCmShows : v.n_shows);  // one past
  const int64_t s = s_first + threadIdx.x;
  const bool mine = s < v.n_shows;
  const int e0 = mine ? v.entry_offsets[s] : 0, e1 = mine ? v.entry_offsets[s + 1] : 0;
  const int tile_begin = v.entry_offsets[s_first], tile_end = v.entry_offsets[s_last];
  int planned = 0, completed = 0, no_launch = 0, abort_ = 0, dn = 0;
  double sum = 0.0;  // delays.reduce((a, b) => a + b, 0): left to right
  unsigned long long carries = 0;  // which of the first 64 entries carry an issue (status !== 'Completed' && primaryIssue)
  // pass 1.  The rows of the CTA's shows are one contiguous range of the entry columns, taken in rounds of kCmChunk:
  //   phase A, row-parallel (consecutive threads read consecutive rows: coalesced): the strings are compared as
  //            aligned 32-bit words (exact match) and every row becomes one byte, its delay one double;
  //   phase B, show-parallel: a thread walks the rows of its show in entry order from shared memory.
  if (threadIdx.x == 0) sm_queued = 0;  // (a CTA whose shows have no rows at all never enters the loop)
  __syncthreads();
  for (int c0 = tile_begin; c0 < tile_end; c0 += kCmChunk) {
    const int c1 = tile_end - c0 > kCmChunk ? c0 + kCmChunk : tile_end;
    if (threadIdx.x == 0) sm_queued = 0;
    __syncthreads();
    for (int e = c0 + (int)threadIdx.x; e < c1; e += kCmShows) {
      const int pb = v.planned.offsets[e], pn = v.planned.offsets[e + 1] - pb;
      const int sb = v.status.offsets[e], sn = v.status.offsets[e + 1] - sb;
      const int in = v.primary_issue.offsets[e + 1] - v.primary_issue.offsets[e];
      const bool dvalid = v.delay_valid[e] != 0;
      uint32_t xp[1], xs[3];
      fetch_words_raw<1>(v.planned.data + pb, pn == 3 ? 3 : 0, xp);
      fetch_words_raw<3>(v.status.data + sb, (sn == 9 || sn == 5) ? sn : 0, xs);
      const bool nine = sn == 9, five = sn == 5;
      const bool comp = nine && xs[0] == lit_word("Completed", 0) && xs[1] == lit_word("Completed", 1) &&
                        (xs[2] & 0xFFu) == lit_word("Completed", 2);
      const bool nol = nine && xs[0] == lit_word("No-launch", 0) && xs[1] == lit_word("No-launch", 1) &&
                       (xs[2] & 0xFFu) == lit_word("No-launch", 2);
      const bool abo = five && xs[0] == lit_word("Abort", 0) && (xs[1] & 0xFFu) == lit_word("Abort", 1);
      uint32_t code = (pn == 3 && (xp[0] & 0xFFFFFFu) == lit_word("Yes", 0)) ? kCmPlanned : 0;
      code |= comp ? kCmCompleted : 0;
      code |= nol ? kCmNoLaunch : 0;
      code |= abo ? kCmAbort : 0;
      code |= dvalid ? kCmDelay : 0;  // typeof v === 'number'
      code |= (!comp && in > 0) ? kCmCarries : 0;
      sm_code[e - c0] = (uint8_t)code;
      sm_delay[e - c0] = dvalid ? v.delay_sec[e] : 0.0;
      if (code & kCmCarries) {
        sm_queue[atomicAdd(&sm_queued, 1)] = (uint16_t)(e - c0);
        sm_print[e - c0] = (uint16_t)(((in > 255 ? 255 : in) << 8) | v.primary_issue.data[v.primary_issue.offsets[e]]);
      }
    }
    __syncthreads();
    const int b0 = e0 > c0 ? e0 : c0, b1 = e1 < c1 ? e1 : c1;
    for (int e = b0; e < b1; ++e) {
      const uint32_t code = sm_code[e - c0];
      sm_show[e - c0] = (uint8_t)threadIdx.x;
      planned += code & kCmPlanned;
      completed += (code >> 1) & 1;
      no_launch += (code >> 2) & 1;
      abort_ += (code >> 3) & 1;
      if (code & kCmDelay) {
        sum = sum + sm_delay[e - c0];
        ++dn;
      }
      if ((code & kCmCarries) && e - e0 < 64) carries |= 1ull << (e - e0);
    }
    __syncthreads();
  }
  Ranked top[3] = {{-1, 0, 0}, {-1, 0, 0}, {-1, 0, 0}};
  unsigned int created = 0;  // string keys created so far
  auto enter = [&](int e, int count, const uint8_t* p, int n) {  // a distinct issue, met first in row e
    Ranked r{e, count, 0};
    uint32_t index;
    if (array_index_key(p, n, &index)) r.order = index;
    else r.order = (1ull << 32) + created++;
    // insert into the three-slot ranking
    if (top[2].first < 0 || ranks_before(r, top[2])) {
      top[2] = r;
      if (top[1].first < 0 || ranks_before(top[2], top[1])) {
        const Ranked t = top[1]; top[1] = top[2]; top[2] = t;
        if (top[0].first < 0 || ranks_before(top[1], top[0])) {
          const Ranked u = top[0]; top[0] = top[1]; top[1] = u;
        }
      }
    }
  };
  // pass 2: the first carrier of every distinct issue counts the later ones and enters the ranking.
  if (tile_end - tile_begin <= kCmChunk) {
    // The usual case — all rows of the CTA's shows were one round, and their codes are still in shared memory:
    //   phase C, carrier-parallel (a lane per queued row, all lanes busy): compare the row's issue with the other
    //            carriers of its show — an equal one before it: not the first; equal ones after it: counted;
    //   phase D, show-parallel: a thread ranks the first carriers of its show in entry order.
    const int c0 = tile_begin;
    const int queued = sm_queued;
    for (int qi = threadIdx.x; qi < queued; qi += kCmShows) {
      const int e = c0 + sm_queue[qi];
      const int64_t sh = s_first + sm_show[e - c0];
      const int r0 = v.entry_offsets[sh], r1 = v.entry_offsets[sh + 1];
      const int b = v.primary_issue.offsets[e], n = v.primary_issue.offsets[e + 1] - b;
      const uint8_t* p = v.primary_issue.data + b;
      const uint32_t print = sm_print[e - c0];
      int count = 1;
      for (int j = r0; j < r1; ++j) {
        if (j == e || !(sm_code[j - c0] & kCmCarries) || sm_print[j - c0] != print) continue;
        const int jb = v.primary_issue.offsets[j], jn = v.primary_issue.offsets[j + 1] - jb;
        if (jn != n || !words_equal(v.primary_issue.data + jb, p, n)) continue;
        if (j < e) { count = 0; break; }
        ++count;
      }
      sm_rank[e - c0] = (uint16_t)(count > 0xFFFF ? 0xFFFF : count);
    }
    __syncthreads();
    if (mine) {
      for (int e = e0; e < e1; ++e) {
        if (!(sm_code[e - c0] & kCmCarries) || sm_rank[e - c0] == 0) continue;
        const int b = v.primary_issue.offsets[e], n = v.primary_issue.offsets[e + 1] - b;
        enter(e, sm_rank[e - c0], v.primary_issue.data + b, n);
      }
    }
  } else if (mine) {
    // shows longer than a round: the same, thread per show on the columns themselves
    auto carries_issue = [&](int e) -> bool {
      if (e - e0 < 64) return (carries >> (e - e0)) & 1ull;
      const int sb = v.status.offsets[e], sn = v.status.offsets[e + 1] - sb;
      return !equals_exact(v.status.data + sb, sn, "Completed") && v.primary_issue.offsets[e + 1] > v.primary_issue.offsets[e];
    };
    for (int e = e0; e < e1; ++e) {
      if (!carries_issue(e)) continue;
      const int b = v.primary_issue.offsets[e], n = v.primary_issue.offsets[e + 1] - b;
      const uint8_t* p = v.primary_issue.data + b;
      bool first = true;
      for (int j = e0; j < e && first; ++j) {
        if (!carries_issue(j)) continue;
        const int jb = v.primary_issue.offsets[j], jn = v.primary_issue.offsets[j + 1] - jb;
        if (jn == n && words_equal(v.primary_issue.data + jb, p, n)) first = false;
      }
      if (!first) continue;
      int count = 1;
      for (int j = e + 1; j < e1; ++j) {
        if (!carries_issue(j)) continue;
        const int jb = v.primary_issue.offsets[j], jn = v.primary_issue.offsets[j + 1] - jb;
        count += (jn == n && words_equal(v.primary_issue.data + jb, p, n));
      }
      enter(e, count, p, n);
    }
  }
  if (!mine) return;
  int rate = 0;
  if (planned) {
    const double q = ((double)completed / (double)planned) * 100.0;  // two IEEE operations, as written in the source
    const double r = floor(q);
    rate = (int)((q - r >= 0.5) ? r + 1.0 : r);  // Math.round: ties toward +inf
  }
  out[PIE_CM_SUCCESS_RATE * stride + s] = rate;
  out[PIE_CM_COMPLETED * stride + s] = completed;
  out[PIE_CM_NO_LAUNCH * stride + s] = no_launch;
  out[PIE_CM_ABORT * stride + s] = abort_;
  out[PIE_CM_TOP0 * stride + s] = top[0].first;
  out[PIE_CM_TOP1 * stride + s] = top[1].first;
  out[PIE_CM_TOP2 * stride + s] = top[2].first;
  __align__(16) char buf[PIE_CM_TEXT];
#pragma unroll
  for (int i = 0; i < PIE_CM_TEXT; ++i) buf[i] = 0;
  int len;
  if (dn) {
    len = js_to_fixed2(sum / (double)dn, buf);
  } else {
    buf[0] = '0'; buf[1] = '.'; buf[2] = '0'; buf[3] = '0';
    len = 4;
  }
  out[PIE_CM_AVG_LEN * stride + s] = len;
  uint4* dst = reinterpret_cast<uint4*>(text + PIE_CM_TEXT * s);  // 32-byte slots of a 16-byte aligned buffer
  const uint4* src = reinterpret_cast<const uint4*>(buf);
  dst[0] = src[0];
  dst[1] = src[1];
}

cudaError_t launch_compute_metrics(const pie_archive_view& v, int32_t* metrics_i32, uint8_t* avg_delay_text,
                                   int64_t stride, cudaStream_t stream) {
  if (v.n_shows == 0) return cudaSuccess;
  compute_metrics_kernel<<<(unsigned)((v.n_shows + kCmShows - 1) / kCmShows), kCmShows, 0, stream>>>(v, metrics_i32,
                                                                                                    avg_delay_text, stride);
  g_launches += 1;
  return cudaGetLastError();
}

}  // namespace pie
