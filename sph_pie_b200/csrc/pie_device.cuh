// Device-side helpers shared by the archive kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sph_pie_b200.h"

#define PIE_SM_COUNT_B200 148

namespace pie {

// ---- entry classification code (1 byte per entry) ------------------------------------------
// bits 0-1 status: 0 other, 1 completed, 2 no-launch, 3 abort   (public/app.js:3907-3914)
// bit  2   launched == 'yes'                                   (:3915)
// bits 3-6 primary issue: 0 none, 1..10 = PRIMARY_ISSUES[k-1]   (:3921-3925)
// bit  7   Number.isFinite(delaySec)                            (:3918)
constexpr uint32_t kStatusMask = 0x3;
constexpr uint32_t kLaunchedBit = 0x4;
constexpr uint32_t kIssueShift = 3;
constexpr uint32_t kIssueMask = 0xF;
constexpr uint32_t kDelayBit = 0x80;

__device__ __forceinline__ uint8_t ascii_lower(uint8_t c) {
  return (c >= 'A' && c <= 'Z') ? (uint8_t)(c + 32) : c;
}

// s[0..n) equals lit (lower-case ASCII literal of length L) under ASCII case folding.
// Equivalent to JS `String(s).toLowerCase() === lit` for the literals used on this path: the only
// non-ASCII code point that lower-cases to an ASCII letter is U+212A (-> 'k'), and none of
// 'completed', 'no-launch', 'abort', 'yes', 'no' contains 'k'.
template <int L>
__device__ __forceinline__ bool equals_lower_ascii(const uint8_t* __restrict__ s, int n, const char (&lit)[L]) {
  if (n != L - 1) return false;
#pragma unroll
  for (int i = 0; i < L - 1; ++i)
    if (ascii_lower(s[i]) != (uint8_t)lit[i]) return false;
  return true;
}

template <int L>
__device__ __forceinline__ bool equals_exact(const uint8_t* __restrict__ s, int n, const char (&lit)[L]) {
  if (n != L - 1) return false;
#pragma unroll
  for (int i = 0; i < L - 1; ++i)
    if (s[i] != (uint8_t)lit[i]) return false;
  return true;
}

// ---- word-at-a-time string access -----------------------------------------------------------
// Strings are compared 4 bytes at a time: the NW aligned 32-bit words that contain bytes
// [p, p+n) are loaded (never a word that holds no byte of the string, so no over-read beyond the
// 4-byte-aligned end of the buffer), funnel-shifted to string alignment and zero-padded past n.
// Requires n >= 1 and n <= 4*NW.  Buffers must start 4-byte aligned in memory (any cudaMalloc /
// torch allocation does) — the aligned word containing the first / last byte is then in bounds.
template <int NW>
__device__ __forceinline__ void fetch_words(const uint8_t* __restrict__ p, int n, uint32_t (&x)[NW]) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uint32_t* __restrict__ w = reinterpret_cast<const uint32_t*>(a & ~static_cast<uintptr_t>(3));
  const uint32_t sh = static_cast<uint32_t>(a & 3) * 8;
  const int last = n > 0 ? static_cast<int>(((a & 3) + n - 1) >> 2) : -1;  // n <= 0: no load at all
  uint32_t raw[NW + 1];
#pragma unroll
  for (int k = 0; k <= NW; ++k) raw[k] = (k <= last) ? __ldg(w + k) : 0u;
#pragma unroll
  for (int k = 0; k < NW; ++k) {
    const uint32_t v = __funnelshift_r(raw[k], raw[k + 1], sh);
    const int nb = n - 4 * k;
    x[k] = nb >= 4 ? v : (nb <= 0 ? 0u : (v & ((1u << (8 * nb)) - 1u)));
  }
}

template <int L>
__host__ __device__ constexpr uint32_t lit_word(const char (&s)[L], int k);  // defined below

// Same fetch without the zero padding: bytes past n are whatever follows in the loaded words.
// For comparisons against a literal of exactly n bytes through literal-specific masks.
template <int NW>
__device__ __forceinline__ void fetch_words_raw(const uint8_t* __restrict__ p, int n, uint32_t (&x)[NW]) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uint32_t* __restrict__ w = reinterpret_cast<const uint32_t*>(a & ~static_cast<uintptr_t>(3));
  const uint32_t sh = static_cast<uint32_t>(a & 3) * 8;
  const int last = n > 0 ? static_cast<int>(((a & 3) + n - 1) >> 2) : -1;  // n <= 0: no load at all
  uint32_t raw[NW + 1];
#pragma unroll
  for (int k = 0; k <= NW; ++k) raw[k] = (k <= last) ? __ldg(w + k) : 0u;
#pragma unroll
  for (int k = 0; k < NW; ++k) x[k] = __funnelshift_r(raw[k], raw[k + 1], sh);
}

// Mask word k for comparing against literal s: 0xFF per literal byte (0xDF for a lower-case letter
// when case_insensitive: the byte may differ from the literal in bit 5 only, i.e. be its upper
// case), 0x00 past the literal's end.
template <int L>
__host__ __device__ constexpr uint32_t lit_mask(const char (&s)[L], int k, bool case_insensitive) {
  uint32_t v = 0;
  for (int j = 0; j < 4; ++j) {
    const int i = 4 * k + j;
    if (i < L - 1) {
      const bool letter = s[i] >= 'a' && s[i] <= 'z';
      v |= static_cast<uint32_t>((case_insensitive && letter) ? 0xDFu : 0xFFu) << (8 * j);
    }
  }
  return v;
}

// x[0..NW) (raw words of a string of EXACTLY L-1 bytes) equals the lower-case literal `s` under
// ASCII case folding.  One LOP3 per word: ((x ^ lit) & mask), OR-reduced.
template <int NW, int L>
__device__ __forceinline__ bool words_equal_ci(const uint32_t (&x)[NW], const char (&s)[L]) {
  static_assert(4 * NW >= L - 1, "literal longer than the fetched words");
  uint32_t diff = 0;
#pragma unroll
  for (int k = 0; k < NW; ++k) diff |= (x[k] ^ lit_word(s, k)) & lit_mask(s, k, true);
  return diff == 0;
}

// ASCII upper -> lower on 4 packed bytes, other bytes unchanged (no cross-byte carries).
__device__ __forceinline__ uint32_t lower4(uint32_t w) {
  const uint32_t hept = w & 0x7f7f7f7fu;
  const uint32_t ge_a = hept + 0x3f3f3f3fu;  // bit 7 set iff (b & 0x7f) >= 'A'
  const uint32_t gt_z = hept + 0x25252525u;  // bit 7 set iff (b & 0x7f) >  'Z'
  const uint32_t upper = ge_a & ~gt_z & ~w & 0x80808080u;
  return w | (upper >> 2);
}

// little-endian pack of literal bytes [4k, 4k+4), zero padded
template <int L>
__host__ __device__ constexpr uint32_t lit_word(const char (&s)[L], int k) {
  uint32_t v = 0;
  for (int j = 0; j < 4; ++j) {
    const int i = 4 * k + j;
    if (i < L - 1) v |= static_cast<uint32_t>(static_cast<unsigned char>(s[i])) << (8 * j);
  }
  return v;
}

// a / b for a, b small non-negative integers held in doubles, given y = RN(1/b):
// q0 = RN(a*y); r = a - b*q0 (exact in one FMA); q = RN(q0 + r*y) is the correctly rounded
// quotient (Markstein).  Used only for b <= kFastDivMax, where it is verified EXHAUSTIVELY against
// IEEE division by pie_selftest_fast_div (tests/test_gpu_parity.py).
constexpr int kFastDivMax = 4096;
__device__ __forceinline__ double div_by_shared_reciprocal(double a, double b, double y) {
  const double q0 = a * y;
  const double r = fma(-b, q0, a);
  return fma(r, y, q0);
}

// Length in bytes of an ECMAScript WhiteSpace/LineTerminator code point starting at s[i] (UTF-8),
// 0 if s[i] does not start one.  Set: TAB LF VT FF CR SP, U+00A0, U+1680, U+2000-200A, U+2028,
// U+2029, U+202F, U+205F, U+3000, U+FEFF  (String.prototype.trim).
__device__ __forceinline__ int js_ws_len_at(const uint8_t* __restrict__ s, int i, int n) {
  uint8_t c = s[i];
  if (c == 0x20 || (c >= 0x09 && c <= 0x0D)) return 1;
  if (c < 0xC2) return 0;
  if (c == 0xC2) return (i + 1 < n && s[i + 1] == 0xA0) ? 2 : 0;
  if (i + 2 >= n) return 0;
  uint8_t d = s[i + 1], e = s[i + 2];
  if (c == 0xE1) return (d == 0x9A && e == 0x80) ? 3 : 0;                       // U+1680
  if (c == 0xE2) {
    if (d == 0x80) return ((e >= 0x80 && e <= 0x8A) || e == 0xA8 || e == 0xA9 || e == 0xAF) ? 3 : 0;
    if (d == 0x81) return (e == 0x9F) ? 3 : 0;                                  // U+205F
    return 0;
  }
  if (c == 0xE3) return (d == 0x80 && e == 0x80) ? 3 : 0;                       // U+3000
  if (c == 0xEF) return (d == 0xBB && e == 0xBF) ? 3 : 0;                       // U+FEFF
  return 0;
}

// Same, for a code point that ENDS at s[end-1].
__device__ __forceinline__ int js_ws_len_before(const uint8_t* __restrict__ s, int begin, int end) {
  uint8_t c = s[end - 1];
  if (c == 0x20 || (c >= 0x09 && c <= 0x0D)) return 1;
  if (c < 0x80) return 0;
  if (end - begin >= 2 && s[end - 2] == 0xC2 && c == 0xA0) return 2;
  if (end - begin >= 3) {
    int l = js_ws_len_at(s, end - 3, end);
    return l == 3 ? 3 : 0;
  }
  return 0;
}

// Math.max / Math.min on finite doubles: +0 is larger than -0.
__device__ __forceinline__ double js_max(double a, double b) {
  if (a > b) return a;
  if (b > a) return b;
  return (__double2hiint(a) < 0) ? b : a;  // equal (or both zero): prefer the one without sign bit
}
__device__ __forceinline__ double js_min(double a, double b) {
  if (a < b) return a;
  if (b < a) return b;
  return (__double2hiint(a) < 0) ? a : b;
}

// Order-preserving map from finite doubles to signed 64-bit integers: a < b  <=>  key(a) < key(b),
// and key(-0.0) = -1 < key(+0.0) = 0 — exactly Math.max / Math.min's ordering — so max/min become
// branch-free integer max/min.  The map is an involution on the low 63 bits.
constexpr long long kKeyLowest = (long long)0x8000000000000000ull;   // below every key
constexpr long long kKeyHighest = (long long)0x7FFFFFFFFFFFFFFFull;  // above every finite key
__device__ __forceinline__ long long ordered_key(double x) {
  const long long b = __double_as_longlong(x);
  return b ^ ((b >> 63) & 0x7FFFFFFFFFFFFFFFLL);
}
__device__ __forceinline__ double from_ordered_key(long long k) {
  return __longlong_as_double(k ^ ((k >> 63) & 0x7FFFFFFFFFFFFFFFLL));
}

__device__ __forceinline__ bool is_finite_f64(double x) {
  return (__double2hiint(x) & 0x7ff00000) != 0x7ff00000;
}

__device__ __forceinline__ double quiet_nan() { return __longlong_as_double(0x7ff8000000000000LL); }

// ---------------------------------------------------------------------------------------------
// ECMA-262 21.4.1.32 date-time string `${date}T${time}` (public/app.js:4122-4124).
// returns 1 parsed (ms in *out), 0 -> NaN (illegal element values; the chain continues),
// -1 unsupported (outside the specified grammar: V8's legacy parser would decide).
__device__ __forceinline__ int two_digits(const uint8_t* s) {
  uint32_t a = s[0] - '0', b = s[1] - '0';
  return (a > 9 || b > 9) ? -1 : (int)(a * 10 + b);
}

__device__ __forceinline__ int64_t days_from_civil(int64_t y, int m, int d) {
  y -= m <= 2;
  const int64_t era = (y >= 0 ? y : y - 399) / 400;
  const int64_t yoe = y - era * 400;
  const int64_t doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
  const int64_t doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
  return era * 146097 + doe - 719468;
}

__device__ inline int parse_show_date_time(const uint8_t* ds, int dn, const uint8_t* ts, int tn, int64_t tz_off_ms,
                                    double* out) {
  const uint8_t dflt[5] = {'0', '0', ':', '0', '0'};
  if (tn == 0) { ts = dflt; tn = 5; }
  if (dn != 10 || ds[4] != '-' || ds[7] != '-') return -1;
  const int y1 = two_digits(ds), y2 = two_digits(ds + 2), mo = two_digits(ds + 5), d = two_digits(ds + 8);
  if (y1 < 0 || y2 < 0 || mo < 0 || d < 0) return -1;
  if (tn < 5 || ts[2] != ':') return -1;
  const int h = two_digits(ts), mi = two_digits(ts + 3);
  if (h < 0 || mi < 0) return -1;
  int pos = 5, sec = 0, ms = 0;
  if (pos < tn && ts[pos] == ':') {
    if (pos + 3 > tn) return -1;
    sec = two_digits(ts + pos + 1);
    if (sec < 0) return -1;
    pos += 3;
    if (pos < tn && ts[pos] == '.') {
      if (pos + 4 > tn) return -1;
      const uint32_t a = ts[pos + 1] - '0', b = ts[pos + 2] - '0', c = ts[pos + 3] - '0';
      if (a > 9 || b > 9 || c > 9) return -1;
      ms = (int)(a * 100 + b * 10 + c);
      pos += 4;
    }
  }
  bool has_off = false;
  int64_t off_ms = 0;
  if (pos < tn) {
    if (ts[pos] == 'Z' && pos + 1 == tn) {
      has_off = true;
    } else if ((ts[pos] == '+' || ts[pos] == '-') && pos + 6 == tn && ts[pos + 3] == ':') {
      const int oh = two_digits(ts + pos + 1), om = two_digits(ts + pos + 4);
      if (oh < 0 || om < 0) return -1;
      if (oh > 23 || om > 59) return 0;
      off_ms = (int64_t)(oh * 60 + om) * 60000 * (ts[pos] == '-' ? -1 : 1);
      has_off = true;
    } else {
      return -1;
    }
  }
  const int year = y1 * 100 + y2;
  if (mo < 1 || mo > 12 || d < 1 || d > 31 || h > 24 || mi > 59 || sec > 59) return 0;
  if (h == 24 && (mi || sec || ms)) return 0;
  const bool leap = (year % 4 == 0) && (year % 100 != 0 || year % 400 == 0);
  const int dim = (mo == 2) ? (leap ? 29 : 28) : ((mo == 4 || mo == 6 || mo == 9 || mo == 11) ? 30 : 31);
  if (d > dim) return -1;  // e.g. Feb 30: engines disagree (V8 rolls over, others NaN)
  const int64_t local = ((days_from_civil(year, mo, d) * 24 + h) * 60 + mi) * 60000 + sec * 1000 + ms;
  *out = (double)(local - (has_off ? off_ms : tz_off_ms));
  return 1;
}

}  // namespace pie
