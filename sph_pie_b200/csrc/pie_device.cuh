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

__device__ __forceinline__ bool is_finite_f64(double x) {
  return (__double2hiint(x) & 0x7ff00000) != 0x7ff00000;
}

__device__ __forceinline__ double quiet_nan() { return __longlong_as_double(0x7ff8000000000000LL); }

}  // namespace pie
