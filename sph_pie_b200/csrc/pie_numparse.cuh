// JSON number text -> IEEE-754 binary64, correctly rounded (what JSON.parse / StringToNumber give), host + device.
//
//   1. exact cases by one IEEE operation (Clinger): a significand below 2^53 times or over a power of ten that is
//      itself exact (10^22 and below);
//   2. otherwise the Eisel-Lemire algorithm (Lemire, "Number parsing at a gigabyte per second", SPE 2021), restated
//      from the paper: the significand times a 128-bit truncated power of five (pow5_128_table.h, generated from
//      the definitions by scripts/gen_pow5_128.py), one or two 64x64 multiplications, then rounding to even;
//   3. the rare inputs the algorithm cannot decide (a product whose low bits are all ones outside the exact range),
//      and more than 19 significant digits whose truncation matters, are REPORTED (kNumUndecided): the caller
//      fails loudly rather than guess.
// Every function is __host__ __device__ so that the same code is unit-tested on the CPU against Python's float()
// (tests/native/numparse_host.cpp) and runs in the ingest kernels.
#pragma once
#include <stdint.h>

#include "pow5_128_table.h"

#if defined(__CUDACC__)
#define PIE_NP_HD __host__ __device__ __forceinline__
#else
#define PIE_NP_HD inline
#endif

namespace pie {

enum NumParse : int { kNumOk = 0, kNumSyntax = 1, kNumUndecided = 2 };

struct Pow5Table {
  const uint64_t (*t)[2];  // [PIE_POW5_128_N] {high, low}
};

PIE_NP_HD uint64_t np_mul64(uint64_t a, uint64_t b, uint64_t* hi) {
#if defined(__CUDA_ARCH__)
  *hi = __umul64hi(a, b);
  return a * b;
#else
  const unsigned __int128 p = (unsigned __int128)a * b;
  *hi = (uint64_t)(p >> 64);
  return (uint64_t)p;
#endif
}
PIE_NP_HD int np_clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __clzll((long long)x);
#else
  return __builtin_clzll(x);
#endif
}
PIE_NP_HD double np_bits_to_double(uint64_t bits) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)bits);
#else
  double d;
  __builtin_memcpy(&d, &bits, 8);
  return d;
#endif
}

// w * 10^q (w != 0) -> (mantissa field incl. no hidden bit, biased exponent); false if undecided
PIE_NP_HD bool eisel_lemire(uint64_t w, int64_t q, const Pow5Table& tab, uint64_t* out_bits) {
  if (q < PIE_POW5_128_MIN_Q) { *out_bits = 0; return true; }                      // underflows to 0
  if (q > PIE_POW5_128_MAX_Q) { *out_bits = 0x7FF0000000000000ull; return true; }  // overflows to Infinity
  const int lz = np_clz64(w);
  w <<= lz;
  const uint64_t* p5 = tab.t[q - PIE_POW5_128_MIN_Q];
  // 128-bit product of w and the high half of the power; refined with the low half when the 55 bits we need
  // could still be off by one
  uint64_t hi, lo = np_mul64(w, p5[0], &hi);
  const uint64_t precision_mask = 0xFFFFFFFFFFFFFFFFull >> 55;
  if ((hi & precision_mask) == precision_mask) {
    uint64_t hi2;
    np_mul64(w, p5[1], &hi2);
    lo += hi2;
    if (hi2 > lo) ++hi;
  }
  if (lo == 0xFFFFFFFFFFFFFFFFull && !(q >= -27 && q <= 55)) return false;  // the truncated power may have decided it
  const int upperbit = (int)(hi >> 63);
  uint64_t mantissa = hi >> (upperbit + 64 - 52 - 3);
  // binary exponent of the product: floor(log2(10^q)) + 63 (fixed-point constant log2(10) * 2^16), the position of
  // the product's top bit, the shift that normalised w, and the exponent bias
  int64_t power2 = (((152170 + 65536) * q) >> 16) + 63 + upperbit - lz + 1023;
  if (power2 <= 0) {  // subnormal (or zero)
    if (-power2 + 1 >= 64) { *out_bits = 0; return true; }
    mantissa >>= -power2 + 1;
    mantissa += (mantissa & 1);
    mantissa >>= 1;
    const uint64_t e = (mantissa < (1ull << 52)) ? 0 : 1;
    *out_bits = (e << 52) | (mantissa & ((1ull << 52) - 1));
    return true;
  }
  // exactly halfway between two doubles can only happen for small powers of five: then round to even
  if (lo <= 1 && q >= -4 && q <= 23 && (mantissa & 3) == 1) {
    if ((mantissa << (upperbit + 64 - 52 - 3)) == hi) mantissa &= ~1ull;
  }
  mantissa += (mantissa & 1);
  mantissa >>= 1;
  if (mantissa >= (2ull << 52)) {
    mantissa = 1ull << 52;
    ++power2;
  }
  mantissa &= ~(1ull << 52);
  if (power2 >= 0x7FF) { *out_bits = 0x7FF0000000000000ull; return true; }
  *out_bits = ((uint64_t)power2 << 52) | mantissa;
  return true;
}

// Byte source over memory (the ingest kernel has its own, register-buffered one): peek() is the current byte or -1 at
// the end, next() steps over it.
struct MemSource {
  const uint8_t* s;
  int64_t n, i;
  PIE_NP_HD int peek() const { return i < n ? (int)s[i] : -1; }
  PIE_NP_HD void next() { ++i; }
};

PIE_NP_HD bool np_is_digit(int c) { return (unsigned)(c - '0') <= 9u; }

// Parses a JSON number (ECMA-404: -? (0 | [1-9][0-9]*) (. [0-9]+)? ([eE] [+-]? [0-9]+)?) from `src`, which is left on
// the first byte that is not part of the number (the caller checks what follows).  *value receives the correctly
// rounded double.  kCompute = false only checks the grammar (values the caller throws away).
template <bool kCompute, class Src>
PIE_NP_HD int parse_json_number_from(Src& src, const Pow5Table& tab, double* value) {
  bool neg = false;
  if (src.peek() == '-') { neg = true; src.next(); }
  if (!np_is_digit(src.peek())) return kNumSyntax;
  uint64_t w = 0;       // up to 19 significant digits
  int digits = 0;       // significant digits taken into w
  int64_t exp10 = 0;    // value = w * 10^exp10 (before the explicit exponent)
  bool dropped_nonzero = false;
  if (src.peek() == '0') {
    src.next();
    if (np_is_digit(src.peek())) return kNumSyntax;  // no leading zeros
  } else {
    for (int c = src.peek(); np_is_digit(c); src.next(), c = src.peek()) {
      if (!kCompute) continue;
      if (digits < 19) { w = w * 10 + (uint64_t)(c - '0'); ++digits; }
      else { if (exp10 < 1000000) ++exp10; dropped_nonzero |= (c != '0'); }
    }
  }
  if (src.peek() == '.') {
    src.next();
    if (!np_is_digit(src.peek())) return kNumSyntax;
    for (int c = src.peek(); np_is_digit(c); src.next(), c = src.peek()) {
      if (!kCompute) continue;
      if (digits < 19) {
        if (w != 0 || c != '0') { w = w * 10 + (uint64_t)(c - '0'); ++digits; }  // leading zeros are not significant
        if (exp10 > -1000000) --exp10;
      } else {
        dropped_nonzero |= (c != '0');
      }
    }
  }
  if (src.peek() == 'e' || src.peek() == 'E') {
    src.next();
    bool eneg = false;
    if (src.peek() == '+' || src.peek() == '-') { eneg = src.peek() == '-'; src.next(); }
    if (!np_is_digit(src.peek())) return kNumSyntax;
    int64_t e = 0;
    for (int c = src.peek(); np_is_digit(c); src.next(), c = src.peek())
      if (e < 10000000) e = e * 10 + (c - '0');  // saturates: anything this large over- or underflows anyway
    exp10 += eneg ? -e : e;
  }
  if (!kCompute) return kNumOk;
  const uint64_t sign = neg ? 0x8000000000000000ull : 0;
  if (w == 0) { *value = np_bits_to_double(sign); return kNumOk; }
  if (!dropped_nonzero && w < (1ull << 53) && exp10 >= -22 && exp10 <= 22) {
    // one IEEE operation on exact operands is correctly rounded (the powers up to 10^22 are exact doubles)
    double p = 1.0;
    const int64_t a = exp10 < 0 ? -exp10 : exp10;
    if (a & 1) p *= 1e1;
    if (a & 2) p *= 1e2;
    if (a & 4) p *= 1e4;
    if (a & 8) p *= 1e8;
    if (a & 16) p *= 1e16;
    double d = (double)w;
    d = exp10 < 0 ? d / p : d * p;
    *value = neg ? -d : d;
    return kNumOk;
  }
  uint64_t bits;
  if (!eisel_lemire(w, exp10, tab, &bits)) return kNumUndecided;
  if (dropped_nonzero) {  // the true value lies in (w, w+1) * 10^exp10: decided iff both ends round alike
    uint64_t bits2;
    if (!eisel_lemire(w + 1, exp10, tab, &bits2) || bits2 != bits) return kNumUndecided;
  }
  *value = np_bits_to_double(bits | sign);
  return kNumOk;
}

// The same over s[0..n): *used = how many bytes the number took.
PIE_NP_HD int parse_json_number(const uint8_t* s, int64_t n, const Pow5Table& tab, double* value, int64_t* used) {
  MemSource src{s, n, 0};
  const int rc = parse_json_number_from<true>(src, tab, value);
  *used = src.i;
  return rc;
}

}  // namespace pie
