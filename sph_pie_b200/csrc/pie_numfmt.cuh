// Number::toString(x) (ECMA-262 6.1.6.1.20, radix 10) for IEEE-754 binary64, host + device.
//
// Digits: the shortest decimal that round-trips, closest to x among the shortest — computed with
// the Ryu algorithm (Ulf Adams, PLDI 2018), restated here from the paper's description:
// interval [vm, vp] around 4*m2 scaled by a power of 10 through 128-bit power-of-5 tables
// (ryu_tables.h, generated from the definitions by scripts/gen_ryu_tables.py), then digits are
// dropped while the interval still identifies x.
// Notation: ECMA-262's rules (fixed for 1e-7 <= |x| < 1e21, exponent form otherwise).
//
// Every function is __host__ __device__ so the same code is unit-tested on the CPU
// (tests/native/numfmt_host.cpp against printf/strtod and Python's repr) and runs in the
// export kernels.  This is used for entry.delaySec in buildTableRow -> csvEscape(String(value))
// (reference server/webhookDispatcher.js:301, :333).
#pragma once
#include <math.h>
#include <stdint.h>

#include "ryu_tables.h"

#if defined(__CUDACC__)
#define PIE_HD __host__ __device__ __forceinline__
#else
#define PIE_HD inline
#endif

namespace pie {

struct RyuTables {
  const uint64_t (*pow5_inv)[2];  // [PIE_RYU_POW5_INV_SPLIT_N]
  const uint64_t (*pow5)[2];      // [PIE_RYU_POW5_SPLIT_N]
};

PIE_HD uint64_t umul128(uint64_t a, uint64_t b, uint64_t* hi) {
#if defined(__CUDA_ARCH__)
  *hi = __umul64hi(a, b);
  return a * b;
#else
  const unsigned __int128 p = (unsigned __int128)a * b;
  *hi = (uint64_t)(p >> 64);
  return (uint64_t)p;
#endif
}

PIE_HD uint64_t shiftright128(uint64_t lo, uint64_t hi, uint32_t dist) {  // 0 < dist < 64
  return (hi << (64 - dist)) | (lo >> dist);
}

// (m * mul) >> j for a 128-bit mul = {lo, hi}, j >= 64, m < 2^55
PIE_HD uint64_t mul_shift_64(uint64_t m, const uint64_t* mul, int32_t j) {
  uint64_t high1, high0;
  const uint64_t low1 = umul128(m, mul[1], &high1);
  umul128(m, mul[0], &high0);
  const uint64_t sum = high0 + low1;
  if (sum < high0) ++high1;
  return shiftright128(sum, high1, (uint32_t)(j - 64));
}

PIE_HD uint32_t pow5_factor(uint64_t v) {
  uint32_t c = 0;
  while (v > 0 && v % 5 == 0) { v /= 5; ++c; }
  return c;
}
PIE_HD bool multiple_of_pow5(uint64_t v, uint32_t p) { return pow5_factor(v) >= p; }
PIE_HD bool multiple_of_pow2(uint64_t v, uint32_t p) { return (v & ((1ull << p) - 1)) == 0; }

PIE_HD int32_t pow5bits(int32_t e) { return (int32_t)(((uint32_t)e * 1217359u) >> 19) + 1; }   // bit length of 5^e
PIE_HD int32_t log10_pow2(int32_t e) { return (int32_t)(((uint32_t)e * 78913u) >> 18); }       // floor(log10(2^e))
PIE_HD int32_t log10_pow5(int32_t e) { return (int32_t)(((uint32_t)e * 732923u) >> 20); }      // floor(log10(5^e))

PIE_HD uint32_t decimal_length9(uint32_t v) {  // v < 10^9
  uint32_t n = 1;
  uint32_t p = 10;
  while (n < 9 && v >= p) { p *= 10; ++n; }
  return n;
}

PIE_HD uint32_t decimal_length17(uint64_t v) {  // v < 10^17
  uint32_t n = 1;
  uint64_t p = 10;
  while (n < 17 && v >= p) { p *= 10; ++n; }
  return n;
}

// Shortest decimal of a finite, non-zero double given as (mantissa field, exponent field):
// |x| = out_mantissa * 10^out_exponent.
PIE_HD void shortest_decimal(uint64_t ieee_mantissa, uint32_t ieee_exponent, const RyuTables& t,
                             uint64_t* out_mantissa, int32_t* out_exponent) {
  // small integers (|x| < 2^53 and integral): exact digits, trailing zeros moved to the exponent
  if (ieee_exponent != 0) {
    const int32_t e2i = (int32_t)ieee_exponent - 1023 - 52;
    if (e2i <= 0 && e2i >= -52) {
      const uint64_t m2i = (1ull << 52) | ieee_mantissa;
      const uint64_t mask = (1ull << -e2i) - 1;
      if ((m2i & mask) == 0) {
        uint64_t m = m2i >> -e2i;
        int32_t e = 0;
        while (m % 10 == 0) { m /= 10; ++e; }
        *out_mantissa = m;
        *out_exponent = e;
        return;
      }
    }
  }
  // short exact decimals (halves, quarters, eighths ... of moderate size): x = odd * 2^-f is exactly the
  // decimal (odd * 5^f) * 10^-f.  With at most 15 significant digits that decimal is THE shortest
  // round-trip string: two different decimals of <= 15 digits never round to the same double
  // (DBL_DIG), so no shorter one can sit in x's rounding interval; and it ends in no zero (odd * 5^f
  // is odd).  Skips the general path's digit-by-digit removal of ~15 trailing zeros.
  if (ieee_exponent != 0) {
    const int32_t e2i = (int32_t)ieee_exponent - 1023 - 52;
    if (e2i < 0 && e2i >= -62) {
      const uint64_t m2i = (1ull << 52) | ieee_mantissa;
#if defined(__CUDA_ARCH__)
      const int32_t tz = __ffsll((long long)m2i) - 1;
#else
      const int32_t tz = __builtin_ctzll(m2i);
#endif
      const int32_t f = -e2i - tz;  // fractional bits of x (> 0: integers were handled above)
      if (f > 0 && f <= 10) {
        uint64_t p5 = 1;
        for (int32_t i = 0; i < f; ++i) p5 *= 5;
        const uint64_t odd = m2i >> tz;
        if (odd <= 999999999999999ull / p5) {
          *out_mantissa = odd * p5;
          *out_exponent = -f;
          return;
        }
      }
    }
  }
  int32_t e2;
  uint64_t m2;
  if (ieee_exponent == 0) {
    e2 = 1 - 1023 - 52 - 2;
    m2 = ieee_mantissa;
  } else {
    e2 = (int32_t)ieee_exponent - 1023 - 52 - 2;
    m2 = (1ull << 52) | ieee_mantissa;
  }
  const bool accept_bounds = (m2 & 1) == 0;  // round-to-even: the interval is closed for even m2
  const uint64_t mv = 4 * m2;
  const uint32_t mm_shift = (ieee_mantissa != 0 || ieee_exponent <= 1) ? 1u : 0u;  // lower gap is half at a power of 2

  uint64_t vr, vp, vm;
  int32_t e10;
  bool vm_trailing_zeros = false, vr_trailing_zeros = false;
  if (e2 >= 0) {
    const int32_t q = log10_pow2(e2) - (e2 > 3);
    e10 = q;
    const int32_t k = PIE_RYU_POW5_INV_BITCOUNT + pow5bits(q) - 1;
    const int32_t i = -e2 + q + k;
    const uint64_t* mul = t.pow5_inv[q];
    vr = mul_shift_64(4 * m2, mul, i);
    vp = mul_shift_64(4 * m2 + 2, mul, i);
    vm = mul_shift_64(4 * m2 - 1 - mm_shift, mul, i);
    if (q <= 21) {  // only then can 10^q divide one of the three numbers
      const uint32_t mv_mod5 = (uint32_t)(mv % 5);
      if (mv_mod5 == 0) vr_trailing_zeros = multiple_of_pow5(mv, (uint32_t)q);
      else if (accept_bounds) vm_trailing_zeros = multiple_of_pow5(mv - 1 - mm_shift, (uint32_t)q);
      else vp -= multiple_of_pow5(mv + 2, (uint32_t)q);
    }
  } else {
    const int32_t q = log10_pow5(-e2) - (-e2 > 1);
    e10 = q + e2;
    const int32_t i = -e2 - q;
    const int32_t k = pow5bits(i) - PIE_RYU_POW5_BITCOUNT;
    const int32_t j = q - k;
    const uint64_t* mul = t.pow5[i];
    vr = mul_shift_64(4 * m2, mul, j);
    vp = mul_shift_64(4 * m2 + 2, mul, j);
    vm = mul_shift_64(4 * m2 - 1 - mm_shift, mul, j);
    if (q <= 1) {
      vr_trailing_zeros = true;  // mv = 4*m2 always has two trailing 0 bits
      if (accept_bounds) vm_trailing_zeros = mm_shift == 1;
      else --vp;
    } else if (q < 63) {
      vr_trailing_zeros = multiple_of_pow2(mv, (uint32_t)q);
    }
  }

  int32_t removed = 0;
  uint8_t last_removed = 0;
  uint64_t output;
  if (vm_trailing_zeros || vr_trailing_zeros) {  // rare: exact ties / exactly representable bounds
    while (vp / 10 > vm / 10) {
      vm_trailing_zeros &= (vm % 10 == 0);
      vr_trailing_zeros &= (last_removed == 0);
      last_removed = (uint8_t)(vr % 10);
      vr /= 10; vp /= 10; vm /= 10;
      ++removed;
    }
    if (vm_trailing_zeros) {
      while (vm % 10 == 0) {
        vr_trailing_zeros &= (last_removed == 0);
        last_removed = (uint8_t)(vr % 10);
        vr /= 10; vp /= 10; vm /= 10;
        ++removed;
      }
    }
    if (vr_trailing_zeros && last_removed == 5 && vr % 2 == 0) last_removed = 4;  // exactly half: round to even
    output = vr + (((vr == vm && (!accept_bounds || !vm_trailing_zeros)) || last_removed >= 5) ? 1 : 0);
  } else {
    bool round_up = false;
    if (vp / 100 > vm / 100) {
      round_up = (vr % 100) >= 50;
      vr /= 100; vp /= 100; vm /= 100;
      removed += 2;
    }
    while (vp / 10 > vm / 10) {
      round_up = (vr % 10) >= 5;
      vr /= 10; vp /= 10; vm /= 10;
      ++removed;
    }
    output = vr + ((vr == vm || round_up) ? 1 : 0);
  }
  *out_mantissa = output;
  *out_exponent = e10 + removed;
}

constexpr int kMaxNumberChars = 32;  // longest Number::toString output is 25 chars

// Writes Number::toString(x) to buf (no terminator) and returns its length (<= 25).
PIE_HD int js_number_to_string(double x, char* buf, const RyuTables& t) {
  uint64_t bits;
#if defined(__CUDA_ARCH__)
  bits = (uint64_t)__double_as_longlong(x);
#else
  __builtin_memcpy(&bits, &x, 8);
#endif
  const bool neg = (bits >> 63) != 0;
  const uint64_t mant = bits & ((1ull << 52) - 1);
  const uint32_t expo = (uint32_t)((bits >> 52) & 0x7FF);
  int n = 0;
  if (expo == 0x7FF) {
    if (mant != 0) { buf[0] = 'N'; buf[1] = 'a'; buf[2] = 'N'; return 3; }
    if (neg) buf[n++] = '-';
    const char inf[8] = {'I', 'n', 'f', 'i', 'n', 'i', 't', 'y'};
    for (int i = 0; i < 8; ++i) buf[n++] = inf[i];
    return n;
  }
  if (expo == 0 && mant == 0) { buf[0] = '0'; return 1; }  // +0 and -0 are both "0"
  uint64_t m = 0;
  int32_t e = 0;
  // Numbers people type — integers and decimals with a few fractional digits (12, 3.5, 0.25, 3.7, 12.34):
  // the smallest d <= 6 with fl(n / 10^d) == |x| for n = rint(|x| * 10^d).  IEEE division is correctly
  // rounded, so the equality says the decimal n * 10^-d reads back as x.  With n < 10^15 it is the ONLY
  // decimal of <= 15 significant digits that does (two such decimals never share a double: DBL_DIG),
  // so it is the shortest round-trip string and the closest — exactly what the general algorithm below
  // would find after stripping up to 16 digits one 64-bit division at a time.  (A true candidate n* is
  // never missed: |x| * 10^d then lies within an ulp of the integer n*, so rint returns it.)
  bool found = false;
  {
    const double ax = neg ? -x : x;
    if (ax < 1e15) {
      double p = 1.0;
      for (int d = 0; d <= 6; ++d) {
        const double scaled = ax * p;
        if (!(scaled < 1e15)) break;
        const double cand = rint(scaled);
        if (cand / p == ax) {
          m = (uint64_t)cand;
          e = -d;
          found = m != 0;
          break;
        }
        p *= 10.0;
      }
    }
  }
  if (!found) shortest_decimal(mant, expo, t, &m, &e);
  // digits of m (< 10^17) through two 32-bit halves: no 64-bit division per digit
  char digits[17];
  const uint32_t hi9 = (uint32_t)(m / 100000000ull);                    // < 10^9
  uint32_t lo8 = (uint32_t)(m - (uint64_t)hi9 * 100000000ull);          // < 10^8
  int k;
  if (hi9 == 0) {
    k = (int)decimal_length9(lo8);
    for (int i = k - 1; i >= 0; --i) { digits[i] = (char)('0' + (lo8 % 10u)); lo8 /= 10u; }
  } else {
    const int kh = (int)decimal_length9(hi9);
    k = kh + 8;
    uint32_t h = hi9;
    for (int i = kh - 1; i >= 0; --i) { digits[i] = (char)('0' + (h % 10u)); h /= 10u; }
    for (int i = k - 1; i >= kh; --i) { digits[i] = (char)('0' + (lo8 % 10u)); lo8 /= 10u; }
  }
  const int pt = k + e;  // value = 0.d1..dk * 10^pt
  if (neg) buf[n++] = '-';
  if (k <= pt && pt <= 21) {  // integer: digits then zeros
    for (int i = 0; i < k; ++i) buf[n++] = digits[i];
    for (int i = k; i < pt; ++i) buf[n++] = '0';
  } else if (0 < pt && pt <= 21) {  // decimal point inside the digits
    for (int i = 0; i < pt; ++i) buf[n++] = digits[i];
    buf[n++] = '.';
    for (int i = pt; i < k; ++i) buf[n++] = digits[i];
  } else if (-6 < pt && pt <= 0) {  // 0.000ddd
    buf[n++] = '0';
    buf[n++] = '.';
    for (int i = 0; i < -pt; ++i) buf[n++] = '0';
    for (int i = 0; i < k; ++i) buf[n++] = digits[i];
  } else {  // d[.ddd]e±x
    int ex = pt - 1;
    buf[n++] = digits[0];
    if (k > 1) {
      buf[n++] = '.';
      for (int i = 1; i < k; ++i) buf[n++] = digits[i];
    }
    buf[n++] = 'e';
    buf[n++] = ex < 0 ? '-' : '+';
    if (ex < 0) ex = -ex;
    if (ex >= 100) buf[n++] = (char)('0' + ex / 100);
    if (ex >= 10) buf[n++] = (char)('0' + (ex / 10) % 10);
    buf[n++] = (char)('0' + ex % 10);
  }
  return n;
}

}  // namespace pie
