// TEST-ONLY stand-in for <cuda_runtime.h>: lets g++ compile the device code of the warp-per-document ingest
// (sph_pie_b200/csrc/pie_json_fast.cuh) for the host, where tests/native/fast_host.cpp runs its 32 lanes as 32 fibers
// that meet at every warp collective.  Nothing in the product includes this.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __restrict__ __restrict
#define __launch_bounds__(...)

struct uint2 { unsigned x, y; };
struct uint3 { unsigned x, y, z; };
struct uint4 { unsigned x, y, z, w; };
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
typedef int cudaError_t;
typedef void* cudaStream_t;

// ---- the warp: 32 fibers, one collective at a time (tests/native/fast_host.cpp) ----------------------------------
extern uint3 threadIdx;                       // set by the scheduler whenever it resumes a lane
extern unsigned long long pie_warp_slot[2][32];  // what the lanes hand in to a collective (double-buffered)
extern int pie_warp_parity[32];
void pie_warp_barrier();                      // returns when every lane has arrived

template <class T>
static inline unsigned long long pie_to_bits(T v) {
  unsigned long long b = 0;
  memcpy(&b, &v, sizeof(T) < 8 ? sizeof(T) : 8);
  return b;
}
template <class T>
static inline T pie_from_bits(unsigned long long b) {
  T v;
  memcpy(&v, &b, sizeof(T));
  return v;
}
// every lane hands in a value, then reads the lanes' values
template <class T>
static inline const unsigned long long* pie_warp_exchange(T v) {
  const int lane = (int)(threadIdx.x & 31);
  const int buf = pie_warp_parity[lane];
  pie_warp_parity[lane] ^= 1;
  pie_warp_slot[buf][lane] = pie_to_bits(v);
  pie_warp_barrier();
  return pie_warp_slot[buf];
}
template <class T>
static inline T __shfl_sync(unsigned, T v, int src) { return pie_from_bits<T>(pie_warp_exchange(v)[src & 31]); }
template <class T>
static inline T __shfl_up_sync(unsigned, T v, unsigned delta) {
  const int lane = (int)(threadIdx.x & 31);
  const unsigned long long* s = pie_warp_exchange(v);
  return lane >= (int)delta ? pie_from_bits<T>(s[lane - delta]) : v;
}
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int mask) {
  const int lane = (int)(threadIdx.x & 31);
  return pie_from_bits<T>(pie_warp_exchange(v)[(lane ^ mask) & 31]);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
  const unsigned long long* s = pie_warp_exchange<unsigned>(pred ? 1u : 0u);
  unsigned m = 0;
  for (int i = 0; i < 32; ++i) m |= (unsigned)(s[i] & 1u) << i;
  return m;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == 0xffffffffu; }
template <class T>
static inline unsigned __match_any_sync(unsigned, T key) {
  const int lane = (int)(threadIdx.x & 31);
  const unsigned long long* s = pie_warp_exchange(key);
  unsigned m = 0;
  for (int i = 0; i < 32; ++i) m |= (unsigned)(s[i] == s[lane]) << i;
  return m;
}
static inline void __syncwarp(unsigned = 0xffffffffu) { pie_warp_barrier(); }

// ---- scalar intrinsics -------------------------------------------------------------------------------------------
template <class T>
static inline T __ldg(const T* p) { return *p; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
  return (unsigned)((((unsigned long long)hi << 32) | lo) >> (sh & 31));
}
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
  const unsigned long long v = ((unsigned long long)b << 32) | a;
  unsigned r = 0;
  for (int i = 0; i < 4; ++i) {
    const unsigned sel = (s >> (4 * i)) & 0xf;
    unsigned byte = (unsigned)(v >> (8 * (sel & 7))) & 0xff;
    if (sel & 8) byte = (byte & 0x80) ? 0xff : 0x00;  // sign replication
    r |= byte << (8 * i);
  }
  return r;
}
static inline long long __double_as_longlong(double d) { long long b; memcpy(&b, &d, 8); return b; }
static inline double __longlong_as_double(long long b) { double d; memcpy(&d, &b, 8); return d; }
static inline int __double2hiint(double d) { return (int)(__double_as_longlong(d) >> 32); }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) {
  return (unsigned long long)(((unsigned __int128)a * b) >> 64);
}
// one OS thread runs all the lanes: plain read-modify-write
static inline unsigned atomicAdd(unsigned* p, unsigned v) { const unsigned o = *p; *p = o + v; return o; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }
static inline unsigned atomicOr(unsigned* p, unsigned v) { const unsigned o = *p; *p = o | v; return o; }
