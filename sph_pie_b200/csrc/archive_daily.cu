// Daily groups and per-day metric summaries on sm_100a.
// Replaces buildArchiveDailyGroups (reference public/app.js:3401-3443, with getShowTimestamp
// :4092-4116 and parseShowDateTime :4118-4126) and the numeric part of
// getOrCreateGroupMetricSummary (:3445-3502).
//
// Pipeline (all on one stream, no host synchronisation):
//   show_day_kernel        timestamp chain -> local-midnight ms and a 32-bit day key per show
//   radix passes           STABLE LSD radix sort of (day key, show index); 8-bit digits; passes whose
//                          digit is identical in every key are skipped on the device
//   group_* kernels        run heads of equal keys -> group ids / offsets / day starts
//   daily_summary_kernel   one thread per (group, metric): left-to-right sum, min, max, count
// The Map-insertion-then-sort of the reference (groups ordered by day, shows inside a group in
// input order) is exactly a stable sort of the shows by day.
#include "pie_device.cuh"
#include "pie_kernels.h"

namespace pie {

constexpr int kTile = 2048;  // shows per CTA in the sort / group kernels (256 threads x 8)
constexpr int kThreads = 256;
constexpr int kItems = kTile / kThreads;
constexpr uint32_t kKeyNone = 0xFFFFFFFFu;
constexpr int64_t kMsPerDay = 86400000LL;
constexpr double kMaxTimeMs = 8.64e15;  // ECMA-262 TimeClip

struct DailyMeta {
  unsigned long long err;  // (show index << 32) | -status, minimum wins; ~0 = no error
  uint32_t or_bits, and_bits;
  uint32_t n_valid;
  uint32_t unsorted;  // some key is smaller than its predecessor: the radix sort has to run
};

struct DailyScratch {
  uint32_t *keys_a, *keys_b;
  int32_t *vals_a, *vals_b;
  uint32_t* hist;       // [256][nblk]
  uint32_t* tile_heads; // [nblk]
  DailyMeta* meta;
};

static inline int64_t n_tiles(int64_t n) { return (n + kTile - 1) / kTile; }
static inline uint64_t align_up(uint64_t x) { return (x + 255) & ~(uint64_t)255; }

uint64_t daily_scratch_bytes(int64_t n_shows) {
  const uint64_t s = (uint64_t)(n_shows > 0 ? n_shows : 1), nb = (uint64_t)n_tiles(s);
  return 4 * align_up(4 * s) + align_up(4 * 256 * nb) + align_up(4 * nb) + align_up(sizeof(DailyMeta));
}

static DailyScratch carve(void* scratch, int64_t n_shows) {
  const uint64_t s = (uint64_t)(n_shows > 0 ? n_shows : 1), nb = (uint64_t)n_tiles(s);
  uint8_t* p = static_cast<uint8_t*>(scratch);
  DailyScratch d;
  d.keys_a = (uint32_t*)p; p += align_up(4 * s);
  d.keys_b = (uint32_t*)p; p += align_up(4 * s);
  d.vals_a = (int32_t*)p; p += align_up(4 * s);
  d.vals_b = (int32_t*)p; p += align_up(4 * s);
  d.hist = (uint32_t*)p; p += align_up(4 * 256 * nb);
  d.tile_heads = (uint32_t*)p; p += align_up(4 * nb);
  d.meta = (DailyMeta*)p;
  return d;
}

__global__ void daily_init_kernel(DailyMeta* meta) {
  meta->err = ~0ull;
  meta->or_bits = 0;
  meta->and_bits = 0xFFFFFFFFu;
  meta->n_valid = 0;
  meta->unsorted = 0;
}

// getShowTimestamp (:4092-4116) -> local midnight (:3412-3414) -> 32-bit day key.  err is 0 or -pie_status.
__device__ __forceinline__ uint32_t show_day_key(const pie_archive_view& v, int64_t s, int64_t tz_off_ms,
                                                 int64_t* start_out, int* err_out) {
  uint32_t key = kKeyNone;
  double ts = quiet_nan();
  int err = 0;
  const double created = v.created_at[s];
  if (is_finite_f64(created)) {
    ts = created;
  } else {
    int parsed = 0;
    if (v.show_date.offsets) {
      const int db = v.show_date.offsets[s], de = v.show_date.offsets[s + 1];
      if (de > db) {
        int tb = 0, te = 0;
        if (v.show_time.offsets) { tb = v.show_time.offsets[s]; te = v.show_time.offsets[s + 1]; }
        parsed = parse_show_date_time(v.show_date.data + db, de - db, v.show_time.data + tb, te - tb, tz_off_ms, &ts);
        if (parsed < 0) err = -PIE_ERR_UNSUPPORTED_DATE;
      }
    }
    if (parsed == 0) {
      const double archived = v.archived_at ? v.archived_at[s] : quiet_nan();
      if (is_finite_f64(archived)) {
        ts = archived;
      } else if (v.entry_ts) {  // smallest finite entry.ts (:4106-4113)
        bool any = false;
        double best = 0.0;
        for (int e = v.entry_offsets[s]; e < v.entry_offsets[s + 1]; ++e) {
          const double t = v.entry_ts[e];
          if (is_finite_f64(t) && (!any || t < best)) { best = t; any = true; }
        }
        if (any) ts = best;
      }
    }
  }
  int64_t start = PIE_DAY_NONE;
  if (!err && is_finite_f64(ts)) {
    if (fabs(ts) > kMaxTimeMs) {
      err = -PIE_ERR_RANGE;  // new Date(ts) is invalid -> toISOString throws (:3415)
    } else {
      const int64_t t = (int64_t)ts;  // TimeClip truncates toward zero
      const int64_t local = t + tz_off_ms;
      int64_t day = local / kMsPerDay;
      if (local % kMsPerDay < 0) day -= 1;  // floor
      start = day * kMsPerDay - tz_off_ms;
      if (start > (int64_t)kMaxTimeMs || start < -(int64_t)kMaxTimeMs) {
        err = -PIE_ERR_RANGE;
        start = PIE_DAY_NONE;
      } else {
        key = (uint32_t)(day + 0x80000000LL);  // |day| <= 1.0e8 + 1
      }
    }
  }
  *start_out = start;
  *err_out = err;
  return key;
}

__global__ void __launch_bounds__(kThreads) show_day_kernel(pie_archive_view v, int64_t tz_off_ms,
                                                            int64_t* __restrict__ show_day_start,
                                                            uint32_t* __restrict__ keys, int32_t* __restrict__ vals,
                                                            DailyMeta* meta) {
  __shared__ uint32_t s_key[kThreads];
  __shared__ uint32_t s_red[4][kThreads / 32];
  uint32_t k_or = 0, k_and = 0xFFFFFFFFu, n_valid = 0, unsorted = 0;
  unsigned long long first_err = ~0ull;
  const int tid = threadIdx.x;
  for (int64_t base = (int64_t)blockIdx.x * kThreads; base < v.n_shows; base += (int64_t)gridDim.x * kThreads) {
    const int64_t s = base + tid;
    uint32_t key = kKeyNone;
    if (s < v.n_shows) {
      int64_t start;
      int err;
      key = show_day_key(v, s, tz_off_ms, &start, &err);
      if (err) {
        const unsigned long long e = ((unsigned long long)s << 32) | (unsigned long long)err;
        first_err = e < first_err ? e : first_err;
      }
      show_day_start[s] = start;
      keys[s] = key;
      vals[s] = (int32_t)s;
      k_or |= key;
      k_and &= key;
      n_valid += (key != kKeyNone);
    }
    // is the archive already ordered by day?  compare with the predecessor's key
    s_key[tid] = key;
    __syncthreads();
    if (s < v.n_shows && s > 0) {
      uint32_t prev;
      if (tid > 0) {
        prev = s_key[tid - 1];
      } else {  // predecessor belongs to another tile: recompute it (1 in 256 shows)
        int64_t st;
        int er;
        prev = show_day_key(v, s - 1, tz_off_ms, &st, &er);
      }
      unsorted |= (key < prev);
    }
    __syncthreads();
  }
  // one set of atomics per CTA
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    k_or |= __shfl_xor_sync(0xFFFFFFFFu, k_or, o);
    k_and &= __shfl_xor_sync(0xFFFFFFFFu, k_and, o);
    n_valid += __shfl_xor_sync(0xFFFFFFFFu, n_valid, o);
    unsorted |= __shfl_xor_sync(0xFFFFFFFFu, unsorted, o);
    const unsigned long long oe = __shfl_xor_sync(0xFFFFFFFFu, first_err, o);
    first_err = oe < first_err ? oe : first_err;
  }
  if ((tid & 31) == 0) {
    s_red[0][tid >> 5] = k_or; s_red[1][tid >> 5] = k_and; s_red[2][tid >> 5] = n_valid; s_red[3][tid >> 5] = unsorted;
    if (first_err != ~0ull) atomicMin(&meta->err, first_err);  // rare
  }
  __syncthreads();
  if (tid == 0) {
    uint32_t a = 0, b = 0xFFFFFFFFu, c = 0, d = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) { a |= s_red[0][w]; b &= s_red[1][w]; c += s_red[2][w]; d |= s_red[3][w]; }
    atomicOr(&meta->or_bits, a);
    atomicAnd(&meta->and_bits, b);
    if (c) atomicAdd(&meta->n_valid, c);
    if (d) atomicOr(&meta->unsorted, 1u);
  }
}

// ---- stable LSD radix sort, 8-bit digits ----------------------------------------------------
__device__ __forceinline__ bool pass_skipped(const DailyMeta* meta, int pass) {
  if (!meta->unsorted) return true;  // already ordered by day: identity permutation, nothing to do
  return (((meta->or_bits ^ meta->and_bits) >> (8 * pass)) & 0xFFu) == 0;
}
// number of executed passes before `pass`, i.e. which buffer currently holds the data
__device__ __forceinline__ int parity_before(const DailyMeta* meta, int pass) {
  int p = 0;
  for (int q = 0; q < pass; ++q) p ^= pass_skipped(meta, q) ? 0 : 1;
  return p;
}

__global__ void __launch_bounds__(kThreads) radix_hist_kernel(DailyScratch d, int64_t n, int pass, int nblk) {
  if (pass_skipped(d.meta, pass)) return;
  const uint32_t* keys = parity_before(d.meta, pass) ? d.keys_b : d.keys_a;
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kTile;
#pragma unroll
  for (int r = 0; r < kItems; ++r) {
    const int64_t i = base + r * kThreads + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> (8 * pass)) & 0xFF], 1u);
  }
  __syncthreads();
  d.hist[(int64_t)threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of `n` uint32 in place by ONE block of 1024 threads: tiles of 4096 elements, each
// thread 4 consecutive words (coalesced 16-byte accesses when a is 16-byte aligned), running carry.
__device__ void block_exclusive_scan_inplace(uint32_t* a, int64_t n, uint32_t* total) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < n; base += 4096) {
    const int64_t i0 = base + 4 * tid;
    uint32_t x[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = (i0 + j < n) ? a[i0 + j] : 0u;
    const uint32_t sum = x[0] + x[1] + x[2] + x[3];
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    const uint32_t carry = carry_s;
    if (wid == 0) {
      const uint32_t w = warp_sums[lane];
      uint32_t wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, wi, o);
        if (lane >= o) wi += t;
      }
      warp_sums[lane] = wi - w;
      if (lane == 31) carry_s = carry + wi;
    }
    __syncthreads();
    uint32_t run = carry + warp_sums[wid] + incl - sum;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (i0 + j < n) a[i0 + j] = run;
      run += x[j];
    }
    __syncthreads();
  }
  if (total && tid == 0) *total = carry_s;
}

__global__ void __launch_bounds__(1024) radix_scan_kernel(DailyScratch d, int pass, int nblk) {
  if (pass_skipped(d.meta, pass)) return;
  block_exclusive_scan_inplace(d.hist, (int64_t)256 * nblk, nullptr);
}

__global__ void __launch_bounds__(kThreads) radix_scatter_kernel(DailyScratch d, int64_t n, int pass, int nblk) {
  if (pass_skipped(d.meta, pass)) return;
  const int par = parity_before(d.meta, pass);
  const uint32_t* kin = par ? d.keys_b : d.keys_a;
  const int32_t* vin = par ? d.vals_b : d.vals_a;
  uint32_t* kout = par ? d.keys_a : d.keys_b;
  int32_t* vout = par ? d.vals_a : d.vals_b;

  __shared__ uint32_t cnt[kThreads / 32][256];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int i = tid; i < (kThreads / 32) * 256; i += kThreads) (&cnt[0][0])[i] = 0;
  __syncthreads();

  // warp `wid` owns kItems*32 consecutive items of the tile; round r covers 32 consecutive items
  const int64_t wbase = (int64_t)blockIdx.x * kTile + (int64_t)wid * (kItems * 32);
  uint32_t key[kItems];
  int32_t val[kItems];
  uint32_t dig[kItems];
#pragma unroll
  for (int r = 0; r < kItems; ++r) {
    const int64_t i = wbase + r * 32 + lane;
    const bool ok = i < n;
    key[r] = ok ? kin[i] : 0;
    val[r] = ok ? vin[i] : 0;
    dig[r] = ok ? ((key[r] >> (8 * pass)) & 0xFF) : 0x100u;
  }
#pragma unroll
  for (int r = 0; r < kItems; ++r) {
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, dig[r]);
    if (dig[r] < 256 && (__ffs(peers) - 1) == lane) cnt[wid][dig[r]] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  {  // thread tid = digit: turn per-warp counts into global start positions, warps in order
    uint32_t run = d.hist[(int64_t)tid * nblk + blockIdx.x];
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) {
      const uint32_t c = cnt[w][tid];
      cnt[w][tid] = run;
      run += c;
    }
  }
  __syncthreads();
  const uint32_t lt = (1u << lane) - 1;
#pragma unroll
  for (int r = 0; r < kItems; ++r) {
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, dig[r]);
    uint32_t pos = 0;
    if (dig[r] < 256) pos = cnt[wid][dig[r]] + __popc(peers & lt);
    __syncwarp();
    if (dig[r] < 256 && (__ffs(peers) - 1) == lane) cnt[wid][dig[r]] += __popc(peers);
    __syncwarp();
    if (dig[r] < 256) {
      kout[pos] = key[r];
      vout[pos] = val[r];
    }
  }
}

// ---- groups ---------------------------------------------------------------------------------
__device__ __forceinline__ bool is_head(const uint32_t* __restrict__ keys, int64_t i) {
  const uint32_t k = keys[i];
  return k != kKeyNone && (i == 0 || keys[i - 1] != k);
}

__global__ void __launch_bounds__(kThreads) group_count_kernel(DailyScratch d, int64_t n) {
  const uint32_t* keys = parity_before(d.meta, 4) ? d.keys_b : d.keys_a;
  const int64_t base = (int64_t)blockIdx.x * kTile;
  uint32_t c = 0;
#pragma unroll
  for (int r = 0; r < kItems; ++r) {
    const int64_t i = base + r * kThreads + threadIdx.x;
    if (i < n) c += is_head(keys, i);
  }
  __shared__ uint32_t ws[kThreads / 32];
#pragma unroll
  for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) t += ws[w];
    d.tile_heads[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(1024) group_scan_kernel(DailyScratch d, int nblk, pie_daily_out out) {
  __shared__ uint32_t total;
  block_exclusive_scan_inplace(d.tile_heads, nblk, &total);
  __syncthreads();
  if (threadIdx.x == 0) {
    *out.n_groups = (int64_t)total;
    out.group_offsets[total] = (int32_t)d.meta->n_valid;
    const unsigned long long err = d.meta->err;
    out.status[0] = (err == ~0ull) ? 0 : -(int32_t)(err & 0xFFFFFFFFull);
    out.status[1] = (err == ~0ull) ? -1 : (int32_t)(err >> 32);
  }
}

__global__ void __launch_bounds__(kThreads) group_write_kernel(DailyScratch d, int64_t n, pie_daily_out out) {
  const int par = parity_before(d.meta, 4);
  const uint32_t* keys = par ? d.keys_b : d.keys_a;
  const int32_t* vals = par ? d.vals_b : d.vals_a;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int64_t tbase = (int64_t)blockIdx.x * kTile + (int64_t)tid * kItems;  // kItems consecutive per thread
  bool head[kItems];
  uint32_t c = 0;
#pragma unroll
  for (int j = 0; j < kItems; ++j) {
    const int64_t i = tbase + j;
    head[j] = (i < n) && is_head(keys, i);
    c += head[j];
  }
  uint32_t incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= o) incl += t;
  }
  __shared__ uint32_t ws[kThreads / 32];
  if (lane == 31) ws[wid] = incl;
  __syncthreads();
  uint32_t run = d.tile_heads[blockIdx.x] + incl - c;
  for (int w = 0; w < wid; ++w) run += ws[w];
#pragma unroll
  for (int j = 0; j < kItems; ++j) {
    const int64_t i = tbase + j;
    if (i < n) {
      const int32_t s = vals[i];
      out.show_order[i] = s;
      if (head[j]) {
        out.group_offsets[run] = (int32_t)i;
        out.group_day_start[run] = out.show_day_start[s];
        run += 1;
      }
    }
  }
}

// metric m of show s as a Number, NaN when the reference's getValue yields null
// (public/app.js:21-86, :3978-3988); validity is then isValidMetricValue == isfinite (:4128-4134).
// m 0..3 -> int planes TOTAL, COMPLETED, NO_LAUNCH, ABORT; m 4..18 -> f64 planes 0..14.
__device__ __forceinline__ double metric_value(const int32_t* __restrict__ si, const double* __restrict__ sf,
                                               int64_t stride, int m, int64_t s) {
  if (m < 4) return (double)si[(int64_t)m * stride + s];
  return sf[(int64_t)(m - 4) * stride + s];
}

// One thread per (daily group, metric); blockIdx.y is the metric.  A group's sum has to run left to right, so a group
// cannot be split further — but an archive with few days and many shows per day (the replicated sample of the ingest
// bench: 2.6 ms with a thread per group) still gets 19x the threads, and on many small groups the 19 short walks of a
// group run side by side instead of one after the other.
__global__ void __launch_bounds__(128) daily_summary_kernel(const int32_t* __restrict__ si, const double* __restrict__ sf,
                                                            int64_t stats_stride, pie_daily_out out) {
  const int64_t n_groups = *out.n_groups;
  const int m = (int)blockIdx.y;
  const double nan = quiet_nan();
  const int64_t plane = (int64_t)PIE_N_METRICS * out.stride;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += (int64_t)gridDim.x * blockDim.x) {
    const int b = out.group_offsets[g], e = out.group_offsets[g + 1];
    double sum = 0.0;
    long long kmin = kKeyHighest, kmax = kKeyLowest;  // Math.min / Math.max as integer min / max
    int n = 0;
    for (int i = b; i < e; ++i) {
      const double x = metric_value(si, sf, stats_stride, m, out.show_order[i]);
      const bool ok = is_finite_f64(x);
      sum = sum + (ok ? x : -0.0);  // left to right, initial 0 (:3481); -0.0 is an exact no-op here
      const long long k = ordered_key(x);
      kmin = (ok && k < kmin) ? k : kmin;
      kmax = (ok && k > kmax) ? k : kmax;
      n += ok;
    }
    const int64_t o = (int64_t)m * out.stride + g;
    out.summary_f64[PIE_DF_AVERAGE * plane + o] = n ? sum / (double)n : nan;
    out.summary_f64[PIE_DF_MIN * plane + o] = n ? from_ordered_key(kmin) : nan;
    out.summary_f64[PIE_DF_MAX * plane + o] = n ? from_ordered_key(kmax) : nan;
    out.summary_count[o] = n;
  }
}

cudaError_t launch_daily_summary(const pie_archive_view& v, const int32_t* si, const double* sf, int64_t stats_stride,
                                 int32_t tz_offset_minutes, const pie_daily_out& out, void* scratch, int sm_count,
                                 cudaStream_t stream) {
  const int64_t n = v.n_shows;
  DailyScratch d = carve(scratch, n);
  const int nblk = (int)n_tiles(n > 0 ? n : 1);
  daily_init_kernel<<<1, 1, 0, stream>>>(d.meta);
  g_launches += 2 + (n > 0 ? 4 + 12 : 0);  // init, group_scan + (show_day, 12 radix, group_count/write, summary)
  if (n > 0) {
    int64_t day_blocks = (n + kThreads - 1) / kThreads;
    if (day_blocks > (int64_t)sm_count * 8) day_blocks = (int64_t)sm_count * 8;
    show_day_kernel<<<(unsigned)day_blocks, kThreads, 0, stream>>>(
        v, (int64_t)tz_offset_minutes * 60000, out.show_day_start, d.keys_a, d.vals_a, d.meta);
    for (int pass = 0; pass < 4; ++pass) {
      radix_hist_kernel<<<nblk, kThreads, 0, stream>>>(d, n, pass, nblk);
      radix_scan_kernel<<<1, 1024, 0, stream>>>(d, pass, nblk);
      radix_scatter_kernel<<<nblk, kThreads, 0, stream>>>(d, n, pass, nblk);
    }
    group_count_kernel<<<nblk, kThreads, 0, stream>>>(d, n);
  } else {
    cudaMemsetAsync(d.tile_heads, 0, sizeof(uint32_t), stream);
  }
  group_scan_kernel<<<1, 1024, 0, stream>>>(d, n > 0 ? nblk : 1, out);
  if (n > 0) {
    group_write_kernel<<<nblk, kThreads, 0, stream>>>(d, n, out);
    int64_t gx = (n + 127) / 128;  // groups <= shows; the grid cannot know n_groups (it is on the device): stride loop
    if (gx > (int64_t)sm_count * 2) gx = (int64_t)sm_count * 2;
    daily_summary_kernel<<<dim3((unsigned)gx, PIE_N_METRICS), 128, 0, stream>>>(si, sf, stats_stride, out);
  }
  return cudaGetLastError();
}

}  // namespace pie
