// JSON ingest: the stored show documents (show_archive.data / shows.data, written by JSON.stringify at
// server/storage/sqlProvider.js:682, :696) parsed on the GPU straight into the columnar archive table — what
// `rows.map(row => this._mapArchiveRow(row)).filter(Boolean)` (sqlProvider.js:230-234, :892-926: JSON.parse, drop
// what is not an object) followed by the projection on the table's schema (columnar.py pack_shows) does.
//
// Two passes with a scan between them, because the caller allocates the table from what pass 1 counted.  Round 2: A WARP
// TAKES A DOCUMENT (pie_json_fast.cuh: structural index by SIMD-in-register byte classes and ballots, member-parallel
// projection) when it is of the shape the provider writes itself, and hands every other document to the walk of round
// 1, in which ONE THREAD WALKS ONE DOCUMENT with a complete ECMA-404 recogniser written as a resumable state machine
// (pie_json_walk.cuh) — the only place that reports errors.  Results are identical whichever takes a document.
//
//   pass 1  ingest_route_kernel                 documents over 16 KB -> the walk's list at once (side stream, see below)
//           ingest_fast_kernel<false, Small>    a warp per document: validates, counts into the document's 104-byte row of
//                                               26 counts, leaves records for pass 2 (where every value is and where it
//                                               goes) in a pool of the scratch area; what it declines -> list
//           ingest_fast_kernel<false, Big>      the same with roomy lists, over that list (documents of 8.5 - 16 KB)
//           ingest_walk_kernel<false>           what both declined (pretty-printed text, entries without a key, ...); and,
//                                               on a stream of the library's own beside the three above, the long documents
//   scan    ingest_scan_*                       exclusive prefix sums of the 26 counts over the documents, in place; totals
//   pass 2  ingest_fast_kernel<true, Small>     documents with records: a scatter (no index, no keys, no grammar; numbers
//                                               are converted here); without (pool full, the last document): parsed again
//           ingest_fast_kernel<true, Big>       the latter for the roomy configuration's documents
//           ingest_walk_kernel<true>            the walk's documents, the same walk with every counter started at its prefix
//
// With pie_debug_ingest_warp_path(0) the walk takes everything, as in round 1: documents handed out by LENGTH CLASS
// (ingest_order_*: a counting sort; lanes that start documents of one length together walk them roughly in step), one
// 96-byte row per entry in pass 2 and ingest_rows_to_columns to turn the rows into the entry columns.
//
// Restrictions that fail loudly (pie_status in the status word, first offending document): a text field that is not
// a string / null, delaySec that is not a number / null, a string with a lone surrogate escape (pack_shows raises
// TypeError for these: the table holds provider-normalised documents); a known key twice in one object
// (JSON.stringify never writes that; JSON.parse would keep the last); nesting deeper than 64; bytes that are not
// UTF-8; a number the Eisel-Lemire parser cannot decide (> 19 significant digits on a rounding boundary).
#include "pie_device.cuh"
#include "pie_kernels.h"
#include "pie_json_walk.cuh"
#include "pie_json_fast.cuh"

#include <cstdlib>
#include <mutex>

namespace pie {

namespace {

using namespace jw;

__device__ const uint64_t g_pow5_dev[PIE_POW5_128_N][2] = PIE_POW5_128_INIT;

constexpr jf::KeyTables key_tables_checked() {
  bool clash = false;
  return jf::make_key_tables(&clash);
}
__device__ const jf::KeyTables g_key_tables = key_tables_checked();

constexpr int kIngestThreads = 128;

struct IngestScratch {
  uint32_t* planes;          // [n_docs][kPlanes]: a document's 26 counts are one 104-byte row (pass 1 writes it, the
  int64_t stride;            //   scans turn it into exclusive prefixes in place, pass 2 reads it) — not 26 scattered words
  unsigned long long* block_sums;  // [kPlanes][nblk] + the arrival counter of the scan
  unsigned long long* err_key;     // min over documents of (doc << 8 | code): the first hard error
  unsigned long long* next_doc;    // [2] the next place of `order` to hand out in pass 1 / pass 2
  uint32_t* buckets;               // [kOrderBuckets + 1] documents per length class, then where each class starts
  int32_t* order;                  // [n_docs] the documents by length class; with the warp path: the documents it declined
  // the warp-cooperative path (pie_json_fast.cuh)
  unsigned long long* next_fast;   // [2] the next document a warp takes in pass 1 / pass 2
  uint32_t* n_slow;                // documents on the list of the thread-per-document walk
  int32_t* order_long;             // [n_docs] documents too long for the warp path: the walk takes them on a second stream
  uint32_t* n_long;                //   WHILE the warp path takes the others (a long document is ~1 ms per KB on one lane)
  unsigned long long* next_long;   // [2]
  int32_t* order_big;              // [n_docs] what the roomy configuration of the warp path declined as well: the walk's list
  uint32_t* n_slow_big;
  unsigned long long* next_big;    // [2] the next place of `order` the roomy configuration takes in pass 1 / pass 2
  uint8_t* route;                  // [n_docs] jf::kRouteFast / kRouteRecords / kRouteFastBig: pass 1 accepted the document on the warp path
  jf::RecCtx rec;                  // the records pass 1 leaves for pass 2 (pie_json_fast.cuh)
};

// Documents are handed to the warps in order of length (to the byte, up to 16 KB; longer ones share a class), 32 neighbours of that order at a time:
// documents of one length are almost always documents of one make (same number of entries, same keys), and 32 lanes
// that start such documents together walk them in step — the same tokens in the same turns — instead of each lane
// paying for the union of 32 unrelated paths.  Measured: 66 -> 51 ms per 2^20 documents.
constexpr int kOrderBuckets = 16384;
constexpr int kOrderShift = 0;

__device__ __forceinline__ uint32_t length_class(const int64_t* __restrict__ doc_offsets, int64_t s) {
  const int64_t c = (doc_offsets[s + 1] - doc_offsets[s]) >> kOrderShift;
  return (uint32_t)(c < kOrderBuckets - 1 ? (c < 0 ? 0 : c) : kOrderBuckets - 1);
}

__global__ void __launch_bounds__(256) ingest_order_count_kernel(const int64_t* __restrict__ doc_offsets, int64_t n_docs,
                                                                 IngestScratch sc) {
  const int64_t s = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (s < n_docs) atomicAdd(sc.buckets + length_class(doc_offsets, s), 1u);
}

// counts -> where each class starts (exclusive scan of kOrderBuckets counters by one CTA of 1024 threads)
__global__ void __launch_bounds__(1024) ingest_order_scan_kernel(IngestScratch sc) {
  __shared__ uint32_t warp_sums[32];
  constexpr int kPer = kOrderBuckets / 1024;
  uint32_t item[kPer], sum = 0;
#pragma unroll
  for (int k = 0; k < kPer; ++k) {
    item[k] = sc.buckets[threadIdx.x * kPer + k];
    sum += item[k];
  }
  uint32_t v = sum;
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) >= d) v += t;
  }
  if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    uint32_t ws = warp_sums[threadIdx.x];
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, ws, d);
      if (threadIdx.x >= d) ws += t;
    }
    warp_sums[threadIdx.x] = ws;
  }
  __syncthreads();
  uint32_t run = ((threadIdx.x >> 5) ? warp_sums[(threadIdx.x >> 5) - 1] : 0) + v - sum;
#pragma unroll
  for (int k = 0; k < kPer; ++k) {
    sc.buckets[threadIdx.x * kPer + k] = run;
    run += item[k];
  }
}

__global__ void __launch_bounds__(256) ingest_order_place_kernel(const int64_t* __restrict__ doc_offsets, int64_t n_docs,
                                                                 IngestScratch sc) {
  const int64_t s = (int64_t)blockIdx.x * 256 + threadIdx.x;
  // longest first: the documents drawn last are the short ones, so the kernel does not end on a few lanes still walking
  // the longest documents
  if (s < n_docs) sc.order[n_docs - 1 - atomicAdd(sc.buckets + length_class(doc_offsets, s), 1u)] = (int32_t)s;
}

constexpr int kScanThreads = 1024;
constexpr int kScanChunk = kScanThreads;  // documents per CTA of the scans: a thread holds one document's row

__global__ void ingest_init_kernel(IngestScratch sc) {
  *sc.err_key = ~0ull;
  sc.next_doc[0] = 0;
  sc.next_doc[1] = 0;
  sc.next_fast[0] = 0;
  sc.next_fast[1] = 0;
  *sc.n_slow = 0;
  *sc.n_slow_big = 0;
  *sc.n_long = 0;
  sc.next_big[0] = 0;
  sc.next_big[1] = 0;
  sc.next_long[0] = 0;
  sc.next_long[1] = 0;
  *sc.rec.cursor = 0;
}

// Which documents are too long for the warp path is known from their lengths alone: they go on a list of their own,
// which the walk takes on a second stream while the warp path works (every other route starts as "declined").
__global__ void __launch_bounds__(256) ingest_route_kernel(const int64_t* __restrict__ doc_offsets, int64_t n_docs, IngestScratch sc) {
  const int64_t s = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (s >= n_docs) return;
  const bool is_long = doc_offsets[s + 1] - doc_offsets[s] + 31 > jf::kFastMaxBytes;  // whatever its alignment
  sc.route[s] = is_long ? jf::kRouteLong : jf::kRouteSlow;
  if (is_long) sc.order_long[atomicAdd(sc.n_long, 1u)] = (int32_t)s;
}

// The warp-cooperative path: every warp takes documents in table order (neighbours in time write neighbouring parts
// of every heap) until none is left.  Pass 1 hands what it declines to a list: that of the roomy configuration
// (kFromList = false -> sc.order), which hands what it declines as well to the thread-per-document walk (sc.order_big).
template <bool kFill, class Caps, bool kFromList>
__global__ void __launch_bounds__(Caps::kWarps * 32) ingest_fast_kernel(const int64_t* __restrict__ doc_offsets,
                                                                        const uint8_t* __restrict__ text, int64_t n_docs,
                                                                        IngestScratch sc, uint8_t* __restrict__ doc_status,
                                                                        IngestOut out) {
  extern __shared__ __align__(16) unsigned char fast_smem[];
  jf::TablePointers& tp = *reinterpret_cast<jf::TablePointers*>(fast_smem);
  jf::WarpShared<Caps>& ws =
      reinterpret_cast<jf::WarpShared<Caps>*>(fast_smem + ((sizeof(jf::TablePointers) + 15) & ~(size_t)15))[threadIdx.x >> 5];
  if (kFill) {
    for (int h = threadIdx.x; h < kHeaps; h += blockDim.x) {
      tp.data[h] = out.data[h];
      tp.off[h] = out.off[h];
    }
  }
  for (int i = threadIdx.x; i < (int)(sizeof(jf::KeyTables) / 8); i += blockDim.x)
    reinterpret_cast<unsigned long long*>(&tp.keys)[i] = reinterpret_cast<const unsigned long long*>(&g_key_tables)[i];
  __syncthreads();
  const Pow5Table pow5{g_pow5_dev};
  const int lane = threadIdx.x & 31;
  const int64_t n_take = kFromList ? (int64_t)*sc.n_slow : n_docs;
  unsigned long long* next = (kFromList ? sc.next_big : sc.next_fast) + (kFill ? 1 : 0);
  // (Pass 2 drawing a document ahead and asking L2 for its text and records was measured: the same 8.3 ms, and 4.2 GB
  // more DRAM reads per 2^20 documents — the prefetched lines are gone again before they are used.  Not shipped.)
  for (;;) {
    unsigned long long drawn = 0;
    if (lane == 0) drawn = atomicAdd(next, 1ull);
    drawn = __shfl_sync(0xffffffffu, drawn, 0);
    if ((int64_t)drawn >= n_take) break;
    const int64_t s = kFromList ? (int64_t)sc.order[drawn] : (int64_t)drawn;
    if (!kFill && !kFromList && sc.route[s] == jf::kRouteLong) continue;  // the walk has it already
    if (kFill) {
      const uint8_t route = sc.route[s];
      if (kFromList) {  // the roomy configuration parses again what only it could take and the pool had no room for
        if (route == jf::kRouteFastBig)
          jf::fast_doc<true, Caps>(ws, tp, text, doc_offsets[s], doc_offsets[s + 1], s, n_docs, sc.planes + s * kPlanes, out, pow5, sc.rec);
      } else if (route == jf::kRouteRecords) {
        jf::fill_records(ws, tp, text, doc_offsets[s], doc_offsets[s + 1], s, sc.planes + s * kPlanes, out, sc.rec, pow5);
      } else if (route == jf::kRouteFast) {
        jf::fast_doc<true, Caps>(ws, tp, text, doc_offsets[s], doc_offsets[s + 1], s, n_docs, sc.planes + s * kPlanes, out, pow5, sc.rec);
      }
    } else {
      int route = jf::fast_doc<false, Caps>(ws, tp, text, doc_offsets[s], doc_offsets[s + 1], s, n_docs, sc.planes + s * kPlanes, out,
                                            pow5, sc.rec);
      if (kFromList && route == jf::kRouteFast) route = jf::kRouteFastBig;
      if (lane == 0) {
        sc.route[s] = (uint8_t)route;
        if (route != jf::kRouteSlow) doc_status[s] = 0;
        else if (kFromList) sc.order_big[atomicAdd(sc.n_slow_big, 1u)] = (int32_t)s;
        else sc.order[atomicAdd(sc.n_slow, 1u)] = (int32_t)s;
      }
    }
  }
}

// Both passes: every lane walks documents until none is left.
template <bool kFill>
__global__ void __launch_bounds__(kIngestThreads, kFill ? 8 : 10) ingest_walk_kernel(const int64_t* __restrict__ doc_offsets,
                                                                      const uint8_t* __restrict__ text, int64_t n_docs,
                                                                      IngestScratch sc, uint8_t* __restrict__ doc_status,
                                                                      IngestOut out, const int32_t* __restrict__ order,
                                                                      const uint32_t* __restrict__ list_len,
                                                                      unsigned long long* __restrict__ next) {
  const Pow5Table pow5{g_pow5_dev};
  // list_len: `order` is a list of some documents (in no particular order) — what the warp path declined, or the long
  // documents; nullptr: all the documents by length class (the walk alone)
  const int64_t n_order = list_len ? (int64_t)*list_len : n_docs;
  uint32_t cnt[kPlanes];
  DocWalker<kFill> w;
  bool active = false, exhausted = false;
  int64_t s = 0;
  for (;;) {
    int r = kDocRunning;
    // the warp (all its lanes are between documents here) draws 32 places of the order at once
    unsigned long long base = 0;
    if ((threadIdx.x & 31) == 0) base = atomicAdd(next, 32ull);
    base = __shfl_sync(0xffffffffu, base, 0);
    const int64_t drawn = (int64_t)base + (threadIdx.x & 31);
    if (!exhausted) {
      s = drawn < n_order ? order[drawn] : n_docs;
      if (s >= n_docs) {
        exhausted = true;
      } else {
        bool walk = true;
        if (kFill) {
#pragma unroll
          for (int p = 0; p < kPlanes; p += 2) {
            const uint2 two = *reinterpret_cast<const uint2*>(sc.planes + s * kPlanes + p);
            cnt[p] = two.x;
            cnt[p + 1] = two.y;
          }
          // the show's row: every text field starts where the previous show's ended (an absent key is '')
#pragma unroll
          for (int h = 0; h < 7; ++h) out.off[h][s] = (int32_t)cnt[h];
          out.entry_offsets[s] = (int32_t)cnt[kPlaneEntries];
          out.crew_list[s] = (int32_t)cnt[kPlaneCrewItems];
          out.created_at[s] = jw_nan();
          out.archived_at[s] = jw_nan();
          if (out.time_val[PIE_TF_UPDATED]) out.time_val[PIE_TF_UPDATED][s] = jw_nan();
          if (out.time_val[PIE_TF_DELETED]) out.time_val[PIE_TF_DELETED][s] = jw_nan();
          if (out.time_kind) *reinterpret_cast<uint32_t*>(out.time_kind + s * PIE_TF_COUNT) = 0u;  // four PIE_TK_ABSENT
          walk = doc_status[s] == 0;  // a dropped row stays the empty show
        } else {
#pragma unroll
          for (int p = 0; p < kPlanes; ++p) cnt[p] = 0;
        }
        if (walk) {
          w.begin(text, doc_offsets[s], doc_offsets[s + 1], s);
          active = true;
        } else {
          r = kDocDropped;
        }
      }
    }
    // a whole document per turn (measured against warp votes on the token kind, warp-uniform string loops and
    // token- / member-sized turns in profiles/ncu_r01_ingest_summary.md: the fewer turns, the faster)
    if (__all_sync(0xffffffffu, exhausted)) break;
    if (active) {
#pragma unroll 1
      do { r = w.step_member(cnt, out, pow5); } while (r == kDocRunning);
    }
    __syncwarp();
    if (r == kDocRunning) continue;
    active = false;
    if (!kFill) {
      if (r != kDocOk) {
#pragma unroll
        for (int p = 0; p < kPlanes; ++p) cnt[p] = 0;
        if (r != kDocDropped) atomicMin(sc.err_key, ((unsigned long long)s << 8) | (unsigned long long)r);
      }
      doc_status[s] = r == kDocOk ? 0 : 1;
#pragma unroll
      for (int p = 0; p < kPlanes; p += 2) *reinterpret_cast<uint2*>(sc.planes + s * kPlanes + p) = make_uint2(cnt[p], cnt[p + 1]);
    } else if (s == n_docs - 1) {  // the terminal offsets: where the last document ended
#pragma unroll
      for (int h = 0; h < 7; ++h) out.off[h][n_docs] = (int32_t)cnt[h];
      out.entry_offsets[n_docs] = (int32_t)cnt[kPlaneEntries];
      out.crew_list[n_docs] = (int32_t)cnt[kPlaneCrewItems];
      out.off[kHeapCrew][cnt[kPlaneCrewItems]] = (int32_t)cnt[kHeapCrew];
      const uint32_t rows = cnt[kPlaneEntries];
#pragma unroll
      for (int h = kHeapEntry0; h < kHeapEntry0 + 14; ++h) out.off[h][rows] = (int32_t)cnt[h];
      out.actions_list[rows] = (int32_t)cnt[kPlaneActionItems];
      out.off[kHeapActions][cnt[kPlaneActionItems]] = (int32_t)cnt[kHeapActions];
    }
  }
}

// entry rows -> the table's entry columns: 256 rows per CTA through shared memory, every access coalesced
__global__ void __launch_bounds__(256) ingest_rows_to_columns_kernel(const EntryRow* __restrict__ rows, int64_t n_entries,
                                                                     IngestOut out) {
  constexpr int kWords = sizeof(EntryRow) / 4, kPitch = kWords + 1;  // 24 words a row; 25 in shared memory: no bank conflicts
  __shared__ uint32_t tile[256 * kPitch];
  const int64_t first = (int64_t)blockIdx.x * 256;
  const int n = n_entries - first < 256 ? (int)(n_entries - first) : 256;
  const uint32_t* src = reinterpret_cast<const uint32_t*>(rows + first);
  for (int i = threadIdx.x; i < n * kWords; i += 256) tile[(i / kWords) * kPitch + i % kWords] = src[i];
  __syncthreads();
  if ((int)threadIdx.x >= n) return;
  const uint32_t* r = tile + threadIdx.x * kPitch;
  const int64_t e = first + threadIdx.x;
#pragma unroll
  for (int h = 0; h < 14; ++h) out.off[kHeapEntry0 + h][e] = (int32_t)r[h];
  out.actions_list[e] = (int32_t)r[14];
  out.delay_sec[e] = __longlong_as_double((long long)(((unsigned long long)r[17] << 32) | r[16]));
  out.entry_ts[e] = __longlong_as_double((long long)(((unsigned long long)r[19] << 32) | r[18]));
  out.delay_valid[e] = (uint8_t)r[20];
}

// the terminal offsets of an empty table
__global__ void ingest_empty_kernel(IngestOut out) {
  for (int h = 0; h < kHeaps; ++h) out.off[h][0] = 0;
  out.entry_offsets[0] = 0;
  out.crew_list[0] = 0;
  out.actions_list[0] = 0;
}

// exclusive scan of every count over the documents: sums of 1024 documents per CTA, scan of the CTA sums (one CTA
// per count), then the documents in place.  A thread holds its document's whole row.
__device__ __forceinline__ void load_row(const IngestScratch& sc, int64_t s, int64_t n, uint32_t (&v)[kPlanes]) {
#pragma unroll
  for (int p = 0; p < kPlanes; p += 2) {
    const uint2 two = s < n ? *reinterpret_cast<const uint2*>(sc.planes + s * kPlanes + p) : make_uint2(0u, 0u);
    v[p] = two.x;
    v[p + 1] = two.y;
  }
}

__global__ void __launch_bounds__(kScanThreads) ingest_scan_sums_kernel(IngestScratch sc, int64_t n, int nblk) {
  __shared__ unsigned long long warp_sums[kPlanes][32];
  const int64_t s = (int64_t)blockIdx.x * kScanChunk + threadIdx.x;
  uint32_t v[kPlanes];
  load_row(sc, s, n, v);
#pragma unroll
  for (int p = 0; p < kPlanes; ++p) {
    unsigned long long x = v[p];
    for (int d = 16; d > 0; d >>= 1) x += __shfl_down_sync(0xffffffffu, x, d);
    if ((threadIdx.x & 31) == 0) warp_sums[p][threadIdx.x >> 5] = x;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    for (int p = 0; p < kPlanes; ++p) {
      unsigned long long x = warp_sums[p][threadIdx.x];
      for (int d = 16; d > 0; d >>= 1) x += __shfl_down_sync(0xffffffffu, x, d);
      if (threadIdx.x == 0) sc.block_sums[(int64_t)p * nblk + blockIdx.x] = x;
    }
  }
}

__global__ void __launch_bounds__(1024) ingest_scan_blocks_kernel(IngestScratch sc, int nblk, int64_t* __restrict__ totals,
                                                                   int32_t* __restrict__ status) {
  const int p = blockIdx.x;
  unsigned long long* b = sc.block_sums + (int64_t)p * nblk;
  __shared__ unsigned long long warp_sums[32];
  __shared__ unsigned long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblk; base += 1024) {
    const int i = base + threadIdx.x;
    const unsigned long long x = i < nblk ? b[i] : 0;
    unsigned long long v = x;
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, v, d);
      if ((threadIdx.x & 31) >= d) v += t;
    }
    if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
      unsigned long long ws = warp_sums[threadIdx.x];
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, ws, d);
        if (threadIdx.x >= d) ws += t;
      }
      warp_sums[threadIdx.x] = ws;
    }
    __syncthreads();
    const unsigned long long before = carry + ((threadIdx.x >> 5) ? warp_sums[(threadIdx.x >> 5) - 1] : 0) + v - x;
    if (i < nblk) b[i] = before;
    __syncthreads();
    if (threadIdx.x == 1023) carry = before + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    totals[p] = (int64_t)carry;
    // offsets are int32: a heap or a row count of 2 GiB and more does not fit (split the batch)
    if (carry > 0x7fffffffull) atomicMin(sc.err_key, (0xffffffffffull << 8) | (unsigned long long)(-PIE_ERR_CAPACITY));
  }
  // the status word, once every plane is through (the last CTA to get here writes it)
  __shared__ bool last;
  if (threadIdx.x == 0) {
    __threadfence();
    // block_sums[kPlanes * nblk] doubles as the arrival counter
    last = atomicAdd(sc.block_sums + (int64_t)kPlanes * nblk, 1ull) == (unsigned long long)(gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    const unsigned long long k = atomicAdd(sc.err_key, 0ull);
    if (k == ~0ull) {
      status[0] = 0;
      status[1] = -1;
    } else {
      status[0] = -(int32_t)(k & 0xff);
      const unsigned long long doc = k >> 8;
      status[1] = doc >= 0x7fffffffull ? -1 : (int32_t)doc;
    }
  }
}

__global__ void __launch_bounds__(kScanThreads) ingest_scan_apply_kernel(IngestScratch sc, int64_t n, int nblk) {
  __shared__ uint32_t warp_sums[kPlanes][32];
  const int64_t s = (int64_t)blockIdx.x * kScanChunk + threadIdx.x;
  uint32_t v[kPlanes], incl[kPlanes];
  load_row(sc, s, n, v);
#pragma unroll
  for (int p = 0; p < kPlanes; ++p) {
    uint32_t x = v[p];
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, x, d);
      if ((threadIdx.x & 31) >= d) x += t;
    }
    incl[p] = x;
    if ((threadIdx.x & 31) == 31) warp_sums[p][threadIdx.x >> 5] = x;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    for (int p = 0; p < kPlanes; ++p) {
      uint32_t ws = warp_sums[p][threadIdx.x];
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, ws, d);
        if (threadIdx.x >= d) ws += t;
      }
      warp_sums[p][threadIdx.x] = ws;
    }
  }
  __syncthreads();
  if (s >= n) return;
#pragma unroll
  for (int p = 0; p < kPlanes; p += 2) {
    uint32_t two[2];
#pragma unroll
    for (int k = 0; k < 2; ++k)
      two[k] = (uint32_t)sc.block_sums[(int64_t)(p + k) * nblk + blockIdx.x] +
               ((threadIdx.x >> 5) ? warp_sums[p + k][(threadIdx.x >> 5) - 1] : 0) + incl[p + k] - v[p + k];
    *reinterpret_cast<uint2*>(sc.planes + s * kPlanes + p) = make_uint2(two[0], two[1]);
  }
}

// persistent: as many CTAs as can be resident (16 per SM at most), fewer for small batches.  (Giving a small batch
// fewer lanes so that each takes several documents was measured and is slower: 2^17 documents 21.9 ms instead of
// ~14 — a lane needs ~3.5 ms per 4 KB document, so a launch is latency-bound until every SM is full.)
unsigned walk_blocks(int64_t n_docs) {
  const int64_t want = (n_docs + kIngestThreads - 1) / kIngestThreads;
  const int64_t cap = (int64_t)sm_count_or_default() * 16;
  return (unsigned)(want < cap ? want : cap);
}

int scan_blocks(int64_t n_docs) { return (int)((n_docs + kScanChunk - 1) / kScanChunk) + (n_docs == 0 ? 1 : 0); }

IngestScratch carve(void* scratch, int64_t n_docs) {
  IngestScratch sc;
  const int64_t stride = ((n_docs > 0 ? n_docs : 1) + 31) & ~(int64_t)31;
  uint8_t* p = (uint8_t*)scratch;
  sc.planes = (uint32_t*)p;
  sc.stride = stride;
  p += (uint64_t)kPlanes * stride * 4;
  sc.block_sums = (unsigned long long*)p;
  p += ((uint64_t)kPlanes * scan_blocks(n_docs) + 1) * 8;
  sc.err_key = (unsigned long long*)p;
  sc.next_doc = sc.err_key + 1;
  p = (uint8_t*)(sc.next_doc + 2);
  sc.buckets = (uint32_t*)p;
  p += 4 * (kOrderBuckets + 2);
  sc.order = (int32_t*)p;
  p += 4 * (uint64_t)stride;
  sc.next_fast = (unsigned long long*)p;
  sc.n_slow = (uint32_t*)(sc.next_fast + 2);
  sc.rec.cursor = sc.next_fast + 3;
  sc.next_big = sc.next_fast + 4;
  sc.n_slow_big = (uint32_t*)(sc.next_fast + 6);
  sc.n_long = (uint32_t*)(sc.next_fast + 7);
  sc.next_long = sc.next_fast + 8;
  p += 96;
  sc.order_big = (int32_t*)p;
  p += 4 * (uint64_t)stride;
  sc.order_long = (int32_t*)p;
  p += 4 * (uint64_t)stride;
  sc.route = p;
  p += (uint64_t)stride;  // a multiple of 32
  sc.rec.doc_rec = (jf::DocRec*)p;
  p += sizeof(jf::DocRec) * (uint64_t)stride;
  sc.rec.pool = (unsigned long long*)p;
  sc.rec.capacity = (unsigned long long)jf::kPoolUnitsPerDoc * (uint64_t)stride;
  return sc;
}

// PIE_INGEST_WARP_PATH=0 keeps every document on the thread-per-document walk (the round-1 pipeline: A/B timing; it is
// the faster one only when a warp's 32 documents are copies of one another, profiles/ncu_r02_ingest_summary.md)
std::atomic<int> g_warp_path{-1};  // -1: not decided yet (the environment decides)
bool warp_path_enabled() {
  int v = g_warp_path.load();
  if (v < 0) {
    const char* e = std::getenv("PIE_INGEST_WARP_PATH");
    v = (e && e[0] == '0') ? 0 : 1;
    g_warp_path.store(v);
  }
  return v != 0;
}
template <class Caps>
constexpr size_t fast_smem_bytes() {
  return ((sizeof(jf::TablePointers) + 15) & ~(size_t)15) + sizeof(jf::WarpShared<Caps>) * Caps::kWarps;
}

// one launch of a configuration of the warp path: as many CTAs as can be resident, fewer for small batches
template <bool kFill, class Caps, bool kFromList>
cudaError_t launch_fast(const pie_json_docs& docs, const IngestScratch& sc, uint8_t* doc_status, const IngestOut& out,
                        cudaStream_t stream) {
  static int per_sm = -1;
  constexpr size_t smem = fast_smem_bytes<Caps>();
  if (per_sm < 0) {
    cudaError_t e = cudaFuncSetAttribute(ingest_fast_kernel<kFill, Caps, kFromList>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int n = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, ingest_fast_kernel<kFill, Caps, kFromList>, Caps::kWarps * 32, smem);
    if (e != cudaSuccess) return e;
    per_sm = n > 0 ? n : 1;
  }
  static const int cap_env = [] {  // experiment knob: fewer resident CTAs per SM (PIE_INGEST_FAST_CTAS)
    const char* e = std::getenv("PIE_INGEST_FAST_CTAS");
    return e ? std::atoi(e) : 0;
  }();
  const int ctas = cap_env > 0 && cap_env < per_sm ? cap_env : per_sm;
  // the roomy configuration does not know how long its list is when it is launched: half a wave of it (it exits at
  // once when the list is empty, which is the usual case)
  const int64_t want = kFromList ? (int64_t)sm_count_or_default() * 2 : (docs.n_docs + Caps::kWarps - 1) / Caps::kWarps;
  const int64_t cap = (int64_t)sm_count_or_default() * ctas;
  const unsigned grid = (unsigned)(want < cap ? (want > 0 ? want : 1) : cap);
  ingest_fast_kernel<kFill, Caps, kFromList><<<grid, Caps::kWarps * 32, smem, stream>>>(docs.offsets, docs.data, docs.n_docs, sc,
                                                                                        doc_status, out);
  ++g_launches;
  return cudaGetLastError();
}

IngestOut make_out(const pie_archive_table& t) {
  IngestOut o;
  o.rows = nullptr;
  const pie_strcol_mut* show_cols[7] = {&t.show_id, &t.show_date, &t.show_time, &t.show_label, &t.lead_pilot, &t.monkey_lead,
                                        &t.show_notes};
  const pie_strcol_mut* entry_cols[14] = {&t.entry_id, &t.unit_id, &t.planned, &t.launched, &t.status, &t.primary_issue,
                                          &t.sub_issue, &t.other_detail, &t.severity, &t.root_cause, &t.operator_name,
                                          &t.battery_id, &t.command_rx, &t.notes};
  for (int h = 0; h < 7; ++h) { o.off[h] = show_cols[h]->offsets; o.data[h] = show_cols[h]->data; }
  o.off[kHeapCrew] = t.crew.items.offsets;
  o.data[kHeapCrew] = t.crew.items.data;
  for (int h = 0; h < 14; ++h) { o.off[kHeapEntry0 + h] = entry_cols[h]->offsets; o.data[kHeapEntry0 + h] = entry_cols[h]->data; }
  o.off[kHeapActions] = t.actions.items.offsets;
  o.data[kHeapActions] = t.actions.items.data;
  o.entry_offsets = t.entry_offsets;
  o.crew_list = t.crew.list_offsets;
  o.actions_list = t.actions.list_offsets;
  o.created_at = t.created_at;
  o.archived_at = t.archived_at;
  o.delay_sec = t.delay_sec;
  o.delay_valid = t.delay_valid;
  o.entry_ts = t.entry_ts;
  o.time_val[PIE_TF_CREATED] = t.created_at;
  o.time_val[PIE_TF_UPDATED] = t.updated_at;
  o.time_val[PIE_TF_ARCHIVED] = t.archived_at;
  o.time_val[PIE_TF_DELETED] = t.deleted_at;
  o.time_kind = t.time_kind;
  o.text = nullptr;  // set by the launcher
  return o;
}

}  // namespace

int ingest_set_warp_path(int on) {
  const int old = warp_path_enabled() ? 1 : 0;
  if (on >= 0) g_warp_path.store(on ? 1 : 0);
  return old;
}
cudaError_t ingest_read_declined(const void* scratch, int64_t n_docs, unsigned int* out, cudaStream_t stream) {
  IngestScratch sc = carve(const_cast<void*>(scratch), n_docs);
  unsigned int twice = 0, too_long = 0;  // declined by both configurations; never offered to them
  cudaError_t e = cudaMemcpyAsync(&twice, sc.n_slow_big, sizeof(unsigned int), cudaMemcpyDeviceToHost, stream);
  if (e != cudaSuccess) return e;
  e = cudaMemcpyAsync(&too_long, sc.n_long, sizeof(unsigned int), cudaMemcpyDeviceToHost, stream);
  if (e != cudaSuccess) return e;
  e = cudaStreamSynchronize(stream);
  *out = twice + too_long;
  return e;
}

uint64_t ingest_scratch_bytes(int64_t n_docs) {
  const int64_t stride = ((n_docs > 0 ? n_docs : 1) + 31) & ~(int64_t)31;
  return (uint64_t)kPlanes * stride * 4 + ((uint64_t)kPlanes * scan_blocks(n_docs) + 1) * 8 + 64 + 4 * (kOrderBuckets + 2) +
         12 * (uint64_t)stride + 96 + (uint64_t)stride + (sizeof(jf::DocRec) + 8ull * jf::kPoolUnitsPerDoc) * (uint64_t)stride;
}

// The walk of the long documents runs beside the warp path on a stream of the library's own: fork after the routes
// are known, join before what follows reads the planes / the table.
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  int device = -1;
};
SideStream g_side;
std::mutex g_side_mutex;

cudaError_t side_fork(cudaStream_t main, cudaStream_t* side) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (!g_side.stream || g_side.device != dev) {
    if (g_side.stream) {
      cudaStreamDestroy(g_side.stream);
      cudaEventDestroy(g_side.fork);
      cudaEventDestroy(g_side.join);
      g_side = SideStream();
    }
    e = cudaStreamCreateWithFlags(&g_side.stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) return e;
    e = cudaEventCreateWithFlags(&g_side.fork, cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
    e = cudaEventCreateWithFlags(&g_side.join, cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
    g_side.device = dev;
  }
  e = cudaEventRecord(g_side.fork, main);
  if (e != cudaSuccess) return e;
  e = cudaStreamWaitEvent(g_side.stream, g_side.fork, 0);
  *side = g_side.stream;
  return e;
}
cudaError_t side_join(cudaStream_t main) {
  cudaError_t e = cudaEventRecord(g_side.join, g_side.stream);
  if (e != cudaSuccess) return e;
  return cudaStreamWaitEvent(main, g_side.join, 0);
}

void ingest_release() {  // pie_release(): the side stream and its events
  std::lock_guard<std::mutex> lock(g_side_mutex);
  if (g_side.stream) {
    cudaStreamDestroy(g_side.stream);
    cudaEventDestroy(g_side.fork);
    cudaEventDestroy(g_side.join);
  }
  g_side = SideStream();
}

cudaError_t launch_ingest_measure(const pie_json_docs& docs, void* scratch, uint8_t* doc_status, int64_t* totals,
                                  int32_t* status, cudaStream_t stream) {
  const int64_t n = docs.n_docs;
  IngestScratch sc = carve(scratch, n);
  const int nblk = scan_blocks(n);
  cudaError_t e = cudaMemsetAsync(sc.block_sums, 0, ((uint64_t)kPlanes * nblk + 1) * 8, stream);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(sc.buckets, 0, 4 * (kOrderBuckets + 2), stream);
  if (e != cudaSuccess) return e;
  ingest_init_kernel<<<1, 1, 0, stream>>>(sc);
  ++g_launches;
  if (n > 0) {
    if (warp_path_enabled()) {
      // the documents too long for the warp path are known from their lengths: the walk takes them on the side stream
      // while the warp path takes every document it can decide — first with the lists nearly every document fits, then,
      // what that declined, with the roomy ones; what is declined twice goes on sc.order_big for a second walk
      std::lock_guard<std::mutex> lock(g_side_mutex);
      ingest_route_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(docs.offsets, n, sc);
      cudaStream_t side = nullptr;
      e = side_fork(stream, &side);
      if (e != cudaSuccess) return e;
      ingest_walk_kernel<false><<<walk_blocks(n), kIngestThreads, 0, side>>>(docs.offsets, docs.data, n, sc, doc_status,
                                                                            IngestOut{}, sc.order_long, sc.n_long, sc.next_long);
      e = launch_fast<false, jf::CapsSmall, false>(docs, sc, doc_status, IngestOut{}, stream);
      if (e != cudaSuccess) return e;
      e = launch_fast<false, jf::CapsBig, true>(docs, sc, doc_status, IngestOut{}, stream);
      if (e != cudaSuccess) return e;
      ingest_walk_kernel<false><<<walk_blocks(n), kIngestThreads, 0, stream>>>(docs.offsets, docs.data, n, sc, doc_status,
                                                                              IngestOut{}, sc.order_big, sc.n_slow_big, sc.next_doc);
      g_launches += 3;
      e = side_join(stream);
      if (e != cudaSuccess) return e;
    } else {
      const unsigned doc_blocks = (unsigned)((n + 255) / 256);
      ingest_order_count_kernel<<<doc_blocks, 256, 0, stream>>>(docs.offsets, n, sc);
      ingest_order_scan_kernel<<<1, 1024, 0, stream>>>(sc);
      ingest_order_place_kernel<<<doc_blocks, 256, 0, stream>>>(docs.offsets, n, sc);
      g_launches += 3;
      ingest_walk_kernel<false><<<walk_blocks(n), kIngestThreads, 0, stream>>>(docs.offsets, docs.data, n, sc, doc_status,
                                                                              IngestOut{}, sc.order, nullptr, sc.next_doc);
      ++g_launches;
    }
    ingest_scan_sums_kernel<<<nblk, kScanThreads, 0, stream>>>(sc, n, nblk);
    ++g_launches;
  }
  ingest_scan_blocks_kernel<<<kPlanes, 1024, 0, stream>>>(sc, nblk, totals, status);
  ++g_launches;
  if (n > 0) {
    ingest_scan_apply_kernel<<<nblk, kScanThreads, 0, stream>>>(sc, n, nblk);
    ++g_launches;
  }
  return cudaGetLastError();
}

uint64_t ingest_fill_scratch_bytes(int64_t n_entries) { return sizeof(EntryRow) * (uint64_t)(n_entries > 0 ? n_entries : 1); }

cudaError_t launch_ingest_fill(const pie_json_docs& docs, const void* scratch, const uint8_t* doc_status,
                               const pie_archive_table& table, void* fill_scratch, cudaStream_t stream) {
  const int64_t n = docs.n_docs;
  IngestScratch sc = carve(const_cast<void*>(scratch), n);
  cudaError_t e = cudaMemsetAsync(sc.next_doc + 1, 0, 8, stream);  // pass 2 may run more than once per pass 1
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(sc.next_fast + 1, 0, 8, stream);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(sc.next_big + 1, 0, 8, stream);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(sc.next_long + 1, 0, 8, stream);
  if (e != cudaSuccess) return e;
  IngestOut out = make_out(table);
  out.text = docs.data;
  if (n > 0 && warp_path_enabled()) {
    // both kernels write the table's columns directly (the walk's entry rows are for when it takes every document)
    out.rows = nullptr;
    std::lock_guard<std::mutex> lock(g_side_mutex);
    cudaStream_t side = nullptr;
    e = side_fork(stream, &side);
    if (e != cudaSuccess) return e;
    ingest_walk_kernel<true><<<walk_blocks(n), kIngestThreads, 0, side>>>(docs.offsets, docs.data, n, sc, const_cast<uint8_t*>(doc_status),
                                                                         out, sc.order_long, sc.n_long, sc.next_long + 1);
    e = launch_fast<true, jf::CapsSmall, false>(docs, sc, const_cast<uint8_t*>(doc_status), out, stream);
    if (e != cudaSuccess) return e;
    e = launch_fast<true, jf::CapsBig, true>(docs, sc, const_cast<uint8_t*>(doc_status), out, stream);
    if (e != cudaSuccess) return e;
    ingest_walk_kernel<true><<<walk_blocks(n), kIngestThreads, 0, stream>>>(docs.offsets, docs.data, n, sc, const_cast<uint8_t*>(doc_status),
                                                                           out, sc.order_big, sc.n_slow_big, sc.next_doc + 1);
    ++g_launches;
    e = side_join(stream);
    if (e != cudaSuccess) return e;
  } else if (n > 0) {
    out.rows = static_cast<EntryRow*>(fill_scratch);
    ingest_walk_kernel<true><<<walk_blocks(n), kIngestThreads, 0, stream>>>(docs.offsets, docs.data, n, sc, const_cast<uint8_t*>(doc_status),
                                                                           out, sc.order, nullptr, sc.next_doc + 1);
    if (table.n_entries > 0) {
      ingest_rows_to_columns_kernel<<<(unsigned)((table.n_entries + 255) / 256), 256, 0, stream>>>(out.rows, table.n_entries, out);
      ++g_launches;
    }
  } else {
    ingest_empty_kernel<<<1, 1, 0, stream>>>(out);
  }
  ++g_launches;
  return cudaGetLastError();
}

}  // namespace pie
