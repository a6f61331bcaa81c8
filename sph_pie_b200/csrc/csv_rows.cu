// Export rows on sm_100a.
// Replaces buildTableRow + csvEscape + buildCsvRow (reference server/webhookDispatcher.js:276-342,
// twin public/app.js:5582-5612, :6025-6034) mapped over every entry of every show — what
// dispatchShowEvent puts in csv.rows (:571) and exportShowAsCsv joins with '\n' (public/app.js:5567).
//
// Output = one string column: row i is out_data[row_offsets[i] .. row_offsets[i+1]-1) and is followed
// by one '\n', so a show's CSV body is a single contiguous slice.
//
// ONE pass over the inputs.  A CTA takes a tile of kRowsPerTile consecutive entries:
//   1. every thread measures its row (escaped length of 24 cells; Number::toString for delaySec)
//   2. block scan -> tile total; decoupled look-back over the tile totals -> the tile's byte offset
//   3. every thread writes its row into SHARED memory at the same 16-byte phase as the global
//      destination; the tile is flushed with 16-byte coalesced stores
// Rows of the tile that do not fit the shared buffer (very long free text) are written straight
// to global memory by the same code.
#include "pie_device.cuh"
#include "pie_kernels.h"
#include "pie_numfmt.cuh"

namespace pie {

__device__ const uint64_t d_pow5_inv[PIE_RYU_POW5_INV_SPLIT_N][2] = PIE_RYU_POW5_INV_SPLIT_INIT;
__device__ const uint64_t d_pow5[PIE_RYU_POW5_SPLIT_N][2] = PIE_RYU_POW5_SPLIT_INIT;

constexpr int kRowsPerTile = 128;
constexpr int kTileBytes = 40 * 1024;  // shared staging buffer (rows of ~190 B -> ~24 KB per tile)

constexpr unsigned long long kStatusShift = 62;
constexpr unsigned long long kValueMask = (1ull << kStatusShift) - 1;
constexpr unsigned long long kAggregate = 1ull << kStatusShift;
constexpr unsigned long long kPrefix = 2ull << kStatusShift;

struct CsvScratch {
  unsigned long long* tile_state;  // [n_tiles] packed (status, value); zeroed before launch
  unsigned int* tile_counter;      // [1] dynamic tile ids; zeroed before launch
  int32_t* entry_show;             // [n_entries]
};

static inline uint64_t align256(uint64_t x) { return (x + 255) & ~(uint64_t)255; }
__host__ __device__ static inline int64_t csv_tiles(int64_t n_entries) { return (n_entries + kRowsPerTile - 1) / kRowsPerTile; }

uint64_t csv_scratch_bytes(int64_t n_entries) {
  const uint64_t e = (uint64_t)(n_entries > 0 ? n_entries : 1);
  return align256(8 * (uint64_t)csv_tiles(e)) + 256 + align256(4 * e);
}
uint64_t csv_scratch_zero_bytes(int64_t n_entries) {  // leading part that must be zero at launch
  const uint64_t e = (uint64_t)(n_entries > 0 ? n_entries : 1);
  return align256(8 * (uint64_t)csv_tiles(e)) + 256;
}
static CsvScratch carve_csv(void* scratch, int64_t n_entries) {
  const uint64_t e = (uint64_t)(n_entries > 0 ? n_entries : 1);
  uint8_t* p = static_cast<uint8_t*>(scratch);
  CsvScratch s;
  s.tile_state = (unsigned long long*)p; p += align256(8 * (uint64_t)csv_tiles(e));
  s.tile_counter = (unsigned int*)p; p += 256;
  s.entry_show = (int32_t*)p;
  return s;
}

// show index of every entry (rows of show s are entry_offsets[s] .. entry_offsets[s+1])
__global__ void __launch_bounds__(256) expand_entry_show_kernel(pie_archive_view v, int32_t* __restrict__ entry_show) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= v.n_shows) return;
  for (int e = v.entry_offsets[s]; e < v.entry_offsets[s + 1]; ++e) entry_show[e] = (int32_t)s;
}

// ---- sinks: the row builder is written once and instantiated for measuring and for writing ----
struct SizeSink {
  uint32_t n = 0;
  __device__ __forceinline__ void put(uint8_t) { n += 1; }
  __device__ __forceinline__ void skip(uint32_t k) { n += k; }
  static constexpr bool kMeasureOnly = true;
};
struct ByteSink {  // shared or global memory, byte granular
  uint8_t* p;
  __device__ __forceinline__ void put(uint8_t c) { *p++ = c; }
  static constexpr bool kMeasureOnly = false;
};

__device__ __forceinline__ bool csv_special(uint8_t c) { return c == '"' || c == ',' || c == '\n' || c == '\r'; }

// csvEscape(str) (:332-338): quote iff the string contains " , \n or \r; double the quotes.
template <typename Sink>
__device__ __forceinline__ void emit_cell(Sink& out, const uint8_t* __restrict__ s, int n) {
  bool quote = false;
  uint32_t dq = 0;
  for (int i = 0; i < n; ++i) {
    const uint8_t c = s[i];
    quote |= csv_special(c);
    dq += (c == '"');
  }
  if constexpr (Sink::kMeasureOnly) {
    out.skip((uint32_t)n + (quote ? 2u + dq : 0u));
  } else {
    if (quote) out.put('"');
    for (int i = 0; i < n; ++i) {
      const uint8_t c = s[i];
      if (quote && c == '"') out.put('"');
      out.put(c);
    }
    if (quote) out.put('"');
  }
}

template <typename Sink>
__device__ __forceinline__ void emit_strcol(Sink& out, const pie_strcol& c, int64_t i) {
  const int b = c.offsets[i], e = c.offsets[i + 1];
  emit_cell(out, c.data + b, e - b);
}

// Array.prototype.join('|') then csvEscape of the joined string (crew :284, actions :298)
template <typename Sink>
__device__ __forceinline__ void emit_joined(Sink& out, const pie_strlistcol& c, int64_t i) {
  const int l0 = c.list_offsets[i], l1 = c.list_offsets[i + 1];
  if (l1 <= l0) return;
  const int b = c.items.offsets[l0], e = c.items.offsets[l1];
  const uint8_t* __restrict__ s = c.items.data;
  bool quote = false;
  uint32_t dq = 0;
  for (int k = b; k < e; ++k) {
    const uint8_t ch = s[k];
    quote |= csv_special(ch);
    dq += (ch == '"');
  }
  if constexpr (Sink::kMeasureOnly) {
    out.skip((uint32_t)(e - b) + (uint32_t)(l1 - l0 - 1) + (quote ? 2u + dq : 0u));
  } else {
    if (quote) out.put('"');
    for (int l = l0; l < l1; ++l) {
      if (l > l0) out.put('|');
      for (int k = c.items.offsets[l]; k < c.items.offsets[l + 1]; ++k) {
        const uint8_t ch = s[k];
        if (quote && ch == '"') out.put('"');
        out.put(ch);
      }
    }
    if (quote) out.put('"');
  }
}

// One CSV row: EXPORT_COLUMNS order (:15-19), cells joined by ',' (:341), then '\n'.
template <typename Sink>
__device__ __forceinline__ void emit_row(Sink& out, const pie_archive_view& v, int64_t e, int64_t s) {
  emit_strcol(out, v.show_id, s);      out.put(',');
  emit_strcol(out, v.show_date, s);    out.put(',');
  emit_strcol(out, v.show_time, s);    out.put(',');
  emit_strcol(out, v.show_label, s);   out.put(',');
  emit_joined(out, v.crew, s);         out.put(',');
  emit_strcol(out, v.lead_pilot, s);   out.put(',');
  emit_strcol(out, v.monkey_lead, s);  out.put(',');
  emit_strcol(out, v.show_notes, s);   out.put(',');
  emit_strcol(out, v.entry_id, e);     out.put(',');
  emit_strcol(out, v.unit_id, e);      out.put(',');
  emit_strcol(out, v.planned, e);      out.put(',');
  emit_strcol(out, v.launched, e);     out.put(',');
  emit_strcol(out, v.status, e);       out.put(',');
  // entry.status === 'Completed' (strict, case-sensitive, :293-297) blanks the five issue cells
  const int sb = v.status.offsets[e], sn = v.status.offsets[e + 1] - sb;
  const bool completed = equals_exact(v.status.data + sb, sn, "Completed");
  if (!completed) emit_strcol(out, v.primary_issue, e);
  out.put(',');
  if (!completed) emit_strcol(out, v.sub_issue, e);
  out.put(',');
  if (!completed) emit_strcol(out, v.other_detail, e);
  out.put(',');
  if (!completed) emit_strcol(out, v.severity, e);
  out.put(',');
  if (!completed) emit_strcol(out, v.root_cause, e);
  out.put(',');
  emit_joined(out, v.actions, e);      out.put(',');
  emit_strcol(out, v.operator_name, e); out.put(',');
  emit_strcol(out, v.battery_id, e);   out.put(',');
  if (v.delay_valid[e]) {  // delaySec === null || undefined ? '' : delaySec, then String() (:301, :333)
    char buf[kMaxNumberChars];
    const RyuTables t{d_pow5_inv, d_pow5};
    const int n = js_number_to_string(v.delay_sec[e], buf, t);
    if constexpr (Sink::kMeasureOnly) out.skip((uint32_t)n);
    else for (int i = 0; i < n; ++i) out.put((uint8_t)buf[i]);
  }
  out.put(',');
  emit_strcol(out, v.command_rx, e);   out.put(',');
  emit_strcol(out, v.notes, e);
  out.put('\n');
}

__global__ void __launch_bounds__(kRowsPerTile) csv_rows_kernel(pie_archive_view v, CsvScratch sc,
                                                                int64_t* __restrict__ row_offsets,
                                                                uint8_t* __restrict__ out_data, uint64_t capacity,
                                                                unsigned long long* __restrict__ total_out) {
  extern __shared__ __align__(16) uint8_t s_tile[];  // kTileBytes + 16
  __shared__ uint32_t s_warp[kRowsPerTile / 32];
  __shared__ unsigned int s_tile_id;
  __shared__ unsigned long long s_base;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

  if (tid == 0) s_tile_id = atomicAdd(sc.tile_counter, 1u);  // tiles start in id order: look-back cannot deadlock
  __syncthreads();
  const int64_t tile = s_tile_id;
  const int64_t e = tile * kRowsPerTile + tid;
  const bool have = e < v.n_entries;
  const int64_t s = have ? sc.entry_show[e] : 0;

  // 1. measure
  uint32_t len = 0;
  if (have) {
    SizeSink sz;
    emit_row(sz, v, e, s);
    len = sz.n;
  }
  // 2. block exclusive scan of the row lengths
  uint32_t incl = len;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  uint32_t warp_base = 0, tile_total = 0;
#pragma unroll
  for (int w = 0; w < kRowsPerTile / 32; ++w) {
    if (w < wid) warp_base += s_warp[w];
    tile_total += s_warp[w];
  }
  const uint32_t local = warp_base + incl - len;  // byte offset of this row inside the tile

  // decoupled look-back over the tile totals (warp 0)
  if (wid == 0) {
    volatile unsigned long long* state = sc.tile_state;
    if (lane == 0) {
      __threadfence();
      state[tile] = (tile == 0 ? kPrefix : kAggregate) | (unsigned long long)tile_total;
    }
    unsigned long long exclusive = 0;
    int64_t idx = tile - 1;
    while (idx >= 0) {
      const int64_t j = idx - lane;
      unsigned long long st;
      do {
        st = 2ull << kStatusShift;  // before tile 0: an empty prefix
        if (j >= 0) st = state[j];
      } while (__any_sync(0xFFFFFFFFu, (st >> kStatusShift) == 0));
      const uint32_t is_prefix = __ballot_sync(0xFFFFFFFFu, (st >> kStatusShift) == 2);
      const int stop = is_prefix ? (__ffs(is_prefix) - 1) : 32;  // nearest tile that already knows its prefix
      unsigned long long part = (lane <= stop) ? (st & kValueMask) : 0ull;
#pragma unroll
      for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, o);
      exclusive += part;
      if (is_prefix) break;
      idx -= 32;
    }
    if (lane == 0) {
      if (tile > 0) {
        __threadfence();
        state[tile] = kPrefix | (exclusive + tile_total);
      }
      s_base = exclusive;
      if (tile == csv_tiles(v.n_entries) - 1) {
        *total_out = exclusive + tile_total;
        row_offsets[v.n_entries] = (int64_t)(exclusive + tile_total);
      }
    }
  }
  __syncthreads();
  const unsigned long long base = s_base;
  if (have) row_offsets[e] = (int64_t)(base + local);
  if (out_data == nullptr || base + tile_total > capacity) return;  // size-only call, or caller's buffer too small

  // 3. write the rows
  const uint32_t pad = (uint32_t)((reinterpret_cast<uintptr_t>(out_data) + base) & 15);
  const bool staged = (pad + tile_total) <= (uint32_t)kTileBytes + 16u;
  if (have) {
    ByteSink w{staged ? (s_tile + pad + local) : (out_data + base + local)};
    emit_row(w, v, e, s);
  }
  if (!staged) return;
  __syncthreads();
  // flush s_tile[pad .. pad+tile_total) -> out_data[base ..): smem and global share the 16-byte phase
  uint8_t* __restrict__ dst = out_data + base - pad;  // dst + k <-> s_tile + k
  const uint32_t end = pad + tile_total;
  const uint32_t body_begin = pad ? 16u : 0u, body_end = end & ~15u;
  if (body_end > body_begin) {
    for (uint32_t k = body_begin + 16u * tid; k < body_end; k += 16u * kRowsPerTile)
      *reinterpret_cast<uint4*>(dst + k) = *reinterpret_cast<const uint4*>(s_tile + k);
    for (uint32_t k = pad + tid; k < body_begin && k < end; k += kRowsPerTile) dst[k] = s_tile[k];
    for (uint32_t k = body_end + tid; k < end; k += kRowsPerTile) dst[k] = s_tile[k];
  } else {
    for (uint32_t k = pad + tid; k < end; k += kRowsPerTile) dst[k] = s_tile[k];
  }
}

cudaError_t launch_csv_rows(const pie_archive_view& v, int64_t* row_offsets, uint8_t* out_data, uint64_t capacity,
                            unsigned long long* total_out, void* scratch, cudaStream_t stream) {
  CsvScratch sc = carve_csv(scratch, v.n_entries);
  cudaError_t err = cudaMemsetAsync(scratch, 0, csv_scratch_zero_bytes(v.n_entries), stream);
  if (err != cudaSuccess) return err;
  if (v.n_entries == 0) {
    err = cudaMemsetAsync(total_out, 0, 8, stream);
    if (err != cudaSuccess) return err;
    return cudaMemsetAsync(row_offsets, 0, 8, stream);
  }
  static int configured_device = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_device != dev) {
    err = cudaFuncSetAttribute(csv_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileBytes + 16);
    if (err != cudaSuccess) return err;
    configured_device = dev;
  }
  expand_entry_show_kernel<<<(unsigned)((v.n_shows + 255) / 256), 256, 0, stream>>>(v, sc.entry_show);
  csv_rows_kernel<<<(unsigned)csv_tiles(v.n_entries), kRowsPerTile, kTileBytes + 16, stream>>>(
      v, sc, row_offsets, out_data, capacity, total_out);
  g_launches += 2;
  return cudaGetLastError();
}

}  // namespace pie
