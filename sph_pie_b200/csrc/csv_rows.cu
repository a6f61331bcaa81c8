// Export rows on sm_100a.
// Replaces buildTableRow + csvEscape + buildCsvRow (reference server/webhookDispatcher.js:276-342,
// twin public/app.js:5582-5612, :6025-6034) mapped over every entry of every show — what
// dispatchShowEvent puts in csv.rows (:571) and exportShowAsCsv joins with '\n' (public/app.js:5567).
//
// Output = one string column: row i is out_data[row_offsets[i] .. row_offsets[i+1]-1) and is followed
// by one '\n', so a show's CSV body is a single contiguous slice.
//
// ONE pass over the inputs.  A CTA takes a tile of kRowsPerTile consecutive entries:
//   1. every thread measures its row: escaped length of the 24 cells (word-wise scan for the four
//      characters that force quoting), Number::toString(delaySec) formatted once
//   2. block scan -> tile total; decoupled look-back over the tile totals -> the tile's byte offset
//   3. every thread writes its row into SHARED memory at the same 16-byte phase as the global
//      destination (word-wise copy through a byte-stream writer); the tile is flushed with 16-byte
//      coalesced stores
// Tiles whose rows do not fit the shared buffer (very long free text) are written straight to
// global memory by the same code.
#include "pie_device.cuh"
#include "pie_kernels.h"
#include "pie_numfmt.cuh"

namespace pie {

__device__ const uint64_t d_pow5_inv[PIE_RYU_POW5_INV_SPLIT_N][2] = PIE_RYU_POW5_INV_SPLIT_INIT;
__device__ const uint64_t d_pow5[PIE_RYU_POW5_SPLIT_N][2] = PIE_RYU_POW5_SPLIT_INIT;

constexpr int kRowsPerTile = 128;
constexpr int kTileBytes = 40 * 1024;  // shared staging buffer (rows of ~280 B -> ~35 KB per tile)

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
__host__ __device__ static inline int64_t csv_tiles(int64_t n_entries) {
  return (n_entries + kRowsPerTile - 1) / kRowsPerTile;
}

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

// ---- word-wise scanning and copying -----------------------------------------------------------
// A cell is measured by scanning the ALIGNED 32-bit words it touches for the four characters that
// force quoting (SIMD-in-register zero-byte test), and written by re-aligning those words with a
// funnel shift into a byte-stream writer that emits aligned 32-bit stores.  Per-byte loops remain
// only for cells that do need quoting (rare) and for the 0..3 bytes at the two ends of a row.

// != 0 iff some byte of v is zero (exact as a boolean)
__device__ __forceinline__ uint32_t zero_byte_flags(uint32_t v) { return (v - 0x01010101u) & ~v & 0x80808080u; }

__device__ __forceinline__ uint32_t special_flags(uint32_t x) {  // " , \n \r   (csvEscape, :334)
  return zero_byte_flags(x ^ 0x22222222u) | zero_byte_flags(x ^ 0x2C2C2C2Cu) | zero_byte_flags(x ^ 0x0A0A0A0Au) |
         zero_byte_flags(x ^ 0x0D0D0D0Du);
}

__device__ __forceinline__ bool has_special(const uint8_t* __restrict__ p, int n) {
  if (n <= 0) return false;
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uint32_t* __restrict__ w = reinterpret_cast<const uint32_t*>(a & ~static_cast<uintptr_t>(3));
  const uint32_t lead = static_cast<uint32_t>(a & 3);
  const int nw = static_cast<int>((lead + n + 3) >> 2);  // aligned words that hold bytes of the cell
  const uint32_t tail = (lead + n) & 3u;
  uint32_t flags = 0;
  for (int k = 0; k < nw; ++k) {
    uint32_t x = __ldg(w + k);
    if (k == 0) x &= 0xFFFFFFFFu << (8 * lead);             // bytes before the cell -> 0 (not special)
    if (k == nw - 1 && tail) x &= (1u << (8 * tail)) - 1u;  // bytes after the cell  -> 0
    flags |= special_flags(x);
  }
  return flags != 0;
}

__device__ __forceinline__ uint32_t count_quotes(const uint8_t* __restrict__ p, int n) {
  uint32_t c = 0;
  for (int i = 0; i < n; ++i) c += (p[i] == '"');
  return c;
}

// Byte-stream writer: bytes are collected in a 64-bit accumulator and leave as aligned 32-bit stores.
// The first word of a row may start mid-word (its low `lead` bytes belong to the previous row, which
// another thread writes): that word and the last partial word are stored byte by byte.
struct StreamWriter {
  uint8_t* p;  // aligned address of the next word to store
  unsigned long long acc;
  uint32_t fill;  // bytes pending in acc (including `lead` placeholders before the first flush)
  uint32_t lead;  // placeholder bytes of the first word; 0 once the first word has been stored

  __device__ __forceinline__ void init(uint8_t* start) {
    lead = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(start) & 3);
    p = start - lead;
    acc = 0;
    fill = lead;
  }
  __device__ __forceinline__ void flush_word() {
    const uint32_t v = static_cast<uint32_t>(acc);
    if (lead) {
      for (uint32_t b = lead; b < 4; ++b) p[b] = static_cast<uint8_t>(v >> (8 * b));
      lead = 0;
    } else {
      *reinterpret_cast<uint32_t*>(p) = v;
    }
    p += 4;
    acc >>= 32;
    fill -= 4;
  }
  // k (1..4) low bytes of w; the other bytes of w must be zero
  __device__ __forceinline__ void append(uint32_t w, uint32_t k) {
    acc |= static_cast<unsigned long long>(w) << (8 * fill);
    fill += k;
    if (fill >= 4) flush_word();
  }
  __device__ __forceinline__ void put(uint8_t c) { append(c, 1); }
  __device__ __forceinline__ void finish() {
    for (uint32_t b = lead; b < fill; ++b) p[b] = static_cast<uint8_t>(acc >> (8 * b));
  }
};

// copy s[0..n) (no quoting needed) followed by the separator byte
__device__ __forceinline__ void copy_plain(StreamWriter& out, const uint8_t* __restrict__ s, int n, uint8_t sep) {
  if (n > 0) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(s);
    const uint32_t* __restrict__ w = reinterpret_cast<const uint32_t*>(a & ~static_cast<uintptr_t>(3));
    const uint32_t sh = static_cast<uint32_t>(a & 3) * 8;
    const int last = static_cast<int>(((a & 3) + n - 1) >> 2);  // last aligned word holding cell bytes
    uint32_t cur = __ldg(w);
    int k = 0;
    for (; n >= 4; n -= 4, ++k) {  // full words of the cell
      const uint32_t nxt = (k + 1 <= last) ? __ldg(w + k + 1) : 0u;
      out.append(__funnelshift_r(cur, nxt, sh), 4);
      cur = nxt;
    }
    if (n > 0) {  // 1..3 trailing bytes; the separator rides in the same word
      const uint32_t nxt = (k + 1 <= last) ? __ldg(w + k + 1) : 0u;
      const uint32_t x = __funnelshift_r(cur, nxt, sh) & ((1u << (8 * n)) - 1u);
      out.append(x | (static_cast<uint32_t>(sep) << (8 * n)), static_cast<uint32_t>(n) + 1);
      return;
    }
  }
  out.put(sep);
}

// csvEscape of a cell that needs quotes (rare): byte-wise
__device__ __forceinline__ void copy_quoted_bytes(StreamWriter& out, const uint8_t* __restrict__ s, int n) {
  for (int i = 0; i < n; ++i) {
    const uint8_t c = s[i];
    if (c == '"') out.put('"');
    out.put(c);
  }
}

struct RowPlan {
  uint32_t len;    // bytes of the row including the trailing '\n'
  uint32_t quote;  // bit c: cell c (EXPORT_COLUMNS index) is wrapped in quotes; bit 31: status === 'Completed'
};

__device__ __forceinline__ void measure_cell(RowPlan& r, int col, const uint8_t* __restrict__ s, int n) {
  if (has_special(s, n)) {
    r.quote |= 1u << col;
    r.len += 2u + count_quotes(s, n);
  }
  r.len += static_cast<uint32_t>(n);
}

__device__ __forceinline__ void write_cell(StreamWriter& out, bool quote, const uint8_t* __restrict__ s, int n,
                                           uint8_t sep) {
  if (!quote) {
    copy_plain(out, s, n, sep);
  } else {
    out.put('"');
    copy_quoted_bytes(out, s, n);
    out.put('"');
    out.put(sep);
  }
}

// entry.status === 'Completed' (strict, case-sensitive, :293-297) blanks the five issue cells
__device__ __forceinline__ bool status_is_completed(const pie_archive_view& v, int64_t e) {
  const int b = v.status.offsets[e], n = v.status.offsets[e + 1] - b;
  if (n != 9) return false;
  uint32_t x[3];
  fetch_words_raw<3>(v.status.data + b, 9, x);
  return x[0] == lit_word("Completed", 0) && x[1] == lit_word("Completed", 1) &&
         (x[2] & 0xFFu) == lit_word("Completed", 2);
}

// The 24 cells of a row, in EXPORT_COLUMNS order (:15-19), as a table the kernels LOOP over: calling
// 24 inlined cell handlers twice made a 30 000-instruction kernel that thrashed the instruction cache.
enum : uint8_t { kCellString = 0, kCellJoined = 1, kCellNumber = 2 };
struct CellDesc {
  const int32_t* offsets;       // string column / items of a list column
  const uint8_t* data;
  const int32_t* list_offsets;  // kCellJoined only
  uint8_t kind;
  uint8_t per_entry;            // row index is the entry (1) or its show (0)
  uint8_t blank_if_completed;   // :293-297
};
struct RowTable {
  CellDesc cell[PIE_N_EXPORT_COLUMNS];
};

static RowTable make_row_table(const pie_archive_view& v) {
  RowTable t;
  auto str = [](const pie_strcol& c, int per_entry, int blank = 0) {
    return CellDesc{c.offsets, c.data, nullptr, kCellString, (uint8_t)per_entry, (uint8_t)blank};
  };
  auto lst = [](const pie_strlistcol& c, int per_entry) {
    return CellDesc{c.items.offsets, c.items.data, c.list_offsets, kCellJoined, (uint8_t)per_entry, 0};
  };
  t.cell[0] = str(v.show_id, 0);      t.cell[1] = str(v.show_date, 0);     t.cell[2] = str(v.show_time, 0);
  t.cell[3] = str(v.show_label, 0);   t.cell[4] = lst(v.crew, 0);          t.cell[5] = str(v.lead_pilot, 0);
  t.cell[6] = str(v.monkey_lead, 0);  t.cell[7] = str(v.show_notes, 0);    t.cell[8] = str(v.entry_id, 1);
  t.cell[9] = str(v.unit_id, 1);      t.cell[10] = str(v.planned, 1);      t.cell[11] = str(v.launched, 1);
  t.cell[12] = str(v.status, 1);      t.cell[13] = str(v.primary_issue, 1, 1);
  t.cell[14] = str(v.sub_issue, 1, 1);  t.cell[15] = str(v.other_detail, 1, 1);
  t.cell[16] = str(v.severity, 1, 1);   t.cell[17] = str(v.root_cause, 1, 1);
  t.cell[18] = lst(v.actions, 1);     t.cell[19] = str(v.operator_name, 1); t.cell[20] = str(v.battery_id, 1);
  t.cell[21] = CellDesc{nullptr, nullptr, nullptr, kCellNumber, 1, 0};
  t.cell[22] = str(v.command_rx, 1);  t.cell[23] = str(v.notes, 1);
  return t;
}

// Row = cells joined by ',' (:341), then '\n'.
// num / num_len: Number::toString(delaySec), produced once here and reused by write_row.
__device__ __forceinline__ RowPlan measure_row(const pie_archive_view& v, const RowTable& tab, int64_t e, int64_t s,
                                               char* num, int* num_len) {
  RowPlan r{24u, 0u};  // 23 commas + '\n'
  const bool completed = status_is_completed(v, e);
  if (completed) r.quote |= 1u << 31;  // remembered for write_row
  *num_len = 0;
#pragma unroll 1
  for (int col = 0; col < PIE_N_EXPORT_COLUMNS; ++col) {
    const CellDesc& d = tab.cell[col];
    const int64_t i = d.per_entry ? e : s;
    if (d.blank_if_completed && completed) continue;
    if (d.kind == kCellString) {
      const int b = d.offsets[i];
      measure_cell(r, col, d.data + b, d.offsets[i + 1] - b);
    } else if (d.kind == kCellJoined) {
      const int l0 = d.list_offsets[i], l1 = d.list_offsets[i + 1];
      if (l1 > l0) {
        const int b = d.offsets[l0], n = d.offsets[l1] - b;
        measure_cell(r, col, d.data + b, n);  // '|' is not special: quotes iff some item needs them
        r.len += static_cast<uint32_t>(l1 - l0 - 1);
      }
    } else if (v.delay_valid[e]) {  // delaySec === null || undefined ? '' : delaySec, then String() (:301, :333)
      const RyuTables t{d_pow5_inv, d_pow5};
      const int nl = js_number_to_string(v.delay_sec[e], num, t);
      *num_len = nl;
      r.len += static_cast<uint32_t>(nl);
    }
  }
  return r;
}

__device__ __forceinline__ void write_row(StreamWriter& out, const RowTable& tab, int64_t e, int64_t s, uint32_t q,
                                          const char* num, int num_len) {
  const bool completed = (q >> 31) != 0;
#pragma unroll 1
  for (int col = 0; col < PIE_N_EXPORT_COLUMNS; ++col) {
    const CellDesc& d = tab.cell[col];
    const int64_t i = d.per_entry ? e : s;
    const uint8_t sep = (col == PIE_N_EXPORT_COLUMNS - 1) ? (uint8_t)'\n' : (uint8_t)',';
    const bool quote = (q >> col) & 1u;
    if (d.blank_if_completed && completed) {
      out.put(sep);
    } else if (d.kind == kCellString) {
      const int b = d.offsets[i];
      write_cell(out, quote, d.data + b, d.offsets[i + 1] - b, sep);
    } else if (d.kind == kCellJoined) {
      const int l0 = d.list_offsets[i], l1 = d.list_offsets[i + 1];
      if (quote) out.put('"');
      for (int l = l0; l < l1; ++l) {
        const int b = d.offsets[l], n = d.offsets[l + 1] - b;
        const bool last_item = (l + 1 == l1);
        if (quote) {
          copy_quoted_bytes(out, d.data + b, n);
          if (!last_item) out.put('|');
        } else {
          copy_plain(out, d.data + b, n, last_item ? sep : (uint8_t)'|');
        }
      }
      if (quote) out.put('"');
      if (quote || l1 <= l0) out.put(sep);
    } else {
      for (int k = 0; k < num_len; ++k) out.put(static_cast<uint8_t>(num[k]));
      out.put(sep);
    }
  }
  out.finish();
}

__global__ void __launch_bounds__(kRowsPerTile) csv_rows_kernel(pie_archive_view v, const __grid_constant__ RowTable tab,
                                                                CsvScratch sc,
                                                                int64_t* __restrict__ row_offsets,
                                                                uint8_t* __restrict__ out_data, uint64_t capacity,
                                                                unsigned long long* __restrict__ total_out) {
  extern __shared__ __align__(16) uint8_t s_tile[];  // kTileBytes + 16
  __shared__ char s_num[kRowsPerTile][kMaxNumberChars];
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

  // 1. measure (and format delaySec once)
  uint32_t len = 0, qmask = 0;
  int num_len = 0;
  if (have) {
    const RowPlan plan = measure_row(v, tab, e, s, s_num[tid], &num_len);
    len = plan.len;
    qmask = plan.quote;
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
    StreamWriter w;
    w.init(staged ? (s_tile + pad + local) : (out_data + base + local));
    write_row(w, tab, e, s, qmask, s_num[tid], num_len);
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
      v, make_row_table(v), sc, row_offsets, out_data, capacity, total_out);
  g_launches += 2;
  return cudaGetLastError();
}

}  // namespace pie
