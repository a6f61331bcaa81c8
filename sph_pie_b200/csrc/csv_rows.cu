// Export rows on sm_100a.
// Replaces buildTableRow + csvEscape + buildCsvRow (reference server/webhookDispatcher.js:276-342,
// twin public/app.js:5582-5612, :6025-6034) mapped over every entry of every show — what
// dispatchShowEvent puts in csv.rows (:571) and exportShowAsCsv joins with '\n' (public/app.js:5567).
//
// Output = one string column: row i is out_data[row_offsets[i] .. row_offsets[i+1]-1) and is followed
// by one '\n', so a show's CSV body is a single contiguous slice.
//
// ONE pass over the inputs.  A CTA takes a tile of kRows consecutive entries = 24 x
// kRows CELLS, and works cell-parallel, column-major: a warp handles ONE column for 32 consecutive
// rows, so its offset loads are coalesced, its cell lengths are alike (no divergence), and every
// thread has independent loads in flight (a thread per row walked 24 dependent cells serially).
//   1. measure  every cell: escaped length (word-wise scan for the characters that force quotes);
//               Number::toString(delaySec) is formatted once into shared memory
//   2. scan     per row over its 24 cells (shared memory), block scan over the rows -> tile total;
//               the tile's aggregate is published for the decoupled look-back
//   3. write    every cell into the SHARED tile buffer through a byte-stream writer (aligned
//               32-bit stores; only the 0..3 bytes at a cell's two ends are byte stores)
//   4. look-back over the tile totals -> the tile's global byte offset (overlaps 3 in other CTAs)
//   5. flush    16-byte coalesced stores; the shared buffer is re-aligned to the destination with a
//               funnel shift, so step 3 does not have to wait for the offset
// A tile that does not fit the shared buffer (very long free text) is written straight to global
// memory by the same cell code, after its look-back.
#include "pie_device.cuh"
#include "pie_kernels.h"
#include "pie_numfmt.cuh"

namespace pie {

__device__ const uint64_t d_pow5_inv[PIE_RYU_POW5_INV_SPLIT_N][2] = PIE_RYU_POW5_INV_SPLIT_INIT;
__device__ const uint64_t d_pow5[PIE_RYU_POW5_SPLIT_N][2] = PIE_RYU_POW5_SPLIT_INIT;

#ifndef PIE_CSV_ROWS
#define PIE_CSV_ROWS 128
#endif
#ifndef PIE_CSV_THREADS
#define PIE_CSV_THREADS 512
#endif
#ifndef PIE_CSV_MIN_BLOCKS
#define PIE_CSV_MIN_BLOCKS 3
#endif
#ifndef PIE_CSV_TILE_KB
#define PIE_CSV_TILE_KB 48
#endif
constexpr int kRows = PIE_CSV_ROWS;                 // rows (entries) per tile; multiple of 32
constexpr int kThreads = PIE_CSV_THREADS;           // multiple of kRows
constexpr int kCols = PIE_N_EXPORT_COLUMNS;         // 24
constexpr int kCells = kRows * kCols;
constexpr int kTileBytes = PIE_CSV_TILE_KB * 1024;  // shared tile buffer (rows of ~280 B -> ~36 KB per 128 rows)
static_assert(kRows % 32 == 0 && kThreads % kRows == 0 && kCells % kThreads == 0, "tile shape");

constexpr unsigned long long kStatusShift = 62;
constexpr unsigned long long kValueMask = (1ull << kStatusShift) - 1;
constexpr unsigned long long kAggregate = 1ull << kStatusShift;
constexpr unsigned long long kPrefix = 2ull << kStatusShift;

struct CsvScratch {
  unsigned long long* tile_state;  // [n_tiles] packed (status, value); zeroed before launch
  unsigned int* tile_counter;      // [1] dynamic tile ids; zeroed before launch
  unsigned int* col_dirty;         // [24] column c holds at least one " , \n \r somewhere; zeroed before launch
  int32_t* entry_show;             // [n_entries]
};

static inline uint64_t align256(uint64_t x) { return (x + 255) & ~(uint64_t)255; }
__host__ __device__ static inline int64_t csv_tiles(int64_t n_entries) { return (n_entries + kRows - 1) / kRows; }

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
  s.tile_counter = (unsigned int*)p;
  s.col_dirty = (unsigned int*)(p + 64);
  p += 256;
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
// != 0 iff some byte of v is zero (exact as a boolean)
__device__ __forceinline__ uint32_t zero_byte_flags(uint32_t v) { return (v - 0x01010101u) & ~v & 0x80808080u; }

__device__ __forceinline__ uint32_t special_flags(uint32_t x) {  // " , \n \r   (csvEscape, :334)
  return zero_byte_flags(x ^ 0x22222222u) | zero_byte_flags(x ^ 0x2C2C2C2Cu) | zero_byte_flags(x ^ 0x0A0A0A0Au) |
         zero_byte_flags(x ^ 0x0D0D0D0Du);
}

// does s[0..n) contain a character that forces quoting?  Scans the ALIGNED words the cell touches.
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

// number of '"' in s[0..n), four bytes at a time over the aligned words the cell touches
__device__ __forceinline__ uint32_t count_quotes(const uint8_t* __restrict__ p, int n) {
  if (n <= 0) return 0;
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uint32_t* __restrict__ w = reinterpret_cast<const uint32_t*>(a & ~static_cast<uintptr_t>(3));
  const uint32_t lead = static_cast<uint32_t>(a & 3);
  const int nw = static_cast<int>((lead + n + 3) >> 2);
  const uint32_t tail = (lead + n) & 3u;
  uint32_t c = 0;
  for (int k = 0; k < nw; ++k) {
    uint32_t x = __ldg(w + k);
    if (k == 0) x &= 0xFFFFFFFFu << (8 * lead);
    if (k == nw - 1 && tail) x &= (1u << (8 * tail)) - 1u;
    c += __popc(__vcmpeq4(x, 0x22222222u)) >> 3;  // 0xFF per byte equal to '"'
  }
  return c;
}

// Byte-stream writer: bytes are collected in a 64-bit accumulator and leave as aligned 32-bit stores.
// A cell may start and end mid-word; the neighbouring bytes of those words belong to other cells,
// which other threads write, so the first and the last partial word are stored byte by byte.
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
  if (n <= 0) return;
  const uintptr_t a = reinterpret_cast<uintptr_t>(s);
  const uint32_t* __restrict__ w = reinterpret_cast<const uint32_t*>(a & ~static_cast<uintptr_t>(3));
  const uint32_t sh = static_cast<uint32_t>(a & 3) * 8;
  const int last = static_cast<int>(((a & 3) + n - 1) >> 2);
  uint32_t cur = __ldg(w);
  for (int k = 0; n > 0; n -= 4, ++k) {
    const uint32_t nxt = (k + 1 <= last) ? __ldg(w + k + 1) : 0u;
    const uint32_t m = n >= 4 ? 4u : static_cast<uint32_t>(n);
    uint32_t x = __funnelshift_r(cur, nxt, sh);
    if (m < 4) x &= (1u << (8 * m)) - 1u;
    cur = nxt;
    if (zero_byte_flags(x ^ 0x22222222u) == 0) {  // no '"' in these bytes: whole word at once
      out.append(x, m);
    } else {
      for (uint32_t b = 0; b < m; ++b) {
        const uint8_t c = static_cast<uint8_t>(x >> (8 * b));
        if (c == '"') out.put('"');
        out.put(c);
      }
    }
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

// The 24 cells of a row, in EXPORT_COLUMNS order (:15-19)
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
  CellDesc cell[kCols];
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

// ---- pre-pass: which columns can need quoting at all? -----------------------------------------------
// Most columns of an archive (ids, dates, enumerations, names) never contain " , \n or \r.  One
// streaming pass over every column's byte heap (16 bytes per thread and step, HBM speed) sets a
// per-column flag; the row kernel then skips the per-cell scan for clean columns altogether.
__global__ void __launch_bounds__(256) column_dirty_kernel(const __grid_constant__ RowTable tab, int64_t n_shows,
                                                           int64_t n_entries, unsigned int* __restrict__ col_dirty) {
  const int col = blockIdx.y;
  const CellDesc& d = tab.cell[col];
  if (d.kind == kCellNumber) return;
  const int64_t n = d.per_entry ? n_entries : n_shows;
  int64_t first = 0, last = n;  // rows of the string column that hold this cell's bytes
  if (d.kind == kCellJoined) {
    first = d.list_offsets[0];
    last = d.list_offsets[n];
  }
  const int64_t b0 = d.offsets[first], b1 = d.offsets[last];
  if (b1 <= b0) return;
  const uintptr_t a0 = reinterpret_cast<uintptr_t>(d.data + b0), a1 = reinterpret_cast<uintptr_t>(d.data + b1);
  const uintptr_t w0 = a0 & ~static_cast<uintptr_t>(15), w1 = (a1 + 15) & ~static_cast<uintptr_t>(15);
  const int64_t chunks = static_cast<int64_t>((w1 - w0) >> 4);
  uint32_t flags = 0;
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < chunks; c += (int64_t)gridDim.x * blockDim.x) {
    const uintptr_t a = w0 + 16 * (uintptr_t)c;
    // Only words that hold heap bytes are loaded; the edge words are masked to the heap's bytes.
    uint32_t x[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uintptr_t wa = a + 4 * k;
      uint32_t v = 0;
      if (wa + 4 > a0 && wa < a1) {
        v = __ldg(reinterpret_cast<const uint32_t*>(wa));
        if (wa < a0) v &= 0xFFFFFFFFu << (8 * (uint32_t)(a0 - wa));
        if (wa + 4 > a1) v &= (1u << (8 * (uint32_t)(a1 - wa))) - 1u;
      }
      x[k] = v;
    }
    flags |= special_flags(x[0]) | special_flags(x[1]) | special_flags(x[2]) | special_flags(x[3]);
  }
  if (__any_sync(0xFFFFFFFFu, flags != 0) && (threadIdx.x & 31) == 0) atomicOr(&col_dirty[col], 1u);
}

// ---- the kernel ---------------------------------------------------------------------------------
// Work item = (row, group of kGroupCols consecutive columns); thread tid owns row tid % kRows, group
// tid / kRows, in BOTH the measuring and the writing phase, so the quote flags stay in registers and
// one byte-stream writer emits the whole group (byte stores only at the group's two ends).  A warp =
// one group x 32 consecutive rows: it walks the same column at the same time.
constexpr int kGroups = 4;
constexpr int kGroupCols = kCols / kGroups;   // 6
constexpr int kWorkers = kRows * kGroups;     // 512 worker threads ...
constexpr int kCtaThreads = kWorkers + 32;    // ... + one warp that runs the decoupled look-back
static_assert(kGroups * kGroupCols == kCols && kThreads == kWorkers, "tile shape");
constexpr uint32_t kCompletedBit = 1u << 31;

struct CsvSmem {
  uint32_t group[kGroups][kRows];  // phase 1: bytes of the group (with its separators); phase 2: start in the row
  uint32_t row_start[kRows];       // byte offset of the row inside the tile
  char num[kRows][kMaxNumberChars];
  uint8_t num_len[kRows];
  uint32_t warp_sum[kRows / 32];
  uint32_t col_dirty[kCols];
  uint32_t tile_total;
  unsigned int tile_id;
  unsigned long long base;
};

__global__ void __launch_bounds__(kCtaThreads, PIE_CSV_MIN_BLOCKS) csv_rows_kernel(pie_archive_view v, const __grid_constant__ RowTable tab,
                                                                  CsvScratch sc, int64_t* __restrict__ row_offsets,
                                                                  uint8_t* __restrict__ out_data, uint64_t capacity,
                                                                  unsigned long long bias,
                                                                  unsigned long long* __restrict__ total_out) {
  extern __shared__ __align__(16) uint8_t s_dyn[];
  uint8_t* s_tile = s_dyn;                                               // kTileBytes + 32
  CsvSmem& sm = *reinterpret_cast<CsvSmem*>(s_dyn + kTileBytes + 32);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const bool worker = tid < kWorkers;

  if (tid == 0) sm.tile_id = atomicAdd(sc.tile_counter, 1u);  // tiles start in id order: look-back cannot deadlock
  if (tid < kCols) sm.col_dirty[tid] = sc.col_dirty[tid];
  __syncthreads();
  const int64_t tile = sm.tile_id;
  const int64_t e0 = tile * kRows;
  const int rows = (v.n_entries - e0 < kRows) ? (int)(v.n_entries - e0) : kRows;
  const int g = tid / kRows, r = tid % kRows;
  const bool have = worker && r < rows;
  const int64_t e = e0 + r;
  const int64_t show = (have && g < 2) ? sc.entry_show[e] : 0;  // groups 0-1 hold the show-level columns 0..7

  // ---- 1a. locate the group's cells: every offset load is issued before any is consumed (a tile is
  // latency-bound: per column a dependent chain offsets -> bytes; walking 6 columns one after the
  // other cost 6 such chains per phase)
  int cb[kGroupCols], cn[kGroupCols], citems[kGroupCols];
  {
    int a0[kGroupCols], a1[kGroupCols];
#pragma unroll
    for (int k = 0; k < kGroupCols; ++k) {
      const CellDesc& d = tab.cell[g * kGroupCols + k];
      const int64_t i = d.per_entry ? e : show;
      const int32_t* __restrict__ p = (d.kind == kCellJoined) ? d.list_offsets : d.offsets;
      a0[k] = 0;
      a1[k] = 0;
      if (have && d.kind != kCellNumber) {
        a0[k] = p[i];
        a1[k] = p[i + 1];
      }
    }
#pragma unroll
    for (int k = 0; k < kGroupCols; ++k) {
      const CellDesc& d = tab.cell[g * kGroupCols + k];
      cb[k] = a0[k];
      cn[k] = a1[k] - a0[k];
      citems[k] = 1;
      if (d.kind == kCellJoined) {  // Array.prototype.join('|') (crew :284, actions :298): items are contiguous
        citems[k] = cn[k];
        cb[k] = 0;
        cn[k] = 0;
        if (have && citems[k] > 0) {
          cb[k] = d.offsets[a0[k]];
          cn[k] = d.offsets[a1[k]] - cb[k];
        }
      }
    }
  }

  // ---- 1b. measure the group
  uint32_t glen = 0, qmask = 0;
  if (have) {
    bool completed = false;  // entry.status === 'Completed' (:293-297); status is column 12 = group 2, k = 0
    if (g == 2 && cn[0] == 9) {
      uint32_t x[3];
      fetch_words_raw<3>(tab.cell[12].data + cb[0], 9, x);
      completed = x[0] == lit_word("Completed", 0) && x[1] == lit_word("Completed", 1) &&
                  (x[2] & 0xFFu) == lit_word("Completed", 2);
    }
    if (completed) qmask |= kCompletedBit;
#pragma unroll 1  // generic body (unrolled it is specialised per column: 28k instructions, icache-bound)
    for (int k = 0; k < kGroupCols; ++k) {
      const int col = g * kGroupCols + k;
      const CellDesc& d = tab.cell[col];
      uint32_t len = 0;
      if (d.kind == kCellNumber) {  // delaySec === null || undefined ? '' : delaySec, then String() (:301, :333)
        int nl = 0;
        if (v.delay_valid[e]) {
          const RyuTables t{d_pow5_inv, d_pow5};
          nl = js_number_to_string(v.delay_sec[e], sm.num[r], t);
        }
        sm.num_len[r] = (uint8_t)nl;
        len = (uint32_t)nl;
      } else if (!(d.blank_if_completed && completed)) {
        len = (uint32_t)cn[k] + (citems[k] > 1 ? (uint32_t)(citems[k] - 1) : 0u);  // '|' between items: not special
        if (sm.col_dirty[col] && has_special(d.data + cb[k], cn[k])) {  // csvEscape (:332-338)
          qmask |= 1u << k;
          len += 2u + count_quotes(d.data + cb[k], cn[k]);
        }
      }
      glen += len + 1u;  // + ',' (or the final '\n')
    }
  }
  if (worker) sm.group[g][r] = glen;
  __syncthreads();

  // ---- 2. per-row scan over the groups, then block scan over the rows
  uint32_t row_len = 0;
  if (tid < kRows) {
    if (tid < rows) {
      uint32_t run = 0;
#pragma unroll
      for (int k = 0; k < kGroups; ++k) {
        const uint32_t x = sm.group[k][tid];
        sm.group[k][tid] = run;
        run += x;
      }
      row_len = run;
    }
    uint32_t incl = row_len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) sm.warp_sum[wid] = incl;
    sm.row_start[tid] = incl - row_len;  // completed below with the preceding warps' sums
  }
  __syncthreads();
  if (tid < kRows) {
    uint32_t before = 0;
    for (int w = 0; w < wid; ++w) before += sm.warp_sum[w];
    sm.row_start[tid] += before;
    if (tid == kRows - 1) {
      const uint32_t total = sm.row_start[tid] + row_len;
      sm.tile_total = total;
      __threadfence();
      reinterpret_cast<volatile unsigned long long*>(sc.tile_state)[tile] =
          (tile == 0 ? kPrefix : kAggregate) | (unsigned long long)total;  // published for later tiles
    }
  }
  __syncthreads();
  const uint32_t tile_total = sm.tile_total;
  const bool write = out_data != nullptr;
  const bool staged = tile_total <= (uint32_t)kTileBytes;

  // ---- 4. decoupled look-back, by the extra warp — WHILE the workers write the cells of a staged
  // tile (step 3 does not need the offset); before step 3 when the tile goes straight to global memory.
  auto look_back = [&]() {
    volatile unsigned long long* state = sc.tile_state;
    unsigned long long exclusive = 0;
    int64_t idx = tile - 1;
    while (idx >= 0) {
      const int64_t j = idx - lane;
      unsigned long long st;
      unsigned ns = 32;
      for (;;) {
        st = 2ull << kStatusShift;  // before tile 0: an empty prefix
        if (j >= 0) st = state[j];
        if (!__any_sync(0xFFFFFFFFu, (st >> kStatusShift) == 0)) break;
        __nanosleep(ns);  // the tiles we wait for are still measuring: do not hammer L2 / the issue slots
        if (ns < 1024) ns <<= 1;
      }
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
      sm.base = exclusive;
      if (tile == csv_tiles(v.n_entries) - 1) {
        *total_out = exclusive + tile_total;
        row_offsets[v.n_entries] = (int64_t)(bias + exclusive + tile_total);
      }
    }
  };
  const bool overlap = write && staged;
  if (!overlap) {
    if (!worker) look_back();
    __syncthreads();
  }
  if (!write) {
    if (tid < rows) row_offsets[e0 + tid] = (int64_t)(bias + sm.base + sm.row_start[tid]);
    return;
  }
  const bool fits = staged || (sm.base + tile_total <= capacity);  // direct writes respect the caller's capacity

  // ---- 3. write the group
  if (!worker) {
    if (overlap) look_back();
  } else if (have && fits) {
    StreamWriter out;
    out.init((staged ? s_tile : (out_data + sm.base)) + sm.row_start[r] + sm.group[g][r]);
    const bool completed = (qmask & kCompletedBit) != 0;
#pragma unroll 1
    for (int k = 0; k < kGroupCols; ++k) {
      const int col = g * kGroupCols + k;
      const CellDesc& d = tab.cell[col];
      const uint8_t sep = (col == kCols - 1) ? (uint8_t)'\n' : (uint8_t)',';
      const bool quote = (qmask >> k) & 1u;
      if (d.kind == kCellNumber) {
        const int nl = sm.num_len[r];
        for (int j = 0; j < nl; ++j) out.put((uint8_t)sm.num[r][j]);
        out.put(sep);
      } else if (d.blank_if_completed && completed) {
        out.put(sep);
      } else if (d.kind == kCellString) {
        if (!quote) {
          copy_plain(out, d.data + cb[k], cn[k], sep);
        } else {
          out.put('"');
          copy_quoted_bytes(out, d.data + cb[k], cn[k]);
          out.put('"');
          out.put(sep);
        }
      } else {
        const int64_t i = d.per_entry ? e : show;
        const int l0 = citems[k] > 0 ? d.list_offsets[i] : 0, l1 = l0 + citems[k];
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
      }
    }
    out.finish();
  }
  if (staged) __syncthreads();  // the tile is complete in shared memory and its offset is known
  const unsigned long long base = sm.base;
  if (tid < rows) row_offsets[e0 + tid] = (int64_t)(bias + base + sm.row_start[tid]);
  if (!staged || base + tile_total > capacity) return;

  // ---- 5. flush s_tile[0 .. tile_total) -> out_data[base ..) with 16-byte stores.  Global chunk k
  // starts at the first 16-byte boundary >= out_data+base, i.e. at tile offset head + 16k, which has
  // an arbitrary phase in shared memory: read 5 aligned words and funnel-shift.
  uint8_t* __restrict__ dst = out_data + base;
  const uint32_t head_raw = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15)) & 15u;
  const uint32_t head = head_raw < tile_total ? head_raw : tile_total;
  for (uint32_t k = tid; k < head; k += kCtaThreads) dst[k] = s_tile[k];
  const uint32_t n_chunks = (tile_total - head) >> 4;
  const uint32_t sh = (head & 3u) * 8u;
  const uint32_t* __restrict__ sw = reinterpret_cast<const uint32_t*>(s_tile) + (head >> 2);
  for (uint32_t k = tid; k < n_chunks; k += kCtaThreads) {
    const uint32_t* w = sw + 4 * k;
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4];  // w[4] stays inside the +32 slack
    uint4 o;
    o.x = __funnelshift_r(w0, w1, sh);
    o.y = __funnelshift_r(w1, w2, sh);
    o.z = __funnelshift_r(w2, w3, sh);
    o.w = __funnelshift_r(w3, w4, sh);
    *reinterpret_cast<uint4*>(dst + head + 16u * k) = o;
  }
  for (uint32_t k = head + 16u * n_chunks + tid; k < tile_total; k += kCtaThreads) dst[k] = s_tile[k];
}

cudaError_t launch_csv_rows(const pie_archive_view& v, int64_t* row_offsets, uint8_t* out_data, uint64_t capacity,
                            unsigned long long bias, unsigned long long* total_out, void* scratch,
                            cudaStream_t stream) {
  CsvScratch sc = carve_csv(scratch, v.n_entries);
  cudaError_t err = cudaMemsetAsync(scratch, 0, csv_scratch_zero_bytes(v.n_entries), stream);
  if (err != cudaSuccess) return err;
  if (v.n_entries == 0) {
    err = cudaMemsetAsync(total_out, 0, 8, stream);
    if (err != cudaSuccess) return err;
    return cudaMemcpyAsync(row_offsets, total_out, 8, cudaMemcpyDeviceToDevice, stream);  // 0; the caller adds its bias
  }
  const int smem = kTileBytes + 32 + (int)sizeof(CsvSmem);
  static int configured_device = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_device != dev) {
    err = cudaFuncSetAttribute(csv_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err != cudaSuccess) return err;
    configured_device = dev;
  }
  expand_entry_show_kernel<<<(unsigned)((v.n_shows + 255) / 256), 256, 0, stream>>>(v, sc.entry_show);
  const RowTable tab = make_row_table(v);
  {
    int64_t blocks = (v.n_entries * 3 + 255) / 256;  // ~ a 16-byte chunk per thread for the widest heaps
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    column_dirty_kernel<<<dim3((unsigned)blocks, kCols), 256, 0, stream>>>(tab, v.n_shows, v.n_entries, sc.col_dirty);
  }
  csv_rows_kernel<<<(unsigned)csv_tiles(v.n_entries), kCtaThreads, smem, stream>>>(v, tab, sc, row_offsets,
                                                                                out_data, capacity, bias, total_out);
  g_launches += 3;
  return cudaGetLastError();
}

}  // namespace pie
