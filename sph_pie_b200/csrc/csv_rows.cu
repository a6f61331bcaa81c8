// Export rows on sm_100a.
// Replaces buildTableRow + csvEscape + buildCsvRow (reference server/webhookDispatcher.js:276-342,
// twin public/app.js:5582-5612, :6025-6034) mapped over every entry of every show — what
// dispatchShowEvent puts in csv.rows (:571) and exportShowAsCsv joins with '\n' (public/app.js:5567).
//
// The same kernel, with another row format, writes JSON.stringify(buildArchiveEntryPayload(show, entry)) per
// entry (:315-330, the bodies dispatchShowEvent('show.archived') posts, :520-540): see RowTable below.
//
// Output = one string column: row i is out_data[row_offsets[i] .. row_offsets[i+1]-1) and is followed
// by one '\n', so a show's CSV body is a single contiguous slice (and the payload output is JSON Lines).
//
// ONE pass over the inputs, by a PERSISTENT, warp-specialised kernel.  A CTA loops over
// tiles of kRows consecutive entries (claimed from a counter, so tile ids start in order):
//   producer warp   one tile AHEAD of the workers.  Everything a tile reads is a handful of contiguous
//                   ranges — per column the bytes of its heap and the slice of its offsets array, plus
//                   delaySec / validity / the rows' show indices — so the warp resolves the ~50 range
//                   ends (the only dependent global loads of the kernel) and issues one TMA bulk copy
//                   (cp.async.bulk, completion on an mbarrier) per range into a two-stage shared-memory
//                   pool; the < 16 bytes past a range's last full 16-byte chunk are copied as words so
//                   nothing is read past an array's end.
//   16 worker warps never touch global memory for input:
//     cells  every cell becomes a plain (src, len) pair in the stage: show-level cells are normalised
//            once per show of the tile, not once per row; cells that need csvEscape or
//            Array.join('|') are materialised in a bump area behind the staged bytes;
//            Number::toString (Ryu) output lives next to it; cells blanked by status === 'Completed'
//            get len 0.
//     scan   per-row and per-tile sizes; the tile's aggregate is published.
//     write  thread (row, group of 6 consecutive columns) streams its cells into the shared output tile
//            through a byte accumulator (aligned 32-bit stores; byte stores only for the group's first
//            and last word).  A warp = 32 consecutive rows on the same column at every step, so the
//            cells it handles together are alike in length.
//     flush  16-byte coalesced stores; the tile is re-aligned to the destination with a funnel shift.
//   look-back warp  decoupled look-back over the published tile totals while the workers write.
// A tile that does not fit (very long free text, more than kMaxTileShows shows) takes a slow path:
// a warp per row, lanes striding over the bytes of a cell, straight from / to global memory.
#include "pie_device.cuh"
#include "pie_kernels.h"
#include "pie_numfmt.cuh"
#ifdef PIE_CSV_PROFILE
#include <cstdio>
#endif

namespace pie {

__device__ const uint64_t d_pow5_inv[PIE_RYU_POW5_INV_SPLIT_N][2] = PIE_RYU_POW5_INV_SPLIT_INIT;
__device__ const uint64_t d_pow5[PIE_RYU_POW5_SPLIT_N][2] = PIE_RYU_POW5_SPLIT_INIT;

#ifndef PIE_CSV_ROWS
#define PIE_CSV_ROWS 160
#endif
#ifndef PIE_CSV_MIN_BLOCKS
#define PIE_CSV_MIN_BLOCKS 1
#endif
#ifndef PIE_CSV_OUT_KB
#define PIE_CSV_OUT_KB 52
#endif
#ifndef PIE_CSV_STAGE_KB
#define PIE_CSV_STAGE_KB 58
#endif
constexpr int kRows = PIE_CSV_ROWS;            // rows (entries) per tile; multiple of 32
constexpr int kCols = PIE_N_EXPORT_COLUMNS;    // 24
constexpr int kGroups = 4;
constexpr int kGroupCols = kCols / kGroups;    // 6
constexpr int kWorkers = kRows * kGroups;      // worker threads: (row, group of 6 columns)
constexpr int kWorkerWarps = kWorkers / 32;
constexpr int kProducerWarp = kWorkerWarps;    // stages the next tile
constexpr int kLookbackWarp = kWorkerWarps + 1;
constexpr int kNumberWarp0 = kWorkerWarps + 2;  // Number::toString for the NEXT tile: a row per lane
constexpr int kNumberWarps = kRows / 32;
constexpr int kCtaThreads = kWorkers + 64 + 32 * kNumberWarps;
constexpr int kOutBytes = PIE_CSV_OUT_KB * 1024;      // shared output tile (rows of ~280 B -> ~45 KB per 160 rows)
constexpr int kStageBytes = PIE_CSV_STAGE_KB * 1024;  // one stage: column bytes + offset arrays + bump area
constexpr int kNumBytes = kRows * kMaxNumberChars;    // Number::toString output, behind the stage
constexpr int kStageStride = kStageBytes + kNumBytes + 16;
constexpr int kMaxTileShows = kRows;           // shows a tile may span on the fast path
constexpr int kCellStride = kCols + 1;         // padded: lanes = consecutive rows hit distinct banks
static_assert(kRows % 32 == 0 && kGroups * kGroupCols == kCols, "tile shape");
static_assert(kStageBytes + kNumBytes <= 65536, "cell sources are 16-bit offsets into a stage");
static_assert(kStageStride % 16 == 0, "stages are 16-byte aligned");

constexpr unsigned long long kStatusShift = 62;
constexpr unsigned long long kValueMask = (1ull << kStatusShift) - 1;
constexpr unsigned long long kAggregate = 1ull << kStatusShift;
constexpr unsigned long long kPrefix = 2ull << kStatusShift;

struct CsvScratch {
  unsigned long long* tile_state;  // [n_tiles] packed (status, value); zeroed before launch
  unsigned int* tile_counter;      // [1] dynamic tile ids; zeroed before launch
  unsigned int* slow_tiles;        // [1] tiles that took the slow path (diagnostics); zeroed before launch
  unsigned int* tile_rows;         // [1] rows per tile of this launch (plan_tile_rows), 32 .. kRows
  unsigned int* col_dirty;         // [24] the sample of column c met a byte to escape (plan_tile_rows only); zeroed before launch
  unsigned int* col_dirty_chunks;  // [24] how many of its sampled 16-byte chunks did; zeroed before launch
  unsigned int* col_sampled_chunks;  // [24] how many chunks were sampled; zeroed before launch
  int32_t* entry_show;             // [n_entries]
};

static inline uint64_t align256(uint64_t x) { return (x + 255) & ~(uint64_t)255; }
constexpr uint64_t kCtlBytes = 2048;  // tile counter, slow-tile counter, column flags; developer counters from byte 256
constexpr int kMinTileRows = 32;  // a launch on wide rows uses tiles of fewer rows (plan_tile_rows)
// tiles a batch can have at most (the tile-state array is sized for it)
__host__ __device__ static inline int64_t csv_tiles(int64_t n_entries) { return (n_entries + kMinTileRows - 1) / kMinTileRows; }

uint64_t csv_scratch_bytes(int64_t n_entries) {
  const uint64_t e = (uint64_t)(n_entries > 0 ? n_entries : 1);
  return align256(8 * (uint64_t)csv_tiles(e)) + kCtlBytes + align256(4 * e);
}
uint64_t csv_scratch_zero_bytes(int64_t n_entries) {  // leading part that must be zero at launch
  const uint64_t e = (uint64_t)(n_entries > 0 ? n_entries : 1);
  return align256(8 * (uint64_t)csv_tiles(e)) + kCtlBytes;
}
static CsvScratch carve_csv(void* scratch, int64_t n_entries) {
  const uint64_t e = (uint64_t)(n_entries > 0 ? n_entries : 1);
  uint8_t* p = static_cast<uint8_t*>(scratch);
  CsvScratch s;
  s.tile_state = (unsigned long long*)p; p += align256(8 * (uint64_t)csv_tiles(e));
  s.tile_counter = (unsigned int*)p;
  s.slow_tiles = (unsigned int*)(p + 16);
  s.tile_rows = (unsigned int*)(p + 32);
  s.col_dirty = (unsigned int*)(p + 64);
  s.col_dirty_chunks = (unsigned int*)(p + 160);
  s.col_sampled_chunks = (unsigned int*)(p + 768);  // behind the developer counters (bytes 256 .. 767)
  p += kCtlBytes;
  s.entry_show = (int32_t*)p;
  return s;
}

static int g_force_slow = 0;  // tests: every tile through the slow path
int csv_set_force_slow(int on) {
  const int old = g_force_slow;
  if (on >= 0) g_force_slow = on ? 1 : 0;
  return old;
}
cudaError_t csv_read_slow_tiles(const void* scratch, int64_t n_entries, unsigned int* out, cudaStream_t stream) {
  CsvScratch sc = carve_csv(const_cast<void*>(scratch), n_entries);
#ifdef PIE_CSV_PROFILE
  {
    unsigned long long ph[64];
    cudaMemcpy(ph, sc.tile_counter + 64, sizeof(ph), cudaMemcpyDeviceToHost);
    const long long tiles = (n_entries + kRows - 1) / kRows;  // of full tiles (a launch on wide rows has more)
    for (int i = 0; i < 12; ++i)
      fprintf(stderr, "phase %2d: cycles per tile, first thread of group 0..3: %7llu %7llu %7llu %7llu\n", i,
              ph[i] / tiles, ph[16 + i] / tiles, ph[32 + i] / tiles, ph[48 + i] / tiles);
  }
#endif
  cudaError_t err = cudaMemcpyAsync(out, sc.slow_tiles, 4, cudaMemcpyDeviceToHost, stream);
  if (err != cudaSuccess) return err;
  return cudaStreamSynchronize(stream);
}

// show index of every entry (rows of show s are entry_offsets[s] .. entry_offsets[s+1])
__global__ void __launch_bounds__(256) expand_entry_show_kernel(pie_archive_view v, int32_t* __restrict__ entry_show) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= v.n_shows) return;
  for (int e = v.entry_offsets[s]; e < v.entry_offsets[s + 1]; ++e) entry_show[e] = (int32_t)s;
}

// ---- word-wise scanning ---------------------------------------------------------------------------
// != 0 iff some byte of v is zero (exact as a boolean)
__device__ __forceinline__ uint32_t zero_byte_flags(uint32_t v) { return (v - 0x01010101u) & ~v & 0x80808080u; }

// != 0 iff some byte of x is below 0x20 (exact as a boolean)
__device__ __forceinline__ uint32_t control_byte_flags(uint32_t x) { return (x - 0x20202020u) & ~x & 0x80808080u; }

// Which bytes make a cell "special"?
//   kEscapeCsv   " , \n \r  force quotes, '"' is doubled            (csvEscape, webhookDispatcher.js:332-338)
//   kEscapeJson  " \ and U+0000..U+001F are escaped                  (JSON.stringify, ECMA-262 QuoteJSONString)
enum : uint8_t { kEscapeCsv = 0, kEscapeJson = 1 };
template <bool kJson>
__device__ __forceinline__ uint32_t special_flags(uint32_t x) {
  if (kJson) return zero_byte_flags(x ^ 0x22222222u) | zero_byte_flags(x ^ 0x5C5C5C5Cu) | control_byte_flags(x);
  return zero_byte_flags(x ^ 0x22222222u) | zero_byte_flags(x ^ 0x2C2C2C2Cu) | zero_byte_flags(x ^ 0x0A0A0A0Au) |
         zero_byte_flags(x ^ 0x0D0D0D0Du);
}
template <bool kJson>
__device__ __forceinline__ bool is_special_byte(uint8_t c) {
  if (kJson) return c == '"' || c == '\\' || c < 0x20;
  return c == '"' || c == ',' || c == '\n' || c == '\r';
}
// bytes a special byte ADDS to the output: CSV doubles '"' (the two enclosing quotes are counted per cell);
// JSON: \" \\ \b \t \n \f \r take 2, the other control characters \u00XX take 6
template <bool kJson>
__device__ __forceinline__ uint32_t extra_bytes(uint8_t c) {
  if (!kJson) return c == '"';
  if (c == '"' || c == '\\') return 1;
  if (c >= 0x20) return 0;
  return (c == 8 || c == 9 || c == 10 || c == 12 || c == 13) ? 1u : 5u;
}
// writes the escaped form of c at dst, returns its length (1, 2 or 6)
template <bool kJson>
__device__ __forceinline__ uint32_t put_escaped(uint8_t* dst, uint8_t c) {
  if (!kJson) {
    if (c == '"') {
      dst[0] = '"';
      dst[1] = '"';
      return 2;
    }
    dst[0] = c;
    return 1;
  }
  if (c == '"' || c == '\\') {
    dst[0] = '\\';
    dst[1] = c;
    return 2;
  }
  if (c >= 0x20) {
    dst[0] = c;
    return 1;
  }
  const char short_form = c == 8 ? 'b' : c == 9 ? 't' : c == 10 ? 'n' : c == 12 ? 'f' : c == 13 ? 'r' : 0;
  dst[0] = '\\';
  if (short_form) {
    dst[1] = (uint8_t)short_form;
    return 2;
  }
  dst[1] = 'u';
  dst[2] = '0';
  dst[3] = '0';
  dst[4] = (uint8_t)('0' + (c >> 4));
  dst[5] = (uint8_t)((c & 15) < 10 ? '0' + (c & 15) : 'a' + (c & 15) - 10);
  return 6;
}

// A row is kCols cells, each followed by ONE separator byte.  The two row formats of the path:
//   CSV      buildCsvRow(buildTableRow(show, entry)): the 24 EXPORT_COLUMNS (webhookDispatcher.js:15-19), ',' between
//            them, '\n' after the last.
//   payload  JSON.stringify(buildArchiveEntryPayload(show, entry)) (webhookDispatcher.js:315-330, the body
//            dispatchShowEvent posts per entry of an archived show, :527-540): 12 values between literal key text;
//            a literal's last byte rides as its separator, and the three booleans select between two literals
//            that also carry the next key.
enum : uint8_t { kCellString = 0, kCellJoined = 1, kCellNumber = 2, kCellLiteral = 3, kCellYesNo = 4 };
constexpr int kLiteralBytes = 256;  // literal text of a row format; staged at the start of every stage
constexpr int kMaxShowSlots = 8;    // show-level value cells of a row format
struct CellDesc {
  const int32_t* offsets;       // string column / items of a list column
  const uint8_t* data;
  const int32_t* list_offsets;  // kCellJoined only
  uint8_t kind;
  uint8_t per_entry;            // row index is the entry (1) or its show (0)
  uint8_t blank_if_completed;   // :293-297
  uint8_t sep;                  // the byte that follows the cell
  int8_t show_slot;             // show-level value cell: its row in the per-tile show-cell table; else -1
  uint8_t pad_[3];
  uint16_t lit, lit_len;        // kCellLiteral: its text in RowTable::literals; kCellYesNo: the text for true
  uint16_t lit_no, lit_no_len;  // kCellYesNo: the text for false
};
struct RowTable {
  CellDesc cell[kCols];
  signed char owned[4][6];          // value cells of the ENTRY level that worker group g prepares (-1 ends)
  uint8_t show_col[kMaxShowSlots];  // column of show slot i
  uint8_t n_show_slots;
  int8_t status_col;                // the column `blank_if_completed` looks at, -1 if the format has none
  uint8_t json;                     // escape mode of the value cells
  uint8_t has_number;               // the format has a delaySec cell (number warps, delay ranges)
  uint8_t literals[kLiteralBytes];
};

static void finish_row_table(RowTable& t) {
  t.n_show_slots = 0;
  t.has_number = 0;
  for (int c = 0; c < kCols; ++c) {
    if (t.cell[c].kind == kCellNumber) t.has_number = 1;
    t.cell[c].show_slot = -1;
    if (!t.cell[c].per_entry && (t.cell[c].kind == kCellString || t.cell[c].kind == kCellJoined)) {
      t.cell[c].show_slot = (int8_t)t.n_show_slots;
      t.show_col[t.n_show_slots++] = (uint8_t)c;
    }
  }
}

static RowTable make_csv_table(const pie_archive_view& v) {
  RowTable t{};
  auto str = [](const pie_strcol& c, int per_entry, int blank = 0) {
    CellDesc d{};
    d.offsets = c.offsets; d.data = c.data; d.kind = kCellString; d.per_entry = (uint8_t)per_entry;
    d.blank_if_completed = (uint8_t)blank;
    return d;
  };
  auto lst = [](const pie_strlistcol& c, int per_entry) {
    CellDesc d{};
    d.offsets = c.items.offsets; d.data = c.items.data; d.list_offsets = c.list_offsets; d.kind = kCellJoined;
    d.per_entry = (uint8_t)per_entry;
    return d;
  };
  t.cell[0] = str(v.show_id, 0);      t.cell[1] = str(v.show_date, 0);     t.cell[2] = str(v.show_time, 0);
  t.cell[3] = str(v.show_label, 0);   t.cell[4] = lst(v.crew, 0);          t.cell[5] = str(v.lead_pilot, 0);
  t.cell[6] = str(v.monkey_lead, 0);  t.cell[7] = str(v.show_notes, 0);    t.cell[8] = str(v.entry_id, 1);
  t.cell[9] = str(v.unit_id, 1);      t.cell[10] = str(v.planned, 1);      t.cell[11] = str(v.launched, 1);
  t.cell[12] = str(v.status, 1);      t.cell[13] = str(v.primary_issue, 1, 1);
  t.cell[14] = str(v.sub_issue, 1, 1);  t.cell[15] = str(v.other_detail, 1, 1);
  t.cell[16] = str(v.severity, 1, 1);   t.cell[17] = str(v.root_cause, 1, 1);
  t.cell[18] = lst(v.actions, 1);     t.cell[19] = str(v.operator_name, 1); t.cell[20] = str(v.battery_id, 1);
  t.cell[21] = CellDesc{};
  t.cell[21].kind = kCellNumber;
  t.cell[21].per_entry = 1;
  t.cell[22] = str(v.command_rx, 1);  t.cell[23] = str(v.notes, 1);
  for (int c = 0; c < kCols; ++c) t.cell[c].sep = (c == kCols - 1) ? (uint8_t)'\n' : (uint8_t)',';
  // four value cells per worker group, the ones that usually need a scan or more (primary issue, actions, other
  // detail, notes) on different groups; delaySec (21, formatted by the number warps, possibly still in flight)
  // last on its group
  const signed char owned[4][6] = {
      {8, 13, 16, 21, -1, -1}, {9, 14, 17, 18, -1, -1}, {10, 11, 12, 15, -1, -1}, {19, 20, 22, 23, -1, -1}};
  for (int g = 0; g < 4; ++g)
    for (int k = 0; k < 6; ++k) t.owned[g][k] = owned[g][k];
  t.status_col = 12;
  t.json = 0;
  finish_row_table(t);
  return t;
}

// {"showDate":"..","showTime":"..","showNumber":"..","leadPilot":"..","monkeyLead":"..","operator":"..",
//  "monkeyId":"..","planned":b,"launched":b,"commandReceived":b,"primaryIssue":"..","subIssue":".."}\n
// — the property order of the object literal at webhookDispatcher.js:316-329, as JSON.stringify emits it.
static RowTable make_payload_table(const pie_archive_view& v) {
  RowTable t{};
  int used = 0;
  auto lit_at = [&](const char* text, uint16_t& off, uint16_t& len) {  // text without its last byte; returns that byte
    int n = 0;
    while (text[n]) ++n;
    off = (uint16_t)used;
    len = (uint16_t)(n - 1);
    for (int i = 0; i < n - 1; ++i) t.literals[used++] = (uint8_t)text[i];
    return (uint8_t)text[n - 1];
  };
  int c = 0;
  auto literal = [&](const char* text) {
    CellDesc d{};
    d.kind = kCellLiteral;
    d.per_entry = 1;
    d.sep = lit_at(text, d.lit, d.lit_len);
    t.cell[c++] = d;
  };
  auto value = [&](const pie_strcol& col, int per_entry, char sep) {
    CellDesc d{};
    d.offsets = col.offsets; d.data = col.data; d.kind = kCellString; d.per_entry = (uint8_t)per_entry;
    d.sep = (uint8_t)sep;
    t.cell[c++] = d;
  };
  auto yes_no = [&](const pie_strcol& col, const char* if_true, const char* if_false) {  // both end in the same byte
    CellDesc d{};
    d.offsets = col.offsets; d.data = col.data; d.kind = kCellYesNo; d.per_entry = 1;
    d.sep = lit_at(if_true, d.lit, d.lit_len);
    lit_at(if_false, d.lit_no, d.lit_no_len);
    t.cell[c++] = d;
  };
  literal("{\"showDate\":\"");        value(v.show_date, 0, '"');
  literal(",\"showTime\":\"");        value(v.show_time, 0, '"');
  literal(",\"showNumber\":\"");      value(v.show_label, 0, '"');
  literal(",\"leadPilot\":\"");       value(v.lead_pilot, 0, '"');
  literal(",\"monkeyLead\":\"");      value(v.monkey_lead, 0, '"');
  literal(",\"operator\":\"");        value(v.operator_name, 1, '"');
  literal(",\"monkeyId\":\"");        value(v.unit_id, 1, '"');
  literal(",\"planned\"");            literal(":");
  yes_no(v.planned, "true,\"launched\":", "false,\"launched\":");
  yes_no(v.launched, "true,\"commandReceived\":", "false,\"commandReceived\":");
  yes_no(v.command_rx, "true,\"primaryIssue\":\"", "false,\"primaryIssue\":\"");
  value(v.primary_issue, 1, '"');
  literal(",\"subIssue\":\"");        value(v.sub_issue, 1, '"');
  literal("}");                        literal("\n");
  // c == 24 by construction; value cells of the entry level: 11, 13 | 16, 17 | 18, 19 | 21
  const signed char owned[4][6] = {
      {11, 13, -1, -1, -1, -1}, {16, 17, -1, -1, -1, -1}, {18, 19, -1, -1, -1, -1}, {21, -1, -1, -1, -1, -1}};
  for (int g = 0; g < 4; ++g)
    for (int k = 0; k < 6; ++k) t.owned[g][k] = owned[g][k];
  t.status_col = -1;
  t.json = 1;
  finish_row_table(t);
  return t;
}
static_assert(kCols == 24, "the row formats above are laid out for 24 cells");

// ---- rows per tile of a launch ------------------------------------------------------------------------
// A tile's column bytes, offset slices and output must fit the shared-memory stage and output tile; tiles
// that do not take the slow path, which is ~20x slower.  So a batch of wide rows (long free text) is cut into
// tiles of fewer rows: from the AVERAGE bytes per row of the batch (one thread reads the ~50 end offsets),
// with a margin; a tile that is still too large goes the slow way.
__global__ void plan_tile_rows_kernel(const __grid_constant__ RowTable tab, int64_t n_shows, int64_t n_entries,
                                      int stage_bytes, int out_bytes, const unsigned int* __restrict__ col_dirty,
                                      const unsigned int* __restrict__ col_dirty_chunks,
                                      const unsigned int* __restrict__ col_sampled_chunks,
                                      unsigned int* __restrict__ tile_rows) {
  // one warp, lane c = column c (the end offsets of the columns are read in parallel), sums by shuffles
  const int c = threadIdx.x;
  double entry_bytes = 0, show_bytes = 0, literal_bytes = 0, arrays_per_row = 0;
  double bump_bytes = 0;  // cells that may be written out again (escaped / joined), per row
  const double E = (double)(n_entries > 0 ? n_entries : 1), S = (double)(n_shows > 0 ? n_shows : 1);
  if (c < kCols) {
    const CellDesc& d = tab.cell[c];
    if (d.kind == kCellLiteral) literal_bytes = d.lit_len;
    if (d.kind == kCellYesNo) literal_bytes = d.lit_no_len;
    if (d.kind == kCellString || d.kind == kCellJoined || d.kind == kCellYesNo) {
      const int64_t n = d.per_entry ? n_entries : n_shows;
      int64_t first = 0, last = n;
      double items = 0;
      if (d.kind == kCellJoined) {
        first = d.list_offsets[0];
        last = d.list_offsets[n];
        items = (double)(last - first);
      }
      const double bytes = (double)(d.offsets[last] - d.offsets[first]);
      // A cell that holds a byte to escape, and a list cell of several items, is written out a second time (a
      // little longer) in the bump area.  List columns: all of them.  Other columns: the share of their (sampled)
      // 16-byte chunks that hold such a byte, times 3 (a cell is a few chunks) — generous, because falling off the
      // fast path costs far more than a smaller tile; a column whose sample was clean still gets 2 %.
      double copy_share = 0;
      if (d.kind == kCellJoined) copy_share = 1;
      else if (d.kind == kCellString) {
        const double sampled = (double)col_sampled_chunks[c];
        copy_share = 0.02;
        if (col_dirty[c]) copy_share = sampled < 64 ? 1.0 : fmin(1.0, 3.0 * (double)col_dirty_chunks[c] / sampled + 0.02);
      }
      if (d.per_entry) {
        entry_bytes = bytes;
        arrays_per_row = 4.0 + 4.0 * items / E;
        bump_bytes = copy_share * (1.1 * bytes / E + 4.0);
      } else {
        show_bytes = bytes;
        arrays_per_row = (4.0 + 4.0 * items / S) * S / E;
        bump_bytes = copy_share * (1.1 * bytes / S + 4.0) * S / E;
      }
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    entry_bytes += __shfl_xor_sync(0xFFFFFFFFu, entry_bytes, o);
    show_bytes += __shfl_xor_sync(0xFFFFFFFFu, show_bytes, o);
    literal_bytes += __shfl_xor_sync(0xFFFFFFFFu, literal_bytes, o);
    arrays_per_row += __shfl_xor_sync(0xFFFFFFFFu, arrays_per_row, o);
    bump_bytes += __shfl_xor_sync(0xFFFFFFFFu, bump_bytes, o);
  }
  if (c != 0) return;
  arrays_per_row += 4 + 9;  // the rows' show indices, delaySec + validity
  // staged per row: the entry-level bytes, the row's share of its show's bytes and of the offset slices, the bump
  // area; 10 % on top for the 16-byte rounding of ~50 ranges and rows longer than the average
  const double in_per_row = 1.10 * (entry_bytes / E + show_bytes / E + arrays_per_row + bump_bytes);
  // written per row: every cell (a show's cells are repeated on each of its rows), separators, literal text; quotes
  // and escapes add, blanked cells subtract: 10 %
  const double out_per_row = 1.10 * (entry_bytes / E + show_bytes / S + kCols + literal_bytes);
  double rows = fmin((stage_bytes - kLiteralBytes - 2048) / in_per_row, out_bytes / out_per_row);
  int r = rows >= kRows ? kRows : (int)rows;
  r = (r / 32) * 32;
  if (r < kMinTileRows) r = kMinTileRows;
  *tile_rows = (unsigned int)r;
}

// ---- how much of a column needs escaping?  (an estimate, for the tile plan only) ---------------------------------
// The row kernel decides per cell from a bitmask of the staged bytes (build_special_mask below), so nothing has to
// be known about a column beforehand; only plan_tile_rows wants to know roughly which share of a column's cells will
// be written out a second time (escaped) in the bump area.  A sample is enough for that: up to kSampleChunks evenly
// spaced 16-byte chunks of every column heap (~1.5 MB read in all, instead of a sweep over every byte).
constexpr int kSampleChunks = 4096;
template <bool kJson>
__global__ void __launch_bounds__(256) column_sample_kernel(const __grid_constant__ RowTable tab, int64_t n_shows,
                                                            int64_t n_entries, unsigned int* __restrict__ col_dirty,
                                                            unsigned int* __restrict__ col_dirty_chunks,
                                                            unsigned int* __restrict__ col_sampled_chunks) {
  const int col = blockIdx.y;
  const CellDesc& d = tab.cell[col];
  if (d.kind != kCellString && d.kind != kCellJoined) return;
  const int64_t n = d.per_entry ? n_entries : n_shows;
  int64_t first = 0, last = n;  // rows of the string column that hold this cell's bytes
  if (d.kind == kCellJoined) {
    first = d.list_offsets[0];
    last = d.list_offsets[n];
  }
  const int64_t b0 = d.offsets[first], b1 = d.offsets[last];
  if (b1 <= b0) return;
  const uintptr_t a0 = reinterpret_cast<uintptr_t>(d.data + b0), a1 = reinterpret_cast<uintptr_t>(d.data + b1);
  const uintptr_t w0 = (a0 + 15) & ~static_cast<uintptr_t>(15), w1 = a1 & ~static_cast<uintptr_t>(15);
  if (w1 <= w0) {  // a heap shorter than a chunk: "dirty" is the safe answer
    if (blockIdx.x == 0 && threadIdx.x == 0) { col_dirty[col] = 1; col_dirty_chunks[col] = 1; col_sampled_chunks[col] = 1; }
    return;
  }
  const int64_t chunks = static_cast<int64_t>((w1 - w0) >> 4);
  const int64_t samples = chunks < kSampleChunks ? chunks : kSampleChunks;
  const uint4* __restrict__ p = reinterpret_cast<const uint4*>(w0);
  uint32_t dirty_chunks = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < samples; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 x = __ldg(p + (i * chunks) / samples);
    const uint32_t f = special_flags<kJson>(x.x) | special_flags<kJson>(x.y) | special_flags<kJson>(x.z) | special_flags<kJson>(x.w);
    dirty_chunks += (f != 0);
  }
  dirty_chunks = __reduce_add_sync(0xFFFFFFFFu, dirty_chunks);
  if ((threadIdx.x & 31) == 0 && dirty_chunks) {
    atomicOr(&col_dirty[col], 1u);
    atomicAdd(&col_dirty_chunks[col], dirty_chunks);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) col_sampled_chunks[col] = (unsigned int)samples;
}

// ---- PTX: mbarrier, 1-D TMA bulk copy, named barriers ----------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Watchdog for the kernel's spin waits (mbarriers, look-back): the protocol below cannot deadlock by
// construction, but a wait that outlives kWatchdogNs means it did — trap, so that the host sees a launch
// failure instead of a kernel that never returns.
constexpr unsigned long long kWatchdogNs = 10ull * 1000 * 1000 * 1000;
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
struct Watchdog {
  unsigned long long start = 0;
  uint32_t polls = 0;
  __device__ __forceinline__ void poll() {
    if ((++polls & 0x3FFu) != 0) return;
    const unsigned long long now = global_ns();
    if (start == 0) start = now;
    else if (now - start > kWatchdogNs) __trap();
  }
};
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  Watchdog dog;
  while (!mbar_try_wait(bar, parity)) dog.poll();
}
// Same for the warps that run AHEAD of the workers (producer, number warps): their waits are long and not
// on the critical path, so they sleep between polls instead of spinning on the issue slots.
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  Watchdog dog;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(400);
    dog.polls += 0x3Fu;  // a poll here stands for ~0.5 us: check the clock every 16 of them
    dog.poll();
  }
}
// global -> shared, 16-byte aligned on both sides, bytes a multiple of 16; completion on the mbarrier
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
constexpr int kBarWorkers = 1, kBarTotalReady = 2, kBarBaseReady = 3;  // named barriers (0 = __syncthreads)
constexpr int kBarItemsReady = 4, kBarFillsDone = 5;                    // workers <-> number warps
__device__ __forceinline__ void workers_sync() { asm volatile("bar.sync %0, %1;" ::"n"(kBarWorkers), "n"(kWorkers) : "memory"); }
template <int kBar>
__device__ __forceinline__ void bar_arrive_workers_and_lookback() {
  __threadfence_block();
  asm volatile("bar.arrive %0, %1;" ::"n"(kBar), "n"(kWorkers + 32) : "memory");
}
template <int kBar, int kCount>
__device__ __forceinline__ void bar_arrive_n() {
  __threadfence_block();
  asm volatile("bar.arrive %0, %1;" ::"n"(kBar), "n"(kCount) : "memory");
}
template <int kBar, int kCount>
__device__ __forceinline__ void bar_sync_n() {
  asm volatile("bar.sync %0, %1;" ::"n"(kBar), "n"(kCount) : "memory");
}
template <int kBar>
__device__ __forceinline__ void bar_sync_workers_and_lookback() {
  asm volatile("bar.sync %0, %1;" ::"n"(kBar), "n"(kWorkers + 32) : "memory");
}

// ---- shared-memory state ----------------------------------------------------------------------------
// What the producer tells the workers about a staged tile.  "staged address" = byte offset in the stage.
struct StageInfo {
  long long tile;             // < 0: no more tiles
  int32_t rows;
  int32_t show0;              // first show of the tile
  uint32_t n_tile_shows;
  uint32_t slow;              // the tile does not fit the stage: slow path
  uint32_t bump0;             // first free byte behind the staged ranges
  uint32_t heap_end;          // the staged column bytes are [kLiteralBytes, heap_end): what build_special_mask scans
  uint32_t delta[kCols];      // staged address of heap byte b of column c = delta[c] + b
  uint32_t off_base[kCols];   // staged address of the column's first offset word (offsets[e0] / offsets[show0] /
                              // list_offsets[...] for the two list columns)
  uint32_t item_base[kCols];  // list columns: staged address of items.offsets[item_first[c]]
  int32_t item_first[kCols];
  uint32_t show_idx_base;     // entry_show[e0 ..]
  uint32_t delay_base;        // delay_sec[e0 ..]
  uint32_t valid_base;        // delay_valid[e0 ..]
};
constexpr int kMaxFillItems = 2 * kRows;  // per kind and tile; more sends the tile down the slow path
// a cell whose bytes have to be produced in the bump area (see plan_cell / fill_item)
struct FillItem {
  uint32_t src_n;      // src:16 | n:16
  uint32_t dst_items;  // dst:16 | items:15 | special:1
  uint32_t io_delta;   // staged byte address of item_offsets[0] : 16 | low 16 bits of delta
};
struct CsvSmem {
  unsigned long long full[2];                   // producer -> workers: stage s holds a tile
  unsigned long long empty[2];                  // workers + number warps -> producer: stage s may be overwritten
  unsigned long long nums[2];                   // number warps -> workers: the stage's delaySec strings are ready
  unsigned long long base[2];                   // global byte offset of the tile (look-back result), by tile parity
  StageInfo info[2];
  uint32_t cell[kRows * kCellStride];           // (src:16 | len:16 << 16) of cell (r, c) at r*kCellStride + c
  uint32_t shcell[kMaxShowSlots][kMaxTileShows];  // the same for the show-level cells of the tile's shows
  uint32_t special[kStageBytes / 32 + 4];       // bit b: staged byte b of the current tile needs escaping (build_special_mask)
  uint32_t qmask[kRows];                        // slow path: per-row quote masks
  FillItem word_items[kMaxFillItems];           // cells to materialise through the word-wise stream ...
  FillItem quote_items[kMaxFillItems];          // ... and byte by byte ('"' to double)
  uint32_t n_word_items, n_quote_items;
  uint32_t group[kGroups][kRows];               // bytes of a row's group
  uint32_t row_start[2][kRows];                 // byte offset of the row inside the tile, by tile parity
  uint32_t warp_sum[kRows / 32];
  uint32_t tile_total[2];                       // by tile parity (the flush of a tile is deferred by one tile)
  long long cur_tile[2];                        // workers -> look-back warp
  uint32_t bump;                                // next free byte of the current stage's bump area
  uint32_t overflow;                            // the bump area ran out
  uint32_t fill_skip;                           // the tile's queued cells need no bytes (slow path)
  uint32_t done;
  uint8_t num_len[2][kRows];                    // by stage
};
constexpr int kSmemOffStage = kOutBytes + 32;
constexpr int kSmemOffState = kSmemOffStage + 2 * kStageStride;
constexpr int kSmemBytes = kSmemOffState + (int)sizeof(CsvSmem);
static_assert(kSmemOffStage % 16 == 0 && kSmemOffState % 8 == 0, "smem carve-up");
static_assert(kSmemBytes <= 227 * 1024, "shared memory per CTA");

__device__ __forceinline__ uint32_t pack_cell(uint32_t src, uint32_t len) { return (src & 0xFFFFu) | (len << 16); }
__device__ __forceinline__ const int32_t* stage_i32(const uint8_t* stage, uint32_t addr) {
  return reinterpret_cast<const int32_t*>(stage + addr);
}

// does stage[src .. src+n) contain a character that forces quoting?  Aligned words, ends masked.
// 0x80 in every byte of v that is zero — exact per byte (no borrow crosses bytes), so it can be counted
__device__ __forceinline__ uint32_t zero_bytes_exact(uint32_t v) {
  return ~(((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v) & 0x80808080u;
}
// Does stage[src .. src+n) contain a byte that needs escaping, and how many bytes does escaping add?
// Aligned words; the bytes of the two end words that lie outside the cell are replaced by spaces.
template <bool kJson>
__device__ __forceinline__ bool smem_scan_special(const uint8_t* stage, uint32_t src, uint32_t n, uint32_t& extra) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(stage + (src & ~3u));
  const uint32_t lead = src & 3u;
  const int nw = static_cast<int>((lead + n + 3) >> 2);
  const uint32_t tail = (lead + n) & 3u;
  uint32_t flags = 0, q = 0;
  for (int k = 0; k < nw; ++k) {
    uint32_t x = w[k];
    uint32_t keep = 0xFFFFFFFFu;
    if (k == 0) keep &= 0xFFFFFFFFu << (8 * lead);
    if (k == nw - 1 && tail) keep &= (1u << (8 * tail)) - 1u;
    x = (x & keep) | (0x20202020u & ~keep);
    if (kJson) {
      flags |= special_flags<true>(x);
    } else {
      const uint32_t zq = zero_bytes_exact(x ^ 0x22222222u);
      flags |= zq | zero_byte_flags(x ^ 0x2C2C2C2Cu) | zero_byte_flags(x ^ 0x0A0A0A0Au) | zero_byte_flags(x ^ 0x0D0D0D0Du);
      q += __popc(zq);
    }
  }
  if (kJson && flags)  // rare: count byte by byte
    for (uint32_t j = 0; j < n; ++j) q += extra_bytes<true>(stage[src + j]);
  extra = q;
  return flags != 0;
}

// Byte stream into shared memory: bytes collect in an accumulator and leave as aligned 32-bit stores.
// kSharedEdges: the first and the last word are shared with other threads' streams and are stored byte by
// byte; otherwise the stream owns whole words (a 4-byte aligned, padded allocation).
template <bool kSharedEdges>
struct ByteStream {
  uint8_t* op;    // aligned address of the word being filled
  uint32_t lo;    // its bytes so far
  uint32_t fill;  // how many (including `lead` placeholders before the first store)
  uint32_t lead;
  bool shared_word;

  __device__ __forceinline__ void init(uint8_t* start) {
    lead = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(start) & 3u);
    op = start - lead;
    lo = 0;
    fill = lead;
    shared_word = kSharedEdges && lead != 0;
  }
  __device__ __forceinline__ void store_word(uint32_t v) {
    if (kSharedEdges && shared_word) {
      for (uint32_t b = lead; b < 4; ++b) op[b] = static_cast<uint8_t>(v >> (8 * b));
      shared_word = false;
    } else {
      *reinterpret_cast<uint32_t*>(op) = v;
    }
    op += 4;
  }
  // stage[src .. src+n) followed by `nsep` (0 or 1) separator byte(s) `sep`
  __device__ __forceinline__ void append(const uint8_t* stage, uint32_t src, uint32_t n, uint32_t sep, uint32_t nsep) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(stage + (src & ~3u));
    const uint32_t sh = (src & 3u) * 8u;
    const uint32_t osh = 8u * fill, csh = 32u - osh;  // appending 4 bytes leaves `fill` unchanged
    // two source words are kept in flight ahead of the stores (a load is never moved above a store by
    // the compiler: both are shared memory); reads run up to 8 bytes past the cell, inside the stage
    uint32_t cur = 0, nxt = 0;
    if (n) {
      cur = w[0];
      nxt = w[1];
    }
    for (; n >= 4; n -= 4) {
      const uint32_t ahead = w[2];
      ++w;
      const uint32_t x = __funnelshift_r(cur, nxt, sh);
      cur = nxt;
      nxt = ahead;
      store_word(lo | (x << osh));
      lo = __funnelshift_rc(x, 0u, csh);  // clamped: fill == 0 gives 0
    }
    // the 0..3 last bytes and the separator ride in one piece of 0..4 bytes
    uint32_t x = nsep ? sep << (8 * n) : 0u;
    if (n) x |= __funnelshift_r(cur, nxt, sh) & ((1u << (8 * n)) - 1u);
    const uint32_t hi = __funnelshift_rc(x, 0u, csh);
    lo |= x << osh;
    fill += n + nsep;
    if (fill >= 4) {
      store_word(lo);
      lo = hi;
      fill -= 4;
    }
  }
  __device__ __forceinline__ void put(uint32_t byte) { append(nullptr, 0, 0, byte, 1); }
  __device__ __forceinline__ void finish() {
    if (kSharedEdges) {
      for (uint32_t b = shared_word ? lead : 0u; b < fill; ++b) op[b] = static_cast<uint8_t>(lo >> (8 * b));
    } else if (fill) {
      *reinterpret_cast<uint32_t*>(op) = lo;
    }
  }
};

// Rare path: a cell that needs csvEscape (:332-338) and / or Array.prototype.join('|') (:284, :298) is
// written out in the bump area.  It is done in two steps so that no warp pays for its slowest lane:
//   plan_cell  (by the thread that owns the cell) scans the bytes, computes the FINAL length, reserves the
//              bytes and queues the work — the cell table is final right away;
//   fill_*     (one queued item per thread, after a barrier) produces the bytes.  Cells with '"' to double
//              go byte by byte, everything else through the word-wise stream; the two kinds sit in
//              separate queues, taken from opposite ends of the CTA, so a warp runs one kind only.
// stage[src .. src+n) = the cell's staged bytes (for a list: all its items, which are contiguous in the
// heap); items > 1 inserts '|' at the item boundaries, item_offsets[1 ..] in heap coordinates (+ delta =
// staged).
// CSV: a special cell is wrapped in '"' and its '"' are doubled.  JSON: the specials are escaped in place
// (the enclosing quotes belong to the row format).  `extra` = bytes the escaping adds inside the cell.
// Exact per-byte flags (0x80 in every byte of x that needs escaping), for the bitmask below — the *_flags tests
// above are exact only as "any byte of the word".
__device__ __forceinline__ uint32_t control_bytes_exact(uint32_t x) {  // bytes below 0x20
  return ~(((x & 0x7F7F7F7Fu) + 0x60606060u) | x) & 0x80808080u;
}
template <bool kJson>
__device__ __forceinline__ uint32_t special_bytes_exact(uint32_t x) {
  if (kJson) return zero_bytes_exact(x ^ 0x22222222u) | zero_bytes_exact(x ^ 0x5C5C5C5Cu) | control_bytes_exact(x);
  // CSV: '"' ',' LF CR.  On the low seven bits of every byte, (b ^ c) + 0x7F carries into bit 7 unless b == c: an XOR
  // and an ADD per character, the four results ANDed; a byte with its own bit 7 set is never one of the four.
  const uint32_t y = x & 0x7F7F7F7Fu;
  const uint32_t differs = ((y ^ 0x22222222u) + 0x7F7F7F7Fu) & ((y ^ 0x2C2C2C2Cu) + 0x7F7F7F7Fu) &
                           ((y ^ 0x0A0A0A0Au) + 0x7F7F7F7Fu) & ((y ^ 0x0D0D0D0Du) + 0x7F7F7F7Fu);
  return ~(differs | x) & 0x80808080u;
}
// the four 0x80 flags of a word as a nibble (bit j = byte j)
__device__ __forceinline__ uint32_t flags_to_nibble(uint32_t f) { return ((f >> 7) * 0x01020408u) >> 24 & 0xFu; }

// One bit per staged byte of the tile: does it need escaping?  Built by the workers when a stage has landed (a
// 16-byte chunk per thread and step, ~35 KB per tile), it answers "does this cell need csvEscape / JSON escapes" for
// EVERY cell of EVERY column with two loads and a shift.  No column is known
// to be clean beforehand, and no pre-pass over the archive is needed to find out (the round-1 kernel swept every
// column heap once per launch: 1.5 GB of extra reads per step).
template <bool kJson>
__device__ __forceinline__ void build_special_mask(uint32_t* __restrict__ mask, const uint8_t* __restrict__ stage,
                                                   uint32_t heap_end, int tid) {
  const uint32_t first = (uint32_t)kLiteralBytes >> 4, last = (heap_end + 15u) >> 4;  // 16-byte chunks
  uint16_t* m16 = reinterpret_cast<uint16_t*>(mask);
  const uint4* chunk = reinterpret_cast<const uint4*>(stage);
  for (uint32_t i = first + (uint32_t)tid; i < last; i += kWorkers) {
    const uint4 x = chunk[i];
    const uint32_t bits = flags_to_nibble(special_bytes_exact<kJson>(x.x)) | flags_to_nibble(special_bytes_exact<kJson>(x.y)) << 4 |
                          flags_to_nibble(special_bytes_exact<kJson>(x.z)) << 8 | flags_to_nibble(special_bytes_exact<kJson>(x.w)) << 12;
    m16[i] = (uint16_t)bits;
  }
}
// any bit of mask[src .. src+n) set?  (n >= 1)  Cells of up to 32 bytes — nearly all — take two loads and a funnel
// shift whatever their phase against the mask's words: no branch on where the cell happens to start.
__device__ __forceinline__ bool mask_any(const uint32_t* __restrict__ mask, uint32_t src, uint32_t n) {
  uint32_t w = src >> 5;
  const uint32_t sh = src & 31u;
  const uint32_t window = __funnelshift_r(mask[w], mask[w + 1], sh);  // bits src .. src+31
  if (n <= 32u) return (window & (0xFFFFFFFFu >> (32u - n))) != 0;
  uint32_t acc = window;
  n -= 32u;
  src += 32u;
  for (w = src >> 5; n >= 32u; n -= 32u, ++w) acc |= __funnelshift_r(mask[w], mask[w + 1], sh);
  if (n) acc |= __funnelshift_r(mask[w], mask[w + 1], sh) & (0xFFFFFFFFu >> (32u - n));
  return acc != 0;
}

template <bool kJson>
__device__ __forceinline__ uint32_t plan_cell(CsvSmem& sm, uint8_t* stage, const int32_t* item_offsets, uint32_t delta,
                                              uint32_t src, uint32_t n, int items) {
  uint32_t extra = 0;
  const bool special = n > 0 && mask_any(sm.special, src, n);
  if (!special && items <= 1) return pack_cell(src, n);
  if (special) smem_scan_special<kJson>(stage, src, n, extra);  // how many bytes escaping adds (the rare path)
  const uint32_t out_len = n + (items > 1 ? (uint32_t)(items - 1) : 0u) + extra + ((special && !kJson) ? 2u : 0u);
  const uint32_t alloc = (out_len + 3u) & ~3u;      // the word-wise stream may fill its last word
  const uint32_t p = atomicAdd(&sm.bump, alloc);    // stays 4-byte aligned
  const bool bytewise = extra != 0;
  const uint32_t slot = atomicAdd(bytewise ? &sm.n_quote_items : &sm.n_word_items, 1u);
  if (p + alloc > (uint32_t)kStageBytes || slot >= (uint32_t)kMaxFillItems || items > 0x7FFF) {
    sm.overflow = 1;  // benign race: every writer stores 1
    return 0;
  }
  FillItem& f = (bytewise ? sm.quote_items : sm.word_items)[slot];
  f.src_n = src | (n << 16);
  f.dst_items = p | ((uint32_t)items << 16) | ((special && !kJson) ? 0x80000000u : 0u);  // bit 31: wrap in quotes
  f.io_delta = ((uint32_t)(reinterpret_cast<const uint8_t*>(item_offsets) - stage) & 0xFFFFu) | (delta << 16);
  return pack_cell(p, out_len);
}
template <bool kJson>
__device__ __forceinline__ void fill_item(uint8_t* stage, const FillItem& f, bool bytewise) {
  const uint32_t src = f.src_n & 0xFFFFu, n = f.src_n >> 16;
  const uint32_t p = f.dst_items & 0xFFFFu;
  const int items = (int)((f.dst_items >> 16) & 0x7FFFu);
  const bool wrap = (f.dst_items >> 31) != 0;
  const int32_t* item_offsets = reinterpret_cast<const int32_t*>(stage + (f.io_delta & 0xFFFFu));
  const uint32_t delta16 = f.io_delta >> 16;  // staged addresses are < 2^16: the low half of delta is enough
  if (!bytewise) {  // ['"'] item ['|' item]... ['"'] through the word-wise stream
    ByteStream<false> out;
    out.init(stage + p);
    if (wrap) out.put('"');
    uint32_t ib = src;
    for (int it = 0; it < items; ++it) {
      const bool last = it + 1 >= items;
      const uint32_t ie = last ? src + n : ((delta16 + (uint32_t)item_offsets[it + 1]) & 0xFFFFu);
      out.append(stage, ib, ie - ib, last ? (uint32_t)'"' : (uint32_t)'|', (last && !wrap) ? 0u : 1u);
      ib = ie;
    }
    out.finish();
    return;
  }
  uint32_t q = p;  // bytes to escape: one by one
  if (wrap) stage[q++] = '"';
  uint32_t ib = src;
  for (int it = 0; it < items; ++it) {
    const uint32_t ie = (it + 1 < items) ? ((delta16 + (uint32_t)item_offsets[it + 1]) & 0xFFFFu) : src + n;
    for (uint32_t j = ib; j < ie; ++j) q += put_escaped<kJson>(stage + q, stage[j]);
    if (it + 1 < items) stage[q++] = '|';
    ib = ie;
  }
  if (wrap) stage[q++] = '"';
}

// toYesNoBoolean of a string (webhookDispatcher.js:60-77): value.trim().toLowerCase() === 'yes'.  ASCII folding
// is exact here: no non-ASCII code point lower-cases to 'y', 'e' or 's'.  Works on any address space.
__device__ __forceinline__ bool is_yes(const uint8_t* s, int n) {
  int b = 0, e = n;
  while (b < e) {
    const int l = js_ws_len_at(s, b, e);
    if (!l) break;
    b += l;
  }
  while (e > b) {
    const int l = js_ws_len_before(s, b, e);
    if (!l) break;
    e -= l;
  }
  return e - b == 3 && ascii_lower(s[b]) == 'y' && ascii_lower(s[b + 1]) == 'e' && ascii_lower(s[b + 2]) == 's';
}

// ---- slow path: a warp per row, lanes stride over the bytes of a cell, global -> global ------------
struct SlowCell {
  const uint8_t* p;  // cell bytes in the heap (for a list: all items, contiguous)
  int n;
  int l0, items;     // list items (items == 1 for strings)
};
__device__ __forceinline__ SlowCell slow_locate(const CellDesc& d, int64_t i) {
  SlowCell c{nullptr, 0, 0, 1};
  if (d.kind == kCellString || d.kind == kCellYesNo) {
    const int b = d.offsets[i];
    c.p = d.data + b;
    c.n = d.offsets[i + 1] - b;
  } else if (d.kind == kCellJoined) {
    c.l0 = d.list_offsets[i];
    c.items = d.list_offsets[i + 1] - c.l0;
    if (c.items > 0) {
      const int b = d.offsets[c.l0];
      c.p = d.data + b;
      c.n = d.offsets[c.l0 + c.items] - b;
    }
  }
  return c;
}

// Row lengths (sm.group[0][r]) and per-row cell masks (bit c: cell c needs escaping, or — for a yes/no cell — is
// `true`; bit 31: Completed).
template <bool kJson>
__device__ __noinline__ void slow_measure(const pie_archive_view& v, const RowTable& tab, const CsvScratch& sc,
                                          CsvSmem& sm, char* s_num, uint32_t* qmask, uint32_t s, int64_t e0, int rows) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int r = wid; r < rows; r += kWorkerWarps) {
    const int64_t e = e0 + r;
    const int64_t show = sc.entry_show[e];
    bool completed = false;
    if (tab.status_col >= 0) {
      const SlowCell st = slow_locate(tab.cell[tab.status_col], e);
      completed = equals_exact(st.p, st.n, "Completed");
    }
    uint32_t len = 0, qm = 0;
    for (int col = 0; col < kCols; ++col) {
      const CellDesc& d = tab.cell[col];
      uint32_t cl = 0;
      if (d.kind == kCellLiteral) {
        cl = d.lit_len;
      } else if (d.kind == kCellYesNo) {
        const SlowCell c = slow_locate(d, e);
        const bool yes = is_yes(c.p, c.n);  // every lane reads the same few bytes
        cl = yes ? d.lit_len : d.lit_no_len;
        if (yes) qm |= 1u << col;
      } else if (d.kind == kCellNumber) {  // delaySec === null || undefined ? '' : delaySec, then String() (:301, :333)
        int nl = 0;
        if (lane == 0 && v.delay_valid[e]) {
          const RyuTables t{d_pow5_inv, d_pow5};
          nl = js_number_to_string(v.delay_sec[e], s_num + r * kMaxNumberChars, t);
        }
        nl = __shfl_sync(0xFFFFFFFFu, nl, 0);
        if (lane == 0) sm.num_len[s][r] = (uint8_t)nl;
        cl = (uint32_t)nl;
      } else if (!(d.blank_if_completed && completed)) {
        const SlowCell c = slow_locate(d, d.per_entry ? e : show);
        uint32_t sp = 0, extra = 0;
        for (int j = lane; j < c.n; j += 32) {
          const uint8_t ch = c.p[j];
          sp |= is_special_byte<kJson>(ch);
          extra += extra_bytes<kJson>(ch);
        }
        cl = (uint32_t)c.n + (c.items > 1 ? (uint32_t)(c.items - 1) : 0u);
        if (__any_sync(0xFFFFFFFFu, sp)) {
          cl += (kJson ? 0u : 2u) + __reduce_add_sync(0xFFFFFFFFu, extra);
          qm |= 1u << col;
        }
      }
      len += cl + 1u;
    }
    if (lane == 0) {
      sm.group[0][r] = len;
      qmask[r] = qm | (completed ? 1u << 31 : 0u);
    }
  }
}

template <bool kJson>
__device__ __forceinline__ void slow_copy(uint8_t* __restrict__ dst, uint32_t& pos, const uint8_t* __restrict__ s, int n,
                                          bool escape, int lane) {
  if (!escape) {
    for (int j = lane; j < n; j += 32) dst[pos + j] = s[j];
    pos += (uint32_t)n;
    return;
  }
  for (int j0 = 0; j0 < n; j0 += 32) {  // a lane's byte lands after the escaped bytes of the lanes before it
    const int j = j0 + lane;
    const uint8_t ch = j < n ? s[j] : 0;
    const uint32_t mine = j < n ? 1u + extra_bytes<kJson>(ch) : 0u;
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= o) incl += t;
    }
    if (j < n) put_escaped<kJson>(dst + pos + incl - mine, ch);
    pos += __shfl_sync(0xFFFFFFFFu, incl, 31);
  }
}

template <bool kJson>
__device__ __noinline__ void slow_write(const RowTable& tab, const CsvScratch& sc, CsvSmem& sm, const char* s_num,
                                        const uint32_t* qmask, uint32_t s, uint32_t par, int64_t e0, int rows,
                                        uint8_t* __restrict__ out) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int r = wid; r < rows; r += kWorkerWarps) {
    const int64_t e = e0 + r;
    const int64_t show = sc.entry_show[e];
    const uint32_t qm = qmask[r];
    const bool completed = (qm >> 31) != 0;
    uint8_t* dst = out + sm.row_start[par][r];
    uint32_t pos = 0;
    for (int col = 0; col < kCols; ++col) {
      const CellDesc& d = tab.cell[col];
      if (d.kind == kCellLiteral || d.kind == kCellYesNo) {
        const bool yes = d.kind == kCellLiteral || ((qm >> col) & 1u);
        const uint32_t off = yes ? d.lit : d.lit_no, n = yes ? d.lit_len : d.lit_no_len;
        for (uint32_t j = lane; j < n; j += 32) dst[pos + j] = tab.literals[off + j];
        pos += n;
      } else if (d.kind == kCellNumber) {
        const int nl = sm.num_len[s][r];
        if (lane < nl) dst[pos + lane] = (uint8_t)s_num[r * kMaxNumberChars + lane];
        pos += (uint32_t)nl;
      } else if (!(d.blank_if_completed && completed)) {
        const SlowCell c = slow_locate(d, d.per_entry ? e : show);
        const bool escape = (qm >> col) & 1u;
        const bool wrap = escape && !kJson;
        if (wrap) {
          if (lane == 0) dst[pos] = '"';
          ++pos;
        }
        if (c.items <= 1) {
          slow_copy<kJson>(dst, pos, c.p, c.n, escape, lane);
        } else {
          for (int it = 0; it < c.items; ++it) {
            const int b = d.offsets[c.l0 + it], n = d.offsets[c.l0 + it + 1] - b;
            slow_copy<kJson>(dst, pos, d.data + b, n, escape, lane);
            if (it + 1 < c.items) {
              if (lane == 0) dst[pos] = '|';
              ++pos;
            }
          }
        }
        if (wrap) {
          if (lane == 0) dst[pos] = '"';
          ++pos;
        }
      }
      if (lane == 0) dst[pos] = d.sep;
      ++pos;
    }
  }
}

// ---- producer: plan and issue the ranges of one tile -------------------------------------------------
struct Range {
  uintptr_t begin, end;  // byte addresses in global memory
};
struct RangePlan {
  uintptr_t lo, hi;  // [lo, hi): whole 16-byte chunks, by TMA
  uint32_t span;     // bytes of the stage the range occupies (a multiple of 16)
  uint32_t bulk;
};
__device__ __forceinline__ RangePlan plan_range(const Range& r) {
  RangePlan p;
  p.lo = r.begin & ~static_cast<uintptr_t>(15);  // not before the allocation: an address that is not 16-byte
                                                 // aligned cannot be the first byte of one
  p.hi = r.end & ~static_cast<uintptr_t>(15);    // full chunks only: nothing is read past the range's last word
  const bool any = r.end > r.begin;
  p.bulk = (any && p.hi > p.lo) ? (uint32_t)(p.hi - p.lo) : 0u;
  p.span = any ? (uint32_t)(((r.end + 15) & ~static_cast<uintptr_t>(15)) - p.lo) : 0u;
  return p;
}
__device__ __forceinline__ uint32_t warp_exclusive_scan(uint32_t x, uint32_t& total, int lane) {
  uint32_t incl = x;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= o) incl += t;
  }
  total = __shfl_sync(0xFFFFFFFFu, incl, 31);
  return incl - x;
}
// copies the range into stage + region; returns nothing: the staged address of r.begin is region + (begin - lo)
__device__ __forceinline__ void issue_range(const Range& r, const RangePlan& p, uint8_t* stage, uint32_t region, uint32_t bar) {
  if (!p.span) return;
  if (p.bulk) tma_bulk_g2s(smem_u32(stage + region), reinterpret_cast<const void*>(p.lo), p.bulk, bar);
  // The range's last, partial chunk (or a range inside one chunk): the <= 4 aligned words that hold
  // bytes of it — never a word past the one that holds the range's last byte.
  const uintptr_t t0 = p.bulk ? p.hi : (r.begin & ~static_cast<uintptr_t>(3));
  uint32_t x[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) x[k] = (t0 + 4 * k < r.end) ? __ldg(reinterpret_cast<const uint32_t*>(t0) + k) : 0u;
  uint32_t* dst = reinterpret_cast<uint32_t*>(stage + region + (uint32_t)(t0 - p.lo));
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (t0 + 4 * k < r.end) dst[k] = x[k];
}

__device__ __forceinline__ void produce_tile(const pie_archive_view& v, const RowTable& tab, const CsvScratch& sc,
                                             StageInfo& info, uint8_t* stage, uint32_t bar, uint32_t empty_bar,
                                             uint32_t empty_parity, int64_t tile, int tile_rows, int force_slow, int lane) {
  const int64_t e0 = tile * tile_rows;
  const int rows = (v.n_entries - e0 < tile_rows) ? (int)(v.n_entries - e0) : tile_rows;
  const int32_t s0 = sc.entry_show[e0], s1 = sc.entry_show[e0 + rows - 1];
  Range rb{0, 0}, ro{0, 0}, ri{0, 0};  // heap bytes, offsets slice, item offsets slice
  uint32_t b0 = 0;
  int32_t l0 = 0;
  const bool has_bytes = lane < kCols && (tab.cell[lane < kCols ? lane : 0].kind == kCellString ||
                                         tab.cell[lane < kCols ? lane : 0].kind == kCellJoined ||
                                         tab.cell[lane < kCols ? lane : 0].kind == kCellYesNo);
  if (has_bytes) {
    const CellDesc& d = tab.cell[lane];
    const int64_t i0 = d.per_entry ? e0 : (int64_t)s0, i1 = d.per_entry ? e0 + rows : (int64_t)s1 + 1;
    const int32_t* oarr = (d.kind == kCellJoined) ? d.list_offsets : d.offsets;
    ro = Range{reinterpret_cast<uintptr_t>(oarr + i0), reinterpret_cast<uintptr_t>(oarr + i1 + 1)};
    int32_t f0 = oarr[i0], f1 = oarr[i1];
    if (d.kind == kCellJoined) {
      l0 = f0;
      ri = Range{reinterpret_cast<uintptr_t>(d.offsets + f0), reinterpret_cast<uintptr_t>(d.offsets + f1 + 1)};
      f0 = d.offsets[f0];
      f1 = d.offsets[f1];
    }
    b0 = (uint32_t)f0;
    rb = Range{reinterpret_cast<uintptr_t>(d.data) + (uint32_t)f0, reinterpret_cast<uintptr_t>(d.data) + (uint32_t)f1};
  } else if (lane == kCols) {
    ro = Range{reinterpret_cast<uintptr_t>(sc.entry_show + e0), reinterpret_cast<uintptr_t>(sc.entry_show + e0 + rows)};
  } else if (lane == kCols + 1 && tab.has_number) {
    ro = Range{reinterpret_cast<uintptr_t>(v.delay_sec + e0), reinterpret_cast<uintptr_t>(v.delay_sec + e0 + rows)};
  } else if (lane == kCols + 2 && tab.has_number) {
    ro = Range{reinterpret_cast<uintptr_t>(v.delay_valid + e0), reinterpret_cast<uintptr_t>(v.delay_valid + e0 + rows)};
  }
  const RangePlan pb = plan_range(rb), po = plan_range(ro), pi = plan_range(ri);
  uint32_t tb, to, ti;
  // the first kLiteralBytes of a stage hold the row format's literal text (copied once, at kernel start)
  const uint32_t region_b = (uint32_t)kLiteralBytes + warp_exclusive_scan(pb.span, tb, lane);
  const uint32_t region_o = (uint32_t)kLiteralBytes + tb + warp_exclusive_scan(po.span, to, lane);
  const uint32_t region_i = (uint32_t)kLiteralBytes + tb + to + warp_exclusive_scan(pi.span, ti, lane);
  const uint32_t used = (uint32_t)kLiteralBytes + tb + to + ti;
  uint32_t bulk_total = pb.bulk + po.bulk + pi.bulk;
#pragma unroll
  for (int o = 16; o; o >>= 1) bulk_total += __shfl_xor_sync(0xFFFFFFFFu, bulk_total, o);
  const bool fits = used <= (uint32_t)kStageBytes && (s1 - s0) < kMaxTileShows && !force_slow;

  // Everything above — the range ends, the only dependent global loads of the kernel — was computed while the workers
  // were still busy with what the stage held; only now does the producer need the stage itself.
  mbar_wait_relaxed(empty_bar, empty_parity);
  if (lane < kCols) {
    info.delta[lane] = region_b + (uint32_t)(rb.begin - pb.lo) - b0;
    info.off_base[lane] = region_o + (uint32_t)(ro.begin - po.lo);
    info.item_base[lane] = region_i + (uint32_t)(ri.begin - pi.lo);
    info.item_first[lane] = l0;
  } else if (lane == kCols) {
    info.show_idx_base = region_o + (uint32_t)(ro.begin - po.lo);
  } else if (lane == kCols + 1) {
    info.delay_base = region_o + (uint32_t)(ro.begin - po.lo);
  } else if (lane == kCols + 2) {
    info.valid_base = region_o + (uint32_t)(ro.begin - po.lo);
  }
  if (lane == 0) {
    info.tile = tile;
    info.rows = rows;
    info.show0 = s0;
    info.n_tile_shows = (uint32_t)(s1 - s0 + 1);
    info.slow = fits ? 0u : 1u;
    info.bump0 = (used + 3u) & ~3u;
    info.heap_end = (uint32_t)kLiteralBytes + tb;
    if (fits) mbar_expect_tx(bar, bulk_total);
  }
  __syncwarp();
  if (fits) {
    issue_range(rb, pb, stage, region_b, bar);
    issue_range(ro, po, stage, region_o, bar);
    issue_range(ri, pi, stage, region_i, bar);
  }
  __syncwarp();  // every lane's info fields and word copies are ordered before lane 0's release-arrive
  if (lane == 0) mbar_arrive(bar);
}

// ---- look-back warp ------------------------------------------------------------------------------------
__device__ __forceinline__ void look_back(const CsvScratch& sc, CsvSmem& sm, uint32_t par, int64_t n_tiles,
                                          int64_t n_entries, int64_t* __restrict__ row_offsets, unsigned long long bias,
                                          unsigned long long* __restrict__ total_out, int lane) {
  const int64_t tile = sm.cur_tile[par];
  const uint32_t tile_total = sm.tile_total[par];
  volatile unsigned long long* state = sc.tile_state;
  unsigned long long exclusive = 0;
  int64_t idx = tile - 1;
  while (idx >= 0) {
    const int64_t j = idx - lane;
    unsigned long long st;
    unsigned ns = 32;
    Watchdog dog;
    for (;;) {
      st = 2ull << kStatusShift;  // before tile 0: an empty prefix
      if (j >= 0) st = state[j];
      if (!__any_sync(0xFFFFFFFFu, (st >> kStatusShift) == 0)) break;
      __nanosleep(ns);  // the tiles we wait for are still measuring: do not hammer L2 / the issue slots
      if (ns < 512) ns <<= 1;
      dog.polls += 0x3Fu;
      dog.poll();
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
    // (status, value) travel in one 64-bit word and nothing else is read through it: no fence needed
    if (tile > 0) state[tile] = kPrefix | (exclusive + tile_total);
    sm.base[par] = exclusive;
    if (tile == n_tiles - 1) {
      *total_out = exclusive + tile_total;
      row_offsets[n_entries] = (int64_t)(bias + exclusive + tile_total);
    }
  }
}

// Developer instrumentation (-DPIE_CSV_PROFILE): cycles worker thread 0 spends between phase boundaries,
// summed over tiles into 64-bit counters from byte 256 of the scratch control block, [group][phase].  Compiled out by default.
#ifdef PIE_CSV_PROFILE
#define PIE_PHASE(i)                                                                                   \
  do {                                                                                                 \
    if (r == 0) {                                                                                      \
      const long long now_ = clock64();                                                                \
      atomicAdd(reinterpret_cast<unsigned long long*>(sc.tile_counter + 64) + g * 16 + (i),            \
                (unsigned long long)(now_ - t_phase));                                                 \
      t_phase = now_;                                                                                  \
    }                                                                                                  \
  } while (0)
#else
#define PIE_PHASE(i) do { } while (0)
#endif

// ---- the kernel ---------------------------------------------------------------------------------
// kJson: escape mode of the value cells (the row format itself is the run-time RowTable)
template <bool kJson>
__global__ void __launch_bounds__(kCtaThreads, PIE_CSV_MIN_BLOCKS)
    export_rows_kernel(pie_archive_view v, const __grid_constant__ RowTable tab, CsvScratch sc,
                       int64_t* __restrict__ row_offsets, uint8_t* __restrict__ out_data, uint64_t capacity,
                       unsigned long long bias, unsigned long long* __restrict__ total_out, int force_slow) {
  extern __shared__ __align__(128) uint8_t s_dyn[];
  uint8_t* s_out = s_dyn;  // kOutBytes + 32
  CsvSmem& sm = *reinterpret_cast<CsvSmem*>(s_dyn + kSmemOffState);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int tile_rows = (int)*sc.tile_rows;  // rows per tile of this launch: kRows unless the rows are wide
  const int64_t n_tiles = (v.n_entries + tile_rows - 1) / tile_rows;

  if (tid == 0) {
    mbar_init(smem_u32(&sm.full[0]), 1);
    mbar_init(smem_u32(&sm.full[1]), 1);
    mbar_init(smem_u32(&sm.empty[0]), kWorkerWarps + kNumberWarps);
    mbar_init(smem_u32(&sm.empty[1]), kWorkerWarps + kNumberWarps);
    mbar_init(smem_u32(&sm.nums[0]), kNumberWarps);
    mbar_init(smem_u32(&sm.nums[1]), kNumberWarps);
    mbar_fence_init();
    sm.done = 0;
  }
  // the row format's literal text goes to the head of both stages; the cells that are always the same
  // literal are entered in the cell table once
  for (int i = tid; i < 2 * kLiteralBytes; i += kCtaThreads)
    s_dyn[kSmemOffStage + (i / kLiteralBytes) * kStageStride + (i % kLiteralBytes)] = tab.literals[i % kLiteralBytes];
  for (int i = tid; i < kRows * kCols; i += kCtaThreads) {
    const CellDesc& d = tab.cell[i % kCols];
    if (d.kind == kCellLiteral) sm.cell[(i / kCols) * kCellStride + (i % kCols)] = pack_cell(d.lit, d.lit_len);
  }
  __syncthreads();

  // ================= producer warp =================
  if (wid == kProducerWarp) {
    for (uint32_t it = 0;; ++it) {
      const uint32_t s = it & 1u, ph = (it >> 1) & 1u;
      long long tile = 0;
      if (lane == 0) tile = (long long)atomicAdd(sc.tile_counter, 1u);  // tiles start in id order: look-back cannot deadlock
      tile = __shfl_sync(0xFFFFFFFFu, tile, 0);
      if (tile >= n_tiles) {
        mbar_wait_relaxed(smem_u32(&sm.empty[s]), ph ^ 1u);  // the workers are done with what the stage held
        if (lane == 0) {
          sm.info[s].tile = -1;
          mbar_arrive(smem_u32(&sm.full[s]));
        }
        return;
      }
      produce_tile(v, tab, sc, sm.info[s], s_dyn + kSmemOffStage + s * kStageStride, smem_u32(&sm.full[s]),
                   smem_u32(&sm.empty[s]), ph ^ 1u, tile, tile_rows, force_slow, lane);
    }
  }

  // ================= look-back warp =================
  if (wid == kLookbackWarp) {
    for (uint32_t it = 0;; ++it) {
      bar_sync_workers_and_lookback<kBarTotalReady>();
      if (sm.done) return;
      const uint32_t par = it & 1u;
      look_back(sc, sm, par, n_tiles, v.n_entries, row_offsets, bias, total_out, lane);
      bar_arrive_workers_and_lookback<kBarBaseReady>();
    }
  }

  // ================= number warps =================
  // delaySec === null || undefined ? '' : delaySec, then String() (:301, :333).  Number::toString is a
  // long chain of dependent 64-bit operations for the values that need the general algorithm, and a warp
  // takes as long as its slowest lane: done here, a tile ahead of the workers, it is off their path.
  if (wid >= kNumberWarp0) {
    const int row = tid - kNumberWarp0 * 32;
    for (uint32_t it = 0;; ++it) {
      const uint32_t s = it & 1u, ph = (it >> 1) & 1u;
      uint8_t* stage = s_dyn + kSmemOffStage + s * kStageStride;
      const StageInfo& info = sm.info[s];
      if (it > 0) {
        // The cells the workers queued for the PREVIOUS tile (csvEscape / join / JSON escapes) get their
        // bytes here, one item per thread, the two kinds from opposite ends — while the workers add up
        // lengths, publish and flush; they need the bytes only when they write.
        bar_sync_n<kBarItemsReady, kWorkers + 32 * kNumberWarps>();
        if (!sm.fill_skip && !sm.overflow) {
          uint8_t* prev = s_dyn + kSmemOffStage + (s ^ 1u) * kStageStride;
          const uint32_t n_word = sm.n_word_items, n_quote = sm.n_quote_items;
          for (uint32_t i = (uint32_t)row; i < n_word; i += 32 * kNumberWarps) fill_item<kJson>(prev, sm.word_items[i], false);
          for (uint32_t i = 32 * kNumberWarps - 1 - (uint32_t)row; i < n_quote; i += 32 * kNumberWarps)
            fill_item<kJson>(prev, sm.quote_items[i], true);
        }
        bar_arrive_n<kBarFillsDone, kWorkers + 32 * kNumberWarps>();
      }
      mbar_wait_relaxed(smem_u32(&sm.full[s]), ph);
      if (info.tile < 0) return;
      double value = 0.0;
      bool valid = false;
      const bool staged = info.slow == 0;  // read before the stage is released: info is rewritten two tiles on
      if (staged && tab.has_number && row < info.rows) {
        valid = stage[info.valid_base + row] != 0;
        value = *reinterpret_cast<const double*>(stage + info.delay_base + 8 * row);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&sm.empty[s]));  // the staged ranges are not read again
      int nl = 0;
      if (valid) {
        const RyuTables t{d_pow5_inv, d_pow5};
        nl = js_number_to_string(value, reinterpret_cast<char*>(stage + kStageBytes) + row * kMaxNumberChars, t);
      }
      if (staged) sm.num_len[s][row] = (uint8_t)nl;  // a slow tile formats its numbers itself
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&sm.nums[s]));
    }
  }

  // ================= workers =================
  const bool write = out_data != nullptr;
  const int g = tid / kRows, r = tid % kRows;
  uint32_t* qmask = sm.qmask;

  // Row threads: worker threads kRows .. 2*kRows-1 (warps that hold no word-stream fill items) own a row each
  // for the length bookkeeping.
  const int rt = tid - kRows;
  const bool row_thread = rt >= 0 && rt < kRows;
  // row_len of every row thread -> sm.row_start[par], sm.tile_total[par]; the aggregate is published
  auto scan_rows_and_publish = [&](int64_t tile, uint32_t row_len, uint32_t par) {
    if (row_thread) {
      uint32_t incl = row_len;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += t;
      }
      if (lane == 31) sm.warp_sum[rt >> 5] = incl;
      sm.row_start[par][rt] = incl - row_len;  // completed below with the preceding warps' sums
    }
    workers_sync();
    if (row_thread) {
      uint32_t before = 0;
      for (int w = 0; w < (rt >> 5); ++w) before += sm.warp_sum[w];
      sm.row_start[par][rt] += before;
      if (rt == kRows - 1) {
        const uint32_t total = sm.row_start[par][rt] + row_len;
        sm.tile_total[par] = total;
        sm.cur_tile[par] = tile;
        // (status, value) travel in one 64-bit word and nothing else is read through it: no fence needed
        reinterpret_cast<volatile unsigned long long*>(sc.tile_state)[tile] =
            (tile == 0 ? kPrefix : kAggregate) | (unsigned long long)total;
      }
    }
    workers_sync();
  };

  // The flush of a tile is deferred until the next tile has been measured: its look-back (a walk over
  // the totals of the ~150 tiles in flight on the other SMs) then has a whole tile of slack.
  bool pending = false;
  int64_t p_e0 = 0;
  int p_rows = 0;
  uint32_t p_total = 0, p_par = 0;
  auto finish_pending = [&]() {
    if (!pending) return;
    pending = false;
    bar_sync_workers_and_lookback<kBarBaseReady>();  // every chunk is written and the tile's offset is known
    const unsigned long long base = sm.base[p_par];
    if (tid < p_rows) row_offsets[p_e0 + tid] = (int64_t)(bias + base + sm.row_start[p_par][tid]);
    if (!write || base + p_total > capacity) return;
    // s_out[0 .. p_total) -> out_data[base ..) with 16-byte stores.  Global chunk k starts at the first
    // 16-byte boundary >= out_data+base, i.e. at tile offset head + 16k, which has an arbitrary phase in
    // shared memory: read 5 aligned words and funnel-shift.
    uint8_t* __restrict__ dst = out_data + base;
    const uint32_t head_raw = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15)) & 15u;
    const uint32_t head = head_raw < p_total ? head_raw : p_total;
    for (uint32_t k = tid; k < head; k += kWorkers) dst[k] = s_out[k];
    const uint32_t n_chunks = (p_total - head) >> 4;
    const uint32_t sh = (head & 3u) * 8u;
    // The 5 words a chunk needs start at word (head >> 2) + 4k: two aligned 128-bit loads (conflict-free:
    // consecutive lanes read consecutive 16-byte units) cover them; which 5 of the 8 words is uniform.
    const uint32_t m = (head >> 2) & 3u;
    const uint4* __restrict__ sq = reinterpret_cast<const uint4*>(s_out) + (head >> 4);
    for (uint32_t k = tid; k < n_chunks; k += kWorkers) {
      const uint4 a = sq[k], b = sq[k + 1];  // b of the last chunk stays inside the +32 slack
      uint32_t w0, w1, w2, w3, w4;
      if (m == 0) { w0 = a.x; w1 = a.y; w2 = a.z; w3 = a.w; w4 = b.x; }
      else if (m == 1) { w0 = a.y; w1 = a.z; w2 = a.w; w3 = b.x; w4 = b.y; }
      else if (m == 2) { w0 = a.z; w1 = a.w; w2 = b.x; w3 = b.y; w4 = b.z; }
      else { w0 = a.w; w1 = b.x; w2 = b.y; w3 = b.z; w4 = b.w; }
      uint4 o;
      o.x = __funnelshift_r(w0, w1, sh);
      o.y = __funnelshift_r(w1, w2, sh);
      o.z = __funnelshift_r(w2, w3, sh);
      o.w = __funnelshift_r(w3, w4, sh);
      *reinterpret_cast<uint4*>(dst + head + 16u * k) = o;
    }
    for (uint32_t k = head + 16u * n_chunks + tid; k < p_total; k += kWorkers) dst[k] = s_out[k];
  };

#ifdef PIE_CSV_PROFILE
  long long t_phase = clock64();
#endif
  for (uint32_t it = 0;; ++it) {
    const uint32_t s = it & 1u, ph = (it >> 1) & 1u, par = it & 1u;
    uint8_t* stage = s_dyn + kSmemOffStage + s * kStageStride;
    char* s_num = reinterpret_cast<char*>(stage + kStageBytes);
    const StageInfo& info = sm.info[s];
    mbar_wait(smem_u32(&sm.full[s]), ph);  // the staged ranges have landed, info is visible
    PIE_PHASE(0);  // waiting for the producer
    const int64_t tile = info.tile;
    if (tile < 0) {
      finish_pending();
      if (tid == 0) sm.done = 1;
      bar_arrive_workers_and_lookback<kBarTotalReady>();
      return;
    }
    const int64_t e0 = tile * tile_rows;
    const int rows = info.rows;
    const bool have = r < rows;
    bool slow = info.slow != 0;  // uniform
    bool published = false;
    bool items_signalled = false;

    uint32_t cells[kGroupCols];  // the six cells this thread writes, (src:16 | len:16 << 16)
#pragma unroll
    for (int k = 0; k < kGroupCols; ++k) cells[k] = 0;
    uint32_t my_start = 0;   // where this thread's group starts in the tile image
    uint32_t tile_total = 0;
    if (!slow) {
      if (tid == 0) {
        sm.bump = info.bump0;
        sm.overflow = 0;
        sm.n_word_items = 0;
        sm.n_quote_items = 0;
        sm.fill_skip = 0;
      }
      // which staged bytes need escaping: one bit each, for every column (see build_special_mask)
      build_special_mask<kJson>(sm.special, stage, info.heap_end, tid);
      workers_sync();  // the mask is complete; also: every worker has left the previous tile's write phase (cell table)
      PIE_PHASE(1);
      // ---- cells 1. show-level cells once per show of the tile; entry-level cells: thread (r, g) takes
      // the value cells tab.owned[g] (the expensive ones — Array.join, free text — on different groups)
      {
        const uint32_t ns = info.n_tile_shows;
        for (uint32_t idx = tid; idx < ns * tab.n_show_slots; idx += kWorkers) {
          const uint32_t slot = idx / ns, i = idx - slot * ns;
          const uint32_t col = tab.show_col[slot];
          const int32_t* o = stage_i32(stage, info.off_base[col]);
          const int32_t f0 = o[i], f1 = o[i + 1];
          int items = 1;
          uint32_t b = (uint32_t)f0, n = (uint32_t)(f1 - f0);
          const int32_t* io = nullptr;
          if (tab.cell[col].kind == kCellJoined) {  // crew
            items = f1 - f0;
            b = 0;
            n = 0;
            io = stage_i32(stage, info.item_base[col]) + (f0 - info.item_first[col]);
            if (items > 0) {
              b = (uint32_t)io[0];
              n = (uint32_t)io[items] - b;
            }
          }
          const uint32_t src = (info.delta[col] + b) & 0xFFFFu;
          sm.shcell[slot][i] = plan_cell<kJson>(sm, stage, io, info.delta[col], src, n, items);
        }
      }
      PIE_PHASE(2);  // show-level cells
      if (have) {
        // entry.status === 'Completed' (:293-297) blanks the five issue cells, which sit on several groups:
        // every thread reads the row's status itself (three shared-memory words)
        bool completed = false;
        if (tab.status_col >= 0) {
          const int32_t* o = stage_i32(stage, info.off_base[tab.status_col]);
          const int32_t f0 = o[r];
          if (o[r + 1] - f0 == 9) {
            const uint32_t src = (info.delta[tab.status_col] + (uint32_t)f0) & 0xFFFFu;
            const uint32_t* w = reinterpret_cast<const uint32_t*>(stage + (src & ~3u));
            const uint32_t sh = (src & 3u) * 8u;
            const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];  // the 9 bytes lie inside these three words
            completed = __funnelshift_r(w0, w1, sh) == lit_word("Completed", 0) &&
                        __funnelshift_r(w1, w2, sh) == lit_word("Completed", 1) &&
                        (__funnelshift_r(w2, 0u, sh) & 0xFFu) == lit_word("Completed", 2);
          }
        }
        uint32_t* row_cells = sm.cell + r * kCellStride;
#pragma unroll 1
        for (int k = 0; k < kGroupCols; ++k) {
          const int col = tab.owned[g][k];
          if (col < 0) break;
          const CellDesc& d = tab.cell[col];
          uint32_t c = 0;
          if (d.kind == kCellNumber) {  // formatted by the number warps
            mbar_wait(smem_u32(&sm.nums[s]), ph);
            c = pack_cell((uint32_t)kStageBytes + (uint32_t)(r * kMaxNumberChars), (uint32_t)sm.num_len[s][r]);
          } else {
            const int32_t* o = stage_i32(stage, info.off_base[col]);
            const int32_t f0 = o[r], f1 = o[r + 1];
            int items = 1;
            uint32_t b = (uint32_t)f0, n = (uint32_t)(f1 - f0);
            const int32_t* io = nullptr;
            if (d.kind == kCellJoined) {  // actions
              items = f1 - f0;
              b = 0;
              n = 0;
              io = stage_i32(stage, info.item_base[col]) + (f0 - info.item_first[col]);
              if (items > 0) {
                b = (uint32_t)io[0];
                n = (uint32_t)io[items] - b;
              }
            }
            const uint32_t src = (info.delta[col] + b) & 0xFFFFu;
            if (d.kind == kCellYesNo) {  // toYesNoBoolean picks one of two literals
              c = is_yes(stage + src, (int)n) ? pack_cell(d.lit, d.lit_len) : pack_cell(d.lit_no, d.lit_no_len);
            } else if (!(d.blank_if_completed && completed)) {
              c = plan_cell<kJson>(sm, stage, io, info.delta[col], src, n, items);
            }
          }
          row_cells[col] = c;
        }
      }
      PIE_PHASE(3);  // this thread's entry-level cells
      workers_sync();
      PIE_PHASE(4);  // waiting for the other warps' cells
      // ---- lengths.  (The queued cells get their bytes from the number warps meanwhile.)
      bar_arrive_n<kBarItemsReady, kWorkers + 32 * kNumberWarps>();  // the number warps fill the queued cells
      items_signalled = true;
      PIE_PHASE(9);
      slow = sm.overflow != 0;  // the bump area or a fill queue ran out (uniform: written before the barrier)
      if (!slow) {
        // ---- lengths.  Thread (r, g) collects the six cells it is going to write — its show's cells from the
        // per-tile table, the entry's from the cell table — in registers, and publishes the bytes of its group; then
        // every thread adds up its row, and every warp (32 consecutive rows of one group) scans its rows itself.
        uint32_t glen = 0;
        if (have) {
          const int show_i = stage_i32(stage, info.show_idx_base)[r] - info.show0;
          const uint32_t* row_cells = sm.cell + r * kCellStride + g * kGroupCols;
#pragma unroll
          for (int k = 0; k < kGroupCols; ++k) {
            const int slot = tab.cell[g * kGroupCols + k].show_slot;
            cells[k] = slot >= 0 ? sm.shcell[slot][show_i] : row_cells[k];
            glen += (cells[k] >> 16) + 1u;  // + the cell's separator byte
          }
        }
        sm.group[g][r] = glen;
        workers_sync();
        uint32_t row_len = 0;
#pragma unroll
        for (int gg = 0; gg < kGroups; ++gg) {
          const uint32_t x = sm.group[gg][r];
          row_len += x;
          if (gg < g) my_start += x;
        }
        uint32_t incl = row_len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
          if (lane >= o) incl += t;
        }
        const int blk = r >> 5;  // a warp = 32 consecutive rows of one group (kRows is a multiple of 32)
        if (g == 0 && lane == 31) sm.warp_sum[blk] = incl;
        workers_sync();
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kRows / 32; ++w) {
          const uint32_t x = sm.warp_sum[w];
          total += x;
          if (w < blk) before += x;
        }
        const uint32_t row_start = before + incl - row_len;
        my_start += row_start;
        if (g == 0) sm.row_start[par][r] = row_start;  // for the row offsets (flush) — and the slow path
        if (tid == 0) {
          sm.tile_total[par] = total;
          sm.cur_tile[par] = tile;
          // (status, value) travel in one 64-bit word and nothing else is read through it: no fence needed
          reinterpret_cast<volatile unsigned long long*>(sc.tile_state)[tile] =
              (tile == 0 ? kPrefix : kAggregate) | (unsigned long long)total;
        }
        tile_total = total;
        PIE_PHASE(5);  // cells 2 + scans
        published = true;
        if (write && total > (uint32_t)kOutBytes) slow = true;  // uniform
      }
    }

    if (slow) {
      // ---- slow path (uniform for the CTA).  If the fast path already published this tile's total, the
      // measure below recomputes the same row lengths; only the quote masks are new.
      finish_pending();
      if (tid == 0) {
        atomicAdd(sc.slow_tiles, 1u);
        sm.fill_skip = 1;
      }
      // the number warps go through "items ready / fills done" once per tile whatever path it takes
      if (!items_signalled) bar_arrive_n<kBarItemsReady, kWorkers + 32 * kNumberWarps>();
      bar_sync_n<kBarFillsDone, kWorkers + 32 * kNumberWarps>();
      workers_sync();
      slow_measure<kJson>(v, tab, sc, sm, s_num, qmask, s, e0, rows);
      workers_sync();
      if (!published) scan_rows_and_publish(tile, (row_thread && rt < rows) ? sm.group[0][rt] : 0u, par);
      tile_total = sm.tile_total[par];
      bar_arrive_workers_and_lookback<kBarTotalReady>();
      if (lane == 0) mbar_arrive(smem_u32(&sm.empty[s]));  // nothing of the stage's staged ranges is read any more
      bar_sync_workers_and_lookback<kBarBaseReady>();
      const unsigned long long base = sm.base[par];
      if (tid < rows) row_offsets[e0 + tid] = (int64_t)(bias + base + sm.row_start[par][tid]);
      if (write && base + tile_total <= capacity)
        slow_write<kJson>(tab, sc, sm, s_num, qmask, s, par, e0, rows, out_data + base);
      workers_sync();  // qmask / row lengths are free again
      continue;
    }

    bar_arrive_workers_and_lookback<kBarTotalReady>();  // this tile's look-back starts now ...
    finish_pending();                                   // ... while the previous tile leaves s_out
    PIE_PHASE(6);  // wait for the previous tile's offset + its flush
    bar_sync_n<kBarFillsDone, kWorkers + 32 * kNumberWarps>();  // the queued cells have their bytes
    PIE_PHASE(10);
    if (write) {
      // ---- write.  Thread (r, g) streams its 6 consecutive cells into the shared output tile.  Lanes of
      // a warp = 32 consecutive rows on the SAME column at every step, so cell lengths (and with them the
      // trip counts of the word loop) are alike.  Bytes collect in a (lo, hi) accumulator and leave as
      // aligned 32-bit stores; only the group's first and last word, which it shares with its neighbours,
      // are stored byte by byte.
      workers_sync();  // the previous tile has left s_out
      PIE_PHASE(8);    // ... waiting for that
      if (have) {
        ByteStream<true> out;
        out.init(s_out + my_start);
#pragma unroll
        for (int k = 0; k < kGroupCols; ++k) {
          out.append(stage, cells[k] & 0xFFFFu, cells[k] >> 16, tab.cell[g * kGroupCols + k].sep, 1u);
        }
        out.finish();
      }
    }
    PIE_PHASE(7);  // write
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&sm.empty[s]));  // the stage may be refilled (two tiles ahead)
    pending = true;
    p_e0 = e0;
    p_rows = rows;
    p_total = tile_total;
    p_par = par;
  }
}

template <bool kJson>
static cudaError_t launch_rows(const pie_archive_view& v, const RowTable& tab, int64_t* row_offsets, uint8_t* out_data,
                               uint64_t capacity, unsigned long long bias, unsigned long long* total_out, void* scratch,
                               cudaStream_t stream) {
  CsvScratch sc = carve_csv(scratch, v.n_entries);
  cudaError_t err = cudaMemsetAsync(scratch, 0, csv_scratch_zero_bytes(v.n_entries), stream);
  if (err != cudaSuccess) return err;
  if (v.n_entries == 0) {
    err = cudaMemsetAsync(total_out, 0, 8, stream);
    if (err != cudaSuccess) return err;
    return cudaMemcpyAsync(row_offsets, total_out, 8, cudaMemcpyDefault, stream);  // 0; the caller adds its bias
  }
  static int configured_device = -1, resident_ctas = 0;  // per instantiation
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_device != dev) {
    err = cudaFuncSetAttribute(export_rows_kernel<kJson>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (err != cudaSuccess) return err;
    err = cudaFuncSetAttribute(export_rows_kernel<kJson>, cudaFuncAttributePreferredSharedMemoryCarveout,
                               cudaSharedmemCarveoutMaxShared);
    if (err != cudaSuccess) return err;
    int per_sm = 0, sms = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, export_rows_kernel<kJson>, kCtaThreads, kSmemBytes);
    if (err != cudaSuccess) return err;
    err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (err != cudaSuccess) return err;
    resident_ctas = (per_sm > 0 ? per_sm : 1) * sms;  // persistent: one wave of CTAs loops over the tiles
    configured_device = dev;
  }
  expand_entry_show_kernel<<<(unsigned)((v.n_shows + 255) / 256), 256, 0, stream>>>(v, sc.entry_show);
  // the tile plan wants a rough idea of how many cells will be escaped: a sample of every column, not a sweep
  column_sample_kernel<kJson><<<dim3(kSampleChunks / 256, kCols), 256, 0, stream>>>(tab, v.n_shows, v.n_entries, sc.col_dirty,
                                                                                 sc.col_dirty_chunks, sc.col_sampled_chunks);
  plan_tile_rows_kernel<<<1, 32, 0, stream>>>(tab, v.n_shows, v.n_entries, kStageBytes, kOutBytes, sc.col_dirty,
                                              sc.col_dirty_chunks, sc.col_sampled_chunks, sc.tile_rows);
  const int64_t tiles = csv_tiles(v.n_entries);
  const unsigned grid = (unsigned)(tiles < resident_ctas ? tiles : resident_ctas);
  export_rows_kernel<kJson><<<grid, kCtaThreads, kSmemBytes, stream>>>(v, tab, sc, row_offsets, out_data, capacity, bias,
                                                                      total_out, g_force_slow);
  g_launches += 4;
  return cudaGetLastError();
}

cudaError_t launch_csv_rows(const pie_archive_view& v, int64_t* row_offsets, uint8_t* out_data, uint64_t capacity,
                            unsigned long long bias, unsigned long long* total_out, void* scratch,
                            cudaStream_t stream) {
  return launch_rows<false>(v, make_csv_table(v), row_offsets, out_data, capacity, bias, total_out, scratch, stream);
}

cudaError_t launch_payload_rows(const pie_archive_view& v, int64_t* row_offsets, uint8_t* out_data, uint64_t capacity,
                                unsigned long long bias, unsigned long long* total_out, void* scratch,
                                cudaStream_t stream) {
  return launch_rows<true>(v, make_payload_table(v), row_offsets, out_data, capacity, bias, total_out, scratch, stream);
}

}  // namespace pie
