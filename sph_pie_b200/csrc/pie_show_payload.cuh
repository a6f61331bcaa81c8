// Device code of the schemaVersion 2 show payload (show_payload.cu holds the kernels and the launcher; the same code
// runs on the CPU in tests/native/payload_host.cpp, its 32 lanes as fibers): JSON.stringify of the object
// dispatchShowEvent builds for every event but 'show.archived' (reference server/webhookDispatcher.js:545-584) —
// buildShowSummary (:472-488) twice, the table / csv / message views of every entry (buildTableRow :276-305,
// buildCsvRow :340-342) and the entries as the provider stores them (sqlProvider.js:384-409) — one document per show:
//
//   <head>"table":{"columns":[..24 names..],"rows":[[24 values],...]},"csv":{"header":[..],"rows":["<csv row>",...]},
//   "message":{"show":<summary>,"entries":[{24 members},...]},"show":<summary>,"entries":[{17 members},...]<tail>
//
// A WARP PER SHOW, two passes over the same emitter (Emit<kWrite>): the first only adds up the document's length, the
// second writes.  Every piece is produced by the whole warp, a byte per lane.
//
// What it does about the 2.2 s per 2^20 shows of its first version (one dependent global load after the other, per cell;
// r2, last session — parity green on B200 and on the CPU, the time is in the driver's round-end bench line):
//  * stage_show: everything a show's document reads — the show's slice of each of the 23 string heaps and of their
//    offsets arrays, the two list-offset slices — is ~50 contiguous ranges of ~2-5 KB in all.  The warp resolves the
//    range ends lane-parallel (one round of dependent global loads instead of one per cell), copies the ranges into
//    its shared-memory stage as aligned 32-bit words and builds a REBASED pie_archive_view whose pointers lead into the
//    stage, so that the emitter below — unchanged, it only sees a view — finds every cell in shared memory (each cell is
//    visited four times: table row, csv row, message entry, stored entry).  A show that does not fit the stage is
//    emitted from global memory through the caller's view.
//  * Number::toString of every entry's delaySec and ts is computed ONCE, a lane per entry, and kept in the stage (the
//    document spells delaySec four times).
//  * the common cell — one item, no byte that needs an escape, short enough for one round — is one ballot and one
//    store for the quotes, the bytes and the separator together; member names are emitted with their punctuation.
#pragma once
#include <stdint.h>

#include "../../include/sph_pie_b200.h"
#include "pie_device.cuh"
#include "pie_numfmt.cuh"

namespace pie {
namespace sp {

static __device__ const uint64_t d_pow5_inv[PIE_RYU_POW5_INV_SPLIT_N][2] = PIE_RYU_POW5_INV_SPLIT_INIT;
static __device__ const uint64_t d_pow5[PIE_RYU_POW5_SPLIT_N][2] = PIE_RYU_POW5_SPLIT_INIT;

constexpr uint32_t kFullMask = 0xFFFFFFFFu;

// EXPORT_COLUMNS (webhookDispatcher.js:15-19) — also the key order of buildTableRow's object (:279-304)
__device__ const char kColumnsJson[] =
    "[\"showId\",\"showDate\",\"showTime\",\"showLabel\",\"crew\",\"leadPilot\",\"monkeyLead\",\"showNotes\",\"entryId\","
    "\"unitId\",\"planned\",\"launched\",\"status\",\"primaryIssue\",\"subIssue\",\"otherDetail\",\"severity\",\"rootCause\","
    "\"actions\",\"operator\",\"batteryId\",\"delaySec\",\"commandRx\",\"notes\"]";
constexpr char kColumnNames[24][16] = {
    "showId", "showDate", "showTime", "showLabel", "crew", "leadPilot", "monkeyLead", "showNotes", "entryId", "unitId",
    "planned", "launched", "status", "primaryIssue", "subIssue", "otherDetail", "severity", "rootCause", "actions",
    "operator", "batteryId", "delaySec", "commandRx", "notes"};
// the provider-normalised entry (sqlProvider.js:386-408), in its key order; -1 = ts, -2 = actions, -3 = delaySec
constexpr char kEntryKeys[17][16] = {"id", "ts", "unitId", "planned", "launched", "status", "primaryIssue", "subIssue",
                                     "otherDetail", "severity", "rootCause", "actions", "operator", "batteryId",
                                     "delaySec", "commandRx", "notes"};
__device__ const int kEntryCols[17] = {0, -1, 1, 2, 3, 4, 5, 6, 7, 8, 9, -2, 10, 11, -3, 12, 13};  // index into entry_col()

// a member's name as it stands in an object: `"name":`, with the comma that separates it from the member before it
template <int N>
struct MemberNames {
  char s[N][24];
  int n[N];
};
template <int N>
__host__ __device__ constexpr MemberNames<N> make_member_names(const char (&names)[N][16]) {
  MemberNames<N> m{};
  for (int i = 0; i < N; ++i) {
    int k = 0;
    if (i) m.s[i][k++] = ',';
    m.s[i][k++] = '"';
    for (int j = 0; names[i][j]; ++j) m.s[i][k++] = names[i][j];
    m.s[i][k++] = '"';
    m.s[i][k++] = ':';
    m.n[i] = k;
  }
  return m;
}
__device__ const MemberNames<24> kRowMembers = make_member_names<24>(kColumnNames);
__device__ const MemberNames<17> kEntryMembers = make_member_names<17>(kEntryKeys);

struct Cell {
  const uint8_t* p;  // the cell's bytes; for a list all its items, which are contiguous in the heap
  int n;
  const int32_t* item_offsets;  // lists: offsets of the items (item_offsets[0] .. item_offsets[items]), else nullptr
  int items;
  const uint8_t* heap;
};
__device__ __forceinline__ Cell str_cell(const pie_strcol& c, int64_t i) {
  const int b = c.offsets[i];
  return Cell{c.data + b, c.offsets[i + 1] - b, nullptr, 1, c.data};
}
__device__ __forceinline__ Cell list_cell(const pie_strlistcol& c, int64_t i) {
  const int l0 = c.list_offsets[i], items = c.list_offsets[i + 1] - l0;
  Cell x{nullptr, 0, c.items.offsets + l0, items, c.items.data};
  if (items > 0) {
    const int b = c.items.offsets[l0];
    x.p = c.items.data + b;
    x.n = c.items.offsets[l0 + items] - b;
  }
  return x;
}

// JSON.stringify's escape of one byte of a well-formed UTF-8 string (QuoteJSONString, ECMA-262 25.5.2.3)
__device__ __forceinline__ int json_len(uint8_t c) {
  if (c == '"' || c == '\\') return 2;
  if (c >= 0x20) return 1;
  return (c == 8 || c == 9 || c == 10 || c == 12 || c == 13) ? 2 : 6;
}
__device__ __forceinline__ void json_put(uint8_t* dst, uint8_t c) {
  if (c == '"' || c == '\\') { dst[0] = '\\'; dst[1] = c; return; }
  if (c >= 0x20) { dst[0] = c; return; }
  const char s = c == 8 ? 'b' : c == 9 ? 't' : c == 10 ? 'n' : c == 12 ? 'f' : c == 13 ? 'r' : 0;
  dst[0] = '\\';
  if (s) { dst[1] = (uint8_t)s; return; }
  dst[1] = 'u'; dst[2] = '0'; dst[3] = '0';
  dst[4] = (uint8_t)('0' + (c >> 4));
  dst[5] = (uint8_t)((c & 15) < 10 ? '0' + (c & 15) : 'a' + (c & 15) - 10);
}

// ---- a warp's stage ------------------------------------------------------------------------------------------------
constexpr int kStageWords = 1536;  // 6 KB: the ranges of one show (the bench's shows need 2.3 KB at the median, 4.7 KB at most)
constexpr int kNumSlots = 32;      // entries whose numbers are kept
constexpr int kStageRanges = 25;   // 23 string heaps + crew.list_offsets + actions.list_offsets

struct WarpStage {
  pie_archive_view v;  // the caller's view with the staged columns' pointers rebased into `words`
  uint32_t words[kStageWords];
  char num[2][kNumSlots][kMaxNumberChars];  // [0]: String(delaySec), [1]: String(ts) of entry e0 + i
  uint8_t num_len[2][kNumSlots];
  uint8_t num_finite[2][kNumSlots];
  int num_entries;  // 0: nothing kept (more than kNumSlots entries)
};

template <class T>
__device__ __forceinline__ const T* shfl_ptr(const T* p, int src) {
  return reinterpret_cast<const T*>(__shfl_sync(kFullMask, (unsigned long long)reinterpret_cast<uintptr_t>(p), src));
}

// Stages show s (see the head of the file).  Warp-uniform result: false = the show does not fit, st is not to be used.
__device__ __forceinline__ bool stage_show(WarpStage& st, const pie_archive_view& v, int64_t s, int lane) {
  const int64_t e0 = v.entry_offsets[s], e1 = v.entry_offsets[s + 1];
  // my range: rows [r0, r1] of an offsets array, and (string heaps) the bytes between the offsets at its two ends
  const int32_t* off = nullptr;
  const uint8_t* data = nullptr;
  int64_t r0 = 0, r1 = -1;
  if (lane < 7) {
    const pie_strcol* c = &v.show_id + lane;  // show_id .. show_notes are consecutive members of one type
    off = c->offsets; data = c->data; r0 = s; r1 = s + 1;
  } else if (lane == 7) {
    off = v.crew.items.offsets; data = v.crew.items.data;
    r0 = v.crew.list_offsets[s]; r1 = v.crew.list_offsets[s + 1];
  } else if (lane < 22) {
    const pie_strcol* c = &v.entry_id + (lane - 8);  // entry_id .. notes
    off = c->offsets; data = c->data; r0 = e0; r1 = e1;
  } else if (lane == 22) {
    off = v.actions.items.offsets; data = v.actions.items.data;
    r0 = v.actions.list_offsets[e0]; r1 = v.actions.list_offsets[e1];
  } else if (lane == 23) {
    off = v.crew.list_offsets; r0 = s; r1 = s + 1;
  } else if (lane == 24) {
    off = v.actions.list_offsets; r0 = e0; r1 = e1;
  }
  int noff = 0, nbytes = 0, shift = 0, b0 = 0;
  if (lane < kStageRanges) {
    noff = (int)(r1 - r0 + 1);
    if (data) {
      b0 = off[r0];
      nbytes = off[r1] - b0;
      shift = (int)(reinterpret_cast<uintptr_t>(data + b0) & 3);
    }
  }
  const bool bad = noff < 0 || nbytes < 0;
  const int nwords = nbytes > 0 ? (shift + nbytes + 3) >> 2 : 0;  // the aligned words that hold a byte of the range
  const int mine = bad ? 0 : noff + nwords;
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(kFullMask, incl, o);
    if (lane >= o) incl += t;
  }
  const int total = __shfl_sync(kFullMask, incl, 31);
  if (__any_sync(kFullMask, bad) || total > kStageWords) return false;
  const int start = incl - mine;
  if (lane == 0) st.v = v;
  // the copies: range by range, the whole warp
  const uint32_t* src_words = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(data + b0) & ~static_cast<uintptr_t>(3));
#pragma unroll 1
  for (int h = 0; h < kStageRanges; ++h) {
    const int32_t* h_off = shfl_ptr(off, h) + __shfl_sync(kFullMask, (long long)r0, h);
    const uint32_t* h_src = shfl_ptr(src_words, h);
    const int h_noff = __shfl_sync(kFullMask, noff, h), h_nwords = __shfl_sync(kFullMask, nwords, h);
    const int h_start = __shfl_sync(kFullMask, start, h);
    for (int i = lane; i < h_noff; i += 32) st.words[h_start + i] = (uint32_t)h_off[i];
    for (int i = lane; i < h_nwords; i += 32) st.words[h_start + h_noff + i] = h_src[i];
  }
  __syncwarp();  // lane 0's copy of the view is there
  if (lane < kStageRanges) {
    const int32_t* new_off = reinterpret_cast<const int32_t*>(&st.words[start]) - r0;
    const uint8_t* new_data = reinterpret_cast<const uint8_t*>(&st.words[start + noff]) + shift - b0;
    pie_strcol* c = nullptr;
    if (lane < 7) c = &st.v.show_id + lane;
    else if (lane == 7) c = &st.v.crew.items;
    else if (lane < 22) c = &st.v.entry_id + (lane - 8);
    else if (lane == 22) c = &st.v.actions.items;
    if (c) { c->offsets = new_off; c->data = new_data; }
    else if (lane == 23) st.v.crew.list_offsets = new_off;
    else st.v.actions.list_offsets = new_off;
  }
  // the numbers of the entries, a lane each
  const int n = (int)(e1 - e0);
  if (lane == 0) st.num_entries = n <= kNumSlots ? n : 0;
  if (n <= kNumSlots && lane < n) {
    const RyuTables t{d_pow5_inv, d_pow5};
    const int64_t e = e0 + lane;
    int len = 0;
    bool fin = false;
    if (v.delay_valid[e]) {
      const double x = v.delay_sec[e];
      fin = is_finite_f64(x);
      len = js_number_to_string(x, st.num[0][lane], t);
    }
    st.num_len[0][lane] = (uint8_t)len;
    st.num_finite[0][lane] = fin ? 1 : 0;
    const double ts = v.entry_ts ? v.entry_ts[e] : quiet_nan();
    fin = is_finite_f64(ts);
    len = fin ? js_number_to_string(ts, st.num[1][lane], t) : 0;
    st.num_len[1][lane] = (uint8_t)len;
    st.num_finite[1][lane] = fin ? 1 : 0;
  }
  __syncwarp();
  return true;
}

#define PIE_LIT(em, s) (em).lit(s, (int)sizeof(s) - 1)

template <bool kWrite>
struct Emit {
  uint8_t* out;  // the document (kWrite)
  uint64_t pos;
  int lane;

  __device__ __forceinline__ void lit(const char* s, int n) {
    if (kWrite)
      for (int i = lane; i < n; i += 32) out[pos + i] = (uint8_t)s[i];
    pos += (uint64_t)n;
  }
  __device__ __forceinline__ void ch(char c) {
    if (kWrite && lane == 0) out[pos] = (uint8_t)c;
    ++pos;
  }
  __device__ __forceinline__ void raw(const uint8_t* p, int n) {
    if (kWrite)
      for (int i = lane; i < n; i += 32) out[pos + i] = p[i];
    pos += (uint64_t)n;
  }
  // bytes p[0..n) JSON-escaped, no quotes; kCsvQuoted: they sit inside a csvEscape'd cell that is quoted, so a '"' was
  // doubled first ("" -> \"\")
  template <bool kCsvQuoted>
  __device__ __forceinline__ void escaped(const uint8_t* p, int n) {
    for (int j0 = 0; j0 < n; j0 += 32) {
      const int j = j0 + lane;
      const uint8_t c = j < n ? p[j] : 0;
      const bool dq = kCsvQuoted && c == '"';
      const uint32_t mine = j < n ? (dq ? 4u : (uint32_t)json_len(c)) : 0u;
      if (!__any_sync(kFullMask, mine > 1u)) {  // plain bytes: they go where they are
        if (kWrite && j < n) out[pos + lane] = c;
        pos += (uint64_t)(n - j0 < 32 ? n - j0 : 32);
        continue;
      }
      uint32_t incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(kFullMask, incl, o);
        if (lane >= o) incl += t;
      }
      if (kWrite && j < n) {
        uint8_t* dst = out + pos + incl - mine;
        if (dq) { dst[0] = '\\'; dst[1] = '"'; dst[2] = '\\'; dst[3] = '"'; }
        else json_put(dst, c);
      }
      pos += __shfl_sync(kFullMask, incl, 31);
    }
  }
  __device__ __forceinline__ void jstr(const uint8_t* p, int n) {
    ch('"');
    escaped<false>(p, n);
    ch('"');
  }
  // x || '' of a string column as a JSON string; lists: Array.join('|') (buildTableRow :284, :298).  `sep` (0 = none)
  // is the character that follows the value.
  __device__ __forceinline__ void jcell(const Cell& c, char sep) {
    if (c.items <= 1 && c.n <= 29) {  // the everyday cell: quotes, bytes and separator in one round
      const int n = c.n;
      const bool in = lane >= 1 && lane <= n;
      const uint8_t x = in ? c.p[lane - 1] : 0;
      if (!__any_sync(kFullMask, in && json_len(x) != 1)) {
        if (kWrite) {
          if (in) out[pos + lane] = x;
          else if (lane == 0 || lane == n + 1) out[pos + lane] = '"';
          else if (lane == n + 2 && sep) out[pos + lane] = (uint8_t)sep;
        }
        pos += (uint64_t)(n + 2 + (sep ? 1 : 0));
        return;
      }
    }
    ch('"');
    if (c.items <= 1) {
      escaped<false>(c.p, c.n);
    } else {
      for (int it = 0; it < c.items; ++it) {
        escaped<false>(c.heap + c.item_offsets[it], c.item_offsets[it + 1] - c.item_offsets[it]);
        if (it + 1 < c.items) ch('|');
      }
    }
    if (sep) {
      if (kWrite && lane < 2) out[pos + lane] = lane ? (uint8_t)sep : (uint8_t)'"';
      pos += 2;
    } else {
      ch('"');
    }
  }
  // a list as a JSON array of strings (show.crew, entry.actions as they are stored)
  __device__ __forceinline__ void jarray(const Cell& c) {
    ch('[');
    for (int it = 0; it < c.items; ++it) {
      jstr(c.heap + c.item_offsets[it], c.item_offsets[it + 1] - c.item_offsets[it]);
      if (it + 1 < c.items) ch(',');
    }
    ch(']');
  }
  // csvEscape(value) (:332-338) of a cell, as it appears INSIDE the JSON string that holds the csv row
  __device__ __forceinline__ void csv_cell(const Cell& c, char sep) {
    if (c.items <= 1 && c.n <= 31) {  // the everyday cell: nothing for csvEscape, nothing for JSON
      const int n = c.n;
      const bool in = lane < n;
      const uint8_t x = in ? c.p[lane] : 0;
      if (!__any_sync(kFullMask, in && (x == ',' || json_len(x) != 1))) {  // '"', '\n', '\r' have a JSON escape
        if (kWrite) {
          if (in) out[pos + lane] = x;
          else if (lane == n && sep) out[pos + lane] = (uint8_t)sep;
        }
        pos += (uint64_t)(n + (sep ? 1 : 0));
        return;
      }
    }
    bool special = false;
    for (int j0 = 0; j0 < c.n; j0 += 32) {
      const int j = j0 + lane;
      const uint8_t x = j < c.n ? c.p[j] : 0;
      special |= x == '"' || x == ',' || x == '\n' || x == '\r';
    }
    special = __any_sync(kFullMask, special);
    if (special) PIE_LIT(*this, "\\\"");
    if (c.items <= 1) {
      if (special) escaped<true>(c.p, c.n); else escaped<false>(c.p, c.n);
    } else {
      for (int it = 0; it < c.items; ++it) {
        const uint8_t* p = c.heap + c.item_offsets[it];
        const int n = c.item_offsets[it + 1] - c.item_offsets[it];
        if (special) escaped<true>(p, n); else escaped<false>(p, n);
        if (it + 1 < c.items) ch('|');
      }
    }
    if (special) PIE_LIT(*this, "\\\"");
    if (sep) ch(sep);
  }
  // Number::toString(v), computed by lane 0
  __device__ __forceinline__ void number_text(double v) {
    char buf[kMaxNumberChars];
    int n = 0;
    if (lane == 0) {
      const RyuTables t{d_pow5_inv, d_pow5};
      n = js_number_to_string(v, buf, t);
    }
    n = __shfl_sync(kFullMask, n, 0);
    if (kWrite && lane == 0)
      for (int i = 0; i < n; ++i) out[pos + i] = (uint8_t)buf[i];
    pos += (uint64_t)n;
  }
  // a JS number through JSON.stringify: Number::toString when finite, else null
  __device__ __forceinline__ void number(double v) {
    if (is_finite_f64(v)) number_text(v); else PIE_LIT(*this, "null");
  }
};

__device__ __forceinline__ const pie_strcol& entry_col(const pie_archive_view& v, int k) {
  return *(&v.entry_id + k);  // entry_id .. notes: 14 consecutive members of one type
}

struct PayloadArgs {
  pie_archive_view v;
  const uint8_t* head;
  int head_len;
  const uint8_t* tail;
  int tail_len;
};

// The 24 cells of buildTableRow(show, entry) in EXPORT_COLUMNS order; column 21 (delaySec) is the number.
__device__ __forceinline__ Cell table_cell(const pie_archive_view& v, int64_t s, int64_t e, int col, bool completed) {
  if (col < 8) {
    if (col == 4) return list_cell(v.crew, s);
    return str_cell(*(&v.show_id + (col < 4 ? col : col - 1)), s);  // show_id date time label | lead_pilot monkey_lead notes
  }
  if (col == 18) return list_cell(v.actions, e);
  // entry-level text: 8 entryId .. 17 rootCause are entry columns 0..9, 19 operator, 20 batteryId, 22 commandRx, 23 notes
  const int k = col <= 17 ? col - 8 : col == 19 ? 10 : col == 20 ? 11 : col == 22 ? 12 : 13;
  if (completed && col >= 13 && col <= 17) return Cell{nullptr, 0, nullptr, 1, nullptr};  // :293-297
  return str_cell(entry_col(v, k), e);
}

// what a time field of the show is in the summary: `show.x ?? null` through JSON.stringify
template <bool kWrite>
__device__ __forceinline__ void summary_time(Emit<kWrite>& em, const double* val, const uint8_t* kinds, int64_t s, int f,
                                             int* schema_error) {
  const double x = val ? val[s] : quiet_nan();
  if (is_finite_f64(x)) { em.number(x); return; }
  const int kind = kinds ? kinds[s * PIE_TF_COUNT + f] : PIE_TK_ABSENT;
  if (kind == PIE_TK_TRUE) PIE_LIT(em, "true");
  else if (kind == PIE_TK_FALSE) PIE_LIT(em, "false");
  else {
    if (kind == PIE_TK_STRING || kind == PIE_TK_OTHER) *schema_error = 1;  // the table does not hold the value itself
    PIE_LIT(em, "null");  // null, undefined, and a number that is not finite
  }
}

template <bool kWrite>
__device__ void summary(Emit<kWrite>& em, const pie_archive_view& v, int64_t s, int* schema_error) {
  PIE_LIT(em, "{\"id\":");          em.jcell(str_cell(v.show_id, s), 0);
  PIE_LIT(em, ",\"label\":");       em.jcell(str_cell(v.show_label, s), 0);
  PIE_LIT(em, ",\"date\":");        em.jcell(str_cell(v.show_date, s), 0);
  PIE_LIT(em, ",\"time\":");        em.jcell(str_cell(v.show_time, s), 0);
  PIE_LIT(em, ",\"crew\":");        em.jarray(list_cell(v.crew, s));
  PIE_LIT(em, ",\"leadPilot\":");  em.jcell(str_cell(v.lead_pilot, s), 0);
  PIE_LIT(em, ",\"monkeyLead\":"); em.jcell(str_cell(v.monkey_lead, s), 0);
  PIE_LIT(em, ",\"notes\":");       em.jcell(str_cell(v.show_notes, s), 0);
  PIE_LIT(em, ",\"createdAt\":");  summary_time(em, v.created_at, v.time_kind, s, PIE_TF_CREATED, schema_error);
  PIE_LIT(em, ",\"updatedAt\":");  summary_time(em, v.updated_at, v.time_kind, s, PIE_TF_UPDATED, schema_error);
  PIE_LIT(em, ",\"archivedAt\":"); summary_time(em, v.archived_at, v.time_kind, s, PIE_TF_ARCHIVED, schema_error);
  PIE_LIT(em, ",\"deletedAt\":");  summary_time(em, v.deleted_at, v.time_kind, s, PIE_TF_DELETED, schema_error);
  em.ch('}');
}

// `v`: the view the cells are read through — the warp's staged one (then `st` holds the entries' numbers) or the caller's
template <bool kWrite>
__device__ uint64_t emit_document(const PayloadArgs& a, const pie_archive_view& v, const WarpStage* st, int64_t s, uint8_t* out,
                                  int lane, int* schema_error) {
  Emit<kWrite> em{out, 0, lane};
  const int64_t e0 = v.entry_offsets[s], e1 = v.entry_offsets[s + 1];
  const bool kept = st != nullptr && st->num_entries > 0;  // the numbers of entries e0 .. e1 are in the stage
  // String(delaySec) of a valid delaySec (NaN and Infinity spelled out: csvEscape), or its JSON form (null when not finite)
  auto delay_text = [&](int64_t e, bool json) {
    if (kept) {
      const int i = (int)(e - e0);
      if (json && !st->num_finite[0][i]) PIE_LIT(em, "null");
      else em.raw(reinterpret_cast<const uint8_t*>(st->num[0][i]), st->num_len[0][i]);
    } else if (json) {
      em.number(v.delay_sec[e]);
    } else {
      em.number_text(v.delay_sec[e]);
    }
  };
  em.raw(a.head, a.head_len);
  // ---- table: {columns, rows: tableRows.map(row => EXPORT_COLUMNS.map(column => row[column] ?? ''))}
  PIE_LIT(em, "\"table\":{\"columns\":");
  em.lit(kColumnsJson, (int)sizeof(kColumnsJson) - 1);
  PIE_LIT(em, ",\"rows\":[");
  for (int64_t e = e0; e < e1; ++e) {
    const Cell st_cell = str_cell(v.status, e);
    const bool completed = equals_exact(st_cell.p, st_cell.n, "Completed");
    em.ch('[');
    for (int col = 0; col < 24; ++col) {
      const char sep = col < 23 ? ',' : ']';
      if (col == 21) {  // delaySec === null || undefined ? '' : delaySec
        if (v.delay_valid[e]) delay_text(e, true); else PIE_LIT(em, "\"\"");
        em.ch(sep);
      } else {
        em.jcell(table_cell(v, s, e, col, completed), sep);
      }
    }
    if (e + 1 < e1) em.ch(',');
  }
  // ---- csv: {header, rows: tableRows.map(buildCsvRow)} — every row one JSON string
  PIE_LIT(em, "]},\"csv\":{\"header\":");
  em.lit(kColumnsJson, (int)sizeof(kColumnsJson) - 1);
  PIE_LIT(em, ",\"rows\":[");
  for (int64_t e = e0; e < e1; ++e) {
    const Cell st_cell = str_cell(v.status, e);
    const bool completed = equals_exact(st_cell.p, st_cell.n, "Completed");
    em.ch('"');
    for (int col = 0; col < 24; ++col) {
      const char sep = col < 23 ? ',' : '"';
      if (col == 21) {  // String(delaySec)
        if (v.delay_valid[e]) delay_text(e, false);
        em.ch(sep);
      } else {
        em.csv_cell(table_cell(v, s, e, col, completed), sep);
      }
    }
    if (e + 1 < e1) em.ch(',');
  }
  // ---- message: {show: summary, entries: tableRows}
  PIE_LIT(em, "]},\"message\":{\"show\":");
  summary(em, v, s, schema_error);
  PIE_LIT(em, ",\"entries\":[");
  for (int64_t e = e0; e < e1; ++e) {
    const Cell st_cell = str_cell(v.status, e);
    const bool completed = equals_exact(st_cell.p, st_cell.n, "Completed");
    em.ch('{');
    for (int col = 0; col < 24; ++col) {
      em.lit(kRowMembers.s[col], kRowMembers.n[col]);
      if (col == 21) {
        if (v.delay_valid[e]) delay_text(e, true); else PIE_LIT(em, "\"\"");
      } else {
        em.jcell(table_cell(v, s, e, col, completed), col == 23 ? '}' : 0);
      }
    }
    if (e + 1 < e1) em.ch(',');
  }
  // ---- show: summary, entries: the stored entries (normalizeEntryList keeps them as they are)
  PIE_LIT(em, "]},\"show\":");
  summary(em, v, s, schema_error);
  PIE_LIT(em, ",\"entries\":[");
  for (int64_t e = e0; e < e1; ++e) {
    em.ch('{');
    for (int k = 0; k < 17; ++k) {
      em.lit(kEntryMembers.s[k], kEntryMembers.n[k]);
      const int c = kEntryCols[k];
      if (c == -1) {
        if (kept) {
          const int i = (int)(e - e0);
          if (st->num_finite[1][i]) em.raw(reinterpret_cast<const uint8_t*>(st->num[1][i]), st->num_len[1][i]);
          else PIE_LIT(em, "null");
        } else {
          em.number(v.entry_ts ? v.entry_ts[e] : quiet_nan());
        }
      } else if (c == -2) {
        em.jarray(list_cell(v.actions, e));
      } else if (c == -3) {
        if (v.delay_valid[e]) delay_text(e, true); else PIE_LIT(em, "null");
      } else {
        em.jcell(str_cell(entry_col(v, c), e), k == 16 ? '}' : 0);
      }
    }
    if (e + 1 < e1) em.ch(',');
  }
  em.ch(']');
  em.raw(a.tail, a.tail_len);
  return em.pos;
}

}  // namespace sp
}  // namespace pie
