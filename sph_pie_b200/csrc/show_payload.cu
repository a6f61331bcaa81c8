// The schemaVersion 2 show payload on sm_100a: JSON.stringify of the object dispatchShowEvent builds for every event
// but 'show.archived' (reference server/webhookDispatcher.js:545-584) — buildShowSummary (:472-488) twice, the
// table / csv / message views of every entry (buildTableRow :276-305, buildCsvRow :340-342) and the entries as the
// provider stores them (sqlProvider.js:384-409) — one document per show of a batch:
//
//   <head>"table":{"columns":[..24 names..],"rows":[[24 values],...]},"csv":{"header":[..],"rows":["<csv row>",...]},
//   "message":{"show":<summary>,"entries":[{24 members},...]},"show":<summary>,"entries":[{17 members},...]<tail>
//
// <head> is what precedes "table" (event, schemaVersion, dispatchedAt, target — the caller's strings, serialised by the
// host mirror) and <tail> closes the document (`}` or `,"meta":{...}}`).
//
// A WARP PER SHOW, two passes over the same emitter: the first only adds up the document's length, a device-wide
// exclusive scan places the documents, the second writes.  Every piece (a literal, a JSON-quoted value, a csvEscape'd
// cell inside a JSON string, a number) is produced by the whole warp: a byte per lane, the place of a lane's bytes
// from a warp prefix sum of their escaped lengths.  This is a row of the scope table's "next" section (the last bulk
// pure function of the reference's dispatcher), not part of the timed step; it is written for clarity, not tuned.
#include <cub/device/device_scan.cuh>

#include "pie_device.cuh"
#include "pie_kernels.h"
#include "pie_numfmt.cuh"

namespace pie {

namespace {

static __device__ const uint64_t d_pow5_inv[PIE_RYU_POW5_INV_SPLIT_N][2] = PIE_RYU_POW5_INV_SPLIT_INIT;
static __device__ const uint64_t d_pow5[PIE_RYU_POW5_SPLIT_N][2] = PIE_RYU_POW5_SPLIT_INIT;

// EXPORT_COLUMNS (webhookDispatcher.js:15-19) — also the key order of buildTableRow's object (:279-304)
__device__ const char kColumnsJson[] =
    "[\"showId\",\"showDate\",\"showTime\",\"showLabel\",\"crew\",\"leadPilot\",\"monkeyLead\",\"showNotes\",\"entryId\","
    "\"unitId\",\"planned\",\"launched\",\"status\",\"primaryIssue\",\"subIssue\",\"otherDetail\",\"severity\",\"rootCause\","
    "\"actions\",\"operator\",\"batteryId\",\"delaySec\",\"commandRx\",\"notes\"]";
__device__ const char kColumnNames[24][16] = {
    "showId", "showDate", "showTime", "showLabel", "crew", "leadPilot", "monkeyLead", "showNotes", "entryId", "unitId",
    "planned", "launched", "status", "primaryIssue", "subIssue", "otherDetail", "severity", "rootCause", "actions",
    "operator", "batteryId", "delaySec", "commandRx", "notes"};
// the provider-normalised entry (sqlProvider.js:386-408), in its key order; -1 = ts, -2 = actions, -3 = delaySec
__device__ const char kEntryKeys[17][16] = {"id", "ts", "unitId", "planned", "launched", "status", "primaryIssue", "subIssue",
                                            "otherDetail", "severity", "rootCause", "actions", "operator", "batteryId",
                                            "delaySec", "commandRx", "notes"};
__device__ const int kEntryCols[17] = {0, -1, 1, 2, 3, 4, 5, 6, 7, 8, 9, -2, 10, 11, -3, 12, 13};  // index into entry_col()

__device__ __forceinline__ int cstrlen(const char* s) {
  int n = 0;
  while (s[n]) ++n;
  return n;
}

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
      uint32_t incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += t;
      }
      if (kWrite && j < n) {
        uint8_t* dst = out + pos + incl - mine;
        if (dq) { dst[0] = '\\'; dst[1] = '"'; dst[2] = '\\'; dst[3] = '"'; }
        else json_put(dst, c);
      }
      pos += __shfl_sync(0xFFFFFFFFu, incl, 31);
    }
  }
  __device__ __forceinline__ void jstr(const uint8_t* p, int n) {
    ch('"');
    escaped<false>(p, n);
    ch('"');
  }
  // x || '' of a string column as a JSON string; lists: Array.join('|') (buildTableRow :284, :298)
  __device__ __forceinline__ void jcell(const Cell& c) {
    ch('"');
    if (c.items <= 1) {
      escaped<false>(c.p, c.n);
    } else {
      for (int it = 0; it < c.items; ++it) {
        escaped<false>(c.heap + c.item_offsets[it], c.item_offsets[it + 1] - c.item_offsets[it]);
        if (it + 1 < c.items) ch('|');
      }
    }
    ch('"');
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
  __device__ __forceinline__ void csv_cell(const Cell& c) {
    bool special = false;
    for (int j0 = 0; j0 < c.n; j0 += 32) {
      const int j = j0 + lane;
      const uint8_t x = j < c.n ? c.p[j] : 0;
      special |= x == '"' || x == ',' || x == '\n' || x == '\r';
    }
    special = __any_sync(0xFFFFFFFFu, special);
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
  }
  // a JS number through JSON.stringify: Number::toString when finite, else null
  __device__ __forceinline__ void number(double v) {
    char buf[kMaxNumberChars];
    int n = 0;
    if (lane == 0 && is_finite_f64(v)) {
      const RyuTables t{d_pow5_inv, d_pow5};
      n = js_number_to_string(v, buf, t);
    }
    n = __shfl_sync(0xFFFFFFFFu, n, 0);
    if (n == 0) { PIE_LIT(*this, "null"); return; }
    if (kWrite && lane == 0)
      for (int i = 0; i < n; ++i) out[pos + i] = (uint8_t)buf[i];
    pos += (uint64_t)n;
  }
};

__device__ __forceinline__ const pie_strcol& entry_col(const pie_archive_view& v, int k) {
  const pie_strcol* cols[14] = {&v.entry_id, &v.unit_id, &v.planned, &v.launched, &v.status, &v.primary_issue, &v.sub_issue,
                                &v.other_detail, &v.severity, &v.root_cause, &v.operator_name, &v.battery_id, &v.command_rx,
                                &v.notes};
  return *cols[k];
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
  switch (col) {
    case 0: return str_cell(v.show_id, s);
    case 1: return str_cell(v.show_date, s);
    case 2: return str_cell(v.show_time, s);
    case 3: return str_cell(v.show_label, s);
    case 4: return list_cell(v.crew, s);
    case 5: return str_cell(v.lead_pilot, s);
    case 6: return str_cell(v.monkey_lead, s);
    case 7: return str_cell(v.show_notes, s);
    case 18: return list_cell(v.actions, e);
    default: break;
  }
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
  PIE_LIT(em, "{\"id\":");          em.jcell(str_cell(v.show_id, s));
  PIE_LIT(em, ",\"label\":");       em.jcell(str_cell(v.show_label, s));
  PIE_LIT(em, ",\"date\":");        em.jcell(str_cell(v.show_date, s));
  PIE_LIT(em, ",\"time\":");        em.jcell(str_cell(v.show_time, s));
  PIE_LIT(em, ",\"crew\":");        em.jarray(list_cell(v.crew, s));
  PIE_LIT(em, ",\"leadPilot\":");  em.jcell(str_cell(v.lead_pilot, s));
  PIE_LIT(em, ",\"monkeyLead\":"); em.jcell(str_cell(v.monkey_lead, s));
  PIE_LIT(em, ",\"notes\":");       em.jcell(str_cell(v.show_notes, s));
  PIE_LIT(em, ",\"createdAt\":");  summary_time(em, v.created_at, v.time_kind, s, PIE_TF_CREATED, schema_error);
  PIE_LIT(em, ",\"updatedAt\":");  summary_time(em, v.updated_at, v.time_kind, s, PIE_TF_UPDATED, schema_error);
  PIE_LIT(em, ",\"archivedAt\":"); summary_time(em, v.archived_at, v.time_kind, s, PIE_TF_ARCHIVED, schema_error);
  PIE_LIT(em, ",\"deletedAt\":");  summary_time(em, v.deleted_at, v.time_kind, s, PIE_TF_DELETED, schema_error);
  em.ch('}');
}

template <bool kWrite>
__device__ uint64_t emit_document(const PayloadArgs& a, int64_t s, uint8_t* out, int lane, int* schema_error) {
  const pie_archive_view& v = a.v;
  Emit<kWrite> em{out, 0, lane};
  const int64_t e0 = v.entry_offsets[s], e1 = v.entry_offsets[s + 1];
  em.raw(a.head, a.head_len);
  // ---- table: {columns, rows: tableRows.map(row => EXPORT_COLUMNS.map(column => row[column] ?? ''))}
  PIE_LIT(em, "\"table\":{\"columns\":");
  em.lit(kColumnsJson, (int)sizeof(kColumnsJson) - 1);
  PIE_LIT(em, ",\"rows\":[");
  for (int64_t e = e0; e < e1; ++e) {
    const Cell st = str_cell(v.status, e);
    const bool completed = equals_exact(st.p, st.n, "Completed");
    em.ch('[');
    for (int col = 0; col < 24; ++col) {
      if (col == 21) {  // delaySec === null || undefined ? '' : delaySec
        if (v.delay_valid[e]) em.number(v.delay_sec[e]); else PIE_LIT(em, "\"\"");
      } else {
        em.jcell(table_cell(v, s, e, col, completed));
      }
      if (col < 23) em.ch(',');
    }
    em.ch(']');
    if (e + 1 < e1) em.ch(',');
  }
  // ---- csv: {header, rows: tableRows.map(buildCsvRow)} — every row one JSON string
  PIE_LIT(em, "]},\"csv\":{\"header\":");
  em.lit(kColumnsJson, (int)sizeof(kColumnsJson) - 1);
  PIE_LIT(em, ",\"rows\":[");
  for (int64_t e = e0; e < e1; ++e) {
    const Cell st = str_cell(v.status, e);
    const bool completed = equals_exact(st.p, st.n, "Completed");
    em.ch('"');
    for (int col = 0; col < 24; ++col) {
      if (col == 21) {  // String(delaySec): NaN and Infinity are spelled out here (csvEscape, not JSON.stringify)
        if (v.delay_valid[e]) {
          char buf[kMaxNumberChars];
          int n = 0;
          if (lane == 0) {
            const RyuTables t{d_pow5_inv, d_pow5};
            n = js_number_to_string(v.delay_sec[e], buf, t);
          }
          n = __shfl_sync(0xFFFFFFFFu, n, 0);
          if (kWrite && lane == 0)
            for (int i = 0; i < n; ++i) em.out[em.pos + i] = (uint8_t)buf[i];
          em.pos += (uint64_t)n;
        }
      } else {
        em.csv_cell(table_cell(v, s, e, col, completed));
      }
      if (col < 23) em.ch(',');
    }
    em.ch('"');
    if (e + 1 < e1) em.ch(',');
  }
  // ---- message: {show: summary, entries: tableRows}
  PIE_LIT(em, "]},\"message\":{\"show\":");
  summary(em, v, s, schema_error);
  PIE_LIT(em, ",\"entries\":[");
  for (int64_t e = e0; e < e1; ++e) {
    const Cell st = str_cell(v.status, e);
    const bool completed = equals_exact(st.p, st.n, "Completed");
    em.ch('{');
    for (int col = 0; col < 24; ++col) {
      em.ch('"');
      em.lit(kColumnNames[col], cstrlen(kColumnNames[col]));
      PIE_LIT(em, "\":");
      if (col == 21) {
        if (v.delay_valid[e]) em.number(v.delay_sec[e]); else PIE_LIT(em, "\"\"");
      } else {
        em.jcell(table_cell(v, s, e, col, completed));
      }
      if (col < 23) em.ch(',');
    }
    em.ch('}');
    if (e + 1 < e1) em.ch(',');
  }
  // ---- show: summary, entries: the stored entries (normalizeEntryList keeps them as they are)
  PIE_LIT(em, "]},\"show\":");
  summary(em, v, s, schema_error);
  PIE_LIT(em, ",\"entries\":[");
  for (int64_t e = e0; e < e1; ++e) {
    em.ch('{');
    for (int k = 0; k < 17; ++k) {
      em.ch('"');
      em.lit(kEntryKeys[k], cstrlen(kEntryKeys[k]));
      PIE_LIT(em, "\":");
      const int c = kEntryCols[k];
      if (c == -1) em.number(v.entry_ts ? v.entry_ts[e] : quiet_nan());
      else if (c == -2) em.jarray(list_cell(v.actions, e));
      else if (c == -3) { if (v.delay_valid[e]) em.number(v.delay_sec[e]); else PIE_LIT(em, "null"); }
      else em.jcell(str_cell(entry_col(v, c), e));
      if (k < 16) em.ch(',');
    }
    em.ch('}');
    if (e + 1 < e1) em.ch(',');
  }
  em.ch(']');
  em.raw(a.tail, a.tail_len);
  return em.pos;
}

constexpr int kWarpsPerCta = 4;

__global__ void __launch_bounds__(32 * kWarpsPerCta) payload_measure_kernel(PayloadArgs a, int64_t* __restrict__ doc_len,
                                                                           int32_t* __restrict__ status) {
  const int64_t s = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  if (s >= a.v.n_shows) return;
  int schema_error = 0;
  const uint64_t n = emit_document<false>(a, s, nullptr, threadIdx.x & 31, &schema_error);
  if ((threadIdx.x & 31) == 0) {
    doc_len[s] = (int64_t)n;
    if (schema_error) atomicMin(reinterpret_cast<unsigned int*>(status + 1), (unsigned int)s);
  }
}

__global__ void __launch_bounds__(32 * kWarpsPerCta) payload_write_kernel(PayloadArgs a, const int64_t* __restrict__ doc_offsets,
                                                                         uint8_t* __restrict__ out, uint64_t capacity) {
  const int64_t s = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  if (s >= a.v.n_shows) return;
  const int64_t at = doc_offsets[s], end = doc_offsets[s + 1];
  if ((uint64_t)end > capacity) return;  // nothing past the caller's buffer
  int schema_error = 0;
  emit_document<true>(a, s, out + at, threadIdx.x & 31, &schema_error);
}

__global__ void payload_finish_kernel(int64_t n, const int64_t* __restrict__ doc_len, int64_t* __restrict__ doc_offsets,
                                      unsigned long long* __restrict__ total, int32_t* __restrict__ status) {
  // doc_offsets[0..n) hold the exclusive sums; the terminal offset and the status word are written here
  const int64_t last = n > 0 ? doc_offsets[n - 1] + doc_len[n - 1] : 0;
  doc_offsets[n] = last;
  *total = (unsigned long long)last;
  const unsigned int bad = (unsigned int)status[1];
  if (bad != 0xFFFFFFFFu) status[0] = PIE_ERR_SCHEMA;
  else { status[0] = 0; status[1] = -1; }
}

__global__ void payload_init_kernel(int32_t* status) {
  status[0] = 0;
  status[1] = -1;  // 0xFFFFFFFF: no offending show yet
}

}  // namespace

uint64_t show_payload_scratch_bytes(int64_t n_shows) {
  size_t temp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, temp, (const int64_t*)nullptr, (int64_t*)nullptr, (int)(n_shows > 0 ? n_shows : 1));
  return ((uint64_t)temp + 255) / 256 * 256 + 8ull * (uint64_t)(n_shows > 0 ? n_shows : 1) + 256;
}

cudaError_t launch_show_payloads(const pie_archive_view& v, const uint8_t* head, int head_len, const uint8_t* tail,
                                 int tail_len, int64_t* doc_offsets, uint8_t* out, uint64_t capacity,
                                 unsigned long long* total, int32_t* status, void* scratch, cudaStream_t stream) {
  const int64_t n = v.n_shows;
  PayloadArgs a{v, head, head_len, tail, tail_len};
  size_t temp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, temp, (const int64_t*)nullptr, (int64_t*)nullptr, (int)(n > 0 ? n : 1));
  uint8_t* base = static_cast<uint8_t*>(scratch);
  int64_t* doc_len = reinterpret_cast<int64_t*>(base + ((uint64_t)temp + 255) / 256 * 256);
  payload_init_kernel<<<1, 1, 0, stream>>>(status);
  const unsigned blocks = (unsigned)((n + kWarpsPerCta - 1) / kWarpsPerCta);
  if (n > 0) {
    payload_measure_kernel<<<blocks, 32 * kWarpsPerCta, 0, stream>>>(a, doc_len, status);
    cudaError_t e = cub::DeviceScan::ExclusiveSum(base, temp, doc_len, doc_offsets, (int)n, stream);
    if (e != cudaSuccess) return e;
  }
  payload_finish_kernel<<<1, 1, 0, stream>>>(n, doc_len, doc_offsets, total, status);
  if (n > 0 && out) payload_write_kernel<<<blocks, 32 * kWarpsPerCta, 0, stream>>>(a, doc_offsets, out, capacity);
  g_launches += 2 + (n > 0 ? 2 + (out ? 1 : 0) : 0);
  return cudaGetLastError();
}

}  // namespace pie
