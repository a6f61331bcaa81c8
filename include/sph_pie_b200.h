/* sph_pie_b200.h — C ABI of the B200 (sm_100a) archive-analytics / export-row path.
 *
 * IMPORTANT CONTEXT.  The reference (sphereisaiahmin-dev/sph-pie) is a Node/Express web app with
 * no FFI boundary and no data-parallel hot path (BASELINE.json).  This header is the
 * boundary a maintainer WOULD bind if the reference's only bulk pure functions were moved off the
 * JS heap: each entry point names the JavaScript function(s) it replaces.  INTEGRATION.md shows the
 * N-API stub.  Nothing here accelerates the reference at the data sizes it can reach (<= ~6.5k
 * entries); see DESIGN.md §0.
 *
 * Conventions
 *  - plain C, no torch / CUDA types in signatures; `stream` is a cudaStream_t passed as void*
 *    (NULL = the legacy default stream).
 *  - every function returns PIE_OK (0) or a negative pie_status; pie_last_error() gives the text.
 *  - "host" entry points take HOST pointers and do H2D + kernels + D2H themselves;
 *    "dev" entry points take DEVICE pointers (inputs already resident in HBM) and only enqueue
 *    kernels on `stream`.
 *
 * Data layout ("archive table", Arrow-style struct-of-arrays)
 *  - a string column is (offsets int32[n+1], data uint8[]) holding UTF-8; value i is
 *    data[offsets[i] .. offsets[i+1]).  JS null / undefined / '' are all the empty string, which is
 *    what every function on this path does with them (`x || ''`).
 *  - a string-list column (crew, actions) is list_offsets int32[n+1] into a string column.
 *  - numbers are IEEE-754 binary64 (JS Number); "absent / null" is a separate validity byte for
 *    delaySec (the path distinguishes null from NaN) and NaN for timestamps (the path only asks
 *    Number.isFinite of them).
 *  - entries of show s are rows entry_offsets[s] .. entry_offsets[s+1] of the entry columns.
 *  The schema is the provider-normalised show (server/storage/sqlProvider.js:361-409): every text
 *  field is a string, delaySec is number|null, actions/crew are string arrays.
 */
#ifndef SPH_PIE_B200_H
#define SPH_PIE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PIE_ABI_VERSION 2 /* 2: pie_archive_view / _table end with updated_at, deleted_at, time_kind */
#define PIE_N_ISSUES 10   /* public/app.js:1-13 PRIMARY_ISSUES */
#define PIE_N_METRICS 19  /* public/app.js:21-86 ARCHIVE_METRIC_DEFS (9) + issue:<name> (10), :3955-3994 */
#define PIE_N_EXPORT_COLUMNS 24 /* server/webhookDispatcher.js:15-19 EXPORT_COLUMNS */

typedef enum pie_status {
  PIE_OK = 0,
  PIE_ERR_CUDA = -1,             /* a CUDA call failed; text in pie_last_error() */
  PIE_ERR_INVALID_ARG = -2,      /* NULL pointer / negative size / inconsistent offsets */
  PIE_ERR_RANGE = -3,            /* JS RangeError "Invalid time value" (public/app.js:3415 toISOString) */
  PIE_ERR_UNSUPPORTED_DATE = -4, /* show.date/time not in the ECMA-262 date-time format; V8's legacy
                                    Date.parse fallback is implementation-defined and not provided */
  PIE_ERR_CAPACITY = -5,         /* caller's output buffer too small; required size is reported */
  PIE_ERR_NO_DEVICE = -6,        /* no sm_100 device visible: there is NO CPU fallback */
  PIE_ERR_SCHEMA = -7,           /* JSON ingest: a stored document is not a provider-normalised show (a text field
                                    that is neither string nor null, delaySec that is neither number nor null, a
                                    string holding a lone surrogate): what pack_shows raises TypeError for */
  PIE_ERR_UNSUPPORTED_JSON = -8  /* JSON ingest: valid JSON outside what the kernel decides exactly — a known key
                                    twice in one object, nesting deeper than 64, bytes that are not UTF-8, a number
                                    of more than 19 significant digits on a rounding boundary */
} pie_status;

typedef struct pie_strcol {
  const int32_t* offsets; /* [n + 1], offsets[0] may be non-zero (sliced columns) */
  const uint8_t* data;    /* UTF-8 bytes */
} pie_strcol;

typedef struct pie_strlistcol {
  const int32_t* list_offsets; /* [n + 1] into items */
  pie_strcol items;
} pie_strlistcol;

/* One batch of shows with their entries.  All pointers are host pointers for *_host entry points
 * and device pointers for *_dev entry points.  Columns an entry point does not read may be NULL;
 * each entry point lists what it reads. */
typedef struct pie_archive_view {
  int64_t n_shows;
  int64_t n_entries;
  const int32_t* entry_offsets; /* [n_shows + 1] */

  /* show-level columns, n_shows rows */
  pie_strcol show_id, show_date, show_time, show_label, lead_pilot, monkey_lead, show_notes;
  pie_strlistcol crew;
  const double* created_at;  /* ms since epoch; NaN when absent / not a finite number */
  const double* archived_at; /* same */

  /* entry-level columns, n_entries rows */
  pie_strcol entry_id, unit_id, planned, launched, status, primary_issue, sub_issue, other_detail,
      severity, root_cause, operator_name, battery_id, command_rx, notes;
  pie_strlistcol actions;
  const double* delay_sec;    /* value (may be NaN/Inf when delay_valid) */
  const uint8_t* delay_valid; /* 0 = null/undefined, 1 = a JS number */
  const double* entry_ts;     /* ms since epoch; NaN when absent */

  /* ABI 2 — what the provider's own code asks of a stored document (_getTimestamp, sqlProvider.js:970-985, coerces:
   * null is 0, '12' is 12, an ISO text goes through Date.parse) needs more than "a finite number or not".  All three
   * may be NULL (then every NaN above counts as an absent field). */
  const double* updated_at;   /* show.updatedAt, like created_at */
  const double* deleted_at;   /* show.deletedAt, like created_at */
  const uint8_t* time_kind;   /* [n_shows][PIE_TF_COUNT]: what the field holds when it is not a finite number, PIE_TK_* */
} pie_archive_view;

/* fields of time_kind, and their values.  PIE_TK_STRING: the NaN in the field's value array carries, in its low 51
 * bits, the byte offset (in pie_json_docs.data) of the string's first character — the JSON ingest puts it there, and
 * pie_get_timestamps_dev reads the text back from the documents. */
enum { PIE_TF_CREATED = 0, PIE_TF_UPDATED = 1, PIE_TF_ARCHIVED = 2, PIE_TF_DELETED = 3, PIE_TF_COUNT = 4 };
enum {
  PIE_TK_ABSENT = 0,    /* undefined: no such key */
  PIE_TK_NUMBER = 1,    /* a finite number: the value array holds it */
  PIE_TK_NULL = 2,
  PIE_TK_TRUE = 3,
  PIE_TK_FALSE = 4,
  PIE_TK_STRING = 5,
  PIE_TK_OTHER = 6,     /* an array or an object */
  PIE_TK_NONFINITE = 7  /* a number that is not finite (1e999 in the text; NaN / Infinity in a packed document) */
};

/* ---- planes of the per-show statistics table -------------------------------------------------
 * stats_i32 is int32[PIE_SI_COUNT][stride], stats_f64 is double[PIE_SF_COUNT][stride], plane-major
 * (plane p of show s at p*stride + s), stride >= n_shows.  Fields mirror the object returned by
 * computeArchiveShowStats (public/app.js:3939-3952).  A JS null is encoded as NaN in the f64
 * planes and is also derivable: avg/max are null iff DELAY_COUNT==0, rates iff TOTAL==0. */
enum {
  PIE_SI_TOTAL = 0,        /* totalEntries */
  PIE_SI_COMPLETED = 1,    /* completedCount */
  PIE_SI_NO_LAUNCH = 2,    /* noLaunchCount */
  PIE_SI_ABORT = 3,        /* abortCount */
  PIE_SI_LAUNCHED = 4,     /* launchedCount */
  PIE_SI_DELAY_COUNT = 5,  /* delayValues.length */
  PIE_SI_ISSUE_COUNT0 = 6, /* issueCounts[PRIMARY_ISSUES[k]], k = 0..9 (0 = key absent) */
  PIE_SI_ISSUE_ORDER_LO = 16, /* property insertion order of issueCounts: 4-bit nibbles, nibble j = */
  PIE_SI_ISSUE_ORDER_HI = 17, /* (k+1) of the j-th distinct issue met, 0 ends; LO = bits 0-31, HI = 32-39 */
  PIE_SI_COUNT = 18
};
enum {
  PIE_SF_AVG_DELAY = 0,       /* avgDelaySec */
  PIE_SF_MAX_DELAY = 1,       /* maxDelaySec */
  PIE_SF_COMPLETION_RATE = 2, /* completionRate */
  PIE_SF_LAUNCH_RATE = 3,     /* launchRate */
  PIE_SF_ABORT_RATE = 4,      /* abortRate */
  PIE_SF_ISSUE_RATE0 = 5,     /* issueRates[PRIMARY_ISSUES[k]] */
  PIE_SF_COUNT = 15
};

/* ---- planes of the per-day summary table -----------------------------------------------------
 * One row per local calendar day that has >= 1 show (a "daily group", public/app.js:3401-3443),
 * ascending by day start.  summary_f64 is double[PIE_DF_COUNT][PIE_N_METRICS][stride],
 * summary_count is int32[PIE_N_METRICS][stride]; metric m in the order of ALL metric keys
 * (ARCHIVE_METRIC_DEFS order, then issue:<PRIMARY_ISSUES[k]>).  average/min/max are NaN (JS null)
 * when count == 0 (public/app.js:3480-3484). */
enum { PIE_DF_AVERAGE = 0, PIE_DF_MIN = 1, PIE_DF_MAX = 2, PIE_DF_COUNT = 3 };

#define PIE_DAY_NONE INT64_MIN /* show has no usable timestamp: skipped by buildArchiveDailyGroups */

typedef struct pie_daily_out {
  int64_t stride;            /* capacity in rows of every array below (>= n_shows is always enough) */
  int64_t* show_day_start;   /* [stride] local-midnight ms of each show, PIE_DAY_NONE if skipped */
  int32_t* show_order;       /* [stride] show indices, stably ordered by day start, skipped ones last */
  int64_t* group_day_start;  /* [stride] group.timestamp */
  int32_t* group_offsets;    /* [stride + 1] group g = show_order[group_offsets[g] .. group_offsets[g+1]) */
  double* summary_f64;       /* [PIE_DF_COUNT][PIE_N_METRICS][stride] */
  int32_t* summary_count;    /* [PIE_N_METRICS][stride] */
  int64_t* n_groups;         /* [1] number of groups written */
  int32_t* status;           /* [2] {pie_status, offending show index}: PIE_ERR_RANGE /
                                PIE_ERR_UNSUPPORTED_DATE raised by a show, else {0, -1} */
} pie_daily_out;

/* ---- library ------------------------------------------------------------------------------- */
int pie_abi_version(void);
const char* pie_last_error(void);
/* Select the CUDA device for the calling thread and verify it is sm_100.  PIE_ERR_NO_DEVICE if
 * there is none — callers must fail, not fall back. */
int pie_init(int device);
int pie_device_sm_count(void);
/* Pinned host memory for the *_host entry points (pageable pointers also work, slower). */
void* pie_host_alloc(uint64_t bytes);
void pie_host_free(void* p);
/* Bytes the most recent *_host call copied host->device and device->host (e2e accounting). */
void pie_last_transfer_bytes(uint64_t* h2d_bytes, uint64_t* d2h_bytes);
/* Cumulative number of CUDA kernels this library has launched in this process. */
uint64_t pie_kernel_launch_count(void);
/* The *_host entry points keep their device staging arenas, output buffers, streams, events and pinned bounce memory
 * between calls (grow-only).  pie_release gives all of it back — for a long-lived host process (the Node server a
 * binding would live in) after a large batch; the next call allocates again.  Synchronises the device. */
int pie_release(void);

/* ---- archive statistics: replaces computeArchiveShowStats (public/app.js:3898-3953), called per
 * show from buildArchiveDailyGroups (:3429-3432).  Reads: entry_offsets, status, launched,
 * primary_issue, delay_sec, delay_valid.  One kernel, no scratch. */
int pie_show_stats_dev(const pie_archive_view* dev_view, int32_t* stats_i32, double* stats_f64,
                       int64_t stride, void* stream);
int pie_show_stats_host(const pie_archive_view* host_view, int32_t* stats_i32, double* stats_f64,
                        int64_t stride);

/* ---- daily groups + metric summaries: replaces buildArchiveDailyGroups (public/app.js:3401-3443,
 * with getShowTimestamp :4092-4116 and parseShowDateTime :4118-4126 for ISO strings) and the
 * numeric part of getOrCreateGroupMetricSummary (:3445-3502) for all PIE_N_METRICS metrics.
 * `tz_offset_minutes`: the zone is a fixed offset east of UTC (setHours(0,0,0,0) is local time).
 * Reads: the stats planes produced above, created_at, archived_at, show_date, show_time,
 * entry_offsets, entry_ts.
 * `scratch` (dev variant): device buffer of pie_daily_scratch_bytes(n_shows) bytes. */
uint64_t pie_daily_scratch_bytes(int64_t n_shows);
int pie_daily_summary_dev(const pie_archive_view* dev_view, const int32_t* stats_i32,
                          const double* stats_f64, int64_t stats_stride, int32_t tz_offset_minutes,
                          const pie_daily_out* dev_out, void* scratch, void* stream);
/* Host variant runs show statistics + daily summary in one call (the reference computes both in
 * buildArchiveDailyGroups); stats_* may be NULL if the caller only wants the summaries. */
int pie_archive_analytics_host(const pie_archive_view* host_view, int32_t tz_offset_minutes,
                               int32_t* stats_i32, double* stats_f64, int64_t stats_stride,
                               const pie_daily_out* host_out);

/* ---- export rows: replaces buildTableRow + csvEscape + buildCsvRow (server/webhookDispatcher.js:
 * 276-342; browser twins public/app.js:5582-5612, :6025-6034) mapped over every entry of every show:
 * the strings dispatchShowEvent puts in csv.rows (:571) and exportShowAsCsv joins with '\n'
 * (public/app.js:5558-5570).  delaySec goes through Number::toString (ECMA-262 6.1.6.1.20).
 * Output is one string column: row i = out_data[row_offsets[i] .. row_offsets[i+1]-1), every row is
 * followed by one '\n' (a show's CSV body is one contiguous slice); row_offsets has n_entries+1
 * elements.  Reads every string column, crew, actions, delay_sec, delay_valid, entry_offsets.
 * dev variant: out_data == NULL computes row_offsets and the total only; if out_capacity is too
 * small nothing past it is written and *total_bytes_dev still reports the size needed.
 * Device string heaps must start 16-byte aligned (any cudaMalloc / torch allocation does). */
uint64_t pie_csv_rows_scratch_bytes(int64_t n_entries);
int pie_csv_rows_dev(const pie_archive_view* dev_view, int64_t* row_offsets, uint8_t* out_data,
                     uint64_t out_capacity, uint64_t* total_bytes_dev, void* scratch, void* stream);
/* host variant: out_data == NULL is a size query (fills row_offsets and *total_bytes);
 * PIE_ERR_CAPACITY if out_capacity < *total_bytes (which is set either way). */
int pie_csv_rows_host(const pie_archive_view* host_view, int64_t* row_offsets, uint8_t* out_data,
                      uint64_t out_capacity, uint64_t* total_bytes);
/* The host variant streams the batch in chunks of about this many rows (cut at show boundaries):
 * the upload of chunk c+1 and the download of chunk c-1 overlap the kernels of chunk c.  Returns the
 * previous value; rows <= 0 only queries.  Default 2^20. */
int64_t pie_set_csv_chunk_rows(int64_t rows);

/* ---- one call for the archive workspace: pie_archive_analytics_host + pie_csv_rows_host on ONE upload.
 * The rows' pipeline already brings status, launched, primaryIssue and delaySec to the device chunk by chunk, so the
 * show statistics are computed per chunk on the resident columns, and the daily groups / summaries once at the end;
 * uploads, kernels and downloads of consecutive chunks overlap (PCIe is full duplex).  Arguments and results are
 * those of the two calls it replaces; stats_i32 / stats_f64 may both be NULL.  Errors: what either call returns
 * (a show's PIE_ERR_RANGE / PIE_ERR_UNSUPPORTED_DATE takes precedence over PIE_ERR_CAPACITY). */
int pie_archive_step_host(const pie_archive_view* host_view, int32_t tz_offset_minutes, int32_t* stats_i32,
                          double* stats_f64, int64_t stats_stride, const pie_daily_out* host_out,
                          int64_t* row_offsets, uint8_t* out_data, uint64_t out_capacity, uint64_t* total_bytes);

/* ---- live show metrics: replaces computeMetrics(show) (public/app.js:5024-5047), the header strip of a show
 * (success rate, status counts, average delay, top issues), for every show of the batch.
 * metrics_i32 is int32[PIE_CM_COUNT][stride], plane-major like the statistics tables:
 *   SUCCESS_RATE  Math.round(completed / plannedYes * 100), 0 when no entry has planned === 'Yes'
 *   COMPLETED / NO_LAUNCH / ABORT  entries whose status === 'Completed' / 'No-launch' / 'Abort' (strict)
 *   TOP0..TOP2    topIssues as ENTRY ROW INDICES: the first entry (in the batch's entry numbering) that carries
 *                 the issue string, -1 past the end of the list; the string is primary_issue[that row].  Order =
 *                 Object.entries(issues) (array-index keys first, ascending; then insertion order) stably sorted by
 *                 count, descending — what .sort((a,b)=>b[1]-a[1]).slice(0,3) leaves.  Work is O(k^2) in the number
 *                 k of a show's entries that carry an issue.
 *   AVG_LEN       length of avgDelay, whose characters are avg_delay_text[show * PIE_CM_TEXT .. +AVG_LEN):
 *                 (sum / count).toFixed(2) over the entries whose delaySec is a number (typeof: NaN and +-Infinity
 *                 count), '0.00' when there is none.  toFixed rounds the EXACT binary value, ties up in magnitude.
 * Reads: entry_offsets, planned, status, primary_issue, delay_sec, delay_valid. */
enum {
  PIE_CM_SUCCESS_RATE = 0,
  PIE_CM_COMPLETED = 1,
  PIE_CM_NO_LAUNCH = 2,
  PIE_CM_ABORT = 3,
  PIE_CM_TOP0 = 4,
  PIE_CM_TOP1 = 5,
  PIE_CM_TOP2 = 6,
  PIE_CM_AVG_LEN = 7,
  PIE_CM_COUNT = 8
};
#define PIE_CM_TEXT 32 /* bytes reserved per show for avgDelay (the longest is 25) */
int pie_compute_metrics_dev(const pie_archive_view* dev_view, int32_t* metrics_i32, uint8_t* avg_delay_text,
                            int64_t stride, void* stream);
int pie_compute_metrics_host(const pie_archive_view* host_view, int32_t* metrics_i32, uint8_t* avg_delay_text,
                             int64_t stride);

/* ---- archive entry payloads: replaces JSON.stringify(buildArchiveEntryPayload(show, entry))
 * (server/webhookDispatcher.js:315-330, with toYesNoBoolean :60-77) mapped over every entry of every show — the
 * request bodies dispatchShowEvent('show.archived') posts one by one (:520-540; axios serialises the object with
 * JSON.stringify).  Row i of the output string column is the body of entry i:
 *   {"showDate":"..","showTime":"..","showNumber":"..","leadPilot":"..","monkeyLead":"..","operator":"..",
 *    "monkeyId":"..","planned":true|false,"launched":..,"commandReceived":..,"primaryIssue":"..","subIssue":".."}
 * followed by '\n' (the whole output is JSON Lines).  Strings are escaped as QuoteJSONString does (ECMA-262
 * 25.5.2.3: \" \\ \b \t \n \f \r, other code units below U+0020 as \u00xx; everything else verbatim) — the
 * columns hold well-formed UTF-8, so the lone-surrogate case of JSON.stringify cannot arise.  A boolean is
 * value.trim().toLowerCase() === 'yes' (the columns are strings: sqlProvider.js:361-409).
 * Same calling convention, scratch size (pie_csv_rows_scratch_bytes) and capacity rules as pie_csv_rows_*.
 * Reads: entry_offsets, show_date, show_time, show_label, lead_pilot, monkey_lead, operator_name, unit_id, planned,
 * launched, command_rx, primary_issue, sub_issue; every other column may be NULL. */
int pie_archive_payloads_dev(const pie_archive_view* dev_view, int64_t* row_offsets, uint8_t* out_data,
                             uint64_t out_capacity, uint64_t* total_bytes_dev, void* scratch, void* stream);
int pie_archive_payloads_host(const pie_archive_view* host_view, int64_t* row_offsets, uint8_t* out_data,
                              uint64_t out_capacity, uint64_t* total_bytes);

/* ---- the schemaVersion 2 show payload: replaces JSON.stringify of the object dispatchShowEvent builds for every
 * event but 'show.archived' (server/webhookDispatcher.js:545-584, with buildShowSummary :472-488, normalizeEntryList
 * :460-470, buildTableRow / buildCsvRow :276-342) for every show of a batch — one JSON document per show:
 *   <head>"table":{"columns":[..],"rows":[[24 values],..]},"csv":{"header":[..],"rows":["<csv row>",..]},
 *   "message":{"show":<summary>,"entries":[{24 members},..]},"show":<summary>,"entries":[{..},..]<tail>
 * `head` is the text up to and including the comma before "table" — `{"event":..,"schemaVersion":2,"dispatchedAt":..,
 * "target":{"url":..,"method":..},` — and `tail` what closes the document (`}` or `,"meta":{..}}`): the caller's
 * strings, serialised by the caller (sph_pie_b200/webhook.py does it).  Table rows keep delaySec a JSON number ('' when
 * null; null when it is not finite, as JSON.stringify writes NaN / Infinity); a csv row is one JSON string; the summary's
 * four timestamps are `show.x ?? null` (numbers, booleans or null — a text / array there is PIE_ERR_SCHEMA: the table
 * does not hold it).  `entries` are written in the provider's normalised shape (sqlProvider.js:384-409: id, ts, unitId,
 * planned, launched, status, primaryIssue, subIssue, otherDetail, severity, rootCause, actions, operator, batteryId,
 * delaySec, commandRx, notes) — what normalizeEntryList passes through for a stored show; keys the table does not
 * hold are not carried.
 * Document s is out_data[doc_offsets[s] .. doc_offsets[s+1]); doc_offsets has n_shows + 1 elements.  out_data == NULL
 * sizes only; nothing is written past out_capacity and *total_bytes_dev reports the size needed.  status_dev[2] =
 * {pie_status, first offending show}.  Reads every column incl. created_at .. time_kind.  All pointers device. */
uint64_t pie_show_payloads_scratch_bytes(int64_t n_shows);
int pie_show_payloads_dev(const pie_archive_view* dev_view, const uint8_t* head, int32_t head_len, const uint8_t* tail,
                          int32_t tail_len, int64_t* doc_offsets, uint8_t* out_data, uint64_t out_capacity,
                          uint64_t* total_bytes_dev, int32_t* status_dev, void* scratch, void* stream);

/* Test hooks of the export-row kernel.  A tile (32..160 consecutive rows, planned per launch from the batch's
 * average row width) whose column bytes or CSV do not fit the kernel's shared-memory staging takes a slower
 * warp-per-row path; `on` = 1 forces every tile
 * through it, 0 restores the default, < 0 only queries; returns the previous value.
 * pie_debug_csv_slow_tiles reads how many tiles of the most recent pie_csv_rows_dev launch that used
 * `scratch` took that path (synchronises `stream`). */
int pie_debug_csv_force_slow_path(int on);
int pie_debug_csv_slow_tiles(const void* scratch, int64_t n_entries, uint32_t* slow_tiles, void* stream);

/* Debug knobs of the JSON ingest (tests, A/B timing).  The ingest decides documents of the provider's own shape
 * (JSON.stringify of _normalizeShow's result, sqlProvider.js:361-409) a warp per document and hands every other
 * document to the thread-per-document walk; results are identical either way.
 * pie_debug_ingest_warp_path: 1 / 0 switches the warp path on / off, < 0 only queries; returns the previous value
 * (on unless the environment variable PIE_INGEST_WARP_PATH=0 is set).
 * pie_debug_ingest_declined: how many documents of the last pie_ingest_measure_dev on `scratch` the warp path
 * declined (synchronises `stream`). */
int pie_debug_ingest_warp_path(int on);
int pie_debug_ingest_declined(const void* scratch, int64_t n_docs, uint32_t* declined, void* stream);

/* ---- JSON ingest: replaces `rows.map(row => this._mapArchiveRow(row)).filter(Boolean)` (server/storage/
 * sqlProvider.js:230-234; _mapArchiveRow :892-926 — JSON.parse(row.data), null unless the value is an object) and
 * `rows.map(r => JSON.parse(r.data))` (:78-82), projected on the archive table: the stored `data` texts (written by
 * JSON.stringify(show), :682 / :696) of a batch of rows go in, the columnar table every other entry point reads
 * comes out, without the documents ever existing as JS objects.  One document per show:
 *   docs.data[docs.offsets[s] .. docs.offsets[s+1])  UTF-8 JSON text (ECMA-404), each document < 2 GiB.
 * Projection (what pack_shows does with JSON.parse's result; keys in any order, unknown keys skipped whatever
 * they hold):
 *   show:  id date time label leadPilot monkeyLead notes -> text columns (string; null / absent = '');
 *          crew -> list of strings when it is an array (null elements = ''), else empty;
 *          createdAt archivedAt updatedAt deletedAt -> the number when it is a finite number, else NaN, and
 *          time_kind says what else the field holds (null, a boolean, a string ...);
 *          entries -> one row per element when it is an array (an element that is not an object is a row without
 *          fields), else none.
 *   entry: id unitId planned launched status primaryIssue subIssue otherDetail severity rootCause operator batteryId
 *          commandRx notes -> text; actions -> list of strings; delaySec -> number (delay_valid 1) or null / absent
 *          (0); ts -> finite number or NaN.
 *   Numbers are correctly rounded binary64 (StringToNumber), strings are unescaped to UTF-8 (surrogate-pair
 *   escapes become one 4-byte sequence).
 * doc_status[s]: 0 = a show; 1 = dropped — the text is not JSON, or JSON that is not an object / array (the
 *   reference maps the row to null and filters it out); its table row is the empty show (no entries, '' texts, NaN
 *   times), which every analytics entry point skips; callers that need the reference's row COUNT compact by it.
 * Anything else fails loudly: PIE_ERR_SCHEMA / PIE_ERR_UNSUPPORTED_JSON (see pie_status) / PIE_ERR_CAPACITY (a
 * heap or row count of 2 GiB and more: split the batch), reported with the first offending document.
 *
 * Two device calls, because the caller owns the memory of the table:
 *   1. pie_ingest_measure_dev: walks every document once; totals_dev[PIE_INGEST_TOTALS] receive the bytes of the 23
 *      string heaps, n_entries and the item counts of crew / actions; status_dev[2] = {pie_status, document};
 *      `scratch` (pie_ingest_scratch_bytes(n_docs), ~2.5 KB per document) keeps where each document's part of every
 *      column starts and what the first pass learnt of each document (records: where every value is and where it
 *      goes), so that the second pass does not parse again.
 *   2. the caller reads totals/status, allocates the table (offsets: rows + 1 elements) and calls
 *      pie_ingest_fill_dev with the same docs / scratch / doc_status: the second walk writes everything.
 *      Not to be called when status[0] != 0. */
#define PIE_INGEST_HEAPS 23
enum {
  /* 0..22: bytes of the string heaps in table order: show_id, show_date, show_time, show_label, lead_pilot,
   * monkey_lead, show_notes, crew.items, entry_id, unit_id, planned, launched, status, primary_issue, sub_issue,
   * other_detail, severity, root_cause, operator_name, battery_id, command_rx, notes, actions.items */
  PIE_IT_ENTRIES = 23,      /* n_entries */
  PIE_IT_CREW_ITEMS = 24,   /* rows of crew.items */
  PIE_IT_ACTION_ITEMS = 25, /* rows of actions.items */
  PIE_INGEST_TOTALS = 26
};

typedef struct pie_json_docs {
  int64_t n_docs;
  const int64_t* offsets; /* [n_docs + 1] */
  const uint8_t* data;    /* inside an allocation that starts 8-byte aligned (any cudaMalloc / torch tensor): the
                             kernels read the aligned 8-byte words that hold a document's bytes */
} pie_json_docs;

typedef struct pie_strcol_mut {
  int32_t* offsets;
  uint8_t* data;
} pie_strcol_mut;
typedef struct pie_strlistcol_mut {
  int32_t* list_offsets;
  pie_strcol_mut items;
} pie_strlistcol_mut;
/* pie_archive_view with writable pointers: same fields, same order, same layout — a filled table is read by the
 * other entry points through a cast / field-wise copy. */
typedef struct pie_archive_table {
  int64_t n_shows;
  int64_t n_entries;
  int32_t* entry_offsets;
  pie_strcol_mut show_id, show_date, show_time, show_label, lead_pilot, monkey_lead, show_notes;
  pie_strlistcol_mut crew;
  double* created_at;
  double* archived_at;
  pie_strcol_mut entry_id, unit_id, planned, launched, status, primary_issue, sub_issue, other_detail, severity,
      root_cause, operator_name, battery_id, command_rx, notes;
  pie_strlistcol_mut actions;
  double* delay_sec;
  uint8_t* delay_valid;
  double* entry_ts;
  double* updated_at;  /* ABI 2: may be NULL (not written) */
  double* deleted_at;
  uint8_t* time_kind;
} pie_archive_table;

uint64_t pie_ingest_scratch_bytes(int64_t n_docs);
int pie_ingest_measure_dev(const pie_json_docs* dev_docs, void* scratch, uint8_t* doc_status, int64_t* totals_dev,
                           int32_t* status_dev, void* stream);
/* fill_scratch: device memory of pie_ingest_fill_scratch_bytes(n_entries) bytes, 32-byte aligned (used when the
 * thread-per-document walk takes every document — pie_debug_ingest_warp_path(0): it writes one 96-byte row per entry
 * there and a coalesced pass turns the rows into the entry columns; untouched otherwise).
 * dev_table->n_shows must be n_docs and dev_table->n_entries the measured total. */
uint64_t pie_ingest_fill_scratch_bytes(int64_t n_entries);
int pie_ingest_fill_dev(const pie_json_docs* dev_docs, const void* scratch, const uint8_t* doc_status,
                        const pie_archive_table* dev_table, void* fill_scratch, void* stream);
/* Host variant: uploads the texts, runs both passes, downloads the table.  `host_table` receives pointers into
 * pinned memory OWNED BY THE LIBRARY (valid until the next pie_ingest_host call or pie_ingest_host_release);
 * doc_status is the caller's [n_docs]; totals (may be NULL) as above; *bad_doc (may be NULL) = the offending
 * document of a failing call, else -1. */
int pie_ingest_host(const pie_json_docs* host_docs, pie_archive_table* host_table, uint8_t* doc_status,
                    int64_t* totals, int64_t* bad_doc);
void pie_ingest_host_release(void);

/* ---- the archive workspace from the provider's stored texts in ONE call: pie_ingest_host + pie_archive_step_host
 * without the table ever leaving the device — listArchivedShows (sqlProvider.js:230-234) feeding
 * buildArchiveDailyGroups / getOrCreateGroupMetricSummary (public/app.js:3401-3502) and the CSV rows of every entry
 * (server/webhookDispatcher.js:276-342).  host_docs / doc_status / bad_doc as pie_ingest_host; statistics and daily
 * summaries as pie_archive_analytics_host (stats_* may both be NULL); CSV rows as pie_csv_rows_host, except that the
 * number of rows is an output too: row_offsets has room for row_capacity elements (n_entries + 1 are needed) and
 * out_data for out_capacity bytes.  *n_entries and *total_bytes are always set; row_offsets == out_data == NULL is a
 * size query (analytics are still delivered); PIE_ERR_CAPACITY if either is too small.  A dropped document is an
 * empty show: no rows, skipped by the daily grouping. */
/* A full request on more than this many documents runs in chunks of that many over three streams (the next chunk's text
 * goes up and the previous chunk's rows go down while a chunk is ingested).  Returns the previous value; docs <= 0 only
 * queries.  Default 131072. */
int64_t pie_set_json_chunk_docs(int64_t docs);
int pie_archive_step_json_host(const pie_json_docs* host_docs, int32_t tz_offset_minutes, uint8_t* doc_status,
                               int32_t* stats_i32, double* stats_f64, int64_t stats_stride,
                               const pie_daily_out* host_out, int64_t* row_offsets, int64_t row_capacity,
                               uint8_t* out_data, uint64_t out_capacity, int64_t* n_entries, uint64_t* total_bytes,
                               int64_t* bad_doc);

/* ---- _getTimestamp(value) (server/storage/sqlProvider.js:970-985) for the four time fields of every show of an
 * ingested table: a finite number as it is; else Number(value) when that is finite (null -> 0, true -> 1, '' -> 0,
 * ' 12 ' -> 12, '0x10' -> 16); else Date.parse of a string; else null (NaN here).  Date.parse is provided for the
 * format ECMA-262 specifies (YYYY-MM-DD, and YYYY-MM-DDTHH:mm[:ss[.sss]] with an optional Z / +-HH:mm, local time of a
 * fixed-offset zone otherwise); any other text is PIE_ERR_UNSUPPORTED_DATE (V8's legacy parser is implementation-
 * defined), an array or object in a time field PIE_ERR_SCHEMA (Number([5]) is 5: not restated).  `dev_docs` are the
 * documents the table was ingested from (needed when a field holds a string; may be NULL otherwise -> PIE_ERR_SCHEMA
 * if one does).  status_dev[2] = {pie_status, first offending show}.  Arrays of `out` may be NULL. */
typedef struct pie_doc_times {
  double* created_at; /* [n_shows] each; NaN = null */
  double* updated_at;
  double* archived_at;
  double* deleted_at;
} pie_doc_times;
int pie_get_timestamps_dev(const pie_archive_view* dev_view, const pie_json_docs* dev_docs, int32_t tz_offset_minutes,
                           const pie_doc_times* dev_out, int32_t* status_dev, void* stream);

/* ---- archive maintenance decisions (server/storage/sqlProvider.js).
 * pie_archive_due_dev: which rows of `shows` _archiveDailyShows (:758-816) archives now.  Rows are grouped by
 * show.date.trim() ('__undated__' when it is empty; `show_date` holds '' for a date that is not a string); a group is
 * due when now - earliest >= 12 h, earliest = the smallest `created[s]` of the group, a NaN (null) counting as 0 —
 * _getTimestamp(null) is 0 (:784), so one show without a usable timestamp makes its whole date group due.
 *   created[s]      _getTimestamp(show.createdAt) ?? _getTimestamp(show.updatedAt): pie_get_timestamps_dev's
 *                   created_at where it is not NaN, else its updated_at
 *   doc_status[s]   != 0: the row did not parse to an object and is skipped (:767-772); may be NULL
 *   due[s]          1 = archived now
 *   group_first[s]  the first row of s's date group: the reference archives and dispatches the due rows in the order
 *                   (group_first, s) — groups in order of first appearance (a Map), rows in row order; -1 = skipped
 * Reads: show_date.  scratch: pie_archive_due_scratch_bytes(n_shows) bytes of device memory. */
uint64_t pie_archive_due_scratch_bytes(int64_t n_shows);
int pie_archive_due_dev(const pie_archive_view* dev_view, const uint8_t* doc_status, const double* created, double now_ms,
                        uint8_t* due, int32_t* group_first, void* scratch, void* stream);
/* pie_archive_expired_dev: _purgeExpiredArchives (:863-890): expired[s] = created[s] is not NaN and
 * now >= _addMonths(created[s], 2) (:991-1009): new Date(t) (TimeClip), setMonth(getMonth() + 2) in LOCAL time — a
 * fixed-offset zone — where a day past the end of the target month carries over (31 Dec -> 3 Mar, 2 Mar in a leap
 * year), getTime().  created[s] = _getTimestamp(show?.createdAt) ?? _getTimestamp(row.created_at). */
int pie_archive_expired_dev(const double* created, int64_t n, double now_ms, int32_t tz_offset_minutes, uint8_t* expired,
                            void* stream);

/* ---- self tests (device code paths that replace an IEEE operation by a faster exact sequence) */
/* Compares the shared-reciprocal quotient used for the rate columns with IEEE a/b for every
 * 0 <= a <= b <= max_b (max_b <= 4096, the largest b the kernels use it for); writes the number of
 * bit mismatches (must be 0). */
int pie_selftest_fast_div(int32_t max_b, uint64_t* mismatches);

#ifdef __cplusplus
}
#endif
#endif /* SPH_PIE_B200_H */
