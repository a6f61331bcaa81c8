// C ABI of libsphpie_b200 (include/sph_pie_b200.h).  Host-buffer entry points stage the columns an
// operation reads into a grow-only device arena, run the kernels and copy the results back;
// device entry points only enqueue kernels.  There is no CPU fallback anywhere in this file.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "pie_kernels.h"

namespace {

thread_local char g_err[512] = "";
using pie::g_sm_count;  // set by pie_init; the launchers size their grids from it
std::mutex g_host_mutex;  // host entry points share one arena + stream

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define PIE_CUDA(expr)                                                                        \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      return fail(PIE_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

int ensure_init() {
  if (g_sm_count > 0) return PIE_OK;
  return pie_init(-1);
}

// Grow-only device arena with bump allocation, reset at the start of every host entry point.
struct Arena {
  uint8_t* base = nullptr;
  uint64_t cap = 0, used = 0;
  cudaStream_t stream = nullptr;

  int reserve(uint64_t bytes) {
    if (!stream) PIE_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    used = 0;
    if (bytes <= cap) return PIE_OK;
    if (base) PIE_CUDA(cudaFree(base));
    base = nullptr;
    cap = 0;
    PIE_CUDA(cudaMalloc(&base, bytes));
    cap = bytes;
    return PIE_OK;
  }
  void* take(uint64_t bytes) {
    uint64_t off = (used + 255) & ~(uint64_t)255;
    used = off + bytes;
    return base + off;
  }
};
Arena g_arena;
// The upload helpers below stage into *g_cur on g_cur_stream: the analytics path uses g_arena; the
// pipelined export path flips between two input arenas (all under g_host_mutex).
Arena g_pipe_in[2];
Arena* g_cur = &g_arena;
cudaStream_t g_cur_stream = nullptr;

inline uint64_t pad(uint64_t b) { return ((b + 255) & ~(uint64_t)255) + 256; }

struct StrColPlan {
  const pie_strcol* src;
  pie_strcol* dst;
  int64_t n;
  int32_t first, last;
  const char* name;
};

// The offsets arrays uploaded since the last begin_checks(): validated on the device (validate.cu) before any kernel
// walks them.  Filled by the upload helpers, under g_host_mutex.
struct PendingChecks {
  pie::OffsetsBatch batch;
  const char* names[pie::kMaxOffsetsArrays];
  void clear() { batch.count = 0; }
  void add(const int32_t* dev_off, int64_t n, int32_t first, int32_t last, const char* name) {
    if (n <= 0 || batch.count >= pie::kMaxOffsetsArrays) return;  // (a batch has at most 27 arrays)
    names[batch.count] = name;
    batch.a[batch.count++] = pie::OffsetsArray{dev_off, n, first, last};
  }
};
PendingChecks g_checks;
int32_t* g_check_flags = nullptr;  // device, [3][kMaxOffsetsArrays]: one set for the arena path, one per pipeline slot
unsigned long long* g_maint_err = nullptr;  // device word: the first offending show of pie_get_timestamps_dev

int ensure_check_flags();
// enqueue the check of everything uploaded since clear(), flags -> `host_flags` (any host memory) on `st`
int enqueue_checks(const PendingChecks& c, int set, int32_t* host_flags, cudaStream_t st, bool mapped = false);
// after `st` was synchronised: PIE_ERR_INVALID_ARG naming the first malformed array
int report_checks(const PendingChecks& c, const int32_t* host_flags);

// bytes needed on the device for a string column of n rows
int plan_strcol(StrColPlan& p, const pie_strcol* src, pie_strcol* dst, int64_t n, uint64_t* bytes, const char* name) {
  p.src = src; p.dst = dst; p.n = n; p.name = name;
  if (!src->offsets) return fail(PIE_ERR_INVALID_ARG, "column %s: offsets is NULL", name);
  p.first = src->offsets[0];
  p.last = src->offsets[n];
  if (p.last < p.first) return fail(PIE_ERR_INVALID_ARG, "column %s: offsets decrease", name);
  if (p.last > p.first && !src->data) return fail(PIE_ERR_INVALID_ARG, "column %s: data is NULL", name);
  *bytes += pad(4 * (uint64_t)(n + 1)) + pad((uint64_t)(p.last - p.first));
  return PIE_OK;
}

// first <= off[i] <= off[i+1] <= last for every row: with off[0] = first and off[n] = last that is "never decreases".
// A malformed interior offset would make the kernels read outside the staged range.  Returns the first bad row or -1.
int64_t first_decrease(const int32_t* off, int64_t n) {
  int64_t i = 0;
  for (; i + 1024 <= n; i += 1024) {  // branch-free blocks (vectorised), located only on failure
    int bad = 0;
    for (int64_t k = i; k < i + 1024; ++k) bad |= off[k + 1] < off[k];
    if (bad) break;
  }
  for (; i < n; ++i)
    if (off[i + 1] < off[i]) return i;
  return -1;
}

// entry_offsets is walked by the host itself (chunking, slicing): checked here, on the host (4 bytes per show)
int check_entry_offsets(const int32_t* off, int64_t n) {
  const int64_t bad = (off && n > 0) ? first_decrease(off, n) : -1;
  if (bad >= 0) return fail(PIE_ERR_INVALID_ARG, "column entry_offsets: offsets decrease at row %lld", (long long)bad);
  return PIE_OK;
}

int ensure_check_flags() {
  if (g_check_flags) return PIE_OK;
  PIE_CUDA(cudaMalloc(&g_check_flags, sizeof(int32_t) * 3 * pie::kMaxOffsetsArrays));
  return PIE_OK;
}
// mapped: host_flags is pinned memory the device can write (the pipelines: no copy engine involved)
int enqueue_checks(const PendingChecks& c, int set, int32_t* host_flags, cudaStream_t st, bool mapped) {
  static const bool skip = getenv("PIE_DEBUG_SKIP_OFFSET_CHECK") != nullptr;  // timing experiments only
  memset(host_flags, 0, sizeof(int32_t) * pie::kMaxOffsetsArrays);
  if (skip || c.batch.count == 0) return PIE_OK;
  int rc = ensure_check_flags();
  if (rc) return rc;
  int32_t* d = g_check_flags + set * pie::kMaxOffsetsArrays;
  PIE_CUDA(pie::launch_offsets_check(c.batch, d, st));
  if (mapped) PIE_CUDA(pie::launch_offsets_flags_out(d, host_flags, st));
  else PIE_CUDA(cudaMemcpyAsync(host_flags, d, sizeof(int32_t) * pie::kMaxOffsetsArrays, cudaMemcpyDeviceToHost, st));
  return PIE_OK;
}
int report_checks(const PendingChecks& c, const int32_t* host_flags) {
  for (int i = 0; i < c.batch.count; ++i)
    if (host_flags[i])
      return fail(PIE_ERR_INVALID_ARG, "column %s: offsets decrease or leave the column's heap", c.names[i]);
  return PIE_OK;
}

int upload_strcol(const StrColPlan& p, uint64_t* h2d) {
  int32_t* d_off = (int32_t*)g_cur->take(4 * (uint64_t)(p.n + 1));
  const uint64_t nbytes = (uint64_t)(p.last - p.first);
  uint8_t* d_data = (uint8_t*)g_cur->take(nbytes ? nbytes : 1);
  PIE_CUDA(cudaMemcpyAsync(d_off, p.src->offsets, 4 * (uint64_t)(p.n + 1), cudaMemcpyHostToDevice, g_cur_stream));
  if (nbytes)
    PIE_CUDA(cudaMemcpyAsync(d_data, p.src->data + p.first, nbytes, cudaMemcpyHostToDevice, g_cur_stream));
  p.dst->offsets = d_off;
  p.dst->data = d_data - p.first;  // offsets keep their host values
  g_checks.add(d_off, p.n, p.first, p.last, p.name);
  *h2d += 4 * (uint64_t)(p.n + 1) + nbytes;
  return PIE_OK;
}

template <typename T>
int upload_array(const T* src, int64_t n, const T** dst, uint64_t* h2d) {
  T* d = (T*)g_cur->take(sizeof(T) * (uint64_t)(n > 0 ? n : 1));
  if (n > 0) PIE_CUDA(cudaMemcpyAsync(d, src, sizeof(T) * (uint64_t)n, cudaMemcpyHostToDevice, g_cur_stream));
  *dst = d;
  *h2d += sizeof(T) * (uint64_t)n;
  return PIE_OK;
}

int check_view_common(const pie_archive_view* v) {
  if (!v) return fail(PIE_ERR_INVALID_ARG, "view is NULL");
  if (v->n_shows < 0 || v->n_entries < 0) return fail(PIE_ERR_INVALID_ARG, "negative row count");
  if (v->n_shows > 0x7FFFFFF0LL || v->n_entries > 0x7FFFFFF0LL)
    return fail(PIE_ERR_INVALID_ARG, "batch too large: split into batches of < 2^31 rows");
  if (!v->entry_offsets) return fail(PIE_ERR_INVALID_ARG, "entry_offsets is NULL");
  return PIE_OK;
}

uint64_t g_last_h2d = 0, g_last_d2h = 0;

// Every *_host entry point enqueues copies into the caller's buffers.  Whatever way it is left — an error in the
// middle included — none of them may still be in flight: the caller is free to release its memory when the call
// returns.  (On the success path the stream is already idle and this costs nothing.)
struct StreamDrain {
  const cudaStream_t* stream;  // read at scope exit: the arena creates its stream on first use
  ~StreamDrain() {
    if (*stream) cudaStreamSynchronize(*stream);
  }
};

// Grow-only device buffer for outputs whose size is only known after a device pass (CSV bytes):
// growing it must not move the inputs already staged in an arena.
struct OutBuffer {
  uint8_t* base = nullptr;
  uint64_t cap = 0;
  int ensure(uint64_t bytes) {
    if (bytes <= cap) return PIE_OK;
    if (base) PIE_CUDA(cudaFree(base));
    base = nullptr;
    cap = 0;
    PIE_CUDA(cudaMalloc(&base, bytes));
    cap = bytes;
    return PIE_OK;
  }
};

struct ColumnRef {
  const pie_strcol* src;
  pie_strcol* dst;
  int64_t n;
  const char* name;
};
struct ListRef {
  const pie_strlistcol* src;
  pie_strlistcol* dst;
  int64_t n;
  const char* name;
};

// Row formats of the export path (pie_kernels.h launchers)
enum RowFormat { kFormatCsv = 0, kFormatPayload = 1 };
// does the format read string column i of the table below?  (CSV: all of them)
static bool format_reads(RowFormat f, int i) {
  if (f == kFormatCsv) return true;
  // payload: show_date, show_time, show_label, lead_pilot, monkey_lead | unit_id, planned, launched, primary_issue,
  // sub_issue, operator_name, command_rx
  static const bool payload[21] = {false, true, true, true, true, true, false, false, true, true, true,
                                   false, true, true, false, false, false, true, false, true, false};
  return payload[i];
}

// Stage every column the format reads (CSV: all 21 string columns, 2 list columns, delaySec).
int upload_export_view(const pie_archive_view* hv, pie_archive_view* dv, uint64_t extra_bytes, uint64_t* h2d,
                       RowFormat format) {
  const int64_t S = hv->n_shows, E = hv->n_entries;
  const bool csv = format == kFormatCsv;
  ColumnRef cols[] = {
      {&hv->show_id, &dv->show_id, S, "show_id"},           {&hv->show_date, &dv->show_date, S, "show_date"},
      {&hv->show_time, &dv->show_time, S, "show_time"},     {&hv->show_label, &dv->show_label, S, "show_label"},
      {&hv->lead_pilot, &dv->lead_pilot, S, "lead_pilot"},  {&hv->monkey_lead, &dv->monkey_lead, S, "monkey_lead"},
      {&hv->show_notes, &dv->show_notes, S, "show_notes"},  {&hv->entry_id, &dv->entry_id, E, "entry_id"},
      {&hv->unit_id, &dv->unit_id, E, "unit_id"},           {&hv->planned, &dv->planned, E, "planned"},
      {&hv->launched, &dv->launched, E, "launched"},        {&hv->status, &dv->status, E, "status"},
      {&hv->primary_issue, &dv->primary_issue, E, "primary_issue"},
      {&hv->sub_issue, &dv->sub_issue, E, "sub_issue"},     {&hv->other_detail, &dv->other_detail, E, "other_detail"},
      {&hv->severity, &dv->severity, E, "severity"},        {&hv->root_cause, &dv->root_cause, E, "root_cause"},
      {&hv->operator_name, &dv->operator_name, E, "operator_name"},
      {&hv->battery_id, &dv->battery_id, E, "battery_id"},  {&hv->command_rx, &dv->command_rx, E, "command_rx"},
      {&hv->notes, &dv->notes, E, "notes"}};
  ListRef lists[] = {{&hv->crew, &dv->crew, S, "crew"}, {&hv->actions, &dv->actions, E, "actions"}};
  constexpr int kCols = sizeof(cols) / sizeof(cols[0]);
  StrColPlan plans[kCols], item_plans[2];
  pie_strcol item_src[2];   // the item rows this batch's lists point at (lists may be slices)
  int32_t item_first[2];
  uint64_t bytes = extra_bytes + pad(4 * (uint64_t)(S + 1)) + pad(8 * (uint64_t)E) + pad((uint64_t)E);
  int rc;
  for (int i = 0; i < kCols; ++i)
    if (format_reads(format, i) &&
        (rc = plan_strcol(plans[i], cols[i].src, cols[i].dst, cols[i].n, &bytes, cols[i].name)))
      return rc;
  for (int i = 0; csv && i < 2; ++i) {
    const pie_strlistcol* l = lists[i].src;
    if (!l->list_offsets || !l->items.offsets)
      return fail(PIE_ERR_INVALID_ARG, "column %s: list_offsets / items.offsets is NULL", lists[i].name);
    item_first[i] = l->list_offsets[0];
    const int64_t n_items = (int64_t)l->list_offsets[lists[i].n] - item_first[i];
    if (n_items < 0) return fail(PIE_ERR_INVALID_ARG, "column %s: list_offsets decrease", lists[i].name);
    item_src[i].offsets = l->items.offsets + item_first[i];
    item_src[i].data = l->items.data;
    bytes += pad(4 * (uint64_t)(lists[i].n + 1));
    if ((rc = plan_strcol(item_plans[i], &item_src[i], &lists[i].dst->items, n_items, &bytes, lists[i].name))) return rc;
  }
  if (csv && E > 0 && (!hv->delay_sec || !hv->delay_valid))
    return fail(PIE_ERR_INVALID_ARG, "delay_sec/delay_valid is NULL");
  if ((rc = check_entry_offsets(hv->entry_offsets, S))) return rc;
  if ((rc = g_cur->reserve(bytes))) return rc;
  if (!g_cur_stream) g_cur_stream = g_cur->stream;
  dv->n_shows = S;
  dv->n_entries = E;
  if ((rc = upload_array(hv->entry_offsets, S + 1, &dv->entry_offsets, h2d))) return rc;
  for (int i = 0; i < kCols; ++i)
    if (format_reads(format, i) && (rc = upload_strcol(plans[i], h2d))) return rc;
  if (!csv) return PIE_OK;
  for (int i = 0; i < 2; ++i) {
    if ((rc = upload_array(lists[i].src->list_offsets, lists[i].n + 1, &lists[i].dst->list_offsets, h2d))) return rc;
    g_checks.add(lists[i].dst->list_offsets, lists[i].n, item_first[i], item_first[i] + (int32_t)item_plans[i].n, lists[i].name);
    if ((rc = upload_strcol(item_plans[i], h2d))) return rc;
    lists[i].dst->items.offsets -= item_first[i];  // list offsets keep their host (absolute) values
  }
  if ((rc = upload_array(hv->delay_sec, E, &dv->delay_sec, h2d))) return rc;
  if ((rc = upload_array(hv->delay_valid, E, &dv->delay_valid, h2d))) return rc;
  return PIE_OK;
}

int check_export_view_dev(const pie_archive_view* v, RowFormat format) {
  const pie_strcol* cols[] = {&v->show_id, &v->show_date, &v->show_time, &v->show_label, &v->lead_pilot, &v->monkey_lead,
                              &v->show_notes, &v->entry_id, &v->unit_id, &v->planned, &v->launched, &v->status,
                              &v->primary_issue, &v->sub_issue, &v->other_detail, &v->severity, &v->root_cause,
                              &v->operator_name, &v->battery_id, &v->command_rx, &v->notes};
  for (int i = 0; i < 21; ++i)
    if (format_reads(format, i) && !cols[i]->offsets)  // data may be NULL for a column of empty strings
      return fail(PIE_ERR_INVALID_ARG, "a string column this row format reads is NULL");
  if (format != kFormatCsv) return PIE_OK;
  if (!v->crew.items.offsets || !v->actions.items.offsets || !v->crew.list_offsets || !v->actions.list_offsets)
    return fail(PIE_ERR_INVALID_ARG, "list_offsets / items of crew or actions is NULL");
  if (v->n_entries > 0 && (!v->delay_sec || !v->delay_valid)) return fail(PIE_ERR_INVALID_ARG, "delay_sec/delay_valid is NULL");
  return PIE_OK;
}

static cudaError_t launch_rows(RowFormat format, const pie_archive_view& v, int64_t* row_offsets, uint8_t* out_data,
                               uint64_t capacity, unsigned long long bias, unsigned long long* total_out, void* scratch,
                               cudaStream_t stream) {
  return format == kFormatCsv
             ? pie::launch_csv_rows(v, row_offsets, out_data, capacity, bias, total_out, scratch, stream)
             : pie::launch_payload_rows(v, row_offsets, out_data, capacity, bias, total_out, scratch, stream);
}

}  // namespace

extern "C" {

int pie_abi_version(void) { return PIE_ABI_VERSION; }

const char* pie_last_error(void) { return g_err; }

int pie_init(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    return fail(PIE_ERR_NO_DEVICE, "no CUDA device visible (%s); this library has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  }
  if (device >= 0) PIE_CUDA(cudaSetDevice(device));
  int dev = 0;
  PIE_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  PIE_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10 || prop.minor != 0)  // the `a` targets are not forward compatible: sm_103 has no image here
    return fail(PIE_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", dev, prop.major,
                prop.minor);
  g_sm_count = prop.multiProcessorCount;
  return PIE_OK;
}

int pie_device_sm_count(void) { return g_sm_count; }

void* pie_host_alloc(uint64_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    fail(PIE_ERR_CUDA, "cudaHostAlloc(%llu) failed", (unsigned long long)bytes);
    return nullptr;
  }
  return p;
}

void pie_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

/* bytes moved by the most recent *_host call on this process (for bench.py's e2e accounting) */
void pie_last_transfer_bytes(uint64_t* h2d, uint64_t* d2h) {
  if (h2d) *h2d = g_last_h2d;
  if (d2h) *d2h = g_last_d2h;
}

uint64_t pie_kernel_launch_count(void) { return pie::g_launches; }

int pie_show_stats_dev(const pie_archive_view* v, int32_t* stats_i32, double* stats_f64, int64_t stride,
                       void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  if ((rc = check_view_common(v))) return rc;
  if (!stats_i32 || !stats_f64) return fail(PIE_ERR_INVALID_ARG, "output is NULL");
  if (stride < v->n_shows) return fail(PIE_ERR_INVALID_ARG, "stride < n_shows");
  if (v->n_entries > 0 && (!v->status.offsets || !v->launched.offsets || !v->primary_issue.offsets || !v->delay_sec ||
                           !v->delay_valid))
    return fail(PIE_ERR_INVALID_ARG, "show stats reads status, launched, primary_issue, delay_sec, delay_valid");
  PIE_CUDA(pie::launch_show_stats(*v, stats_i32, stats_f64, stride, g_sm_count, (cudaStream_t)stream));
  return PIE_OK;
}

uint64_t pie_daily_scratch_bytes(int64_t n_shows) { return pie::daily_scratch_bytes(n_shows); }

int pie_daily_summary_dev(const pie_archive_view* v, const int32_t* stats_i32, const double* stats_f64,
                          int64_t stats_stride, int32_t tz_offset_minutes, const pie_daily_out* out, void* scratch,
                          void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  if ((rc = check_view_common(v))) return rc;
  if (!stats_i32 || !stats_f64 || !out || !scratch) return fail(PIE_ERR_INVALID_ARG, "NULL argument");
  if (stats_stride < v->n_shows || out->stride < v->n_shows) return fail(PIE_ERR_INVALID_ARG, "stride < n_shows");
  if (tz_offset_minutes < -24 * 60 || tz_offset_minutes > 24 * 60)
    return fail(PIE_ERR_INVALID_ARG, "tz_offset_minutes out of range");
  if (!out->show_day_start || !out->show_order || !out->group_day_start || !out->group_offsets || !out->summary_f64 ||
      !out->summary_count || !out->n_groups || !out->status)
    return fail(PIE_ERR_INVALID_ARG, "pie_daily_out has a NULL array");
  if (v->n_shows > 0 && !v->created_at) return fail(PIE_ERR_INVALID_ARG, "created_at is NULL");
  PIE_CUDA(pie::launch_daily_summary(*v, stats_i32, stats_f64, stats_stride, tz_offset_minutes, *out, scratch,
                                     g_sm_count, (cudaStream_t)stream));
  return PIE_OK;
}

// Device-side pie_daily_out carved from an arena, and its way back to the caller's host arrays.
static void* alloc_daily_out(Arena& a, int64_t S, int64_t Sc, pie_daily_out* dout) {
  memset(dout, 0, sizeof(*dout));
  dout->stride = Sc;
  dout->show_day_start = (int64_t*)a.take(8ull * Sc);
  dout->group_day_start = (int64_t*)a.take(8ull * Sc);
  dout->show_order = (int32_t*)a.take(4ull * Sc);
  dout->group_offsets = (int32_t*)a.take(4ull * (Sc + 1));
  dout->summary_f64 = (double*)a.take(8ull * PIE_DF_COUNT * PIE_N_METRICS * Sc);
  dout->summary_count = (int32_t*)a.take(4ull * PIE_N_METRICS * Sc);
  dout->n_groups = (int64_t*)a.take(16);
  dout->status = (int32_t*)a.take(16);
  return a.take(pie::daily_scratch_bytes(S));
}
static uint64_t daily_out_bytes(int64_t S, int64_t Sc) {
  return pad(pie::daily_scratch_bytes(S)) + pad(8ull * Sc) * 2 + pad(4ull * Sc) + pad(4ull * (Sc + 1)) +
         pad(8ull * PIE_DF_COUNT * PIE_N_METRICS * Sc) + pad(4ull * PIE_N_METRICS * Sc) + pad(64);
}
// waits for the kernels on `st`, raises what a show raised, then copies the groups out
static int download_daily(const pie_daily_out* hout, const pie_daily_out& dout, int64_t S, int64_t Sc, cudaStream_t st,
                          uint64_t* d2h) {
  PIE_CUDA(cudaMemcpyAsync(hout->n_groups, dout.n_groups, 8, cudaMemcpyDeviceToHost, st));
  PIE_CUDA(cudaMemcpyAsync(hout->status, dout.status, 8, cudaMemcpyDeviceToHost, st));
  PIE_CUDA(cudaStreamSynchronize(st));
  *d2h += 16;
  if (hout->status[0] != 0) {
    const int code = hout->status[0];
    return fail(code, code == PIE_ERR_RANGE ? "RangeError: Invalid time value (show %d)"
                                            : "show %d: date/time is not an ECMA-262 date-time string",
                hout->status[1]);
  }
  const int64_t G = *hout->n_groups;
  if (S > 0) {
    PIE_CUDA(cudaMemcpyAsync(hout->show_day_start, dout.show_day_start, 8 * (uint64_t)S, cudaMemcpyDeviceToHost, st));
    PIE_CUDA(cudaMemcpyAsync(hout->show_order, dout.show_order, 4 * (uint64_t)S, cudaMemcpyDeviceToHost, st));
    *d2h += 12 * (uint64_t)S;
  }
  PIE_CUDA(cudaMemcpyAsync(hout->group_offsets, dout.group_offsets, 4 * (uint64_t)(G + 1), cudaMemcpyDeviceToHost, st));
  *d2h += 4 * (uint64_t)(G + 1);
  if (G > 0) {
    PIE_CUDA(cudaMemcpyAsync(hout->group_day_start, dout.group_day_start, 8 * (uint64_t)G, cudaMemcpyDeviceToHost, st));
    PIE_CUDA(cudaMemcpy2DAsync(hout->summary_f64, 8 * (uint64_t)hout->stride, dout.summary_f64, 8 * (uint64_t)Sc,
                               8 * (uint64_t)G, PIE_DF_COUNT * PIE_N_METRICS, cudaMemcpyDeviceToHost, st));
    PIE_CUDA(cudaMemcpy2DAsync(hout->summary_count, 4 * (uint64_t)hout->stride, dout.summary_count, 4 * (uint64_t)Sc,
                               4 * (uint64_t)G, PIE_N_METRICS, cudaMemcpyDeviceToHost, st));
    *d2h += (8ull + 8ull * PIE_DF_COUNT * PIE_N_METRICS + 4ull * PIE_N_METRICS) * (uint64_t)G;
  }
  return PIE_OK;
}
static int check_daily_out(const pie_daily_out* hout, int64_t S) {
  if (hout->stride < S) return fail(PIE_ERR_INVALID_ARG, "pie_daily_out.stride < n_shows");
  if (!hout->show_day_start || !hout->show_order || !hout->group_day_start || !hout->group_offsets ||
      !hout->summary_f64 || !hout->summary_count || !hout->n_groups || !hout->status)
    return fail(PIE_ERR_INVALID_ARG, "pie_daily_out has a NULL array");
  return PIE_OK;
}

static int analytics_host_locked(const pie_archive_view* hv, int32_t tz_offset_minutes, int32_t* stats_i32,
                                 double* stats_f64, int64_t stats_stride, const pie_daily_out* hout) {
  int rc = ensure_init();
  if (rc) return rc;
  if ((rc = check_view_common(hv))) return rc;
  const int64_t S = hv->n_shows, E = hv->n_entries;
  const bool want_daily = hout != nullptr;
  const bool want_stats = stats_i32 != nullptr || stats_f64 != nullptr;
  if (want_stats && (!stats_i32 || !stats_f64)) return fail(PIE_ERR_INVALID_ARG, "stats_i32 and stats_f64 go together");
  if (want_stats && stats_stride < S) return fail(PIE_ERR_INVALID_ARG, "stats_stride < n_shows");
  if (S > 0 && hv->entry_offsets[S] - hv->entry_offsets[0] != E)
    return fail(PIE_ERR_INVALID_ARG, "entry_offsets span %d rows but n_entries is %lld",
                hv->entry_offsets[S] - hv->entry_offsets[0], (long long)E);
  if (S > 0 && hv->entry_offsets[0] != 0) return fail(PIE_ERR_INVALID_ARG, "entry_offsets[0] must be 0");
  if (E > 0 && (!hv->delay_sec || !hv->delay_valid)) return fail(PIE_ERR_INVALID_ARG, "delay_sec/delay_valid is NULL");
  if (want_daily) {
    if ((rc = check_daily_out(hout, S))) return rc;
    if (S > 0 && !hv->created_at) return fail(PIE_ERR_INVALID_ARG, "created_at is NULL");
  }

  // ---- plan device memory
  uint64_t bytes = 0;
  StrColPlan p_status, p_launched, p_issue, p_date, p_time;
  if ((rc = plan_strcol(p_status, &hv->status, nullptr, E, &bytes, "status"))) return rc;
  if ((rc = plan_strcol(p_launched, &hv->launched, nullptr, E, &bytes, "launched"))) return rc;
  if ((rc = plan_strcol(p_issue, &hv->primary_issue, nullptr, E, &bytes, "primary_issue"))) return rc;
  const bool has_date = want_daily && hv->show_date.offsets;
  const bool has_time = has_date && hv->show_time.offsets;
  if (has_date && (rc = plan_strcol(p_date, &hv->show_date, nullptr, S, &bytes, "show_date"))) return rc;
  if (has_time && (rc = plan_strcol(p_time, &hv->show_time, nullptr, S, &bytes, "show_time"))) return rc;
  if ((rc = check_entry_offsets(hv->entry_offsets, S))) return rc;
  const int64_t Sc = S > 0 ? S : 1;
  bytes += pad(4 * (uint64_t)(S + 1)) + pad(8 * (uint64_t)E) + pad((uint64_t)E);          // offsets, delay, valid
  bytes += pad(4ull * PIE_SI_COUNT * Sc) + pad(8ull * PIE_SF_COUNT * Sc);                  // stats planes
  if (want_daily) {
    bytes += 2 * pad(8 * (uint64_t)Sc) + pad(8 * (uint64_t)E);                             // created, archived, entry_ts
    bytes += daily_out_bytes(S, Sc);
  }
  if ((rc = g_arena.reserve(bytes))) return rc;
  cudaStream_t st = g_arena.stream;
  g_cur = &g_arena;
  g_cur_stream = st;
  StreamDrain drain_on_exit{&g_arena.stream};

  // ---- H2D
  uint64_t h2d = 0, d2h = 0;
  pie_archive_view dv;
  memset(&dv, 0, sizeof(dv));
  dv.n_shows = S;
  dv.n_entries = E;
  g_checks.clear();
  if ((rc = upload_array(hv->entry_offsets, S + 1, &dv.entry_offsets, &h2d))) return rc;
  p_status.dst = &dv.status; p_launched.dst = &dv.launched; p_issue.dst = &dv.primary_issue;
  if ((rc = upload_strcol(p_status, &h2d))) return rc;
  if ((rc = upload_strcol(p_launched, &h2d))) return rc;
  if ((rc = upload_strcol(p_issue, &h2d))) return rc;
  if ((rc = upload_array(hv->delay_sec, E, &dv.delay_sec, &h2d))) return rc;
  if ((rc = upload_array(hv->delay_valid, E, &dv.delay_valid, &h2d))) return rc;
  if (want_daily) {
    if (S > 0 && (rc = upload_array(hv->created_at, S, &dv.created_at, &h2d))) return rc;
    if (hv->archived_at && (rc = upload_array(hv->archived_at, S, &dv.archived_at, &h2d))) return rc;
    if (hv->entry_ts && (rc = upload_array(hv->entry_ts, E, &dv.entry_ts, &h2d))) return rc;
    if (has_date) { p_date.dst = &dv.show_date; if ((rc = upload_strcol(p_date, &h2d))) return rc; }
    if (has_time) { p_time.dst = &dv.show_time; if ((rc = upload_strcol(p_time, &h2d))) return rc; }
  }

  // ---- the uploaded offsets, checked on the device before any kernel walks them
  int32_t check_flags[pie::kMaxOffsetsArrays];
  if ((rc = enqueue_checks(g_checks, 0, check_flags, st))) return rc;
  PIE_CUDA(cudaStreamSynchronize(st));
  if ((rc = report_checks(g_checks, check_flags))) return rc;

  // ---- kernels
  int32_t* d_si = (int32_t*)g_arena.take(4ull * PIE_SI_COUNT * Sc);
  double* d_sf = (double*)g_arena.take(8ull * PIE_SF_COUNT * Sc);
  PIE_CUDA(pie::launch_show_stats(dv, d_si, d_sf, Sc, g_sm_count, st));

  pie_daily_out dout;
  memset(&dout, 0, sizeof(dout));
  if (want_daily) {
    void* dscratch = alloc_daily_out(g_arena, S, Sc, &dout);
    PIE_CUDA(pie::launch_daily_summary(dv, d_si, d_sf, Sc, tz_offset_minutes, dout, dscratch, g_sm_count, st));
  }

  // ---- D2H
  if (want_stats && S > 0) {
    PIE_CUDA(cudaMemcpy2DAsync(stats_i32, 4 * (uint64_t)stats_stride, d_si, 4 * (uint64_t)Sc, 4 * (uint64_t)S,
                               PIE_SI_COUNT, cudaMemcpyDeviceToHost, st));
    PIE_CUDA(cudaMemcpy2DAsync(stats_f64, 8 * (uint64_t)stats_stride, d_sf, 8 * (uint64_t)Sc, 8 * (uint64_t)S,
                               PIE_SF_COUNT, cudaMemcpyDeviceToHost, st));
    d2h += (4ull * PIE_SI_COUNT + 8ull * PIE_SF_COUNT) * (uint64_t)S;
  }
  if (want_daily) {
    g_last_h2d = h2d;
    if ((rc = download_daily(hout, dout, S, Sc, st, &d2h))) {
      g_last_d2h = d2h;
      return rc;
    }
  }
  PIE_CUDA(cudaStreamSynchronize(st));
  g_last_h2d = h2d;
  g_last_d2h = d2h;
  return PIE_OK;
}

int pie_archive_analytics_host(const pie_archive_view* hv, int32_t tz_offset_minutes, int32_t* stats_i32,
                               double* stats_f64, int64_t stats_stride, const pie_daily_out* hout) {
  std::lock_guard<std::mutex> lock(g_host_mutex);
  if (tz_offset_minutes < -24 * 60 || tz_offset_minutes > 24 * 60)
    return fail(PIE_ERR_INVALID_ARG, "tz_offset_minutes out of range");
  if (!hout) return fail(PIE_ERR_INVALID_ARG, "pie_daily_out is NULL (use pie_show_stats_host for stats only)");
  return analytics_host_locked(hv, tz_offset_minutes, stats_i32, stats_f64, stats_stride, hout);
}

int pie_show_stats_host(const pie_archive_view* hv, int32_t* stats_i32, double* stats_f64, int64_t stride) {
  std::lock_guard<std::mutex> lock(g_host_mutex);
  if (!stats_i32 || !stats_f64) return fail(PIE_ERR_INVALID_ARG, "output is NULL");
  return analytics_host_locked(hv, 0, stats_i32, stats_f64, stride, nullptr);
}

int pie_compute_metrics_dev(const pie_archive_view* v, int32_t* metrics_i32, uint8_t* avg_delay_text, int64_t stride,
                            void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  if ((rc = check_view_common(v))) return rc;
  if (!metrics_i32 || !avg_delay_text) return fail(PIE_ERR_INVALID_ARG, "output is NULL");
  if (stride < v->n_shows) return fail(PIE_ERR_INVALID_ARG, "stride < n_shows");
  if (!v->planned.offsets || !v->status.offsets || !v->primary_issue.offsets)
    return fail(PIE_ERR_INVALID_ARG, "computeMetrics reads planned, status and primary_issue: one is NULL");
  if (v->n_entries > 0 && (!v->delay_sec || !v->delay_valid)) return fail(PIE_ERR_INVALID_ARG, "delay_sec/delay_valid is NULL");
  if (reinterpret_cast<uintptr_t>(avg_delay_text) & 15) return fail(PIE_ERR_INVALID_ARG, "avg_delay_text must be 16-byte aligned");
  PIE_CUDA(pie::launch_compute_metrics(*v, metrics_i32, avg_delay_text, stride, (cudaStream_t)stream));
  return PIE_OK;
}

int pie_compute_metrics_host(const pie_archive_view* hv, int32_t* metrics_i32, uint8_t* avg_delay_text, int64_t stride) {
  std::lock_guard<std::mutex> lock(g_host_mutex);
  int rc = ensure_init();
  if (rc) return rc;
  if ((rc = check_view_common(hv))) return rc;
  const int64_t S = hv->n_shows, E = hv->n_entries;
  if (!metrics_i32 || !avg_delay_text) return fail(PIE_ERR_INVALID_ARG, "output is NULL");
  if (stride < S) return fail(PIE_ERR_INVALID_ARG, "stride < n_shows");
  if (S > 0 && (hv->entry_offsets[0] != 0 || hv->entry_offsets[S] != E))
    return fail(PIE_ERR_INVALID_ARG, "entry_offsets must run from 0 to n_entries");
  if (E > 0 && (!hv->delay_sec || !hv->delay_valid)) return fail(PIE_ERR_INVALID_ARG, "delay_sec/delay_valid is NULL");
  uint64_t bytes = 0;
  StrColPlan p_planned, p_status, p_issue;
  if ((rc = plan_strcol(p_planned, &hv->planned, nullptr, E, &bytes, "planned"))) return rc;
  if ((rc = plan_strcol(p_status, &hv->status, nullptr, E, &bytes, "status"))) return rc;
  if ((rc = plan_strcol(p_issue, &hv->primary_issue, nullptr, E, &bytes, "primary_issue"))) return rc;
  if ((rc = check_entry_offsets(hv->entry_offsets, S))) return rc;
  const int64_t Sc = S > 0 ? S : 1;
  bytes += pad(4 * (uint64_t)(S + 1)) + pad(8 * (uint64_t)E) + pad((uint64_t)E);
  bytes += pad(4ull * PIE_CM_COUNT * Sc) + pad((uint64_t)PIE_CM_TEXT * Sc);
  if ((rc = g_arena.reserve(bytes))) return rc;
  cudaStream_t st = g_arena.stream;
  g_cur = &g_arena;
  g_cur_stream = st;
  StreamDrain drain_on_exit{&g_arena.stream};
  uint64_t h2d = 0;
  pie_archive_view dv;
  memset(&dv, 0, sizeof(dv));
  dv.n_shows = S;
  dv.n_entries = E;
  g_checks.clear();
  if ((rc = upload_array(hv->entry_offsets, S + 1, &dv.entry_offsets, &h2d))) return rc;
  p_planned.dst = &dv.planned; p_status.dst = &dv.status; p_issue.dst = &dv.primary_issue;
  if ((rc = upload_strcol(p_planned, &h2d))) return rc;
  if ((rc = upload_strcol(p_status, &h2d))) return rc;
  if ((rc = upload_strcol(p_issue, &h2d))) return rc;
  if ((rc = upload_array(hv->delay_sec, E, &dv.delay_sec, &h2d))) return rc;
  if ((rc = upload_array(hv->delay_valid, E, &dv.delay_valid, &h2d))) return rc;
  int32_t check_flags[pie::kMaxOffsetsArrays];
  if ((rc = enqueue_checks(g_checks, 0, check_flags, st))) return rc;
  PIE_CUDA(cudaStreamSynchronize(st));
  if ((rc = report_checks(g_checks, check_flags))) return rc;
  int32_t* d_out = (int32_t*)g_arena.take(4ull * PIE_CM_COUNT * Sc);
  uint8_t* d_text = (uint8_t*)g_arena.take((uint64_t)PIE_CM_TEXT * Sc);
  PIE_CUDA(pie::launch_compute_metrics(dv, d_out, d_text, Sc, st));
  uint64_t d2h = 0;
  if (S > 0) {
    PIE_CUDA(cudaMemcpy2DAsync(metrics_i32, 4 * (uint64_t)stride, d_out, 4 * (uint64_t)Sc, 4 * (uint64_t)S, PIE_CM_COUNT,
                               cudaMemcpyDeviceToHost, st));
    PIE_CUDA(cudaMemcpyAsync(avg_delay_text, d_text, (uint64_t)PIE_CM_TEXT * (uint64_t)S, cudaMemcpyDeviceToHost, st));
    d2h = (4ull * PIE_CM_COUNT + PIE_CM_TEXT) * (uint64_t)S;
  }
  PIE_CUDA(cudaStreamSynchronize(st));
  g_last_h2d = h2d;
  g_last_d2h = d2h;
  return PIE_OK;
}

uint64_t pie_csv_rows_scratch_bytes(int64_t n_entries) { return pie::csv_scratch_bytes(n_entries); }

static int export_rows_dev(RowFormat format, const pie_archive_view* v, int64_t* row_offsets, uint8_t* out_data,
                           uint64_t out_capacity, uint64_t* total_bytes_dev, void* scratch, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  if ((rc = check_view_common(v))) return rc;
  if ((rc = check_export_view_dev(v, format))) return rc;
  if (!row_offsets || !total_bytes_dev || !scratch) return fail(PIE_ERR_INVALID_ARG, "NULL argument");
  PIE_CUDA(launch_rows(format, *v, row_offsets, out_data, out_data ? out_capacity : 0, 0ull,
                       (unsigned long long*)total_bytes_dev, scratch, (cudaStream_t)stream));
  return PIE_OK;
}

int pie_csv_rows_dev(const pie_archive_view* v, int64_t* row_offsets, uint8_t* out_data, uint64_t out_capacity,
                     uint64_t* total_bytes_dev, void* scratch, void* stream) {
  return export_rows_dev(kFormatCsv, v, row_offsets, out_data, out_capacity, total_bytes_dev, scratch, stream);
}

int pie_archive_payloads_dev(const pie_archive_view* v, int64_t* row_offsets, uint8_t* out_data, uint64_t out_capacity,
                             uint64_t* total_bytes_dev, void* scratch, void* stream) {
  return export_rows_dev(kFormatPayload, v, row_offsets, out_data, out_capacity, total_bytes_dev, scratch, stream);
}

int pie_debug_csv_force_slow_path(int on) { return pie::csv_set_force_slow(on); }

int pie_debug_ingest_warp_path(int on) { return pie::ingest_set_warp_path(on); }

int pie_debug_ingest_declined(const void* scratch, int64_t n_docs, uint32_t* declined, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  if (!scratch || !declined || n_docs < 0) return fail(PIE_ERR_INVALID_ARG, "NULL argument");
  unsigned int n = 0;
  PIE_CUDA(pie::ingest_read_declined(scratch, n_docs, &n, (cudaStream_t)stream));
  *declined = n;
  return PIE_OK;
}

int pie_debug_csv_slow_tiles(const void* scratch, int64_t n_entries, uint32_t* slow_tiles, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  if (!scratch || !slow_tiles || n_entries < 0) return fail(PIE_ERR_INVALID_ARG, "NULL argument");
  unsigned int n = 0;
  PIE_CUDA(pie::csv_read_slow_tiles(scratch, n_entries, &n, (cudaStream_t)stream));
  *slow_tiles = n;
  return PIE_OK;
}

// Rows per pipeline chunk of the host export path (H2D of chunk c+1 and D2H of chunk c-1 overlap the
// kernels of chunk c; PCIe is full duplex).
static int64_t kCsvChunkRows = 1 << 20;  // pie_set_csv_chunk_rows (tests exercise the multi-chunk path)

struct CsvPipeline {
  cudaStream_t h2d = nullptr, cmp = nullptr, d2h = nullptr;
  cudaEvent_t h2d_done[2] = {nullptr, nullptr}, kernel_done[2] = {nullptr, nullptr}, d2h_done[2] = {nullptr, nullptr};
  cudaEvent_t daily_up = nullptr;  // the daily summary's per-batch arrays are on the device
  OutBuffer out[2];
  OutBuffer off[2];  // a chunk's row offsets + total: NOT in the input arena, which is refilled while they download
  unsigned long long* h_total = nullptr;  // pinned
  int init() {
    if (h2d) return PIE_OK;
    PIE_CUDA(cudaStreamCreateWithFlags(&h2d, cudaStreamNonBlocking));
    PIE_CUDA(cudaStreamCreateWithFlags(&cmp, cudaStreamNonBlocking));
    PIE_CUDA(cudaStreamCreateWithFlags(&d2h, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      PIE_CUDA(cudaEventCreateWithFlags(&h2d_done[i], cudaEventDisableTiming));
      PIE_CUDA(cudaEventCreateWithFlags(&kernel_done[i], cudaEventDisableTiming));
      PIE_CUDA(cudaEventCreateWithFlags(&d2h_done[i], cudaEventDisableTiming));
    }
    PIE_CUDA(cudaEventCreateWithFlags(&daily_up, cudaEventDisableTiming));
    PIE_CUDA(cudaHostAlloc(&h_total, 512, cudaHostAllocDefault));  // [0] a chunk's total; +64: its check flags; +256: the step's
    return PIE_OK;
  }
};
static CsvPipeline g_pipe;

// Developer aid (PIE_DEBUG_TIMELINE=1): when each chunk's upload, kernels and download finished, in ms since the call
// began, printed to stderr — the only timeline tool there is without nsys.
struct Timeline {
  bool on = getenv("PIE_DEBUG_TIMELINE") != nullptr;
  cudaEvent_t start = nullptr;
  std::vector<cudaEvent_t> ev;
  std::vector<const char*> what;
  std::vector<int> chunk;
  void begin(cudaStream_t st) {
    if (!on) return;
    cudaEventCreate(&start);
    cudaEventRecord(start, st);
  }
  void mark(cudaStream_t st, const char* w, int k) {
    if (!on) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    ev.push_back(e); what.push_back(w); chunk.push_back(k);
  }
  void report() {
    if (!on) return;
    cudaDeviceSynchronize();
    for (size_t i = 0; i < ev.size(); ++i) {
      float ms = 0;
      cudaEventElapsedTime(&ms, start, ev[i]);
      fprintf(stderr, "[timeline] chunk %2d %-10s %8.2f ms\n", chunk[i], what[i], ms);
      cudaEventDestroy(ev[i]);
    }
    cudaEventDestroy(start);
  }
};

struct CsvChunk {
  int64_t s0, s1, e0, e1;
  std::vector<int32_t> entry_offsets;  // rebased to the chunk
  pie_archive_view host;                // sliced host view
  pie_archive_view dev;
  void* scratch;
  int64_t* d_offsets;
  unsigned long long* d_total;
  PendingChecks checks;  // the chunk's offsets arrays on the device, to be validated before its kernels run
};

static void slice_col(pie_strcol* c, int64_t first) { if (c->offsets) c->offsets += first; }

static void make_chunk_view(const pie_archive_view* hv, CsvChunk* c) {
  c->entry_offsets.resize((size_t)(c->s1 - c->s0 + 1));
  for (int64_t s = c->s0; s <= c->s1; ++s) c->entry_offsets[(size_t)(s - c->s0)] = hv->entry_offsets[s] - (int32_t)c->e0;
  pie_archive_view v = *hv;
  v.n_shows = c->s1 - c->s0;
  v.n_entries = c->e1 - c->e0;
  v.entry_offsets = c->entry_offsets.data();
  pie_strcol* show_cols[] = {&v.show_id, &v.show_date, &v.show_time, &v.show_label, &v.lead_pilot, &v.monkey_lead, &v.show_notes};
  for (pie_strcol* x : show_cols) slice_col(x, c->s0);
  pie_strcol* entry_cols[] = {&v.entry_id, &v.unit_id, &v.planned, &v.launched, &v.status, &v.primary_issue, &v.sub_issue,
                              &v.other_detail, &v.severity, &v.root_cause, &v.operator_name, &v.battery_id, &v.command_rx,
                              &v.notes};
  for (pie_strcol* x : entry_cols) slice_col(x, c->e0);
  if (v.crew.list_offsets) v.crew.list_offsets += c->s0;
  if (v.actions.list_offsets) v.actions.list_offsets += c->e0;
  if (v.delay_sec) v.delay_sec += c->e0;
  if (v.delay_valid) v.delay_valid += c->e0;
  c->host = v;
}

// stage chunk c into input arena `slot` on the H2D stream
static int upload_chunk(CsvChunk* c, int slot, uint64_t* h2d, RowFormat format) {
  g_cur = &g_pipe_in[slot];
  g_cur_stream = g_pipe.h2d;
  if (!g_cur->stream) g_cur->stream = g_pipe.h2d;  // reserve() only creates a stream when there is none
  const int64_t E = c->e1 - c->e0;
  memset(&c->dev, 0, sizeof(c->dev));
  const uint64_t extra = pad(pie::csv_scratch_bytes(E)) + pad(8 * (uint64_t)(E + 1)) + pad(64);
  g_checks.clear();
  int rc = upload_export_view(&c->host, &c->dev, extra + (extra >> 2), h2d, format);  // +25 %: later chunks rarely regrow
  if (rc) return rc;
  c->checks = g_checks;
  c->scratch = g_cur->take(pie::csv_scratch_bytes(E));
  c->d_offsets = nullptr;  // in the slot's offsets buffer (export_rows_host)
  c->d_total = nullptr;
  return PIE_OK;
}

int64_t pie_set_csv_chunk_rows(int64_t rows) {
  std::lock_guard<std::mutex> lock(g_host_mutex);
  const int64_t old = kCsvChunkRows;
  if (rows > 0) kCsvChunkRows = rows;
  return old;
}

// The analytics that ride on the export pipeline (pie_archive_step_host): show statistics per chunk on the columns
// the rows need anyway, the daily summary once at the end.
struct StepAnalytics {
  int32_t tz_offset_minutes;
  int32_t* stats_i32;
  double* stats_f64;
  int64_t stats_stride;
  const pie_daily_out* hout;
};

static int export_rows_host(RowFormat format, const pie_archive_view* hv, int64_t* row_offsets, uint8_t* out_data,
                            uint64_t out_capacity, uint64_t* total_bytes, const StepAnalytics* an = nullptr) {
  std::lock_guard<std::mutex> lock(g_host_mutex);
  int rc = ensure_init();
  if (rc) return rc;
  if ((rc = check_view_common(hv))) return rc;
  if (!row_offsets || !total_bytes) return fail(PIE_ERR_INVALID_ARG, "row_offsets / total_bytes is NULL");
  const int64_t S = hv->n_shows, E = hv->n_entries;
  if (S > 0 && (hv->entry_offsets[0] != 0 || hv->entry_offsets[S] != E))
    return fail(PIE_ERR_INVALID_ARG, "entry_offsets must run from 0 to n_entries");
  if ((rc = check_entry_offsets(hv->entry_offsets, S))) return rc;  // the chunking below walks it
  if ((rc = g_pipe.init())) return rc;
  // Whatever way this function is left, nothing it enqueued may still be writing the caller's buffers (or reading
  // an arena the next call resets): drain the three streams and hand the upload helpers back to the default arena.
  struct PipeDrain {
    ~PipeDrain() {
      cudaStreamSynchronize(g_pipe.h2d);
      cudaStreamSynchronize(g_pipe.cmp);
      cudaStreamSynchronize(g_pipe.d2h);
      g_cur = &g_arena;
      g_cur_stream = g_arena.stream;
    }
  } drain_on_exit;

  // ---- analytics riding along: the per-batch arrays of the daily summary go up first, on the upload stream
  const int64_t Sc = S > 0 ? S : 1;
  uint64_t an_h2d = 0;
  int32_t* d_si = nullptr;
  double* d_sf = nullptr;
  pie_archive_view dv_all;
  pie_daily_out dout;
  void* dscratch = nullptr;
  PendingChecks an_checks;
  an_checks.clear();
  std::function<int()> upload_daily_inputs;
  memset(&dv_all, 0, sizeof(dv_all));
  if (an) {
    if (format != kFormatCsv) return fail(PIE_ERR_INVALID_ARG, "the step rides on the CSV rows");
    if (an->tz_offset_minutes < -24 * 60 || an->tz_offset_minutes > 24 * 60)
      return fail(PIE_ERR_INVALID_ARG, "tz_offset_minutes out of range");
    if (!an->hout) return fail(PIE_ERR_INVALID_ARG, "pie_daily_out is NULL");
    if ((an->stats_i32 != nullptr) != (an->stats_f64 != nullptr))
      return fail(PIE_ERR_INVALID_ARG, "stats_i32 and stats_f64 go together");
    if (an->stats_i32 && an->stats_stride < S) return fail(PIE_ERR_INVALID_ARG, "stats_stride < n_shows");
    if ((rc = check_daily_out(an->hout, S))) return rc;
    if (S > 0 && !hv->created_at) return fail(PIE_ERR_INVALID_ARG, "created_at is NULL");
    uint64_t bytes = 0;
    StrColPlan p_date, p_time;
    const bool has_date = hv->show_date.offsets != nullptr, has_time = has_date && hv->show_time.offsets != nullptr;
    if (has_date && (rc = plan_strcol(p_date, &hv->show_date, nullptr, S, &bytes, "show_date"))) return rc;
    if (has_time && (rc = plan_strcol(p_time, &hv->show_time, nullptr, S, &bytes, "show_time"))) return rc;
    bytes += pad(4 * (uint64_t)(S + 1)) + 2 * pad(8 * (uint64_t)Sc) + pad(8 * (uint64_t)E);
    bytes += pad(4ull * PIE_SI_COUNT * Sc) + pad(8ull * PIE_SF_COUNT * Sc) + daily_out_bytes(S, Sc);
    if ((rc = g_arena.reserve(bytes))) return rc;
    dv_all.n_shows = S;
    dv_all.n_entries = E;
    d_si = (int32_t*)g_arena.take(4ull * PIE_SI_COUNT * Sc);
    d_sf = (double*)g_arena.take(8ull * PIE_SF_COUNT * Sc);
    dscratch = alloc_daily_out(g_arena, S, Sc, &dout);
    // The per-batch arrays the daily summary reads (130 MB per 2^20 shows) are needed only at the very end: they go up
    // BEHIND the last chunk, while the rows are still being computed and downloaded, not in front of the first.
    upload_daily_inputs = [&, p_date, p_time, has_date, has_time]() mutable -> int {
      int rc2;
      g_cur = &g_arena;
      g_cur_stream = g_pipe.h2d;
      g_checks.clear();
      if ((rc2 = upload_array(hv->entry_offsets, S + 1, &dv_all.entry_offsets, &an_h2d))) return rc2;
      if (S > 0 && (rc2 = upload_array(hv->created_at, S, &dv_all.created_at, &an_h2d))) return rc2;
      if (hv->archived_at && (rc2 = upload_array(hv->archived_at, S, &dv_all.archived_at, &an_h2d))) return rc2;
      if (hv->entry_ts && (rc2 = upload_array(hv->entry_ts, E, &dv_all.entry_ts, &an_h2d))) return rc2;
      if (has_date) { p_date.dst = &dv_all.show_date; if ((rc2 = upload_strcol(p_date, &an_h2d))) return rc2; }
      if (has_time) { p_time.dst = &dv_all.show_time; if ((rc2 = upload_strcol(p_time, &an_h2d))) return rc2; }
      an_checks = g_checks;
      PIE_CUDA(cudaEventRecord(g_pipe.daily_up, g_pipe.h2d));
      return PIE_OK;
    };
  }

  // chunks of ~kCsvChunkRows rows, cut at show boundaries.  The pipeline's ends are not overlapped — nothing
  // downloads while the first chunk goes up, nothing uploads while the last comes down — so a large batch starts and
  // ends on smaller chunks (a quarter, a half of the usual size).
  std::vector<CsvChunk> chunks;
  const bool ramp = E >= 4 * kCsvChunkRows;
  for (int64_t s0 = 0; s0 < S || chunks.empty();) {
    CsvChunk c;
    c.s0 = s0;
    c.e0 = S > 0 ? hv->entry_offsets[s0] : 0;
    int64_t want = kCsvChunkRows;
    if (ramp) {
      const int64_t left = E - c.e0;
      if (chunks.empty()) want = kCsvChunkRows / 4;
      else if (chunks.size() == 1) want = kCsvChunkRows / 2;
      else if (left <= kCsvChunkRows / 4 + kCsvChunkRows / 8) want = left;           // the last one
      else if (left <= kCsvChunkRows) want = left - kCsvChunkRows / 4;               // half, then a quarter
      else if (left < 2 * kCsvChunkRows) want = left - kCsvChunkRows * 3 / 4;
    }
    if (want < 1) want = 1;  // (trailing shows without entries: the loop below must still take them)
    int64_t s1 = s0;
    while (s1 < S && hv->entry_offsets[s1] - c.e0 < want) {
      const int64_t step = (S - s1 > 4096 && hv->entry_offsets[s1 + 4096] - c.e0 < want) ? 4096 : 1;
      s1 += step;
    }
    c.s1 = s1;
    c.e1 = S > 0 ? hv->entry_offsets[s1] : 0;
    chunks.push_back(std::move(c));
    s0 = s1;
    if (S == 0) break;
  }
  const int K = (int)chunks.size();
  for (CsvChunk& c : chunks) make_chunk_view(hv, &c);

  uint64_t h2d = 0, d2h = 0;
  unsigned long long bias = 0;
  bool overflow = false;
  Timeline tl;
  tl.begin(g_pipe.h2d);
  if ((rc = upload_chunk(&chunks[0], 0, &h2d, format))) return rc;
  PIE_CUDA(cudaEventRecord(g_pipe.h2d_done[0], g_pipe.h2d));
  tl.mark(g_pipe.h2d, "uploaded", 0);
  for (int k = 0; k < K; ++k) {
    const int slot = k & 1;
    CsvChunk& c = chunks[(size_t)k];
    const int64_t Ec = c.e1 - c.e0;
    PIE_CUDA(cudaStreamWaitEvent(g_pipe.cmp, g_pipe.h2d_done[slot], 0));
    int32_t* chunk_flags = reinterpret_cast<int32_t*>(g_pipe.h_total + 8);
    if ((rc = enqueue_checks(c.checks, 1 + slot, chunk_flags, g_pipe.cmp, true))) return rc;
    if (an && c.s1 > c.s0)  // the chunk's shows: status, launched, primaryIssue and delaySec are resident for the rows
      PIE_CUDA(pie::launch_show_stats(c.dev, d_si + c.s0, d_sf + c.s0, Sc, g_sm_count, g_pipe.cmp));
    // the chunk's row offsets live in the slot's own buffer: chunk k-2's must have left for the host
    if (k >= 2) PIE_CUDA(cudaStreamWaitEvent(g_pipe.cmp, g_pipe.d2h_done[slot], 0));
    if (g_pipe.off[slot].cap < 8 * (uint64_t)(Ec + 1) + 64) {
      if (k >= 2) PIE_CUDA(cudaEventSynchronize(g_pipe.d2h_done[slot]));
      if ((rc = g_pipe.off[slot].ensure(8 * (uint64_t)(Ec + 1) + (uint64_t)(Ec + 1) + 64))) return rc;
    }
    c.d_offsets = (int64_t*)g_pipe.off[slot].base;
    c.d_total = (unsigned long long*)(g_pipe.off[slot].base + ((8 * (uint64_t)(Ec + 1) + 15) & ~(uint64_t)15));
    // pass 1: sizes only (row offsets + total), so the output can be placed and sized exactly.  The total is written
    // straight into mapped pinned memory: an 8-byte device-to-host copy would wait in the copy engine's queue behind
    // the previous chunk's 300 MB download — and with it this chunk's second pass, and the next download.
    PIE_CUDA(launch_rows(format, c.dev, c.d_offsets, nullptr, 0, bias, g_pipe.h_total, c.scratch, g_pipe.cmp));
    if (k + 1 < K) {  // prefetch the next chunk into the other input arena once chunk k-1's kernels have left it
      // (after this chunk's kernels are enqueued: the host's check of the next chunk's offsets overlaps them)
      if (k >= 1) PIE_CUDA(cudaStreamWaitEvent(g_pipe.h2d, g_pipe.kernel_done[slot ^ 1], 0));
      if ((rc = upload_chunk(&chunks[(size_t)k + 1], slot ^ 1, &h2d, format))) return rc;
      PIE_CUDA(cudaEventRecord(g_pipe.h2d_done[slot ^ 1], g_pipe.h2d));
      tl.mark(g_pipe.h2d, "uploaded", k + 1);
    }
    if (an && K == 1 && (rc = upload_daily_inputs())) return rc;                                    // a single chunk
    if (an && K > 1 && k + 1 == K - 1 && (rc = upload_daily_inputs())) return rc;                 // behind the last chunk
    PIE_CUDA(cudaStreamSynchronize(g_pipe.cmp));
    if ((rc = report_checks(c.checks, chunk_flags))) return rc;
    const unsigned long long total = *g_pipe.h_total;
    d2h += 8;
    const bool want_data = out_data != nullptr && !overflow && bias + total <= out_capacity;
    if (out_data && !want_data) overflow = true;
    if (want_data) {
      if (k >= 2) PIE_CUDA(cudaStreamWaitEvent(g_pipe.cmp, g_pipe.d2h_done[slot], 0));  // output slot drained
      if (g_pipe.out[slot].cap < total + 256) {
        if (k >= 2) PIE_CUDA(cudaEventSynchronize(g_pipe.d2h_done[slot]));
        if ((rc = g_pipe.out[slot].ensure(total + (total >> 2) + 256))) return rc;
      }
      // pass 2: write (the chunk's inputs are resident; most of them still in L2)
      PIE_CUDA(launch_rows(format, c.dev, c.d_offsets, g_pipe.out[slot].base, total, bias, c.d_total, c.scratch,
                           g_pipe.cmp));
    }
    PIE_CUDA(cudaEventRecord(g_pipe.kernel_done[slot], g_pipe.cmp));
    tl.mark(g_pipe.cmp, "kernels", k);
    PIE_CUDA(cudaStreamWaitEvent(g_pipe.d2h, g_pipe.kernel_done[slot], 0));
    if (Ec > 0)
      PIE_CUDA(cudaMemcpyAsync(row_offsets + c.e0, c.d_offsets, 8 * (uint64_t)Ec, cudaMemcpyDeviceToHost, g_pipe.d2h));
    d2h += 8 * (uint64_t)Ec;
    if (want_data && total) {
      PIE_CUDA(cudaMemcpyAsync(out_data + bias, g_pipe.out[slot].base, total, cudaMemcpyDeviceToHost, g_pipe.d2h));
      d2h += total;
    }
    if (an && an->stats_i32 && c.s1 > c.s0) {  // the chunk's columns of the statistics planes
      const uint64_t ns = (uint64_t)(c.s1 - c.s0);
      PIE_CUDA(cudaMemcpy2DAsync(an->stats_i32 + c.s0, 4 * (uint64_t)an->stats_stride, d_si + c.s0, 4 * (uint64_t)Sc, 4 * ns,
                                 PIE_SI_COUNT, cudaMemcpyDeviceToHost, g_pipe.d2h));
      PIE_CUDA(cudaMemcpy2DAsync(an->stats_f64 + c.s0, 8 * (uint64_t)an->stats_stride, d_sf + c.s0, 8 * (uint64_t)Sc, 8 * ns,
                                 PIE_SF_COUNT, cudaMemcpyDeviceToHost, g_pipe.d2h));
      d2h += (4ull * PIE_SI_COUNT + 8ull * PIE_SF_COUNT) * ns;
    }
    PIE_CUDA(cudaEventRecord(g_pipe.d2h_done[slot], g_pipe.d2h));
    tl.mark(g_pipe.d2h, "downloaded", k);
    bias += total;
  }
  int an_rc = PIE_OK;
  if (an) {
    PIE_CUDA(cudaStreamWaitEvent(g_pipe.cmp, g_pipe.daily_up, 0));  // the per-batch arrays went up behind the last chunk
    int32_t* an_flags = reinterpret_cast<int32_t*>(g_pipe.h_total + 32);
    if ((rc = enqueue_checks(an_checks, 0, an_flags, g_pipe.cmp))) return rc;
    PIE_CUDA(cudaStreamSynchronize(g_pipe.cmp));
    if ((rc = report_checks(an_checks, an_flags))) return rc;
    PIE_CUDA(pie::launch_daily_summary(dv_all, d_si, d_sf, Sc, an->tz_offset_minutes, dout, dscratch, g_sm_count,
                                       g_pipe.cmp));
    an_rc = download_daily(an->hout, dout, S, Sc, g_pipe.cmp, &d2h);
    h2d += an_h2d;
  }
  PIE_CUDA(cudaStreamSynchronize(g_pipe.d2h));
  PIE_CUDA(cudaStreamSynchronize(g_pipe.cmp));
  tl.report();
  row_offsets[E] = (int64_t)bias;
  *total_bytes = bias;
  g_last_h2d = h2d;
  g_last_d2h = d2h + 8;
  g_cur = &g_arena;
  g_cur_stream = g_arena.stream;
  if (an_rc) return an_rc;  // RangeError / unsupported date raised by a show (pie_last_error has the text)
  if (overflow)
    return fail(PIE_ERR_CAPACITY, "the rows need %llu bytes, the caller's buffer holds %llu", bias,
                (unsigned long long)out_capacity);
  return PIE_OK;
}

int pie_archive_step_host(const pie_archive_view* hv, int32_t tz_offset_minutes, int32_t* stats_i32, double* stats_f64,
                          int64_t stats_stride, const pie_daily_out* host_out, int64_t* row_offsets, uint8_t* out_data,
                          uint64_t out_capacity, uint64_t* total_bytes) {
  const StepAnalytics an{tz_offset_minutes, stats_i32, stats_f64, stats_stride, host_out};
  return export_rows_host(kFormatCsv, hv, row_offsets, out_data, out_capacity, total_bytes, &an);
}

int pie_csv_rows_host(const pie_archive_view* hv, int64_t* row_offsets, uint8_t* out_data, uint64_t out_capacity,
                      uint64_t* total_bytes) {
  return export_rows_host(kFormatCsv, hv, row_offsets, out_data, out_capacity, total_bytes);
}

int pie_archive_payloads_host(const pie_archive_view* hv, int64_t* row_offsets, uint8_t* out_data, uint64_t out_capacity,
                              uint64_t* total_bytes) {
  return export_rows_host(kFormatPayload, hv, row_offsets, out_data, out_capacity, total_bytes);
}

/* ---- JSON ingest ---------------------------------------------------------------------------------------- */
uint64_t pie_ingest_scratch_bytes(int64_t n_docs) { return pie::ingest_scratch_bytes(n_docs > 0 ? n_docs : 0); }

static int check_docs(const pie_json_docs* d) {
  if (!d) return fail(PIE_ERR_INVALID_ARG, "docs is NULL");
  if (d->n_docs < 0 || d->n_docs > 0x7FFFFFF0LL) return fail(PIE_ERR_INVALID_ARG, "n_docs out of range (batches of < 2^31 documents)");
  if (!d->offsets) return fail(PIE_ERR_INVALID_ARG, "docs.offsets is NULL");
  return PIE_OK;
}

int pie_ingest_measure_dev(const pie_json_docs* d, void* scratch, uint8_t* doc_status, int64_t* totals_dev,
                           int32_t* status_dev, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  if ((rc = check_docs(d))) return rc;
  if (!scratch || !doc_status || !totals_dev || !status_dev) return fail(PIE_ERR_INVALID_ARG, "NULL argument");
  if (d->n_docs > 0 && !d->data) return fail(PIE_ERR_INVALID_ARG, "docs.data is NULL");
  if (reinterpret_cast<uintptr_t>(scratch) & 7) return fail(PIE_ERR_INVALID_ARG, "scratch must be 8-byte aligned");
  PIE_CUDA(pie::launch_ingest_measure(*d, scratch, doc_status, totals_dev, status_dev, (cudaStream_t)stream));
  return PIE_OK;
}

static int check_table(const pie_archive_table* t) {
  if (!t) return fail(PIE_ERR_INVALID_ARG, "table is NULL");
  const pie_strcol_mut* cols[23] = {&t->show_id, &t->show_date, &t->show_time, &t->show_label, &t->lead_pilot, &t->monkey_lead,
                                    &t->show_notes, &t->crew.items, &t->entry_id, &t->unit_id, &t->planned, &t->launched,
                                    &t->status, &t->primary_issue, &t->sub_issue, &t->other_detail, &t->severity, &t->root_cause,
                                    &t->operator_name, &t->battery_id, &t->command_rx, &t->notes, &t->actions.items};
  for (int i = 0; i < 23; ++i)
    if (!cols[i]->offsets || !cols[i]->data) return fail(PIE_ERR_INVALID_ARG, "table: string column %d has a NULL pointer", i);
  if (!t->entry_offsets || !t->crew.list_offsets || !t->actions.list_offsets || !t->created_at || !t->archived_at ||
      !t->delay_sec || !t->delay_valid || !t->entry_ts)
    return fail(PIE_ERR_INVALID_ARG, "table: a NULL column");
  return PIE_OK;
}

uint64_t pie_ingest_fill_scratch_bytes(int64_t n_entries) { return pie::ingest_fill_scratch_bytes(n_entries); }

int pie_ingest_fill_dev(const pie_json_docs* d, const void* scratch, const uint8_t* doc_status, const pie_archive_table* t,
                        void* fill_scratch, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  if ((rc = check_docs(d))) return rc;
  if ((rc = check_table(t))) return rc;
  if (!scratch || !doc_status || !fill_scratch) return fail(PIE_ERR_INVALID_ARG, "NULL argument");
  if (t->n_shows != d->n_docs || t->n_entries < 0) return fail(PIE_ERR_INVALID_ARG, "table.n_shows / n_entries do not match the documents");
  if (reinterpret_cast<uintptr_t>(fill_scratch) & 31) return fail(PIE_ERR_INVALID_ARG, "fill_scratch must be 32-byte aligned");
  PIE_CUDA(pie::launch_ingest_fill(*d, scratch, doc_status, *t, fill_scratch, (cudaStream_t)stream));
  return PIE_OK;
}

namespace {
OutBuffer g_ingest_out;           // the device image of the table
OutBuffer g_ingest_rows;          // the entry rows of the second walk
uint8_t* g_ingest_host = nullptr; // its pinned host image (what pie_ingest_host hands out)
uint64_t g_ingest_host_cap = 0;

// lays the table out in one block (the same on the device and on the host): returns the block size
uint64_t layout_table(pie_archive_table* t, uint8_t* base, int64_t n_docs, const int64_t* totals) {
  uint64_t off = 0;
  auto take = [&](uint64_t bytes) {
    uint8_t* p = base + off;
    off += (bytes + 255) & ~(uint64_t)255;
    return p;
  };
  const int64_t E = totals[PIE_IT_ENTRIES];
  pie_strcol_mut* show_cols[7] = {&t->show_id, &t->show_date, &t->show_time, &t->show_label, &t->lead_pilot, &t->monkey_lead,
                                  &t->show_notes};
  pie_strcol_mut* entry_cols[14] = {&t->entry_id, &t->unit_id, &t->planned, &t->launched, &t->status, &t->primary_issue,
                                    &t->sub_issue, &t->other_detail, &t->severity, &t->root_cause, &t->operator_name,
                                    &t->battery_id, &t->command_rx, &t->notes};
  t->n_shows = n_docs;
  t->n_entries = E;
  t->entry_offsets = (int32_t*)take(4 * (uint64_t)(n_docs + 1));
  for (int h = 0; h < 7; ++h) {
    show_cols[h]->offsets = (int32_t*)take(4 * (uint64_t)(n_docs + 1));
    show_cols[h]->data = take((uint64_t)totals[h] + 8);
  }
  t->crew.list_offsets = (int32_t*)take(4 * (uint64_t)(n_docs + 1));
  t->crew.items.offsets = (int32_t*)take(4 * (uint64_t)(totals[PIE_IT_CREW_ITEMS] + 1));
  t->crew.items.data = take((uint64_t)totals[7] + 8);
  t->created_at = (double*)take(8 * (uint64_t)(n_docs + 1));
  t->archived_at = (double*)take(8 * (uint64_t)(n_docs + 1));
  for (int h = 0; h < 14; ++h) {
    entry_cols[h]->offsets = (int32_t*)take(4 * (uint64_t)(E + 1));
    entry_cols[h]->data = take((uint64_t)totals[8 + h] + 8);
  }
  t->actions.list_offsets = (int32_t*)take(4 * (uint64_t)(E + 1));
  t->actions.items.offsets = (int32_t*)take(4 * (uint64_t)(totals[PIE_IT_ACTION_ITEMS] + 1));
  t->actions.items.data = take((uint64_t)totals[22] + 8);
  t->delay_sec = (double*)take(8 * (uint64_t)(E + 1));
  t->delay_valid = take((uint64_t)E + 8);
  t->entry_ts = (double*)take(8 * (uint64_t)(E + 1));
  t->updated_at = (double*)take(8 * (uint64_t)(n_docs + 1));
  t->deleted_at = (double*)take(8 * (uint64_t)(n_docs + 1));
  t->time_kind = take(PIE_TF_COUNT * (uint64_t)(n_docs + 1));
  return off;
}
}  // namespace

void pie_ingest_host_release(void) {
  std::lock_guard<std::mutex> lock(g_host_mutex);
  if (g_ingest_host) cudaFreeHost(g_ingest_host);
  g_ingest_host = nullptr;
  g_ingest_host_cap = 0;
}

// Host texts -> the table in device memory (g_ingest_out): upload, first walk, sizes, second walk.  `extra` more
// bytes are reserved in the arena for what the caller computes on the table.  Leaves the stream running.
static int ingest_to_device(const pie_json_docs* hd, uint64_t extra, uint8_t* doc_status, int64_t* totals, int64_t* bad_doc,
                            pie_archive_table* dt, uint64_t* block_bytes, uint64_t* h2d_out, uint64_t* d2h_out) {
  if (bad_doc) *bad_doc = -1;
  int rc = ensure_init();
  if (rc) return rc;
  if ((rc = check_docs(hd))) return rc;
  const int64_t n = hd->n_docs;
  if (n > 0 && !doc_status) return fail(PIE_ERR_INVALID_ARG, "doc_status is NULL");
  const int64_t first = hd->offsets[0], last = hd->offsets[n];
  if (first < 0 || last < first) return fail(PIE_ERR_INVALID_ARG, "docs.offsets must ascend from a non-negative value");
  if (last > first && !hd->data) return fail(PIE_ERR_INVALID_ARG, "docs.data is NULL");
  for (int64_t s = 0; s < n; ++s) {
    const int64_t len = hd->offsets[s + 1] - hd->offsets[s];
    if (len < 0) return fail(PIE_ERR_INVALID_ARG, "docs.offsets decrease at document %lld", (long long)s);
    if (len >= 0x7FFFFFF0LL) return fail(PIE_ERR_INVALID_ARG, "document %lld is 2 GiB or more", (long long)s);
  }
  const uint64_t text_bytes = (uint64_t)(last - first);
  const uint64_t scratch_bytes = pie::ingest_scratch_bytes(n);
  uint64_t bytes = pad(8 * (uint64_t)(n + 1)) + pad(text_bytes + 16) + pad(scratch_bytes) + pad((uint64_t)n + 1) + pad(8 * PIE_INGEST_TOTALS) +
                   pad(8) + extra;
  if ((rc = g_arena.reserve(bytes))) return rc;
  cudaStream_t st = g_arena.stream;
  g_cur = &g_arena;
  g_cur_stream = st;
  uint64_t h2d = 0, d2h = 0;
  pie_json_docs dd;
  dd.n_docs = n;
  const int64_t* d_off = nullptr;
  if ((rc = upload_array(hd->offsets, n + 1, &d_off, &h2d))) return rc;
  dd.offsets = d_off;
  // the text keeps its host offsets; its first byte lands 8-byte aligned, like every document start the cursor aligns down to
  uint8_t* d_text = (uint8_t*)g_arena.take(text_bytes + 16);
  const uint64_t skew = (uint64_t)first & 7;
  if (text_bytes) PIE_CUDA(cudaMemcpyAsync(d_text + skew, hd->data + first, text_bytes, cudaMemcpyHostToDevice, st));
  h2d += text_bytes;
  dd.data = d_text + skew - first;
  void* d_scratch = g_arena.take(scratch_bytes);
  uint8_t* d_status_bytes = (uint8_t*)g_arena.take((uint64_t)n + 1);
  int64_t* d_totals = (int64_t*)g_arena.take(8 * PIE_INGEST_TOTALS);
  int32_t* d_status = (int32_t*)g_arena.take(8);
  PIE_CUDA(pie::launch_ingest_measure(dd, d_scratch, d_status_bytes, d_totals, d_status, st));
  int32_t status[2];
  PIE_CUDA(cudaMemcpyAsync(totals, d_totals, 8 * PIE_INGEST_TOTALS, cudaMemcpyDeviceToHost, st));
  PIE_CUDA(cudaMemcpyAsync(status, d_status, sizeof(status), cudaMemcpyDeviceToHost, st));
  if (n > 0) PIE_CUDA(cudaMemcpyAsync(doc_status, d_status_bytes, (uint64_t)n, cudaMemcpyDeviceToHost, st));
  PIE_CUDA(cudaStreamSynchronize(st));
  d2h += 8 * PIE_INGEST_TOTALS + sizeof(status) + (uint64_t)n;
  *h2d_out = h2d;
  *d2h_out = d2h;
  if (status[0] != 0) {
    if (bad_doc) *bad_doc = status[1];
    const char* what = status[0] == PIE_ERR_SCHEMA ? "is not a provider-normalised show"
                       : status[0] == PIE_ERR_UNSUPPORTED_JSON ? "is JSON the ingest kernels do not decide"
                                                               : "makes a heap or row count reach 2 GiB: split the batch";
    return fail(status[0], "document %d %s", status[1], what);
  }
  const uint64_t block = layout_table(dt, nullptr, n, totals);
  if ((rc = g_ingest_out.ensure(block ? block : 256))) return rc;
  layout_table(dt, g_ingest_out.base, n, totals);
  if ((rc = g_ingest_rows.ensure(pie::ingest_fill_scratch_bytes(totals[PIE_IT_ENTRIES])))) return rc;
  PIE_CUDA(pie::launch_ingest_fill(dd, d_scratch, d_status_bytes, *dt, g_ingest_rows.base, st));
  *block_bytes = block;
  return PIE_OK;
}

int pie_ingest_host(const pie_json_docs* hd, pie_archive_table* host_table, uint8_t* doc_status, int64_t* totals_out,
                    int64_t* bad_doc) {
  std::lock_guard<std::mutex> lock(g_host_mutex);
  if (!host_table) return fail(PIE_ERR_INVALID_ARG, "host_table is NULL");
  StreamDrain drain_on_exit{&g_arena.stream};
  int64_t totals[PIE_INGEST_TOTALS] = {0};
  pie_archive_table dt, ht;
  uint64_t block = 0, h2d = 0, d2h = 0;
  int rc = ingest_to_device(hd, 0, doc_status, totals, bad_doc, &dt, &block, &h2d, &d2h);
  g_last_h2d = h2d;
  g_last_d2h = d2h;
  if (totals_out) memcpy(totals_out, totals, sizeof(totals));
  if (rc) return rc;
  cudaStream_t st = g_arena.stream;
  if (block > g_ingest_host_cap) {
    if (g_ingest_host) cudaFreeHost(g_ingest_host);
    g_ingest_host = nullptr;
    g_ingest_host_cap = 0;
    PIE_CUDA(cudaHostAlloc((void**)&g_ingest_host, block, cudaHostAllocDefault));
    g_ingest_host_cap = block;
  }
  layout_table(&ht, g_ingest_host, hd->n_docs, totals);
  PIE_CUDA(cudaMemcpyAsync(g_ingest_host, g_ingest_out.base, block, cudaMemcpyDeviceToHost, st));
  PIE_CUDA(cudaStreamSynchronize(st));
  g_last_d2h = d2h + block;
  *host_table = ht;
  return PIE_OK;
}

namespace {
OutBuffer g_json_csv;        // CSV scratch and row offsets of pie_archive_step_json_host
OutBuffer g_json_csv_bytes;  // ... and the CSV bytes (grown without moving what the kernels before wrote)
}


// ---- the same in chunks of documents over three streams ---------------------------------------------------------
// Shows are independent: a chunk of documents is ingested, its statistics and CSV rows are computed and its rows go
// back to the host while the next chunk's text is still on its way up (PCIe is full duplex) — a single batch
// spends 80 ms uploading 4.4 GB before the first kernel runs and 60 ms downloading after the last.  Only the daily
// grouping needs every show: the chunks' tables stay on the device, their show-level columns are joined at the end
// and the summary runs once.
static int64_t kJsonChunkDocs = 131072;  // pie_set_json_chunk_docs (tests exercise the multi-chunk path)

namespace {
struct JsonChunkState {
  OutBuffer text[2];      // the documents of a chunk (two in flight)
  OutBuffer scratch[2];   // ingest scratch + totals + status + doc_status of a chunk
  OutBuffer rows[2];      // the entry rows of the second walk
  OutBuffer csv[2];       // CSV scratch + row offsets + total
  OutBuffer csv_bytes[2];
  std::vector<OutBuffer> tables;  // one per chunk: they stay until the daily summary has run
  OutBuffer joined;       // the show-level columns of the whole batch, planes, daily outputs
  cudaEvent_t text_free[2] = {nullptr, nullptr};
  long long* h_small = nullptr;  // pinned: totals[26] + status[2] + csv total
};
JsonChunkState g_jc;
}  // namespace

int64_t pie_set_json_chunk_docs(int64_t docs) {
  std::lock_guard<std::mutex> lock(g_host_mutex);
  const int64_t old = kJsonChunkDocs;
  if (docs > 0) kJsonChunkDocs = docs;
  return old;
}

static int archive_step_json_chunked(const pie_json_docs* hd, int32_t tz_offset_minutes, uint8_t* doc_status, int32_t* stats_i32,
                                     double* stats_f64, int64_t stats_stride, const pie_daily_out* hout, int64_t* row_offsets,
                                     int64_t row_capacity, uint8_t* out_data, uint64_t out_capacity, int64_t* n_entries,
                                     uint64_t* total_bytes, int64_t* bad_doc) {
  int rc;
  const int64_t S = hd->n_docs, Sc = S > 0 ? S : 1;
  if ((rc = g_pipe.init())) return rc;
  struct PipeDrain {
    ~PipeDrain() {
      cudaStreamSynchronize(g_pipe.h2d);
      cudaStreamSynchronize(g_pipe.cmp);
      cudaStreamSynchronize(g_pipe.d2h);
    }
  } drain_on_exit;
  if (!g_jc.h_small) {
    PIE_CUDA(cudaHostAlloc((void**)&g_jc.h_small, 512, cudaHostAllocDefault));
    for (int i = 0; i < 2; ++i) PIE_CUDA(cudaEventCreateWithFlags(&g_jc.text_free[i], cudaEventDisableTiming));
  }
  const bool want_stats = stats_i32 != nullptr;
  // chunks of documents
  std::vector<int64_t> cut{0};
  while (cut.back() < S) cut.push_back(cut.back() + kJsonChunkDocs < S ? cut.back() + kJsonChunkDocs : S);
  const int K = (int)cut.size() - 1;
  if ((int)g_jc.tables.size() < K) g_jc.tables.resize((size_t)K);
  // whole-batch device arrays: document offsets, planes, doc_status
  const uint64_t bytes = pad(8 * (uint64_t)(S + 1)) + pad(4ull * PIE_SI_COUNT * Sc) + pad(8ull * PIE_SF_COUNT * Sc) + pad(64);
  if ((rc = g_arena.reserve(bytes))) return rc;
  int64_t* d_off = (int64_t*)g_arena.take(8 * (uint64_t)(S + 1));
  int32_t* d_si = (int32_t*)g_arena.take(4ull * PIE_SI_COUNT * Sc);
  double* d_sf = (double*)g_arena.take(8ull * PIE_SF_COUNT * Sc);
  uint64_t h2d = 0, d2h = 0;
  PIE_CUDA(cudaMemcpyAsync(d_off, hd->offsets, 8 * (uint64_t)(S + 1), cudaMemcpyHostToDevice, g_pipe.h2d));
  h2d += 8 * (uint64_t)(S + 1);

  struct Chunk {
    pie_json_docs dd;
    pie_archive_table dt;
    int64_t totals[PIE_INGEST_TOTALS];
    int64_t e_base;
  };
  std::vector<Chunk> ch((size_t)K);
  auto upload = [&](int k) -> int {
    const int slot = k & 1;
    const int64_t first = hd->offsets[cut[(size_t)k]], last = hd->offsets[cut[(size_t)k + 1]];
    const uint64_t nbytes = (uint64_t)(last - first);
    if (k >= 2) PIE_CUDA(cudaStreamWaitEvent(g_pipe.h2d, g_jc.text_free[slot], 0));  // chunk k-2 is through its second walk
    int rc2 = g_jc.text[slot].ensure(nbytes + 64);
    if (rc2) return rc2;
    const uint64_t skew = (uint64_t)first & 7;
    if (nbytes) PIE_CUDA(cudaMemcpyAsync(g_jc.text[slot].base + skew, hd->data + first, nbytes, cudaMemcpyHostToDevice, g_pipe.h2d));
    h2d += nbytes;
    ch[(size_t)k].dd.n_docs = cut[(size_t)k + 1] - cut[(size_t)k];
    ch[(size_t)k].dd.offsets = d_off + cut[(size_t)k];
    ch[(size_t)k].dd.data = g_jc.text[slot].base + skew - first;
    PIE_CUDA(cudaEventRecord(g_pipe.h2d_done[slot], g_pipe.h2d));
    return PIE_OK;
  };

  unsigned long long bias = 0;
  int64_t e_base = 0;
  bool overflow = false;
  if (K > 0 && (rc = upload(0))) return rc;
  for (int k = 0; k < K; ++k) {
    const int slot = k & 1;
    Chunk& c = ch[(size_t)k];
    const int64_t n = c.dd.n_docs, s0 = cut[(size_t)k];
    const uint64_t scratch_bytes = pie::ingest_scratch_bytes(n);
    if (k >= 2) {  // the slot's buffers are chunk k-2's until its rows and doc_status have left for the host
      PIE_CUDA(cudaStreamWaitEvent(g_pipe.cmp, g_pipe.d2h_done[slot], 0));
      PIE_CUDA(cudaEventSynchronize(g_pipe.d2h_done[slot]));  // (a buffer that has to grow is freed: nothing may still read it)
    }
    if ((rc = g_jc.scratch[slot].ensure(pad(scratch_bytes) + pad((uint64_t)n + 1) + pad(8 * PIE_INGEST_TOTALS) + pad(16)))) return rc;
    uint8_t* p = g_jc.scratch[slot].base;
    void* d_scratch = p; p += pad(scratch_bytes);
    uint8_t* d_docstat = p; p += pad((uint64_t)n + 1);
    PIE_CUDA(cudaStreamWaitEvent(g_pipe.cmp, g_pipe.h2d_done[slot], 0));
    // (totals and status go straight into mapped pinned memory: no small copy behind a large download)
    PIE_CUDA(pie::launch_ingest_measure(c.dd, d_scratch, d_docstat, (int64_t*)g_jc.h_small,
                                        (int32_t*)(g_jc.h_small + PIE_INGEST_TOTALS), g_pipe.cmp));
    if (k + 1 < K && (rc = upload(k + 1))) return rc;  // the next chunk's text goes up while this one is walked
    PIE_CUDA(cudaStreamSynchronize(g_pipe.cmp));
    d2h += 8 * PIE_INGEST_TOTALS + 8;
    memcpy(c.totals, g_jc.h_small, sizeof(c.totals));
    const int32_t* status = reinterpret_cast<const int32_t*>(g_jc.h_small + PIE_INGEST_TOTALS);
    if (status[0] != 0) {
      if (bad_doc) *bad_doc = status[1] >= 0 ? s0 + status[1] : -1;
      const char* what = status[0] == PIE_ERR_SCHEMA ? "is not a provider-normalised show"
                         : status[0] == PIE_ERR_UNSUPPORTED_JSON ? "is JSON the ingest kernels do not decide"
                                                                 : "makes a heap or row count reach 2 GiB: split the batch";
      return fail(status[0], "document %lld %s", (long long)(s0 + status[1]), what);
    }
    const int64_t E = c.totals[PIE_IT_ENTRIES];
    c.e_base = e_base;
    const uint64_t block = layout_table(&c.dt, nullptr, n, c.totals);
    if ((rc = g_jc.tables[(size_t)k].ensure(block ? block : 256))) return rc;
    layout_table(&c.dt, g_jc.tables[(size_t)k].base, n, c.totals);
    if ((rc = g_jc.rows[slot].ensure(pie::ingest_fill_scratch_bytes(E)))) return rc;
    PIE_CUDA(pie::launch_ingest_fill(c.dd, d_scratch, d_docstat, c.dt, g_jc.rows[slot].base, g_pipe.cmp));
    PIE_CUDA(cudaEventRecord(g_jc.text_free[slot], g_pipe.cmp));
    pie_archive_view dv;
    memcpy(&dv, &c.dt, sizeof(dv));
    if (n > 0) PIE_CUDA(pie::launch_show_stats(dv, d_si + s0, d_sf + s0, Sc, g_sm_count, g_pipe.cmp));
    const uint64_t csv_scratch = pad(pie::csv_scratch_bytes(E)), off_bytes = pad(8 * (uint64_t)(E + 1));
    if ((rc = g_jc.csv[slot].ensure(csv_scratch + off_bytes + pad(8)))) return rc;
    void* d_cscratch = g_jc.csv[slot].base;
    int64_t* d_rows = (int64_t*)(g_jc.csv[slot].base + csv_scratch);
    unsigned long long* d_total = (unsigned long long*)(g_jc.csv[slot].base + csv_scratch + off_bytes);
    PIE_CUDA(launch_rows(kFormatCsv, dv, d_rows, nullptr, 0, bias, (unsigned long long*)(g_jc.h_small + 32), d_cscratch, g_pipe.cmp));
    PIE_CUDA(cudaStreamSynchronize(g_pipe.cmp));
    d2h += 8;
    const unsigned long long total = (unsigned long long)g_jc.h_small[32];
    const bool fits = out_data && row_offsets && bias + total <= out_capacity && e_base + E + 1 <= row_capacity;
    if (!fits) overflow = true;
    if (fits && !overflow) {
      if ((rc = g_jc.csv_bytes[slot].ensure(total ? total : 256))) return rc;
      PIE_CUDA(launch_rows(kFormatCsv, dv, d_rows, g_jc.csv_bytes[slot].base, total, bias, d_total, d_cscratch, g_pipe.cmp));
    }
    PIE_CUDA(cudaEventRecord(g_pipe.kernel_done[slot], g_pipe.cmp));
    PIE_CUDA(cudaStreamWaitEvent(g_pipe.d2h, g_pipe.kernel_done[slot], 0));
    if (n > 0) PIE_CUDA(cudaMemcpyAsync(doc_status + s0, d_docstat, (uint64_t)n, cudaMemcpyDeviceToHost, g_pipe.d2h));
    d2h += (uint64_t)n;
    if (fits && !overflow) {
      if (E > 0) PIE_CUDA(cudaMemcpyAsync(row_offsets + e_base, d_rows, 8 * (uint64_t)E, cudaMemcpyDeviceToHost, g_pipe.d2h));
      if (total) PIE_CUDA(cudaMemcpyAsync(out_data + bias, g_jc.csv_bytes[slot].base, total, cudaMemcpyDeviceToHost, g_pipe.d2h));
      d2h += 8 * (uint64_t)E + total;
    }
    PIE_CUDA(cudaEventRecord(g_pipe.d2h_done[slot], g_pipe.d2h));
    bias += total;
    e_base += E;
  }
  const int64_t E_all = e_base;
  *n_entries = E_all;
  *total_bytes = bias;

  // ---- the daily groups read show-level columns of every chunk: join them, run the summary once
  uint64_t date_bytes = 0, time_bytes = 0;
  for (const Chunk& c : ch) { date_bytes += (uint64_t)c.totals[1]; time_bytes += (uint64_t)c.totals[2]; }
  if (date_bytes >= 0x7FFFFFF0ull || time_bytes >= 0x7FFFFFF0ull || E_all >= 0x7FFFFFF0LL)
    return fail(PIE_ERR_CAPACITY, "the batch's show dates / entries reach 2 GiB: split the batch");
  const uint64_t jbytes = pad(4 * (uint64_t)(S + 1)) * 3 + pad(8 * (uint64_t)Sc) * 2 + pad(8 * (uint64_t)(E_all + 1)) + pad(date_bytes + 16) +
                          pad(time_bytes + 16) + daily_out_bytes(S, Sc) + pad(64);
  if ((rc = g_jc.joined.ensure(jbytes))) return rc;
  Arena ja;
  ja.base = g_jc.joined.base; ja.cap = g_jc.joined.cap; ja.used = 0;
  int32_t* j_eo = (int32_t*)ja.take(4 * (uint64_t)(S + 1));
  int32_t* j_date_off = (int32_t*)ja.take(4 * (uint64_t)(S + 1));
  int32_t* j_time_off = (int32_t*)ja.take(4 * (uint64_t)(S + 1));
  double* j_created = (double*)ja.take(8 * (uint64_t)Sc);
  double* j_archived = (double*)ja.take(8 * (uint64_t)Sc);
  double* j_ets = (double*)ja.take(8 * (uint64_t)(E_all + 1));
  uint8_t* j_date = (uint8_t*)ja.take(date_bytes + 16);
  uint8_t* j_time = (uint8_t*)ja.take(time_bytes + 16);
  uint64_t db = 0, tb = 0;
  for (int k = 0; k < K; ++k) {
    const Chunk& c = ch[(size_t)k];
    const int64_t n = c.dd.n_docs, s0 = cut[(size_t)k], E = c.totals[PIE_IT_ENTRIES];
    const bool last = k + 1 == K;
    PIE_CUDA(pie::launch_rebase_i32(j_eo + s0, c.dt.entry_offsets, n + (last ? 1 : 0), (int32_t)c.e_base, g_pipe.cmp));
    PIE_CUDA(pie::launch_rebase_i32(j_date_off + s0, c.dt.show_date.offsets, n + (last ? 1 : 0), (int32_t)db, g_pipe.cmp));
    PIE_CUDA(pie::launch_rebase_i32(j_time_off + s0, c.dt.show_time.offsets, n + (last ? 1 : 0), (int32_t)tb, g_pipe.cmp));
    if (n > 0) {
      PIE_CUDA(cudaMemcpyAsync(j_created + s0, c.dt.created_at, 8 * (uint64_t)n, cudaMemcpyDeviceToDevice, g_pipe.cmp));
      PIE_CUDA(cudaMemcpyAsync(j_archived + s0, c.dt.archived_at, 8 * (uint64_t)n, cudaMemcpyDeviceToDevice, g_pipe.cmp));
    }
    if (E > 0) PIE_CUDA(cudaMemcpyAsync(j_ets + c.e_base, c.dt.entry_ts, 8 * (uint64_t)E, cudaMemcpyDeviceToDevice, g_pipe.cmp));
    if (c.totals[1]) PIE_CUDA(cudaMemcpyAsync(j_date + db, c.dt.show_date.data, (uint64_t)c.totals[1], cudaMemcpyDeviceToDevice, g_pipe.cmp));
    if (c.totals[2]) PIE_CUDA(cudaMemcpyAsync(j_time + tb, c.dt.show_time.data, (uint64_t)c.totals[2], cudaMemcpyDeviceToDevice, g_pipe.cmp));
    db += (uint64_t)c.totals[1];
    tb += (uint64_t)c.totals[2];
  }
  if (K == 0) {
    PIE_CUDA(cudaMemsetAsync(j_eo, 0, 4, g_pipe.cmp));
    PIE_CUDA(cudaMemsetAsync(j_date_off, 0, 4, g_pipe.cmp));
    PIE_CUDA(cudaMemsetAsync(j_time_off, 0, 4, g_pipe.cmp));
  }
  pie_archive_view jv;
  memset(&jv, 0, sizeof(jv));
  jv.n_shows = S;
  jv.n_entries = E_all;
  jv.entry_offsets = j_eo;
  jv.show_date.offsets = j_date_off; jv.show_date.data = j_date;
  jv.show_time.offsets = j_time_off; jv.show_time.data = j_time;
  jv.created_at = j_created;
  jv.archived_at = j_archived;
  jv.entry_ts = j_ets;
  pie_daily_out dout;
  void* dscratch = alloc_daily_out(ja, S, Sc, &dout);
  PIE_CUDA(pie::launch_daily_summary(jv, d_si, d_sf, Sc, tz_offset_minutes, dout, dscratch, g_sm_count, g_pipe.cmp));
  if (want_stats && S > 0) {
    PIE_CUDA(cudaMemcpy2DAsync(stats_i32, 4 * (uint64_t)stats_stride, d_si, 4 * (uint64_t)Sc, 4 * (uint64_t)S, PIE_SI_COUNT,
                               cudaMemcpyDeviceToHost, g_pipe.cmp));
    PIE_CUDA(cudaMemcpy2DAsync(stats_f64, 8 * (uint64_t)stats_stride, d_sf, 8 * (uint64_t)Sc, 8 * (uint64_t)S, PIE_SF_COUNT,
                               cudaMemcpyDeviceToHost, g_pipe.cmp));
    d2h += (4ull * PIE_SI_COUNT + 8ull * PIE_SF_COUNT) * (uint64_t)S;
  }
  rc = download_daily(hout, dout, S, Sc, g_pipe.cmp, &d2h);
  PIE_CUDA(cudaStreamSynchronize(g_pipe.d2h));
  PIE_CUDA(cudaStreamSynchronize(g_pipe.cmp));
  if (row_offsets && !overflow) row_offsets[E_all] = (int64_t)bias;
  g_last_h2d = h2d;
  g_last_d2h = d2h + 8;
  if (rc) return rc;
  if (overflow)
    return fail(PIE_ERR_CAPACITY, "CSV needs %llu bytes and %lld row offsets (given %llu and %lld)", bias, (long long)(E_all + 1),
                (unsigned long long)out_capacity, (long long)row_capacity);
  return PIE_OK;
}

int pie_archive_step_json_host(const pie_json_docs* hd, int32_t tz_offset_minutes, uint8_t* doc_status, int32_t* stats_i32,
                               double* stats_f64, int64_t stats_stride, const pie_daily_out* hout, int64_t* row_offsets,
                               int64_t row_capacity, uint8_t* out_data, uint64_t out_capacity, int64_t* n_entries,
                               uint64_t* total_bytes, int64_t* bad_doc) {
  std::lock_guard<std::mutex> lock(g_host_mutex);
  if (tz_offset_minutes < -24 * 60 || tz_offset_minutes > 24 * 60) return fail(PIE_ERR_INVALID_ARG, "tz_offset_minutes out of range");
  if (!hd || !hout || !n_entries || !total_bytes) return fail(PIE_ERR_INVALID_ARG, "NULL argument");
  const int64_t S = hd->n_docs;
  StreamDrain drain_on_exit{&g_arena.stream};
  const bool want_stats = stats_i32 != nullptr || stats_f64 != nullptr;
  if (want_stats && (!stats_i32 || !stats_f64)) return fail(PIE_ERR_INVALID_ARG, "stats_i32 and stats_f64 go together");
  if (want_stats && stats_stride < S) return fail(PIE_ERR_INVALID_ARG, "stats_stride < n_docs");
  int rc;
  if (S >= 0 && (rc = check_daily_out(hout, S))) return rc;
  if (out_data && row_offsets && S > kJsonChunkDocs) {  // a full request on a large batch: the pipelined form
    if (bad_doc) *bad_doc = -1;
    if ((rc = ensure_init())) return rc;
    if ((rc = check_docs(hd))) return rc;
    if (!doc_status) return fail(PIE_ERR_INVALID_ARG, "doc_status is NULL");
    if (hd->offsets[0] < 0) return fail(PIE_ERR_INVALID_ARG, "docs.offsets must ascend from a non-negative value");
    for (int64_t i = 0; i < S; ++i) {
      const int64_t len = hd->offsets[i + 1] - hd->offsets[i];
      if (len < 0) return fail(PIE_ERR_INVALID_ARG, "docs.offsets decrease at document %lld", (long long)i);
      if (len >= 0x7FFFFFF0LL) return fail(PIE_ERR_INVALID_ARG, "document %lld is 2 GiB or more", (long long)i);
    }
    if (hd->offsets[S] > hd->offsets[0] && !hd->data) return fail(PIE_ERR_INVALID_ARG, "docs.data is NULL");
    return archive_step_json_chunked(hd, tz_offset_minutes, doc_status, stats_i32, stats_f64, stats_stride, hout, row_offsets,
                                     row_capacity, out_data, out_capacity, n_entries, total_bytes, bad_doc);
  }
  const int64_t Sc = S > 0 ? S : 1;
  const uint64_t extra = pad(4ull * PIE_SI_COUNT * Sc) + pad(8ull * PIE_SF_COUNT * Sc) + daily_out_bytes(S, Sc) + pad(64);
  int64_t totals[PIE_INGEST_TOTALS] = {0};
  pie_archive_table dt;
  uint64_t block = 0, h2d = 0, d2h = 0;
  rc = ingest_to_device(hd, extra, doc_status, totals, bad_doc, &dt, &block, &h2d, &d2h);
  g_last_h2d = h2d;
  g_last_d2h = d2h;
  if (rc) return rc;
  cudaStream_t st = g_arena.stream;
  const int64_t E = totals[PIE_IT_ENTRIES];
  *n_entries = E;
  // the table the kernels read: pie_archive_table and pie_archive_view share their layout
  static_assert(sizeof(pie_archive_view) == sizeof(pie_archive_table), "the view is the table with const pointers");
  pie_archive_view dv;
  memcpy(&dv, &dt, sizeof(dv));
  int32_t* d_si = (int32_t*)g_arena.take(4ull * PIE_SI_COUNT * Sc);
  double* d_sf = (double*)g_arena.take(8ull * PIE_SF_COUNT * Sc);
  PIE_CUDA(pie::launch_show_stats(dv, d_si, d_sf, Sc, g_sm_count, st));
  pie_daily_out dout;
  void* dscratch = alloc_daily_out(g_arena, S, Sc, &dout);
  PIE_CUDA(pie::launch_daily_summary(dv, d_si, d_sf, Sc, tz_offset_minutes, dout, dscratch, g_sm_count, st));
  // CSV rows: sizes first, then the bytes (row offsets of the size pass are final)
  const uint64_t csv_scratch = pad(pie::csv_scratch_bytes(E)), off_bytes = pad(8 * (uint64_t)(E + 1));
  if ((rc = g_json_csv.ensure(csv_scratch + off_bytes + pad(8)))) return rc;
  void* d_cscratch = g_json_csv.base;
  int64_t* d_rows = (int64_t*)(g_json_csv.base + csv_scratch);
  unsigned long long* d_total = (unsigned long long*)(g_json_csv.base + csv_scratch + off_bytes);
  PIE_CUDA(launch_rows(kFormatCsv, dv, d_rows, nullptr, 0, 0ull, d_total, d_cscratch, st));
  unsigned long long total = 0;
  PIE_CUDA(cudaMemcpyAsync(&total, d_total, 8, cudaMemcpyDeviceToHost, st));
  if (want_stats && S > 0) {
    PIE_CUDA(cudaMemcpy2DAsync(stats_i32, 4 * (uint64_t)stats_stride, d_si, 4 * (uint64_t)Sc, 4 * (uint64_t)S, PIE_SI_COUNT,
                               cudaMemcpyDeviceToHost, st));
    PIE_CUDA(cudaMemcpy2DAsync(stats_f64, 8 * (uint64_t)stats_stride, d_sf, 8 * (uint64_t)Sc, 8 * (uint64_t)S, PIE_SF_COUNT,
                               cudaMemcpyDeviceToHost, st));
    d2h += (4ull * PIE_SI_COUNT + 8ull * PIE_SF_COUNT) * (uint64_t)S;
  }
  rc = download_daily(hout, dout, S, Sc, st, &d2h);  // synchronises: `total` is in
  g_last_d2h = d2h;
  if (rc) return rc;
  *total_bytes = total;
  if (!out_data || !row_offsets || out_capacity < total || row_capacity < E + 1) {
    PIE_CUDA(cudaStreamSynchronize(st));
    if (!out_data && !row_offsets) return PIE_OK;  // a size query
    return fail(PIE_ERR_CAPACITY, "CSV needs %llu bytes and %lld row offsets (given %llu and %lld)", total,
                (long long)(E + 1), (unsigned long long)out_capacity, (long long)row_capacity);
  }
  // the bytes: a second buffer after the offsets (grown without moving what the kernels above wrote)
  OutBuffer& csv_bytes = g_json_csv_bytes;
  if ((rc = csv_bytes.ensure(total ? total : 256))) return rc;
  PIE_CUDA(launch_rows(kFormatCsv, dv, d_rows, csv_bytes.base, total, 0ull, d_total, d_cscratch, st));
  PIE_CUDA(cudaMemcpyAsync(row_offsets, d_rows, 8 * (uint64_t)(E + 1), cudaMemcpyDeviceToHost, st));
  if (total) PIE_CUDA(cudaMemcpyAsync(out_data, csv_bytes.base, total, cudaMemcpyDeviceToHost, st));
  PIE_CUDA(cudaStreamSynchronize(st));
  g_last_d2h = d2h + 8 * (uint64_t)(E + 1) + total;
  return PIE_OK;
}

/* Gives back everything the host entry points keep between calls: the device arenas and output buffers (they only
 * ever grow), the pinned host image of pie_ingest_host, the streams and events.  The next call allocates afresh. */
int pie_release(void) {
  std::lock_guard<std::mutex> lock(g_host_mutex);
  if (g_sm_count <= 0) return PIE_OK;  // never initialised: nothing is held
  PIE_CUDA(cudaDeviceSynchronize());
  auto free_arena = [](Arena& a, bool owns_stream) {
    if (a.base) cudaFree(a.base);
    if (a.stream && owns_stream) cudaStreamDestroy(a.stream);
    a = Arena();
  };
  auto free_out = [](OutBuffer& b) {
    if (b.base) cudaFree(b.base);
    b = OutBuffer();
  };
  free_arena(g_arena, true);
  free_arena(g_pipe_in[0], false);  // they borrow the pipeline's upload stream
  free_arena(g_pipe_in[1], false);
  free_out(g_pipe.out[0]);
  free_out(g_pipe.out[1]);
  free_out(g_pipe.off[0]);
  free_out(g_pipe.off[1]);
  pie::ingest_release();
  free_out(g_ingest_out);
  free_out(g_ingest_rows);
  free_out(g_json_csv);
  free_out(g_json_csv_bytes);
  for (int i = 0; i < 2; ++i) {
    free_out(g_jc.text[i]); free_out(g_jc.scratch[i]); free_out(g_jc.rows[i]); free_out(g_jc.csv[i]); free_out(g_jc.csv_bytes[i]);
    if (g_jc.text_free[i]) cudaEventDestroy(g_jc.text_free[i]);
    g_jc.text_free[i] = nullptr;
  }
  for (OutBuffer& b : g_jc.tables) free_out(b);
  g_jc.tables.clear();
  free_out(g_jc.joined);
  if (g_jc.h_small) cudaFreeHost(g_jc.h_small);
  g_jc.h_small = nullptr;
  if (g_ingest_host) cudaFreeHost(g_ingest_host);
  g_ingest_host = nullptr;
  g_ingest_host_cap = 0;
  if (g_check_flags) cudaFree(g_check_flags);
  g_check_flags = nullptr;
  if (g_maint_err) cudaFree(g_maint_err);
  g_maint_err = nullptr;
  if (g_pipe.h2d) {
    cudaStreamDestroy(g_pipe.h2d);
    cudaStreamDestroy(g_pipe.cmp);
    cudaStreamDestroy(g_pipe.d2h);
    for (int i = 0; i < 2; ++i) {
      cudaEventDestroy(g_pipe.h2d_done[i]);
      cudaEventDestroy(g_pipe.kernel_done[i]);
      cudaEventDestroy(g_pipe.d2h_done[i]);
    }
    cudaEventDestroy(g_pipe.daily_up);
    cudaFreeHost(g_pipe.h_total);
    g_pipe = CsvPipeline();
  }
  g_cur = &g_arena;
  g_cur_stream = nullptr;
  cudaGetLastError();
  return PIE_OK;
}

/* ---- the schemaVersion 2 show payload ------------------------------------------------------------------------ */
uint64_t pie_show_payloads_scratch_bytes(int64_t n_shows) { return pie::show_payload_scratch_bytes(n_shows > 0 ? n_shows : 0); }

int pie_show_payloads_dev(const pie_archive_view* v, const uint8_t* head, int32_t head_len, const uint8_t* tail,
                          int32_t tail_len, int64_t* doc_offsets, uint8_t* out_data, uint64_t out_capacity,
                          uint64_t* total_bytes_dev, int32_t* status_dev, void* scratch, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  if ((rc = check_view_common(v))) return rc;
  if ((rc = check_export_view_dev(v, kFormatCsv))) return rc;  // every column the CSV rows read
  if (!doc_offsets || !total_bytes_dev || !status_dev || !scratch) return fail(PIE_ERR_INVALID_ARG, "NULL argument");
  if (head_len < 0 || tail_len < 0 || (head_len > 0 && !head) || (tail_len > 0 && !tail))
    return fail(PIE_ERR_INVALID_ARG, "head / tail");
  if (v->n_shows > 0 && (!v->created_at || !v->archived_at)) return fail(PIE_ERR_INVALID_ARG, "created_at / archived_at is NULL");
  if (reinterpret_cast<uintptr_t>(scratch) & 255) return fail(PIE_ERR_INVALID_ARG, "scratch must be 256-byte aligned");
  PIE_CUDA(pie::launch_show_payloads(*v, head, head_len, tail, tail_len, doc_offsets, out_data, out_data ? out_capacity : 0,
                                     (unsigned long long*)total_bytes_dev, status_dev, scratch, (cudaStream_t)stream));
  return PIE_OK;
}

/* ---- _getTimestamp of the documents' time fields; archive maintenance decisions ---------------------------- */
int pie_get_timestamps_dev(const pie_archive_view* v, const pie_json_docs* docs, int32_t tz_offset_minutes,
                           const pie_doc_times* out, int32_t* status_dev, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  if (!v || !out || !status_dev) return fail(PIE_ERR_INVALID_ARG, "NULL argument");
  if (v->n_shows < 0 || v->n_shows > 0x7FFFFFF0LL) return fail(PIE_ERR_INVALID_ARG, "n_shows out of range");
  if (tz_offset_minutes < -24 * 60 || tz_offset_minutes > 24 * 60) return fail(PIE_ERR_INVALID_ARG, "tz_offset_minutes out of range");
  if (docs && docs->n_docs > 0 && !docs->data) return fail(PIE_ERR_INVALID_ARG, "docs.data is NULL");
  if (!g_maint_err) {
    std::lock_guard<std::mutex> lock(g_host_mutex);
    if (!g_maint_err) PIE_CUDA(cudaMalloc(&g_maint_err, 8));
  }
  PIE_CUDA(pie::launch_get_timestamps(*v, docs, tz_offset_minutes, *out, status_dev, g_maint_err, (cudaStream_t)stream));
  return PIE_OK;
}

uint64_t pie_archive_due_scratch_bytes(int64_t n_shows) { return pie::archive_due_scratch_bytes(n_shows > 0 ? n_shows : 0); }

int pie_archive_due_dev(const pie_archive_view* v, const uint8_t* doc_status, const double* created, double now_ms,
                        uint8_t* due, int32_t* group_first, void* scratch, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  if (!v || !due || !group_first || !scratch) return fail(PIE_ERR_INVALID_ARG, "NULL argument");
  if (v->n_shows < 0 || v->n_shows > 0x3FFFFFF0LL) return fail(PIE_ERR_INVALID_ARG, "n_shows out of range (batches of < 2^30 shows)");
  if (v->n_shows > 0 && (!created || !v->show_date.offsets)) return fail(PIE_ERR_INVALID_ARG, "created / show_date is NULL");
  if (reinterpret_cast<uintptr_t>(scratch) & 15) return fail(PIE_ERR_INVALID_ARG, "scratch must be 16-byte aligned");
  PIE_CUDA(pie::launch_archive_due(*v, doc_status, created, now_ms, due, group_first, scratch, (cudaStream_t)stream));
  return PIE_OK;
}

int pie_archive_expired_dev(const double* created, int64_t n, double now_ms, int32_t tz_offset_minutes, uint8_t* expired,
                            void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  if (n < 0 || (n > 0 && (!created || !expired))) return fail(PIE_ERR_INVALID_ARG, "NULL argument / negative size");
  if (tz_offset_minutes < -24 * 60 || tz_offset_minutes > 24 * 60) return fail(PIE_ERR_INVALID_ARG, "tz_offset_minutes out of range");
  PIE_CUDA(pie::launch_archive_expired(created, n, now_ms, tz_offset_minutes, expired, (cudaStream_t)stream));
  return PIE_OK;
}

int pie_selftest_fast_div(int32_t max_b, uint64_t* mismatches) {
  int rc = ensure_init();
  if (rc) return rc;
  if (!mismatches || max_b < 1 || max_b > 4096) return fail(PIE_ERR_INVALID_ARG, "max_b must be in 1..4096");
  unsigned long long* d = nullptr;
  PIE_CUDA(cudaMalloc(&d, 8));
  PIE_CUDA(cudaMemset(d, 0, 8));
  PIE_CUDA(pie::launch_selftest_fast_div(max_b, d, nullptr));
  unsigned long long h = 0;
  PIE_CUDA(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
  PIE_CUDA(cudaFree(d));
  *mismatches = h;
  return PIE_OK;
}

}  // extern "C"
