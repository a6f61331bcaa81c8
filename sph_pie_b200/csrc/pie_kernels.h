// Internal launch API between the kernel translation units and the C ABI (pie_capi.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/sph_pie_b200.h"

namespace pie {

// cumulative number of kernels this library has launched (bench.py reports it as gpu_launches)
extern std::atomic<unsigned long long> g_launches;
// SMs of the device pie_init selected (grids are sized from it); 0 before pie_init
extern int g_sm_count;
inline int sm_count_or_default() { return g_sm_count > 0 ? g_sm_count : 148; }

// validate.cu: the offsets arrays of a host batch, checked (and disarmed) on the device
constexpr int kMaxOffsetsArrays = 32;
struct OffsetsArray {
  const int32_t* off;  // device, n + 1 elements
  int64_t n;
  int32_t first, last;  // every offset must lie in [first, last] and never decrease
};
struct OffsetsBatch {
  OffsetsArray a[kMaxOffsetsArrays];
  int count;
};
// flags: device int32[kMaxOffsetsArrays]; flags[i] != 0 afterwards: array i was malformed (and now holds empty rows)
cudaError_t launch_offsets_check(const OffsetsBatch& batch, int32_t* flags, cudaStream_t stream);
// ... and hands the flags to the host through mapped pinned memory: a small device-to-host copy would queue behind
// whatever large download occupies the copy engine
cudaError_t launch_offsets_flags_out(const int32_t* flags, int32_t* mapped_host, cudaStream_t stream);
cudaError_t launch_rebase_i32(int32_t* dst, const int32_t* src, int64_t n, int32_t add, cudaStream_t stream);

// archive_stats.cu
cudaError_t launch_show_stats(const pie_archive_view& dev_view, int32_t* stats_i32, double* stats_f64,
                              int64_t stride, int sm_count, cudaStream_t stream);
cudaError_t launch_selftest_fast_div(int max_b, unsigned long long* d_mismatches, cudaStream_t stream);

// csv_rows.cu
uint64_t csv_scratch_bytes(int64_t n_entries);
// row_offsets receive offset_bias + the offset inside out_data (chunked callers place chunks back to back)
cudaError_t launch_csv_rows(const pie_archive_view& dev_view, int64_t* row_offsets, uint8_t* out_data,
                            uint64_t capacity, unsigned long long offset_bias, unsigned long long* total_out,
                            void* scratch, cudaStream_t stream);

// JSON.stringify(buildArchiveEntryPayload(show, entry)) + '\n' per entry: same kernel, another row format
cudaError_t launch_payload_rows(const pie_archive_view& dev_view, int64_t* row_offsets, uint8_t* out_data,
                                uint64_t capacity, unsigned long long offset_bias, unsigned long long* total_out,
                                void* scratch, cudaStream_t stream);

// debug knobs of the export-row kernel (tests): force every tile through the slow path (on < 0 only
// queries; returns the previous value); number of tiles of the last launch on `scratch` that took it
int csv_set_force_slow(int on);
cudaError_t csv_read_slow_tiles(const void* scratch, int64_t n_entries, unsigned int* out, cudaStream_t stream);

// live_metrics.cu: computeMetrics(show) per show; metrics_i32 is int32[PIE_CM_COUNT][stride], text 32 bytes per show
cudaError_t launch_compute_metrics(const pie_archive_view& dev_view, int32_t* metrics_i32, uint8_t* avg_delay_text,
                                   int64_t stride, cudaStream_t stream);

// json_ingest.cu: stored show documents -> the columnar table, two walks (measure, scan, fill)
uint64_t ingest_scratch_bytes(int64_t n_docs);
cudaError_t launch_ingest_measure(const pie_json_docs& dev_docs, void* scratch, uint8_t* doc_status, int64_t* totals,
                                  int32_t* status, cudaStream_t stream);
// debug knobs (tests, A/B timing): the warp-cooperative path on / off (on < 0 only queries; returns the previous
// value; on unless PIE_INGEST_WARP_PATH=0 is in the environment); how many documents of the last measure on
// `scratch` the warp path declined (they took the thread-per-document walk)
int ingest_set_warp_path(int on);
void ingest_release();  // what the ingest holds beyond the caller's buffers (a stream and two events)
cudaError_t ingest_read_declined(const void* scratch, int64_t n_docs, unsigned int* out, cudaStream_t stream);
uint64_t ingest_fill_scratch_bytes(int64_t n_entries);
cudaError_t launch_ingest_fill(const pie_json_docs& dev_docs, const void* scratch, const uint8_t* doc_status,
                               const pie_archive_table& dev_table, void* fill_scratch, cudaStream_t stream);

// show_payload.cu: the schemaVersion 2 payload document of every show
uint64_t show_payload_scratch_bytes(int64_t n_shows);
cudaError_t launch_show_payloads(const pie_archive_view& dev_view, const uint8_t* head, int head_len, const uint8_t* tail,
                                 int tail_len, int64_t* doc_offsets, uint8_t* out, uint64_t capacity,
                                 unsigned long long* total, int32_t* status, void* scratch, cudaStream_t stream);

// archive_maintenance.cu: _getTimestamp of the documents' time fields; the archive / purge decisions
cudaError_t launch_get_timestamps(const pie_archive_view& dev_view, const pie_json_docs* dev_docs, int32_t tz_offset_minutes,
                                  const pie_doc_times& out, int32_t* status, unsigned long long* err_scratch,
                                  cudaStream_t stream);
uint64_t archive_due_scratch_bytes(int64_t n_shows);
cudaError_t launch_archive_due(const pie_archive_view& dev_view, const uint8_t* doc_status, const double* created, double now_ms,
                               uint8_t* due, int32_t* group_first, void* scratch, cudaStream_t stream);
cudaError_t launch_archive_expired(const double* created, int64_t n, double now_ms, int32_t tz_offset_minutes,
                                   uint8_t* expired, cudaStream_t stream);

// archive_daily.cu
uint64_t daily_scratch_bytes(int64_t n_shows);
cudaError_t launch_daily_summary(const pie_archive_view& dev_view, const int32_t* stats_i32,
                                 const double* stats_f64, int64_t stats_stride, int32_t tz_offset_minutes,
                                 const pie_daily_out& out, void* scratch, int sm_count, cudaStream_t stream);

}  // namespace pie
