/* A plain C99 caller of the C ABI (include/sph_pie_b200.h), with nothing but malloc'd buffers: what the reference's
 * N-API addon does, minus Node.  It hands the provider's stored texts — one
 * JSON.stringify(show) per line, the `data` column of show_archive (reference server/storage/sqlProvider.js:696) — to
 * pie_archive_step_json_host and writes what listArchivedShows (:230-234) + buildArchiveDailyGroups (public/app.js:
 * 3401-3443) + buildCsvRow over every entry (server/webhookDispatcher.js:332-342) would have produced:
 *
 *     c_consumer <documents.jsonl> <tz_offset_minutes> <out_prefix>
 *
 *   <out_prefix>.csv        the CSV rows, one per entry, '\n' after each
 *   <out_prefix>.stats_i32  int32[PIE_SI_COUNT][n_docs], the per-show statistics planes
 *   <out_prefix>.status     uint8[n_docs], 1 = the row is dropped (not JSON / not an object)
 *   stdout                  one line: n_docs n_entries csv_bytes n_groups dropped
 *
 * Exit 3 with the library's message when no sm_100 device is visible (there is no CPU fallback), 2 on any other
 * error.  Built and run by tests/test_host_logic.py (no GPU: the loud failure) and tests/test_gpu_abi_errors.py. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sph_pie_b200.h"

static void* xmalloc(size_t n) {
  void* p = malloc(n ? n : 1);
  if (!p) {
    fprintf(stderr, "c_consumer: out of memory (%zu bytes)\n", n);
    exit(2);
  }
  return p;
}

static int write_file(const char* prefix, const char* ext, const void* data, size_t n) {
  char path[4096];
  FILE* f;
  snprintf(path, sizeof path, "%s.%s", prefix, ext);
  f = fopen(path, "wb");
  if (!f) return -1;
  if (n && fwrite(data, 1, n, f) != n) {
    fclose(f);
    return -1;
  }
  return fclose(f);
}

int main(int argc, char** argv) {
  FILE* f;
  long size;
  uint8_t *raw, *text, *doc_status, *csv;
  int64_t *offsets, *row_offsets, n_docs = 0, n_entries = 0, bad_doc = -1, n_groups = 0, s, dropped = 0;
  uint64_t total_bytes = 0;
  size_t i, w, stride;
  int rc, tz;
  int32_t *stats_i32, daily_status[2] = {0, -1};
  double* stats_f64;
  pie_json_docs docs;
  pie_daily_out daily;

  if (argc != 4) {
    fprintf(stderr, "usage: %s <documents.jsonl> <tz_offset_minutes> <out_prefix>\n", argv[0]);
    return 2;
  }
  if (pie_abi_version() != PIE_ABI_VERSION) {
    fprintf(stderr, "c_consumer: header is ABI %d, library is ABI %d\n", PIE_ABI_VERSION, pie_abi_version());
    return 2;
  }
  tz = atoi(argv[2]);
  f = fopen(argv[1], "rb");
  if (!f) {
    perror(argv[1]);
    return 2;
  }
  fseek(f, 0, SEEK_END);
  size = ftell(f);
  fseek(f, 0, SEEK_SET);
  raw = (uint8_t*)xmalloc((size_t)size + 1);
  if (size && fread(raw, 1, (size_t)size, f) != (size_t)size) {
    perror("read");
    return 2;
  }
  fclose(f);
  if (size && raw[size - 1] != '\n') raw[size++] = '\n'; /* the last line may lack its newline */

  /* the documents, back to back (JSON.stringify never leaves a raw newline inside a document) */
  for (i = 0; i < (size_t)size; ++i) n_docs += raw[i] == '\n';
  offsets = (int64_t*)xmalloc(((size_t)n_docs + 1) * sizeof *offsets);
  text = (uint8_t*)xmalloc((size_t)size + 8); /* malloc is 16-byte aligned: the ABI asks for 8 */
  offsets[0] = 0;
  for (i = 0, w = 0, s = 0; i < (size_t)size; ++i) {
    if (raw[i] == '\n')
      offsets[++s] = (int64_t)w;
    else
      text[w++] = raw[i];
  }
  free(raw);

  rc = pie_init(0);
  if (rc != PIE_OK) {
    fprintf(stderr, "c_consumer: pie_init(0) = %d: %s\n", rc, pie_last_error());
    return rc == PIE_ERR_NO_DEVICE ? 3 : 2;
  }

  stride = n_docs ? (size_t)n_docs : 1;
  doc_status = (uint8_t*)xmalloc(stride);
  stats_i32 = (int32_t*)xmalloc(sizeof(int32_t) * PIE_SI_COUNT * stride);
  stats_f64 = (double*)xmalloc(sizeof(double) * PIE_SF_COUNT * stride);
  daily.stride = (int64_t)stride;
  daily.show_day_start = (int64_t*)xmalloc(sizeof(int64_t) * stride);
  daily.show_order = (int32_t*)xmalloc(sizeof(int32_t) * stride);
  daily.group_day_start = (int64_t*)xmalloc(sizeof(int64_t) * stride);
  daily.group_offsets = (int32_t*)xmalloc(sizeof(int32_t) * (stride + 1));
  daily.summary_f64 = (double*)xmalloc(sizeof(double) * PIE_DF_COUNT * PIE_N_METRICS * stride);
  daily.summary_count = (int32_t*)xmalloc(sizeof(int32_t) * PIE_N_METRICS * stride);
  daily.n_groups = &n_groups;
  daily.status = daily_status;
  docs.n_docs = n_docs;
  docs.offsets = offsets;
  docs.data = text;

  /* 1. how many rows, how many bytes (the analytics are delivered by this call already) */
  rc = pie_archive_step_json_host(&docs, tz, doc_status, stats_i32, stats_f64, (int64_t)stride, &daily, NULL, 0, NULL, 0,
                                  &n_entries, &total_bytes, &bad_doc);
  if (rc != PIE_OK) {
    fprintf(stderr, "c_consumer: size query = %d at document %lld: %s\n", rc, (long long)bad_doc, pie_last_error());
    return 2;
  }
  /* 2. the rows */
  row_offsets = (int64_t*)xmalloc(((size_t)n_entries + 1) * sizeof *row_offsets);
  csv = (uint8_t*)xmalloc((size_t)total_bytes);
  rc = pie_archive_step_json_host(&docs, tz, doc_status, stats_i32, stats_f64, (int64_t)stride, &daily, row_offsets,
                                  n_entries + 1, csv, total_bytes, &n_entries, &total_bytes, &bad_doc);
  if (rc != PIE_OK) {
    fprintf(stderr, "c_consumer: step = %d at document %lld: %s\n", rc, (long long)bad_doc, pie_last_error());
    return 2;
  }
  if (n_entries > 0 && (row_offsets[0] != 0 || (uint64_t)row_offsets[n_entries] != total_bytes)) {
    fprintf(stderr, "c_consumer: row offsets do not span the rows\n");
    return 2;
  }
  for (s = 0; s < n_docs; ++s) dropped += doc_status[s] != 0;
  if (write_file(argv[3], "csv", csv, (size_t)total_bytes) || write_file(argv[3], "status", doc_status, (size_t)n_docs) ||
      write_file(argv[3], "stats_i32", stats_i32, sizeof(int32_t) * PIE_SI_COUNT * stride)) {
    perror(argv[3]);
    return 2;
  }
  printf("%lld %lld %llu %lld %lld\n", (long long)n_docs, (long long)n_entries, (unsigned long long)total_bytes,
         (long long)n_groups, (long long)dropped);
  pie_release();
  return 0;
}
