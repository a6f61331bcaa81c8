// Host build of sph_pie_b200/csrc/pie_json_walk.cuh (the walk the ingest kernels run per document), driven the way
// json_ingest.cu drives it: measure every document, exclusive scan of the 26 count planes, fill.  Test-only: lets the
// recogniser / projection be checked against the oracle on the CPU.  The product has no such path.
#include <stdint.h>
#include <string.h>

#include "../../sph_pie_b200/csrc/pie_json_walk.cuh"

using namespace pie;
using namespace pie::jw;

static const uint64_t kPow5[PIE_POW5_128_N][2] = PIE_POW5_128_INIT;

static IngestOut make_out(const pie_archive_table& t) {
  IngestOut o;
  o.rows = nullptr;
  const pie_strcol_mut* show_cols[7] = {&t.show_id, &t.show_date, &t.show_time, &t.show_label, &t.lead_pilot, &t.monkey_lead,
                                        &t.show_notes};
  const pie_strcol_mut* entry_cols[14] = {&t.entry_id, &t.unit_id, &t.planned, &t.launched, &t.status, &t.primary_issue,
                                          &t.sub_issue, &t.other_detail, &t.severity, &t.root_cause, &t.operator_name,
                                          &t.battery_id, &t.command_rx, &t.notes};
  for (int h = 0; h < 7; ++h) { o.off[h] = show_cols[h]->offsets; o.data[h] = show_cols[h]->data; }
  o.off[kHeapCrew] = t.crew.items.offsets;
  o.data[kHeapCrew] = t.crew.items.data;
  for (int h = 0; h < 14; ++h) { o.off[kHeapEntry0 + h] = entry_cols[h]->offsets; o.data[kHeapEntry0 + h] = entry_cols[h]->data; }
  o.off[kHeapActions] = t.actions.items.offsets;
  o.data[kHeapActions] = t.actions.items.data;
  o.entry_offsets = t.entry_offsets;
  o.crew_list = t.crew.list_offsets;
  o.actions_list = t.actions.list_offsets;
  o.created_at = t.created_at;
  o.archived_at = t.archived_at;
  o.delay_sec = t.delay_sec;
  o.delay_valid = t.delay_valid;
  o.entry_ts = t.entry_ts;
  o.time_val[PIE_TF_CREATED] = t.created_at;
  o.time_val[PIE_TF_UPDATED] = t.updated_at;
  o.time_val[PIE_TF_ARCHIVED] = t.archived_at;
  o.time_val[PIE_TF_DELETED] = t.deleted_at;
  o.time_kind = t.time_kind;
  o.text = nullptr;
  return o;
}

// planes: uint32 [26][n_docs] — counts, replaced by their exclusive prefixes; status = {pie_status, document}
extern "C" void ingest_host_measure(const uint8_t* text, const int64_t* offsets, int64_t n_docs, uint32_t* planes,
                                    uint8_t* doc_status, int64_t* totals, int32_t* status) {
  const Pow5Table pow5{kPow5};
  IngestOut none;
  memset(&none, 0, sizeof(none));
  status[0] = 0;
  status[1] = -1;
  for (int64_t s = 0; s < n_docs; ++s) {
    uint32_t cnt[kPlanes] = {0};
    DocWalker<false> w;
    w.begin(text, offsets[s], offsets[s + 1], s);
    int r;  // this pass by members, the other by tokens: both drivers of the walk are exercised, and must agree
    while ((r = w.step_member(cnt, none, pow5)) == kDocRunning) {}
    if (r != kDocOk) {
      memset(cnt, 0, sizeof(cnt));
      if (r != kDocDropped && status[0] == 0) { status[0] = -r; status[1] = (int32_t)s; }
    }
    doc_status[s] = r == kDocOk ? 0 : 1;
    for (int p = 0; p < kPlanes; ++p) planes[p * n_docs + s] = cnt[p];
  }
  for (int p = 0; p < kPlanes; ++p) {
    uint64_t run = 0;
    for (int64_t s = 0; s < n_docs; ++s) {
      const uint32_t v = planes[p * n_docs + s];
      planes[p * n_docs + s] = (uint32_t)run;
      run += v;
    }
    totals[p] = (int64_t)run;
  }
}

extern "C" void ingest_host_fill(const uint8_t* text, const int64_t* offsets, int64_t n_docs, const uint32_t* planes,
                                 const uint8_t* doc_status, const pie_archive_table* table) {
  const Pow5Table pow5{kPow5};
  IngestOut out = make_out(*table);
  out.text = text;
  if (n_docs == 0) {
    for (int h = 0; h < kHeaps; ++h) out.off[h][0] = 0;
    out.entry_offsets[0] = 0;
    out.crew_list[0] = 0;
    out.actions_list[0] = 0;
    return;
  }
  for (int64_t s = 0; s < n_docs; ++s) {
    uint32_t cnt[kPlanes];
    for (int p = 0; p < kPlanes; ++p) cnt[p] = planes[p * n_docs + s];
    for (int h = 0; h < 7; ++h) out.off[h][s] = (int32_t)cnt[h];
    out.entry_offsets[s] = (int32_t)cnt[kPlaneEntries];
    out.crew_list[s] = (int32_t)cnt[kPlaneCrewItems];
    out.created_at[s] = jw_nan();
    out.archived_at[s] = jw_nan();
    if (out.time_val[PIE_TF_UPDATED]) out.time_val[PIE_TF_UPDATED][s] = jw_nan();
    if (out.time_val[PIE_TF_DELETED]) out.time_val[PIE_TF_DELETED][s] = jw_nan();
    if (out.time_kind) memset(out.time_kind + s * PIE_TF_COUNT, 0, PIE_TF_COUNT);
    if (doc_status[s] == 0) {
      DocWalker<true> w;
      w.begin(text, offsets[s], offsets[s + 1], s);
      while (w.step(cnt, out, pow5) == kDocRunning) {}
    }
    if (s == n_docs - 1) {
      for (int h = 0; h < 7; ++h) out.off[h][n_docs] = (int32_t)cnt[h];
      out.entry_offsets[n_docs] = (int32_t)cnt[kPlaneEntries];
      out.crew_list[n_docs] = (int32_t)cnt[kPlaneCrewItems];
      out.off[kHeapCrew][cnt[kPlaneCrewItems]] = (int32_t)cnt[kHeapCrew];
      const uint32_t rows = cnt[kPlaneEntries];
      for (int h = kHeapEntry0; h < kHeapEntry0 + 14; ++h) out.off[h][rows] = (int32_t)cnt[h];
      out.actions_list[rows] = (int32_t)cnt[kPlaneActionItems];
      out.off[kHeapActions][cnt[kPlaneActionItems]] = (int32_t)cnt[kHeapActions];
    }
  }
}
