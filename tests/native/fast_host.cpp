// Host build of the warp-per-document ingest (sph_pie_b200/csrc/pie_json_fast.cuh) — the DEVICE code itself, compiled
// by g++ against tests/native/cuda_shim/cuda_runtime.h, its 32 lanes run as 32 fibers that meet at every warp
// collective — driven the way json_ingest.cu drives it: pass 1 with the lists nearly every document fits, then the
// roomy ones, then the thread-per-document walk (pie_json_walk.cuh) for what both declined; exclusive scan of the 26
// counts; pass 2 by route (records / parsed again / the walk).  Test-only: lets the recogniser of the warp path be held
// to the oracle on the CPU (tests/test_ingest_cpu.py).  The product has no such path.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <vector>

#include "../../sph_pie_b200/csrc/pie_json_fast.cuh"

using namespace pie;
using namespace pie::jw;

// ---- 32 lanes as fibers -------------------------------------------------------------------------------------------
uint3 threadIdx;
unsigned long long pie_warp_slot[2][32];
int pie_warp_parity[32];

namespace {

constexpr int kLanes = 32;
constexpr size_t kStack = 1 << 20;
ucontext_t g_main, g_fiber[kLanes];
char* g_stacks = nullptr;
bool g_done[kLanes];
int g_arrived = 0;
unsigned g_generation = 0;
int g_current = 0;
void (*g_body)(int lane) = nullptr;

void fiber_entry(int lane) {
  g_body(lane);
  g_done[lane] = true;
  swapcontext(&g_fiber[lane], &g_main);
}

// runs body(lane) on the 32 lanes until every lane has returned
void run_warp(void (*body)(int)) {
  if (!g_stacks) g_stacks = (char*)malloc(kStack * kLanes);
  g_body = body;
  g_arrived = 0;
  for (int l = 0; l < kLanes; ++l) {
    pie_warp_parity[l] = 0;
    g_done[l] = false;
    getcontext(&g_fiber[l]);
    g_fiber[l].uc_stack.ss_sp = g_stacks + kStack * l;
    g_fiber[l].uc_stack.ss_size = kStack;
    g_fiber[l].uc_link = &g_main;
    makecontext(&g_fiber[l], (void (*)())fiber_entry, 1, l);
  }
  for (;;) {
    bool any = false;
    for (int l = 0; l < kLanes; ++l) {
      if (g_done[l]) continue;
      any = true;
      g_current = l;
      threadIdx.x = (unsigned)l;
      swapcontext(&g_main, &g_fiber[l]);
    }
    if (!any) break;
  }
}

}  // namespace

void pie_warp_barrier() {
  const int lane = g_current;
  const unsigned gen = g_generation;
  if (++g_arrived == kLanes) {  // the last lane to arrive lets everybody go
    g_arrived = 0;
    ++g_generation;
    return;
  }
  while (g_generation == gen) {
    swapcontext(&g_fiber[lane], &g_main);  // the scheduler resumes the next lane; it sets threadIdx for us when we return
  }
}

// ---- the driver ---------------------------------------------------------------------------------------------------
namespace {

const uint64_t kPow5[PIE_POW5_128_N][2] = PIE_POW5_128_INIT;

IngestOut make_out(const pie_archive_table& t) {
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

// what a pass hands to the lanes of the current document
struct Job {
  const uint8_t* text;
  int64_t from, to, s, n_docs;
  uint32_t* planes_row;
  IngestOut out;
  jf::RecCtx rc;
  jf::TablePointers tp;
  int mode;    // 0 pass 1 small, 1 pass 1 big, 2 pass 2 small, 3 pass 2 big, 4 pass 2 records
  int result;  // pass 1: the route lane 0 saw
  int lane_results[kLanes];
};
Job g_job;
jf::WarpShared<jf::CapsSmall> g_ws_small;
jf::WarpShared<jf::CapsBig> g_ws_big;

void lane_body(int lane) {
  const Pow5Table pow5{kPow5};
  Job& j = g_job;
  int r = 0;
  if (j.mode == 0) r = jf::fast_doc<false, jf::CapsSmall>(g_ws_small, j.tp, j.text, j.from, j.to, j.s, j.n_docs, j.planes_row, j.out, pow5, j.rc);
  else if (j.mode == 1) r = jf::fast_doc<false, jf::CapsBig>(g_ws_big, j.tp, j.text, j.from, j.to, j.s, j.n_docs, j.planes_row, j.out, pow5, j.rc);
  else if (j.mode == 2) r = jf::fast_doc<true, jf::CapsSmall>(g_ws_small, j.tp, j.text, j.from, j.to, j.s, j.n_docs, j.planes_row, j.out, pow5, j.rc);
  else if (j.mode == 3) r = jf::fast_doc<true, jf::CapsBig>(g_ws_big, j.tp, j.text, j.from, j.to, j.s, j.n_docs, j.planes_row, j.out, pow5, j.rc);
  else jf::fill_records(g_ws_small, j.tp, j.text, j.from, j.to, j.s, j.planes_row, j.out, j.rc, pow5);
  j.lane_results[lane] = r;
}

int run_job(int mode) {
  g_job.mode = mode;
  run_warp(lane_body);
  for (int l = 1; l < kLanes; ++l)
    if (g_job.lane_results[l] != g_job.lane_results[0]) return -1000;  // the lanes must agree on what became of the document
  return g_job.lane_results[0];
}

// kept between the two passes, as the scratch area of pie_ingest_measure_dev is
std::vector<uint8_t> g_route;
std::vector<unsigned long long> g_pool;
std::vector<jf::DocRec> g_doc_rec;
unsigned long long g_cursor = 0;

void init_job(const uint8_t* text, int64_t n_docs, const IngestOut& out, bool pool) {
  memset(&g_job, 0, sizeof(g_job));
  g_job.text = text;
  g_job.n_docs = n_docs;
  g_job.out = out;
  bool clash = false;
  g_job.tp.keys = jf::make_key_tables(&clash);
  for (int h = 0; h < kHeaps; ++h) { g_job.tp.data[h] = out.data[h]; g_job.tp.off[h] = out.off[h]; }
  g_job.rc.pool = pool ? g_pool.data() : nullptr;
  g_job.rc.cursor = &g_cursor;
  g_job.rc.capacity = g_pool.size();
  g_job.rc.doc_rec = g_doc_rec.data();
}

}  // namespace

// rows: uint32 [n_docs][26] — counts, replaced by their exclusive prefixes; status = {pie_status, document};
// routes: uint8 [n_docs] (jf::kRoute*); pool_units_per_doc: 0 = no records (pass 2 parses every document again)
extern "C" int fast_host_measure(const uint8_t* text, const int64_t* offsets, int64_t n_docs, uint32_t* rows, uint8_t* doc_status,
                                 int64_t* totals, int32_t* status, uint8_t* routes, int pool_units_per_doc) {
  const Pow5Table pow5{kPow5};
  g_route.assign((size_t)n_docs + 1, 0);
  g_pool.assign((size_t)pool_units_per_doc * (size_t)(n_docs > 0 ? n_docs : 1), 0ull);
  g_doc_rec.assign((size_t)n_docs + 1, jf::DocRec{});
  g_cursor = 0;
  IngestOut none;
  memset(&none, 0, sizeof(none));
  init_job(text, n_docs, none, pool_units_per_doc > 0);
  status[0] = 0;
  status[1] = -1;
  for (int64_t s = 0; s < n_docs; ++s) {
    uint32_t* row = rows + s * kPlanes;
    memset(row, 0, sizeof(uint32_t) * kPlanes);
    g_job.from = offsets[s];
    g_job.to = offsets[s + 1];
    g_job.s = s;
    g_job.planes_row = row;
    int route = jf::kRouteSlow;
    if (g_job.to - g_job.from + 31 > jf::kFastMaxBytes) {
      route = jf::kRouteLong;
    } else {
      route = run_job(0);
      if (route < 0) return -1;
      if (route == jf::kRouteSlow) {
        memset(row, 0, sizeof(uint32_t) * kPlanes);
        route = run_job(1);
        if (route < 0) return -1;
        if (route == jf::kRouteFast) route = jf::kRouteFastBig;
      }
    }
    if (route == jf::kRouteSlow || route == jf::kRouteLong) {
      uint32_t cnt[kPlanes] = {0};
      DocWalker<false> w;
      w.begin(text, offsets[s], offsets[s + 1], s);
      int r;
      while ((r = w.step_member(cnt, none, pow5)) == kDocRunning) {}
      if (r != kDocOk) {
        memset(cnt, 0, sizeof(cnt));
        if (r != kDocDropped && status[0] == 0) { status[0] = -r; status[1] = (int32_t)s; }
      }
      doc_status[s] = r == kDocOk ? 0 : 1;
      memcpy(row, cnt, sizeof(cnt));
    } else {
      doc_status[s] = 0;
    }
    g_route[s] = (uint8_t)route;
    routes[s] = (uint8_t)route;
  }
  for (int p = 0; p < kPlanes; ++p) {
    uint64_t run = 0;
    for (int64_t s = 0; s < n_docs; ++s) {
      const uint32_t v = rows[s * kPlanes + p];
      rows[s * kPlanes + p] = (uint32_t)run;
      run += v;
    }
    totals[p] = (int64_t)run;
  }
  return 0;
}

extern "C" int fast_host_fill(const uint8_t* text, const int64_t* offsets, int64_t n_docs, uint32_t* rows,
                              const uint8_t* doc_status, const pie_archive_table* table) {
  const Pow5Table pow5{kPow5};
  IngestOut out = make_out(*table);
  out.text = text;
  if (n_docs == 0) {
    for (int h = 0; h < kHeaps; ++h) out.off[h][0] = 0;
    out.entry_offsets[0] = 0;
    out.crew_list[0] = 0;
    out.actions_list[0] = 0;
    return 0;
  }
  init_job(text, n_docs, out, true);
  for (int64_t s = 0; s < n_docs; ++s) {
    uint32_t* row = rows + s * kPlanes;
    g_job.from = offsets[s];
    g_job.to = offsets[s + 1];
    g_job.s = s;
    g_job.planes_row = row;
    const int route = g_route[s];
    if (route == jf::kRouteRecords) {
      if (run_job(4) < -999) return -1;
    } else if (route == jf::kRouteFast) {
      if (run_job(2) != jf::kRouteFast) return -2;
    } else if (route == jf::kRouteFastBig) {
      if (run_job(3) != jf::kRouteFast) return -3;
    } else {
      uint32_t cnt[kPlanes];
      memcpy(cnt, row, sizeof(cnt));
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
        const uint32_t nrows = cnt[kPlaneEntries];
        for (int h = kHeapEntry0; h < kHeapEntry0 + 14; ++h) out.off[h][nrows] = (int32_t)cnt[h];
        out.actions_list[nrows] = (int32_t)cnt[kPlaneActionItems];
        out.off[kHeapActions][cnt[kPlaneActionItems]] = (int32_t)cnt[kHeapActions];
      }
    }
  }
  return 0;
}
