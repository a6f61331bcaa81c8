// Host build of the schemaVersion 2 show payload's DEVICE code (sph_pie_b200/csrc/pie_show_payload.cuh), compiled by g++
// against tests/native/cuda_shim/cuda_runtime.h, its 32 lanes run as 32 fibers that meet at every warp collective (the
// scheduler of tests/native/fast_host.cpp), driven the way show_payload.cu drives it: a warp per show stages the show
// (or not: `use_stage` 0 emits through the caller's view), measures,
// then — after the exclusive sum of the lengths — writes.  Test-only: holds the staged, rebased view and the fast paths
// of the emitter to the oracle on the CPU (tests/test_payload_oracles_cpu.py).  The product has no such path.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include "../../sph_pie_b200/csrc/pie_show_payload.cuh"

using namespace pie;

// ---- 32 lanes as fibers (as in fast_host.cpp) -----------------------------------------------------------------------
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
  if (++g_arrived == kLanes) {
    g_arrived = 0;
    ++g_generation;
    return;
  }
  while (g_generation == gen) swapcontext(&g_fiber[lane], &g_main);
}

// ---- the driver -------------------------------------------------------------------------------------------------------
namespace {

struct Job {
  sp::PayloadArgs a;
  int64_t s;
  uint8_t* out;  // nullptr: measure
  int use_stage;
  uint64_t len[kLanes];
  int staged[kLanes];
  int schema_error[kLanes];
};
Job g_job;
sp::WarpStage g_stage;

void lane_body(int lane) {
  Job& j = g_job;
  int schema_error = 0;
  bool staged = false;
  if (j.use_stage) staged = sp::stage_show(g_stage, j.a.v, j.s, lane);
  const pie_archive_view& v = staged ? g_stage.v : j.a.v;
  const sp::WarpStage* st = staged ? &g_stage : nullptr;
  j.len[lane] = j.out ? sp::emit_document<true>(j.a, v, st, j.s, j.out, lane, &schema_error)
                      : sp::emit_document<false>(j.a, v, st, j.s, nullptr, lane, &schema_error);
  j.staged[lane] = staged;
  j.schema_error[lane] = schema_error;
}

}  // namespace

// doc_offsets: int64 [n_shows + 1]; out == nullptr: lengths only.  staged_shows receives how many shows went through the
// stage.  Returns 0, -1 when the lanes of a warp disagree, -2 when the two passes disagree on a length, -3 when a write
// left the document's bytes; status[2] = {PIE_ERR_SCHEMA or 0, first offending show}.
extern "C" int payload_host(const pie_archive_view* v, const uint8_t* head, int head_len, const uint8_t* tail, int tail_len,
                            int64_t* doc_offsets, uint8_t* out, uint64_t capacity, int use_stage, int64_t* staged_shows,
                            int32_t* status) {
  memset(&g_job, 0, sizeof g_job);
  g_job.a = sp::PayloadArgs{*v, head, head_len, tail, tail_len};
  g_job.use_stage = use_stage;
  status[0] = 0;
  status[1] = -1;
  *staged_shows = 0;
  doc_offsets[0] = 0;
  for (int64_t s = 0; s < v->n_shows; ++s) {
    g_job.s = s;
    g_job.out = nullptr;
    memset(&g_stage, 0xA5, sizeof g_stage);  // nothing may depend on what an earlier show left in the stage
    run_warp(lane_body);
    for (int l = 1; l < kLanes; ++l)
      if (g_job.len[l] != g_job.len[0] || g_job.staged[l] != g_job.staged[0]) return -1;
    doc_offsets[s + 1] = doc_offsets[s] + (int64_t)g_job.len[0];
    *staged_shows += g_job.staged[0];
    for (int l = 0; l < kLanes; ++l)
      if (g_job.schema_error[l] && status[0] == 0) { status[0] = PIE_ERR_SCHEMA; status[1] = (int32_t)s; }
  }
  if (!out) return 0;
  if ((uint64_t)doc_offsets[v->n_shows] > capacity) return 0;
  for (int64_t s = 0; s < v->n_shows; ++s) {
    g_job.s = s;
    g_job.out = out + doc_offsets[s];
    memset(&g_stage, 0x5A, sizeof g_stage);
    run_warp(lane_body);
    for (int l = 0; l < kLanes; ++l)
      if ((int64_t)g_job.len[l] != doc_offsets[s + 1] - doc_offsets[s]) return -2;
  }
  return 0;
}

extern "C" int payload_host_stage_words(void) { return sp::kStageWords; }
