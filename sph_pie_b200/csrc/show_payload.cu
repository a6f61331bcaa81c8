// The schemaVersion 2 show payload on sm_100a: JSON.stringify of the object dispatchShowEvent builds for every event
// but 'show.archived' (reference server/webhookDispatcher.js:545-584), one document per show of a batch.  The device
// code — the document's grammar, the emitter, the shared-memory stage of a show — is pie_show_payload.cuh (which also
// runs on the CPU in tests/native/payload_host.cpp); here are the kernels and the launcher:
//   payload_measure_kernel (a warp per show: the document's length) -> cub::DeviceScan::ExclusiveSum ->
//   payload_finish_kernel -> payload_write_kernel (a warp per show: the bytes).
// This is a row of the scope table's "next" section (the last bulk pure function of the reference's dispatcher), not
// part of the timed step.
#include <cub/device/device_scan.cuh>

#include "pie_kernels.h"
#include "pie_show_payload.cuh"

namespace pie {

namespace {

using sp::PayloadArgs;

constexpr int kWarpsPerCta = 4;

__global__ void __launch_bounds__(32 * kWarpsPerCta) payload_measure_kernel(PayloadArgs a, int64_t* __restrict__ doc_len,
                                                                           int32_t* __restrict__ status) {
  __shared__ sp::WarpStage stage[kWarpsPerCta];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t s = (int64_t)blockIdx.x * kWarpsPerCta + warp;
  if (s >= a.v.n_shows) return;
  int schema_error = 0;
  const bool staged = sp::stage_show(stage[warp], a.v, s, lane);
  const uint64_t n = sp::emit_document<false>(a, staged ? stage[warp].v : a.v, staged ? &stage[warp] : nullptr, s, nullptr, lane,
                                              &schema_error);
  if (lane == 0) {
    doc_len[s] = (int64_t)n;
    if (schema_error) atomicMin(reinterpret_cast<unsigned int*>(status + 1), (unsigned int)s);
  }
}

__global__ void __launch_bounds__(32 * kWarpsPerCta) payload_write_kernel(PayloadArgs a, const int64_t* __restrict__ doc_offsets,
                                                                         uint8_t* __restrict__ out, uint64_t capacity) {
  __shared__ sp::WarpStage stage[kWarpsPerCta];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t s = (int64_t)blockIdx.x * kWarpsPerCta + warp;
  if (s >= a.v.n_shows) return;
  const int64_t at = doc_offsets[s], end = doc_offsets[s + 1];
  if ((uint64_t)end > capacity) return;  // nothing past the caller's buffer
  int schema_error = 0;
  const bool staged = sp::stage_show(stage[warp], a.v, s, lane);
  sp::emit_document<true>(a, staged ? stage[warp].v : a.v, staged ? &stage[warp] : nullptr, s, out + at, lane, &schema_error);
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
