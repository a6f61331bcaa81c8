// Offsets arrays of a batch the host handed in (pie_strcol.offsets, list_offsets): checked ON THE DEVICE, after the
// upload, before any kernel walks them — `first <= off[i] <= off[i+1] <= last` for every row.  The check is one
// streaming pass at HBM speed (the same pass on the host costs ~15 ms per 11 M entries and competes with the PCIe
// transfers for the host's memory bandwidth).  An array that fails is made harmless in place — every row becomes the
// empty string at `first` — so the kernels that follow on the stream read nothing outside the staged heaps; the host
// finds the flag at its next synchronisation point and fails the call with PIE_ERR_INVALID_ARG.
#include "pie_kernels.h"

namespace pie {

namespace {

__global__ void __launch_bounds__(256) offsets_check_kernel(const __grid_constant__ OffsetsBatch batch, int32_t* __restrict__ flags) {
  const OffsetsArray a = batch.a[blockIdx.y];
  int bad = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t x = a.off[i], y = a.off[i + 1];
    bad |= (y < x) | (x < a.first) | (y > a.last);
  }
  if (__any_sync(0xFFFFFFFFu, bad) && (threadIdx.x & 31) == 0) atomicOr(flags + blockIdx.y, 1);
}

__global__ void __launch_bounds__(256) offsets_disarm_kernel(const __grid_constant__ OffsetsBatch batch, const int32_t* __restrict__ flags) {
  if (flags[blockIdx.y] == 0) return;  // the usual case: nothing to do
  const OffsetsArray a = batch.a[blockIdx.y];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= a.n; i += (int64_t)gridDim.x * blockDim.x)
    const_cast<int32_t*>(a.off)[i] = a.first;
}

}  // namespace

namespace {
__global__ void __launch_bounds__(256) rebase_i32_kernel(int32_t* __restrict__ dst, const int32_t* __restrict__ src, int64_t n, int32_t add) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i] + add;
}
}  // namespace

namespace {
// the flags, stored where the host can read them without a copy engine (mapped pinned memory)
__global__ void offsets_flags_out_kernel(const int32_t* __restrict__ flags, int32_t* __restrict__ mapped_host) {
  if (threadIdx.x < kMaxOffsetsArrays) mapped_host[threadIdx.x] = flags[threadIdx.x];
}
}  // namespace

cudaError_t launch_offsets_flags_out(const int32_t* flags, int32_t* mapped_host, cudaStream_t stream) {
  offsets_flags_out_kernel<<<1, kMaxOffsetsArrays, 0, stream>>>(flags, mapped_host);
  g_launches += 1;
  return cudaGetLastError();
}

// dst[i] = src[i] + add: the offsets of a chunk's column moved onto the joined column of a batch
cudaError_t launch_rebase_i32(int32_t* dst, const int32_t* src, int64_t n, int32_t add, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count_or_default() * 8;
  rebase_i32_kernel<<<(unsigned)(blocks > cap ? cap : blocks), 256, 0, stream>>>(dst, src, n, add);
  g_launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_offsets_check(const OffsetsBatch& batch, int32_t* flags, cudaStream_t stream) {
  if (batch.count <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(flags, 0, sizeof(int32_t) * kMaxOffsetsArrays, stream);
  if (e != cudaSuccess) return e;
  int64_t longest = 0;
  for (int i = 0; i < batch.count; ++i) longest = batch.a[i].n > longest ? batch.a[i].n : longest;
  int64_t blocks = (longest + 256 * 8 - 1) / (256 * 8);
  const int64_t cap = (int64_t)sm_count_or_default() * 4;
  blocks = blocks < 1 ? 1 : (blocks > cap ? cap : blocks);
  const dim3 grid((unsigned)blocks, (unsigned)batch.count);
  offsets_check_kernel<<<grid, 256, 0, stream>>>(batch, flags);
  offsets_disarm_kernel<<<grid, 256, 0, stream>>>(batch, flags);
  g_launches += 2;
  return cudaGetLastError();
}

}  // namespace pie
