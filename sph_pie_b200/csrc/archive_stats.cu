// Per-show archive statistics on sm_100a.
// Replaces computeArchiveShowStats (reference public/app.js:3898-3953).
//
// Two kernels:
//   classify_entries_kernel  entry-parallel: reads the status / launched / primaryIssue strings and
//                            delaySec of each entry, writes one classification byte per entry.
//   reduce_shows_kernel      show-parallel: walks each show's codes in entry order, counts, sums
//                            delaySec LEFT TO RIGHT (bit-exact with Array.prototype.reduce), derives
//                            the rates and writes the plane-major statistics table.
#include "pie_device.cuh"
#include "pie_kernels.h"

namespace pie {

unsigned long long g_launches = 0;

// PRIMARY_ISSUES.includes(issue) ? index : 'Other'   (public/app.js:3923).  Returns 1..10.
__device__ __forceinline__ uint32_t issue_code(const uint8_t* __restrict__ s, int n) {
  switch (n) {
    case 13:
      if (equals_exact(s, n, "Tracking lost")) return 1;
      if (equals_exact(s, n, "Command delay")) return 3;
      if (equals_exact(s, n, "Motor or prop")) return 6;
      if (equals_exact(s, n, "Sensor or IMU")) return 7;
      break;
    case 16:
      if (equals_exact(s, n, "Failed to launch")) return 2;
      break;
    case 7:
      if (equals_exact(s, n, "RF link")) return 4;
      if (equals_exact(s, n, "Battery")) return 5;
      break;
    case 24:
      if (equals_exact(s, n, "Software or show control")) return 8;
      break;
    case 14:
      if (equals_exact(s, n, "Operator input")) return 9;
      break;
    default:
      break;
  }
  return 10;  // 'Other' (also the literal "Other")
}

__device__ __forceinline__ uint32_t classify_entry(const pie_archive_view& v, int64_t i) {
  uint32_t code = 0;
  {  // status: String(entry?.status || '').toLowerCase()
    int b = v.status.offsets[i], e = v.status.offsets[i + 1];
    const uint8_t* s = v.status.data + b;
    int n = e - b;
    if (equals_lower_ascii(s, n, "completed")) code = 1;
    else if (equals_lower_ascii(s, n, "no-launch")) code = 2;
    else if (equals_lower_ascii(s, n, "abort")) code = 3;
  }
  {  // launched
    int b = v.launched.offsets[i], e = v.launched.offsets[i + 1];
    if (equals_lower_ascii(v.launched.data + b, e - b, "yes")) code |= kLaunchedBit;
  }
  {  // primaryIssue.trim()
    int b = v.primary_issue.offsets[i], e = v.primary_issue.offsets[i + 1];
    const uint8_t* s = v.primary_issue.data;
    while (b < e) {
      int l = js_ws_len_at(s, b, e);
      if (!l) break;
      b += l;
    }
    while (e > b) {
      int l = js_ws_len_before(s, b, e);
      if (!l) break;
      e -= l;
    }
    if (e > b) code |= issue_code(s + b, e - b) << kIssueShift;
  }
  if (v.delay_valid[i] && is_finite_f64(v.delay_sec[i])) code |= kDelayBit;
  return code;
}

__global__ void __launch_bounds__(256) classify_entries_kernel(pie_archive_view v, uint8_t* __restrict__ codes) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < v.n_entries; i += stride)
    codes[i] = (uint8_t)classify_entry(v, i);
}

__global__ void __launch_bounds__(256) reduce_shows_kernel(pie_archive_view v, const uint8_t* __restrict__ codes,
                                                           int32_t* __restrict__ si, double* __restrict__ sf,
                                                           int64_t stride) {
  const int64_t gstride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < v.n_shows; s += gstride) {
    const int e0 = v.entry_offsets[s], e1 = v.entry_offsets[s + 1];
    int completed = 0, no_launch = 0, abort_n = 0, launched = 0, delay_n = 0;
    int issue_n[PIE_N_ISSUES], issue_first[PIE_N_ISSUES];
#pragma unroll
    for (int k = 0; k < PIE_N_ISSUES; ++k) { issue_n[k] = 0; issue_first[k] = -1; }
    double sum = 0.0, mx = 0.0;
    for (int e = e0; e < e1; ++e) {
      const uint32_t c = codes[e];
      const uint32_t st = c & kStatusMask;
      completed += (st == 1);
      no_launch += (st == 2);
      abort_n += (st == 3);
      launched += (c & kLaunchedBit) != 0;
      const uint32_t iss = (c >> kIssueShift) & kIssueMask;
#pragma unroll
      for (int k = 0; k < PIE_N_ISSUES; ++k) {
        if (iss == (uint32_t)(k + 1)) {
          if (issue_n[k] == 0) issue_first[k] = e - e0;
          issue_n[k] += 1;
        }
      }
      if (c & kDelayBit) {
        const double d = v.delay_sec[e];
        sum = sum + d;  // left to right, initial 0 (public/app.js:3928)
        mx = delay_n ? js_max(mx, d) : d;
        delay_n += 1;
      }
    }
    const int total = e1 - e0;
    const double nan = quiet_nan();
    si[PIE_SI_TOTAL * stride + s] = total;
    si[PIE_SI_COMPLETED * stride + s] = completed;
    si[PIE_SI_NO_LAUNCH * stride + s] = no_launch;
    si[PIE_SI_ABORT * stride + s] = abort_n;
    si[PIE_SI_LAUNCHED * stride + s] = launched;
    si[PIE_SI_DELAY_COUNT * stride + s] = delay_n;
    sf[PIE_SF_DELAY_SUM * stride + s] = sum;
    sf[PIE_SF_AVG_DELAY * stride + s] = delay_n ? sum / (double)delay_n : nan;
    sf[PIE_SF_MAX_DELAY * stride + s] = delay_n ? mx : nan;
    const double dt = (double)total;
    sf[PIE_SF_COMPLETION_RATE * stride + s] = total ? ((double)completed / dt) * 100.0 : nan;
    sf[PIE_SF_LAUNCH_RATE * stride + s] = total ? ((double)launched / dt) * 100.0 : nan;
    sf[PIE_SF_ABORT_RATE * stride + s] = total ? ((double)abort_n / dt) * 100.0 : nan;
#pragma unroll
    for (int k = 0; k < PIE_N_ISSUES; ++k) {
      si[(PIE_SI_ISSUE_COUNT0 + k) * stride + s] = issue_n[k];
      si[(PIE_SI_ISSUE_FIRST0 + k) * stride + s] = issue_first[k];
      sf[(PIE_SF_ISSUE_RATE0 + k) * stride + s] = total ? ((double)issue_n[k] / dt) * 100.0 : nan;
    }
  }
}

static inline int grid_for(int64_t n, int block, int sm_count, int waves_cap) {
  int64_t blocks = (n + block - 1) / block;
  int64_t cap = (int64_t)sm_count * waves_cap;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

cudaError_t launch_show_stats(const pie_archive_view& v, int32_t* si, double* sf, int64_t stride, void* scratch,
                              int sm_count, cudaStream_t stream) {
  uint8_t* codes = static_cast<uint8_t*>(scratch);
  if (v.n_entries > 0) {
    // grid-stride, 8 resident CTAs of 256 threads per SM
    classify_entries_kernel<<<grid_for(v.n_entries, 256, sm_count, 8), 256, 0, stream>>>(v, codes);
    g_launches += 1;
  }
  if (v.n_shows > 0) {
    reduce_shows_kernel<<<grid_for(v.n_shows, 256, sm_count, 8), 256, 0, stream>>>(v, codes, si, sf, stride);
    g_launches += 1;
  }
  return cudaGetLastError();
}

}  // namespace pie
