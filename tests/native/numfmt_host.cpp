// Host build of sph_pie_b200/csrc/pie_numfmt.cuh (the same code the export kernels run), so that
// Number::toString can be checked on the CPU against Python's repr / the oracle.  Test-only.
#include <stdint.h>

#include "../../sph_pie_b200/csrc/pie_numfmt.cuh"

static const uint64_t kInv[PIE_RYU_POW5_INV_SPLIT_N][2] = PIE_RYU_POW5_INV_SPLIT_INIT;
static const uint64_t kPow[PIE_RYU_POW5_SPLIT_N][2] = PIE_RYU_POW5_SPLIT_INIT;

extern "C" void numfmt_host_batch(const double* x, int64_t n, char* out, int32_t* lens) {
  pie::RyuTables t{kInv, kPow};
  for (int64_t i = 0; i < n; ++i) lens[i] = pie::js_number_to_string(x[i], out + i * pie::kMaxNumberChars, t);
}
