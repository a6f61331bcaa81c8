// Host build of sph_pie_b200/csrc/pie_numparse.cuh (the same code the ingest kernels run), so that the number
// parser can be checked on the CPU against Python's float().  Test-only.
#include <stdint.h>

#include "../../sph_pie_b200/csrc/pie_numparse.cuh"

static const uint64_t kPow5[PIE_POW5_128_N][2] = PIE_POW5_128_INIT;

// texts: concatenated number strings, offsets[n+1]; out: value bits, status (0 ok, 1 syntax, 2 undecided), bytes used
extern "C" void numparse_host_batch(const uint8_t* texts, const int64_t* offsets, int64_t n, double* values,
                                    int32_t* status, int64_t* used) {
  pie::Pow5Table tab{kPow5};
  for (int64_t i = 0; i < n; ++i) {
    values[i] = 0;
    used[i] = 0;
    status[i] = pie::parse_json_number(texts + offsets[i], offsets[i + 1] - offsets[i], tab, &values[i], &used[i]);
  }
}
