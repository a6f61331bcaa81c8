// The provider's archive maintenance, as far as it is arithmetic (reference server/storage/sqlProvider.js):
//   _getTimestamp (:970-985)            what a stored document's createdAt / updatedAt / archivedAt / deletedAt count as
//   _archiveDailyShows (:758-816)       which rows of `shows` are archived now: date groups older than 12 h
//   _purgeExpiredArchives (:863-890)    which rows of `show_archive` are past their two months (_addMonths :999-1009)
// on the columnar table the JSON ingest builds.  Small kernels, a thread per show; the value is the semantics —
// Number(null) === 0, Date.setMonth's day overflow, Map insertion order — not the speed.
#include "pie_device.cuh"
#include "pie_kernels.h"
#include "pie_numparse.cuh"

namespace pie {

namespace {

__device__ const uint64_t g_pow5_maint[PIE_POW5_128_N][2] = PIE_POW5_128_INIT;

constexpr double kMaxTime = 8.64e15;  // ECMA-262 TimeClip
constexpr int64_t kDayMs = 86400000LL;

// ---- StringToNumber (ECMA-262 7.1.4.1.1) -----------------------------------------------------------------------
// 1 = a number (finite or not) in *out, 0 = NaN, -1 = a form this kernel does not decide (reported, never guessed):
// a literal longer than 40 characters, a radix literal of more than 64 bits, a number the parser cannot round.
__device__ int string_to_number(const uint8_t* s, int n, double* out) {
  int b = 0, e = n;
  while (b < e) { const int l = js_ws_len_at(s, b, e); if (!l) break; b += l; }
  while (e > b) { const int l = js_ws_len_before(s, b, e); if (!l) break; e -= l; }
  if (b == e) { *out = 0.0; return 1; }  // '' and all-white-space are 0
  if (e - b > 2 && s[b] == '0') {        // 0x / 0o / 0b: no sign, integer digits only
    const int p = s[b + 1] | 0x20;
    const int bits = p == 'x' ? 4 : p == 'o' ? 3 : p == 'b' ? 1 : 0;
    if (bits) {
      uint64_t v = 0;
      int used = 0;
      for (int i = b + 2; i < e; ++i) {
        int c = s[i], d;
        if (c >= '0' && c <= '9') d = c - '0';
        else { c |= 0x20; d = (c >= 'a' && c <= 'f') ? c - 'a' + 10 : 99; }
        if (d >= (1 << bits)) return 0;  // not a digit of this radix: NaN
        if (v == 0 && d == 0) continue;
        used += bits;
        if (used > 64) return -1;
        v = (v << bits) | (uint64_t)d;
      }
      *out = (double)v;  // round to nearest even: what the exact integer rounds to
      return 1;
    }
  }
  bool neg = false;
  int i = b;
  if (s[i] == '+' || s[i] == '-') { neg = s[i] == '-'; ++i; }
  if (e - i == 8 && s[i] == 'I') {  // Infinity
    const char* lit = "Infinity";
    for (int k = 0; k < 8; ++k)
      if (s[i + k] != (uint8_t)lit[k]) return 0;
    *out = neg ? -__longlong_as_double(0x7ff0000000000000LL) : __longlong_as_double(0x7ff0000000000000LL);
    return 1;
  }
  // StrUnsignedDecimalLiteral -> the JSON number grammar the parser knows: leading zeros dropped, a bare '.5' / '5.'
  // completed.  Anything that is not digits / one '.' / an exponent is NaN.
  uint8_t buf[48];
  int m = 0;
  if (e - i > 40 || e - i == 0) return e - i == 0 ? 0 : -1;
  int int_digits = 0, frac_digits = 0;
  bool seen_nonzero = false;
  int j = i;
  for (; j < e && s[j] >= '0' && s[j] <= '9'; ++j) {
    ++int_digits;
    if (s[j] != '0') seen_nonzero = true;
    if (seen_nonzero) buf[m++] = s[j];
  }
  if (m == 0) buf[m++] = '0';
  if (j < e && s[j] == '.') {
    ++j;
    const int dot = m;
    buf[m++] = '.';
    for (; j < e && s[j] >= '0' && s[j] <= '9'; ++j) { buf[m++] = s[j]; ++frac_digits; }
    if (frac_digits == 0) m = dot;  // '5.' is 5
  }
  if (int_digits + frac_digits == 0) return 0;  // '.', 'e5', 'abc'
  if (j < e && (s[j] == 'e' || s[j] == 'E')) {
    buf[m++] = 'e';
    ++j;
    if (j < e && (s[j] == '+' || s[j] == '-')) buf[m++] = s[j++];
    const int before = m;
    for (; j < e && s[j] >= '0' && s[j] <= '9'; ++j) buf[m++] = s[j];
    if (m == before) return 0;
  }
  if (j != e) return 0;
  double v = 0.0;
  int64_t used = 0;
  const int rc = parse_json_number(buf, m, Pow5Table{g_pow5_maint}, &v, &used);
  if (rc == kNumUndecided) return -1;
  if (rc != kNumOk || used != m) return 0;
  *out = neg ? -v : v;
  return 1;
}

// Date.parse for the format ECMA-262 specifies: 1 = ms in *out, 0 = NaN, -1 = another format (V8's legacy parser)
__device__ int date_parse(const uint8_t* s, int n, int64_t tz_off_ms, double* out) {
  if (n == 10) {  // YYYY-MM-DD: a date-only form is UTC
    const uint8_t utc_midnight[6] = {'0', '0', ':', '0', '0', 'Z'};
    return parse_show_date_time(s, 10, utc_midnight, 6, tz_off_ms, out);
  }
  if (n < 16 || s[10] != 'T') return -1;
  return parse_show_date_time(s, 10, s + 11, n - 11, tz_off_ms, out);
}

struct TimesArgs {
  const double* val[PIE_TF_COUNT];
  double* out[PIE_TF_COUNT];
  const uint8_t* kind;
  const uint8_t* text;   // the documents (nullptr: none given)
  int64_t text_begin, text_end;
  int64_t n;
  int64_t tz_off_ms;
};

__global__ void __launch_bounds__(256) get_timestamps_kernel(TimesArgs a, unsigned long long* __restrict__ err) {
  const int64_t s = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (s >= a.n) return;
#pragma unroll 1
  for (int f = 0; f < PIE_TF_COUNT; ++f) {
    if (!a.out[f]) continue;
    const double v = a.val[f] ? a.val[f][s] : quiet_nan();
    double r = quiet_nan();  // null
    int bad = 0;
    if (is_finite_f64(v)) {
      r = v;
    } else {
      const int kind = a.kind ? a.kind[s * PIE_TF_COUNT + f] : PIE_TK_ABSENT;
      if (kind == PIE_TK_NULL || kind == PIE_TK_FALSE) r = 0.0;  // Number(null), Number(false)
      else if (kind == PIE_TK_TRUE) r = 1.0;
      else if (kind == PIE_TK_OTHER) bad = -PIE_ERR_SCHEMA;     // Number([5]) is 5: not restated
      else if (kind == PIE_TK_STRING) {
        const int64_t at = (int64_t)(__double_as_longlong(v) & 0x7ffffffffffffLL);
        if (!a.text || at < a.text_begin || at >= a.text_end) {
          bad = -PIE_ERR_SCHEMA;  // a text timestamp, and the documents were not handed in
        } else {
          const uint8_t* p = a.text + at;
          int n = 0;  // the text is JSON (the ingest walked it): the string ends at its closing quote
          while (at + n < a.text_end && p[n] != '"' && p[n] != '\\' && n < 64) ++n;
          if (at + n >= a.text_end || p[n] != '"') {
            bad = -PIE_ERR_UNSUPPORTED_DATE;  // an escape inside, or longer than anything this decides
          } else {
            double x = 0.0;
            const int num = string_to_number(p, n, &x);
            if (num < 0) bad = -PIE_ERR_UNSUPPORTED_DATE;
            else if (num == 1 && is_finite_f64(x)) r = x;
            else {
              const int d = date_parse(p, n, a.tz_off_ms, &x);
              if (d < 0) bad = -PIE_ERR_UNSUPPORTED_DATE;
              else if (d == 1 && is_finite_f64(x)) r = x;
            }
          }
        }
      }
      // PIE_TK_ABSENT (undefined), PIE_TK_NONFINITE: Number() is NaN / not finite, not a string: null
    }
    if (bad) atomicMin(err, ((unsigned long long)s << 8) | (unsigned long long)bad);
    a.out[f][s] = r;
  }
}

__global__ void maint_status_kernel(unsigned long long* err, int32_t* status, int init) {
  if (init) { *err = ~0ull; return; }
  const unsigned long long k = *err;
  status[0] = k == ~0ull ? 0 : -(int32_t)(k & 0xff);
  status[1] = k == ~0ull ? -1 : (int32_t)(k >> 8);
}

// ---- _archiveDailyShows: date groups ---------------------------------------------------------------------------
struct DueSlot {
  int owner;       // a row whose key the slot stands for, -1 = free
  int first_row;   // the smallest row of the group
  long long earliest;  // ordered_key of the smallest createdAt of the group
};
struct DueScratch {
  DueSlot* slots;
  int32_t* slot_of;
  uint64_t cap;  // a power of two
};

__host__ __device__ inline uint64_t due_capacity(int64_t n) {
  uint64_t c = 64;
  while (c < 2 * (uint64_t)(n > 0 ? n : 1)) c <<= 1;
  return c;
}

// show.date.trim(), or "__undated__" when that is empty (a date that is not a string is '' in the table)
struct DateKey {
  const uint8_t* p;
  int n;
};
__device__ __forceinline__ DateKey date_key(const pie_archive_view& v, int64_t s) {
  static __device__ const uint8_t kUndated[12] = {'_', '_', 'u', 'n', 'd', 'a', 't', 'e', 'd', '_', '_', 0};
  const int b0 = v.show_date.offsets[s], e0 = v.show_date.offsets[s + 1];
  const uint8_t* p = v.show_date.data + b0;
  int b = 0, e = e0 - b0;
  while (b < e) { const int l = js_ws_len_at(p, b, e); if (!l) break; b += l; }
  while (e > b) { const int l = js_ws_len_before(p, b, e); if (!l) break; e -= l; }
  if (b == e) return DateKey{kUndated, 11};
  return DateKey{p + b, e - b};
}
__device__ __forceinline__ bool same_key(const DateKey& a, const DateKey& b) {
  if (a.n != b.n) return false;
  for (int i = 0; i < a.n; ++i)
    if (a.p[i] != b.p[i]) return false;
  return true;
}

__global__ void due_init_kernel(DueScratch sc) {
  const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < sc.cap) sc.slots[i] = DueSlot{-1, 0x7fffffff, kKeyHighest};
}

__global__ void __launch_bounds__(256) due_insert_kernel(pie_archive_view v, const uint8_t* __restrict__ doc_status,
                                                         const double* __restrict__ created, DueScratch sc) {
  const int64_t s = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (s >= v.n_shows) return;
  if (doc_status && doc_status[s]) { sc.slot_of[s] = -1; return; }
  const DateKey key = date_key(v, s);
  uint64_t h = 0xcbf29ce484222325ull;  // FNV-1a: only spreads the keys, equality is decided on the bytes
  for (int i = 0; i < key.n; ++i) h = (h ^ key.p[i]) * 0x100000001b3ull;
  h ^= h >> 29;
  uint64_t at = h & (sc.cap - 1);
  for (;;) {
    int owner = atomicCAS(&sc.slots[at].owner, -1, (int)s);
    if (owner == -1) owner = (int)s;  // ours now
    if (owner == (int)s || same_key(key, date_key(v, owner))) break;
    at = (at + 1) & (sc.cap - 1);
  }
  sc.slot_of[s] = (int32_t)at;
  const double c = created[s];
  const double value = is_finite_f64(c) ? c : 0.0;  // _getTimestamp(null) === 0 (:784)
  atomicMin(&sc.slots[at].earliest, ordered_key(value));
  atomicMin(&sc.slots[at].first_row, (int)s);
}

__global__ void __launch_bounds__(256) due_decide_kernel(int64_t n, DueScratch sc, double now_ms, uint8_t* __restrict__ due,
                                                         int32_t* __restrict__ group_first) {
  const int64_t s = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (s >= n) return;
  const int32_t at = sc.slot_of[s];
  if (at < 0) {
    due[s] = 0;
    group_first[s] = -1;
    return;
  }
  const double earliest = from_ordered_key(sc.slots[at].earliest);
  due[s] = (now_ms - earliest >= 43200000.0) ? 1 : 0;  // AUTO_ARCHIVE_WINDOW_MS (:9, :798)
  group_first[s] = sc.slots[at].first_row;
}

// ---- _purgeExpiredArchives: _addMonths(createdAt, 2) -----------------------------------------------------------
__device__ __forceinline__ void civil_from_days(int64_t z, int64_t* y, int* m, int* d) {
  z += 719468;
  const int64_t era = (z >= 0 ? z : z - 146096) / 146097;
  const int64_t doe = z - era * 146097;
  const int64_t yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
  const int64_t doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
  const int64_t mp = (5 * doy + 2) / 153;
  *d = (int)(doy - (153 * mp + 2) / 5 + 1);
  *m = (int)(mp < 10 ? mp + 3 : mp - 9);
  *y = yoe + era * 400 + (*m <= 2);
}

__global__ void __launch_bounds__(256) expired_kernel(const double* __restrict__ created, int64_t n, double now_ms,
                                                      int64_t tz_off_ms, uint8_t* __restrict__ expired) {
  const int64_t s = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (s >= n) return;
  const double c = created[s];
  bool out = false;
  if (is_finite_f64(c)) {  // _isArchiveExpired: false for a createdAt that is not finite (:992-994)
    double expiry = c;     // _addMonths returns the timestamp itself when new Date(timestamp) is invalid (:1003-1005)
    if (fabs(c) <= kMaxTime) {
      const int64_t t = (int64_t)c;  // TimeClip truncates toward zero
      const int64_t local = t + tz_off_ms;
      int64_t day = local / kDayMs, in_day = local % kDayMs;
      if (in_day < 0) { in_day += kDayMs; day -= 1; }
      int64_t y;
      int m, d;
      civil_from_days(day, &y, &m, &d);
      const int m0 = (m - 1) + 2;  // setMonth(getMonth() + ARCHIVE_RETENTION_MONTHS): MakeDay carries the overflow
      const int64_t ym = y + m0 / 12;
      const int mn = m0 % 12;
      const int64_t new_day = days_from_civil(ym, mn + 1, 1) + d - 1;
      const double v = (double)(new_day * kDayMs + in_day - tz_off_ms);
      expiry = fabs(v) <= kMaxTime ? v : quiet_nan();  // TimeClip of the new time value
    }
    out = now_ms >= expiry;  // false when the expiry is NaN
  }
  expired[s] = out ? 1 : 0;
}

}  // namespace

cudaError_t launch_get_timestamps(const pie_archive_view& v, const pie_json_docs* docs, int32_t tz_offset_minutes,
                                  const pie_doc_times& out, int32_t* status, unsigned long long* err_scratch,
                                  cudaStream_t stream) {
  TimesArgs a;
  a.val[PIE_TF_CREATED] = v.created_at;
  a.val[PIE_TF_UPDATED] = v.updated_at;
  a.val[PIE_TF_ARCHIVED] = v.archived_at;
  a.val[PIE_TF_DELETED] = v.deleted_at;
  a.out[PIE_TF_CREATED] = out.created_at;
  a.out[PIE_TF_UPDATED] = out.updated_at;
  a.out[PIE_TF_ARCHIVED] = out.archived_at;
  a.out[PIE_TF_DELETED] = out.deleted_at;
  a.kind = v.time_kind;
  a.text = docs ? docs->data : nullptr;
  a.text_begin = 0;
  a.text_end = 0;
  a.n = v.n_shows;
  a.tz_off_ms = (int64_t)tz_offset_minutes * 60000;
  maint_status_kernel<<<1, 1, 0, stream>>>(err_scratch, status, 1);
  if (docs && docs->n_docs > 0) {
    // the span of the documents' text: offsets[0] .. offsets[n] live on the device — read by the kernel through a
    // tiny helper would cost a launch; the caller's view holds them, so the bounds are passed as "everything"
    a.text_begin = 0;
    a.text_end = INT64_MAX;
  }
  if (v.n_shows > 0) get_timestamps_kernel<<<(unsigned)((v.n_shows + 255) / 256), 256, 0, stream>>>(a, err_scratch);
  maint_status_kernel<<<1, 1, 0, stream>>>(err_scratch, status, 0);
  g_launches += v.n_shows > 0 ? 3 : 2;
  return cudaGetLastError();
}

uint64_t archive_due_scratch_bytes(int64_t n_shows) {
  return due_capacity(n_shows) * sizeof(DueSlot) + 4 * (uint64_t)(n_shows > 0 ? n_shows : 1) + 256;
}

cudaError_t launch_archive_due(const pie_archive_view& v, const uint8_t* doc_status, const double* created, double now_ms,
                               uint8_t* due, int32_t* group_first, void* scratch, cudaStream_t stream) {
  if (v.n_shows == 0) return cudaSuccess;
  DueScratch sc;
  sc.cap = due_capacity(v.n_shows);
  sc.slots = static_cast<DueSlot*>(scratch);
  sc.slot_of = reinterpret_cast<int32_t*>(static_cast<uint8_t*>(scratch) + sc.cap * sizeof(DueSlot));
  const unsigned blocks = (unsigned)((v.n_shows + 255) / 256);
  due_init_kernel<<<(unsigned)((sc.cap + 255) / 256), 256, 0, stream>>>(sc);
  due_insert_kernel<<<blocks, 256, 0, stream>>>(v, doc_status, created, sc);
  due_decide_kernel<<<blocks, 256, 0, stream>>>(v.n_shows, sc, now_ms, due, group_first);
  g_launches += 3;
  return cudaGetLastError();
}

cudaError_t launch_archive_expired(const double* created, int64_t n, double now_ms, int32_t tz_offset_minutes,
                                   uint8_t* expired, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  expired_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(created, n, now_ms, (int64_t)tz_offset_minutes * 60000, expired);
  g_launches += 1;
  return cudaGetLastError();
}

}  // namespace pie
