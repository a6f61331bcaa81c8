/* pie_oracle.c — plain-C restatement of the reference's archive analytics on the columnar
 * "archive table" layout of include/sph_pie_b200.h.
 *
 * TEST INFRASTRUCTURE ONLY: the checker for the CUDA path and the timed CPU baseline of bench.py.
 * Nothing under sph_pie_b200/ links or loads this file.
 *
 * PARITY STATUS: parity unpinned (see oracle/pie_oracle.py header): no JS engine exists in this
 * image, so this restatement is validated against oracle/pie_oracle.py (a second, independent
 * restatement that works on JSON-shaped objects) and the reference's single fixture, not against
 * outputs of the reference itself.
 *
 * Deliberately structured differently from the CUDA kernels (no classification byte, no radix
 * sort): per-show loops straight over the strings, grouping by a merge sort.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/sph_pie_b200.h"

#include <pthread.h>
#include <unistd.h>

static const char* const kIssues[PIE_N_ISSUES] = { /* public/app.js:1-13 */
    "Tracking lost", "Failed to launch", "Command delay", "RF link", "Battery",
    "Motor or prop", "Sensor or IMU", "Software or show control", "Operator input", "Other"};

static int lower_eq(const uint8_t* s, int n, const char* lit) {
  int l = (int)strlen(lit);
  if (n != l) return 0;
  for (int i = 0; i < n; ++i) {
    uint8_t c = s[i];
    if (c >= 'A' && c <= 'Z') c = (uint8_t)(c + 32);
    if (c != (uint8_t)lit[i]) return 0;
  }
  return 1;
}

/* decode one UTF-8 code point at s[i]; returns its length (>=1) and the code point */
static int utf8_at(const uint8_t* s, int i, int n, uint32_t* cp) {
  uint8_t c = s[i];
  if (c < 0x80) { *cp = c; return 1; }
  if ((c & 0xE0) == 0xC0 && i + 1 < n) { *cp = ((c & 0x1Fu) << 6) | (s[i + 1] & 0x3Fu); return 2; }
  if ((c & 0xF0) == 0xE0 && i + 2 < n) {
    *cp = ((c & 0x0Fu) << 12) | ((s[i + 1] & 0x3Fu) << 6) | (s[i + 2] & 0x3Fu);
    return 3;
  }
  if ((c & 0xF8) == 0xF0 && i + 3 < n) {
    *cp = ((c & 0x07u) << 18) | ((s[i + 1] & 0x3Fu) << 12) | ((s[i + 2] & 0x3Fu) << 6) | (s[i + 3] & 0x3Fu);
    return 4;
  }
  *cp = 0xFFFD;
  return 1;
}

static int is_js_space(uint32_t cp) { /* ECMAScript WhiteSpace + LineTerminator */
  return cp == 9 || cp == 10 || cp == 11 || cp == 12 || cp == 13 || cp == 32 || cp == 0xA0 || cp == 0x1680 ||
         (cp >= 0x2000 && cp <= 0x200A) || cp == 0x2028 || cp == 0x2029 || cp == 0x202F || cp == 0x205F ||
         cp == 0x3000 || cp == 0xFEFF;
}

/* String.prototype.trim on UTF-8: [*b, *e) shrinks to the trimmed range */
static void js_trim(const uint8_t* s, int* b, int* e) {
  int i = *b, end = *e;
  while (i < end) {
    uint32_t cp;
    int l = utf8_at(s, i, end, &cp);
    if (!is_js_space(cp)) break;
    i += l;
  }
  /* walk forward remembering the end of the last non-space code point */
  int last = i, j = i;
  while (j < end) {
    uint32_t cp;
    int l = utf8_at(s, j, end, &cp);
    j += l;
    if (!is_js_space(cp)) last = j;
  }
  *b = i;
  *e = last;
}

static double js_max2(double a, double b) {
  if (a > b) return a;
  if (b > a) return b;
  return signbit(a) ? b : a;
}
static double js_min2(double a, double b) {
  if (a < b) return a;
  if (b < a) return b;
  return signbit(a) ? a : b;
}

/* computeArchiveShowStats, public/app.js:3898-3953, for show s */
static void show_stats_one(const pie_archive_view* v, int64_t s, int32_t* si, double* sf, int64_t stride) {
  int e0 = v->entry_offsets[s], e1 = v->entry_offsets[s + 1];
  int completed = 0, no_launch = 0, abort_n = 0, launched = 0, delay_n = 0;
  int issue_n[PIE_N_ISSUES];
  for (int k = 0; k < PIE_N_ISSUES; ++k) issue_n[k] = 0;
  uint64_t order = 0; /* 4-bit nibbles: (k+1) of the j-th distinct issue met = property insertion order */
  int n_distinct = 0;
  double sum = 0.0, mx = 0.0;
  for (int e = e0; e < e1; ++e) {
    const uint8_t* st = v->status.data + v->status.offsets[e];
    int stn = v->status.offsets[e + 1] - v->status.offsets[e];
    if (lower_eq(st, stn, "completed")) completed++;
    else if (lower_eq(st, stn, "no-launch")) no_launch++;
    else if (lower_eq(st, stn, "abort")) abort_n++;
    if (lower_eq(v->launched.data + v->launched.offsets[e], v->launched.offsets[e + 1] - v->launched.offsets[e], "yes"))
      launched++;
    if (v->delay_valid[e] && isfinite(v->delay_sec[e])) {
      double d = v->delay_sec[e];
      sum = sum + d;
      mx = delay_n ? js_max2(mx, d) : d;
      delay_n++;
    }
    int b = v->primary_issue.offsets[e], en = v->primary_issue.offsets[e + 1];
    js_trim(v->primary_issue.data, &b, &en);
    if (en > b) {
      int k = PIE_N_ISSUES - 1; /* 'Other' */
      for (int q = 0; q < PIE_N_ISSUES; ++q) {
        if ((int)strlen(kIssues[q]) == en - b && memcmp(kIssues[q], v->primary_issue.data + b, (size_t)(en - b)) == 0) {
          k = q;
          break;
        }
      }
      if (issue_n[k] == 0) order |= (uint64_t)(k + 1) << (4 * n_distinct++);
      issue_n[k]++;
    }
  }
  int total = e1 - e0;
  si[PIE_SI_TOTAL * stride + s] = total;
  si[PIE_SI_COMPLETED * stride + s] = completed;
  si[PIE_SI_NO_LAUNCH * stride + s] = no_launch;
  si[PIE_SI_ABORT * stride + s] = abort_n;
  si[PIE_SI_LAUNCHED * stride + s] = launched;
  si[PIE_SI_DELAY_COUNT * stride + s] = delay_n;
  si[PIE_SI_ISSUE_ORDER_LO * stride + s] = (int32_t)(uint32_t)(order & 0xFFFFFFFFu);
  si[PIE_SI_ISSUE_ORDER_HI * stride + s] = (int32_t)(uint32_t)(order >> 32);
  sf[PIE_SF_AVG_DELAY * stride + s] = delay_n ? sum / delay_n : NAN;
  sf[PIE_SF_MAX_DELAY * stride + s] = delay_n ? mx : NAN;
  sf[PIE_SF_COMPLETION_RATE * stride + s] = total ? ((double)completed / total) * 100 : NAN;
  sf[PIE_SF_LAUNCH_RATE * stride + s] = total ? ((double)launched / total) * 100 : NAN;
  sf[PIE_SF_ABORT_RATE * stride + s] = total ? ((double)abort_n / total) * 100 : NAN;
  for (int k = 0; k < PIE_N_ISSUES; ++k) {
    si[(PIE_SI_ISSUE_COUNT0 + k) * stride + s] = issue_n[k];
    sf[(PIE_SF_ISSUE_RATE0 + k) * stride + s] = total ? ((double)issue_n[k] / total) * 100 : NAN;
  }
}

typedef struct {
  const pie_archive_view* v;
  int32_t* si;
  double* sf;
  int64_t stride, s0, s1;
} stats_job;

static void* stats_worker(void* arg) {
  stats_job* j = (stats_job*)arg;
  for (int64_t s = j->s0; s < j->s1; ++s) show_stats_one(j->v, s, j->si, j->sf, j->stride);
  return NULL;
}

/* nthreads <= 1: sequential.  Shows are independent; threads take contiguous, entry-balanced
 * ranges of shows and run the same per-show code. */
int oracle_show_stats(const pie_archive_view* v, int32_t* si, double* sf, int64_t stride, int nthreads) {
  const int64_t S = v->n_shows;
  if (nthreads <= 1 || S < 2 * (int64_t)nthreads) {
    for (int64_t s = 0; s < S; ++s) show_stats_one(v, s, si, sf, stride);
    return 0;
  }
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
  stats_job* jobs = (stats_job*)malloc(sizeof(stats_job) * (size_t)nthreads);
  const int64_t E = v->entry_offsets[S];
  int64_t s0 = 0;
  for (int t = 0; t < nthreads; ++t) {
    int64_t s1 = s0;
    const int64_t target = (int64_t)((double)E * (t + 1) / nthreads);
    if (t == nthreads - 1) s1 = S;
    else while (s1 < S && v->entry_offsets[s1] < target) s1++;
    jobs[t] = (stats_job){v, si, sf, stride, s0, s1};
    pthread_create(&th[t], NULL, stats_worker, &jobs[t]);
    s0 = s1;
  }
  for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  free(th);
  free(jobs);
  return 0;
}

int oracle_max_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

/* ---- getShowTimestamp / parseShowDateTime (public/app.js:4092-4126) --------------------------- */
static int64_t days_from_civil(int64_t y, int m, int d) {
  y -= m <= 2;
  int64_t era = (y >= 0 ? y : y - 399) / 400;
  int64_t yoe = y - era * 400;
  int64_t doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
  int64_t doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
  return era * 146097 + doe - 719468;
}

static int digits(const uint8_t* s, int n, int* out) {
  int v = 0;
  for (int i = 0; i < n; ++i) {
    if (s[i] < '0' || s[i] > '9') return 0;
    v = v * 10 + (s[i] - '0');
  }
  *out = v;
  return 1;
}

/* 1 = parsed, 0 = NaN (illegal element value), -1 = outside the ECMA-262 grammar */
static int parse_iso(const uint8_t* ds, int dn, const uint8_t* ts, int tn, int64_t tz_ms, double* out) {
  char buf[64];
  if (tn == 0) { ts = (const uint8_t*)"00:00"; tn = 5; }
  if (dn + 1 + tn >= (int)sizeof(buf)) return -1;
  memcpy(buf, ds, (size_t)dn);
  buf[dn] = 'T';
  memcpy(buf + dn + 1, ts, (size_t)tn);
  int n = dn + 1 + tn;
  const uint8_t* p = (const uint8_t*)buf;
  int Y, M, D, h, mi, sec = 0, ms = 0;
  if (n < 16) return -1;
  if (!digits(p, 4, &Y) || p[4] != '-' || !digits(p + 5, 2, &M) || p[7] != '-' || !digits(p + 8, 2, &D) ||
      p[10] != 'T' || !digits(p + 11, 2, &h) || p[13] != ':' || !digits(p + 14, 2, &mi))
    return -1;
  int pos = 16;
  if (pos < n && p[pos] == ':') {
    if (pos + 3 > n || !digits(p + pos + 1, 2, &sec)) return -1;
    pos += 3;
    if (pos < n && p[pos] == '.') {
      if (pos + 4 > n || !digits(p + pos + 1, 3, &ms)) return -1;
      pos += 4;
    }
  }
  int has_off = 0;
  int64_t off = 0;
  if (pos < n) {
    if (p[pos] == 'Z' && pos + 1 == n) has_off = 1;
    else if ((p[pos] == '+' || p[pos] == '-') && pos + 6 == n && p[pos + 3] == ':') {
      int oh, om;
      if (!digits(p + pos + 1, 2, &oh) || !digits(p + pos + 4, 2, &om)) return -1;
      if (oh > 23 || om > 59) return 0;
      off = (int64_t)(oh * 60 + om) * 60000 * (p[pos] == '-' ? -1 : 1);
      has_off = 1;
    } else return -1;
  }
  if (M < 1 || M > 12 || D < 1 || D > 31 || h > 24 || mi > 59 || sec > 59) return 0;
  if (h == 24 && (mi || sec || ms)) return 0;
  int leap = (Y % 4 == 0) && (Y % 100 != 0 || Y % 400 == 0);
  int dim = M == 2 ? (leap ? 29 : 28) : ((M == 4 || M == 6 || M == 9 || M == 11) ? 30 : 31);
  if (D > dim) return -1;
  int64_t local = ((days_from_civil(Y, M, D) * 24 + h) * 60 + mi) * 60000 + sec * 1000 + ms;
  *out = (double)(local - (has_off ? off : tz_ms));
  return 1;
}

typedef struct { int64_t day; int32_t show; } day_item;

static void merge_sort(day_item* a, day_item* tmp, int64_t n) { /* stable */
  if (n < 2) return;
  int64_t h = n / 2;
  merge_sort(a, tmp, h);
  merge_sort(a + h, tmp, n - h);
  if (a[h - 1].day <= a[h].day) return; /* halves already in order (archives usually arrive sorted) */
  int64_t i = 0, j = h, k = 0;
  while (i < h && j < n) tmp[k++] = (a[j].day < a[i].day) ? a[j++] : a[i++];
  while (i < h) tmp[k++] = a[i++];
  while (j < n) tmp[k++] = a[j++];
  memcpy(a, tmp, (size_t)n * sizeof(day_item));
}

static double metric_value(const int32_t* si, const double* sf, int64_t stride, int m, int64_t s) {
  switch (m) {
    case 0: return si[PIE_SI_TOTAL * stride + s];
    case 1: return si[PIE_SI_COMPLETED * stride + s];
    case 2: return si[PIE_SI_NO_LAUNCH * stride + s];
    case 3: return si[PIE_SI_ABORT * stride + s];
    case 4: return sf[PIE_SF_AVG_DELAY * stride + s];
    case 5: return sf[PIE_SF_MAX_DELAY * stride + s];
    case 6: return sf[PIE_SF_COMPLETION_RATE * stride + s];
    case 7: return sf[PIE_SF_LAUNCH_RATE * stride + s];
    case 8: return sf[PIE_SF_ABORT_RATE * stride + s];
    default: return sf[(PIE_SF_ISSUE_RATE0 + (m - 9)) * stride + s];
  }
}

/* buildArchiveDailyGroups + getOrCreateGroupMetricSummary (public/app.js:3401-3502).
 * Returns a pie_status; out->status mirrors it. */
int oracle_daily_summary(const pie_archive_view* v, const int32_t* si, const double* sf, int64_t stride,
                         int32_t tz_offset_minutes, const pie_daily_out* out) {
  const int64_t S = v->n_shows, tz = (int64_t)tz_offset_minutes * 60000, DAY = 86400000LL;
  day_item* items = (day_item*)malloc((size_t)(S > 0 ? S : 1) * sizeof(day_item) * 2);
  int64_t nvalid = 0, nskip = 0;
  int32_t* skipped = (int32_t*)malloc((size_t)(S > 0 ? S : 1) * sizeof(int32_t));
  out->status[0] = 0;
  out->status[1] = -1;
  for (int64_t s = 0; s < S; ++s) {
    double ts = NAN;
    if (isfinite(v->created_at[s])) {
      ts = v->created_at[s];
    } else {
      int parsed = 0;
      if (v->show_date.offsets && v->show_date.offsets[s + 1] > v->show_date.offsets[s]) {
        int tb = 0, te = 0;
        if (v->show_time.offsets) { tb = v->show_time.offsets[s]; te = v->show_time.offsets[s + 1]; }
        parsed = parse_iso(v->show_date.data + v->show_date.offsets[s],
                           v->show_date.offsets[s + 1] - v->show_date.offsets[s],
                           v->show_time.offsets ? v->show_time.data + tb : NULL, te - tb, tz, &ts);
        if (parsed < 0) {
          out->status[0] = PIE_ERR_UNSUPPORTED_DATE; out->status[1] = (int32_t)s;
          free(items); free(skipped);
          return PIE_ERR_UNSUPPORTED_DATE;
        }
      }
      if (parsed == 0) {
        if (v->archived_at && isfinite(v->archived_at[s])) ts = v->archived_at[s];
        else if (v->entry_ts) {
          int any = 0;
          double best = 0;
          for (int e = v->entry_offsets[s]; e < v->entry_offsets[s + 1]; ++e)
            if (isfinite(v->entry_ts[e]) && (!any || v->entry_ts[e] < best)) { best = v->entry_ts[e]; any = 1; }
          if (any) ts = best;
        }
      }
    }
    if (!isfinite(ts)) {
      out->show_day_start[s] = PIE_DAY_NONE;
      skipped[nskip++] = (int32_t)s;
      continue;
    }
    int bad = fabs(ts) > 8.64e15;
    int64_t start = 0, day = 0;
    if (!bad) {
      int64_t t = (int64_t)ts, local = t + tz;
      day = local / DAY;
      if (local % DAY < 0) day--;
      start = day * DAY - tz;
      if (start > 8640000000000000LL || start < -8640000000000000LL) bad = 1;
    }
    if (bad) {
      out->status[0] = PIE_ERR_RANGE; out->status[1] = (int32_t)s;
      free(items); free(skipped);
      return PIE_ERR_RANGE;
    }
    out->show_day_start[s] = start;
    items[nvalid].day = day;
    items[nvalid].show = (int32_t)s;
    nvalid++;
  }
  merge_sort(items, items + (S > 0 ? S : 1), nvalid);
  int64_t G = 0;
  for (int64_t i = 0; i < nvalid; ++i) {
    out->show_order[i] = items[i].show;
    if (i == 0 || items[i].day != items[i - 1].day) {
      out->group_offsets[G] = (int32_t)i;
      out->group_day_start[G] = items[i].day * DAY - tz;
      G++;
    }
  }
  out->group_offsets[G] = (int32_t)nvalid;
  for (int64_t i = 0; i < nskip; ++i) out->show_order[nvalid + i] = skipped[i];
  *out->n_groups = G;
  const int64_t plane = (int64_t)PIE_N_METRICS * out->stride;
  for (int64_t g = 0; g < G; ++g) {
    for (int m = 0; m < PIE_N_METRICS; ++m) {
      double sum = 0, mn = 0, mx = 0;
      int n = 0;
      for (int i = out->group_offsets[g]; i < out->group_offsets[g + 1]; ++i) {
        double x = metric_value(si, sf, stride, m, out->show_order[i]);
        if (isfinite(x)) {
          sum = sum + x;
          mn = n ? js_min2(mn, x) : x;
          mx = n ? js_max2(mx, x) : x;
          n++;
        }
      }
      int64_t o = (int64_t)m * out->stride + g;
      out->summary_f64[PIE_DF_AVERAGE * plane + o] = n ? sum / n : NAN;
      out->summary_f64[PIE_DF_MIN * plane + o] = n ? mn : NAN;
      out->summary_f64[PIE_DF_MAX * plane + o] = n ? mx : NAN;
      out->summary_count[o] = n;
    }
  }
  free(items);
  free(skipped);
  return 0;
}


/* ---- export rows: buildTableRow + csvEscape + buildCsvRow (server/webhookDispatcher.js:276-342) -- */
#include <stdio.h>

/* Number::toString via printf/strtod (independent of the product's Ryu).  For p = 1..17 digits:
 * m = the correctly rounded p-digit decimal of |x| (printf is exact), which is the p-digit decimal
 * closest to |x|.  The decimals that read back as x form an interval, so at the first p where any
 * p-digit decimal round-trips it is m itself, or else exactly one of its neighbours m+1 / m-1. */
static int js_number_to_string(double x, char* out) {
  if (isnan(x)) return sprintf(out, "NaN");
  if (isinf(x)) return sprintf(out, x > 0 ? "Infinity" : "-Infinity");
  if (x == 0) return sprintf(out, "0");
  const double ax = fabs(x);
  if (ax < 9007199254740992.0 && ax == floor(ax)) /* integers below 2^53 print exactly */
    return sprintf(out, "%s%llu", x < 0 ? "-" : "", (unsigned long long)ax);
  char buf[48], cand[48];
  unsigned long long best = 0;
  int best_q = 0, found = 0;
  /* short fixed-point decimals (12.5, 0.25, 3.1): the fewest fractional digits d with
   * m = |x|*10^d integral and "m e-d" reading back as x.  m < 2^50 keeps neighbouring d-digit
   * decimals more than an ulp apart, so the round-tripping one is unique (hence the closest). */
  double scale = 1;
  for (int d = 1; d <= 6 && !found; ++d) {
    scale *= 10;
    const double m = ax * scale;
    if (m >= 1125899906842624.0) break;
    if (m == floor(m)) {
      snprintf(cand, sizeof cand, "%llue-%d", (unsigned long long)m, d);
      if (strtod(cand, NULL) == ax) { best = (unsigned long long)m; best_q = -d; found = 1; }
    }
  }
  /* general case: "p digits suffice" is monotone in p -> binary search for the smallest p */
  int lo = 1, hi = 17;
  while (!found) {
    const int p = (lo == hi) ? lo : (lo + hi) / 2;
    snprintf(buf, sizeof buf, "%.*e", p - 1, ax);
    char* e = strchr(buf, 'e');
    unsigned long long m = 0;
    for (char* c = buf; c < e; ++c) if (*c != '.') m = m * 10 + (unsigned long long)(*c - '0');
    const int q = atoi(e + 1) - (p - 1); /* |x| ~ m * 10^q */
    const long long delta[3] = {0, 1, -1};
    int ok = 0;
    unsigned long long okc = 0;
    for (int t = 0; t < 3 && !ok; ++t) {
      if (m == 0 && delta[t] < 0) continue;
      const unsigned long long c = m + (unsigned long long)delta[t];
      snprintf(cand, sizeof cand, "%llue%d", c, q);
      if (strtod(cand, NULL) == ax) { ok = 1; okc = c; }
    }
    if (lo == hi) { best = okc; best_q = q; found = 1; break; } /* p = 17 always round-trips */
    if (ok) hi = p; else lo = p + 1;
  }
  while (best % 10 == 0) { best /= 10; best_q++; } /* 10^p from a carry, trailing zeros */
  char digits[24];
  int k = snprintf(digits, sizeof digits, "%llu", best);
  int n = k + best_q; /* value = 0.d1..dk * 10^n */
  char* o = out;
  if (x < 0) *o++ = '-';
  if (k <= n && n <= 21) {
    memcpy(o, digits, (size_t)k); o += k;
    for (int i = k; i < n; ++i) *o++ = '0';
  } else if (0 < n && n <= 21) {
    memcpy(o, digits, (size_t)n); o += n;
    *o++ = '.';
    memcpy(o, digits + n, (size_t)(k - n)); o += k - n;
  } else if (-6 < n && n <= 0) {
    *o++ = '0'; *o++ = '.';
    for (int i = 0; i < -n; ++i) *o++ = '0';
    memcpy(o, digits, (size_t)k); o += k;
  } else {
    *o++ = digits[0];
    if (k > 1) { *o++ = '.'; memcpy(o, digits + 1, (size_t)(k - 1)); o += k - 1; }
    o += sprintf(o, "e%c%d", n - 1 < 0 ? '-' : '+', abs(n - 1));
  }
  *o = 0;
  return (int)(o - out);
}

typedef struct { uint8_t* p; uint64_t n; uint64_t cap; } csv_out; /* p may be NULL: count only */

static inline void put_bytes(csv_out* o, const uint8_t* s, uint64_t n) {
  if (o->p && o->n + n <= o->cap) {
    uint8_t* d = o->p + o->n;
    if (n <= 16) for (uint64_t i = 0; i < n; ++i) d[i] = s[i];
    else memcpy(d, s, (size_t)n);
  }
  o->n += n;
}
static inline void put_char(csv_out* o, char c) {
  if (o->p && o->n < o->cap) o->p[o->n] = (uint8_t)c;
  o->n += 1;
}
static inline int is_special(uint8_t c) { return c == '"' || c == ',' || c == '\n' || c == '\r'; }

/* csvEscape (:332-338) */
static inline void put_cell(csv_out* o, const uint8_t* s, int n) {
  int quote = 0;
  for (int i = 0; i < n; ++i) quote |= is_special(s[i]);
  if (!quote) { put_bytes(o, s, (uint64_t)n); return; }
  put_char(o, '"');
  for (int i = 0; i < n; ++i) { if (s[i] == '"') put_char(o, '"'); put_char(o, (char)s[i]); }
  put_char(o, '"');
}
static inline void put_col(csv_out* o, const pie_strcol* c, int64_t i) {
  put_cell(o, c->data + c->offsets[i], c->offsets[i + 1] - c->offsets[i]);
}
/* list.join('|') then csvEscape of the joined string: '|' is not special, so the joined string
 * needs quotes iff one of the items holds a special character */
static inline void put_joined(csv_out* o, const pie_strlistcol* c, int64_t i) {
  const int l0 = c->list_offsets[i], l1 = c->list_offsets[i + 1];
  if (l1 <= l0) return;
  const uint8_t* d = c->items.data;
  int quote = 0;
  for (int k = c->items.offsets[l0]; k < c->items.offsets[l1]; ++k) quote |= is_special(d[k]);
  if (quote) put_char(o, '"');
  for (int l = l0; l < l1; ++l) {
    if (l > l0) put_char(o, '|');
    const int b = c->items.offsets[l], e = c->items.offsets[l + 1];
    if (!quote) put_bytes(o, d + b, (uint64_t)(e - b));
    else for (int k = b; k < e; ++k) { if (d[k] == '"') put_char(o, '"'); put_char(o, (char)d[k]); }
  }
  if (quote) put_char(o, '"');
}

static void csv_show_range(const pie_archive_view* v, int64_t s0, int64_t s1, int64_t* row_offsets, csv_out* op) {
  csv_out o = *op;
  for (int64_t s = s0; s < s1; ++s) {
    for (int e = v->entry_offsets[s]; e < v->entry_offsets[s + 1]; ++e) {
      if (row_offsets) row_offsets[e] = (int64_t)o.n;
      put_col(&o, &v->show_id, s); put_char(&o, ',');
      put_col(&o, &v->show_date, s); put_char(&o, ',');
      put_col(&o, &v->show_time, s); put_char(&o, ',');
      put_col(&o, &v->show_label, s); put_char(&o, ',');
      put_joined(&o, &v->crew, s); put_char(&o, ',');
      put_col(&o, &v->lead_pilot, s); put_char(&o, ',');
      put_col(&o, &v->monkey_lead, s); put_char(&o, ',');
      put_col(&o, &v->show_notes, s); put_char(&o, ',');
      put_col(&o, &v->entry_id, e); put_char(&o, ',');
      put_col(&o, &v->unit_id, e); put_char(&o, ',');
      put_col(&o, &v->planned, e); put_char(&o, ',');
      put_col(&o, &v->launched, e); put_char(&o, ',');
      put_col(&o, &v->status, e); put_char(&o, ',');
      int sn = v->status.offsets[e + 1] - v->status.offsets[e];
      int completed = sn == 9 && memcmp(v->status.data + v->status.offsets[e], "Completed", 9) == 0; /* === */
      const pie_strcol* blanked[5] = {&v->primary_issue, &v->sub_issue, &v->other_detail, &v->severity, &v->root_cause};
      for (int k = 0; k < 5; ++k) { if (!completed) put_col(&o, blanked[k], e); put_char(&o, ','); }
      put_joined(&o, &v->actions, e); put_char(&o, ',');
      put_col(&o, &v->operator_name, e); put_char(&o, ',');
      put_col(&o, &v->battery_id, e); put_char(&o, ',');
      if (v->delay_valid[e]) {
        char num[40];
        int n = js_number_to_string(v->delay_sec[e], num);
        put_bytes(&o, (const uint8_t*)num, (uint64_t)n);
      }
      put_char(&o, ',');
      put_col(&o, &v->command_rx, e); put_char(&o, ',');
      put_col(&o, &v->notes, e);
      put_char(&o, '\n');
    }
  }
  *op = o;
}

/* ---- archive entry payloads: JSON.stringify(buildArchiveEntryPayload(show, entry)) --------------------
 * server/webhookDispatcher.js:315-330 (the object literal's property order is the serialisation order),
 * toYesNoBoolean :60-77, QuoteJSONString ECMA-262 25.5.2.3. */
static void put_json_string(csv_out* o, const pie_strcol* c, int64_t i) {
  static const char hex[] = "0123456789abcdef";
  const uint8_t* s = c->data + c->offsets[i];
  const int n = c->offsets[i + 1] - c->offsets[i];
  put_char(o, '"');
  for (int k = 0; k < n; ++k) {
    const uint8_t ch = s[k];
    switch (ch) {
      case '"': put_char(o, '\\'); put_char(o, '"'); break;
      case '\\': put_char(o, '\\'); put_char(o, '\\'); break;
      case '\b': put_char(o, '\\'); put_char(o, 'b'); break;
      case '\t': put_char(o, '\\'); put_char(o, 't'); break;
      case '\n': put_char(o, '\\'); put_char(o, 'n'); break;
      case '\f': put_char(o, '\\'); put_char(o, 'f'); break;
      case '\r': put_char(o, '\\'); put_char(o, 'r'); break;
      default:
        if (ch < 0x20) {
          put_char(o, '\\'); put_char(o, 'u'); put_char(o, '0'); put_char(o, '0');
          put_char(o, hex[ch >> 4]); put_char(o, hex[ch & 15]);
        } else {
          put_char(o, (char)ch);
        }
    }
  }
  put_char(o, '"');
}
static void put_lit(csv_out* o, const char* s) { put_bytes(o, (const uint8_t*)s, strlen(s)); }
static void put_yes_no(csv_out* o, const pie_strcol* c, int64_t i) { /* value.trim().toLowerCase() === 'yes' */
  const uint8_t* s = c->data + c->offsets[i];
  int b = 0, e = c->offsets[i + 1] - c->offsets[i];
  js_trim(s, &b, &e);
  put_lit(o, lower_eq(s + b, e - b, "yes") ? "true" : "false");
}

static void payload_show_range(const pie_archive_view* v, int64_t s0, int64_t s1, int64_t* row_offsets, csv_out* op) {
  csv_out o = *op;
  for (int64_t s = s0; s < s1; ++s) {
    for (int e = v->entry_offsets[s]; e < v->entry_offsets[s + 1]; ++e) {
      if (row_offsets) row_offsets[e] = (int64_t)o.n;
      put_lit(&o, "{\"showDate\":"); put_json_string(&o, &v->show_date, s);
      put_lit(&o, ",\"showTime\":"); put_json_string(&o, &v->show_time, s);
      put_lit(&o, ",\"showNumber\":"); put_json_string(&o, &v->show_label, s);
      put_lit(&o, ",\"leadPilot\":"); put_json_string(&o, &v->lead_pilot, s);
      put_lit(&o, ",\"monkeyLead\":"); put_json_string(&o, &v->monkey_lead, s);
      put_lit(&o, ",\"operator\":"); put_json_string(&o, &v->operator_name, e);
      put_lit(&o, ",\"monkeyId\":"); put_json_string(&o, &v->unit_id, e);
      put_lit(&o, ",\"planned\":"); put_yes_no(&o, &v->planned, e);
      put_lit(&o, ",\"launched\":"); put_yes_no(&o, &v->launched, e);
      put_lit(&o, ",\"commandReceived\":"); put_yes_no(&o, &v->command_rx, e);
      put_lit(&o, ",\"primaryIssue\":"); put_json_string(&o, &v->primary_issue, e);
      put_lit(&o, ",\"subIssue\":"); put_json_string(&o, &v->sub_issue, e);
      put_lit(&o, "}\n");
    }
  }
  *op = o;
}

typedef void (*row_range_fn)(const pie_archive_view*, int64_t, int64_t, int64_t*, csv_out*);

/* Fills row_offsets[n_entries+1]; writes rows (each followed by '\n') into out_data when it is not
 * NULL and large enough; *total receives the size needed. */
static int rows_single(row_range_fn fn, const pie_archive_view* v, int64_t* row_offsets, uint8_t* out_data,
                       uint64_t capacity, uint64_t* total) {
  csv_out o = {out_data, 0, capacity};
  fn(v, 0, v->n_shows, row_offsets, &o);
  row_offsets[v->n_entries] = (int64_t)o.n;
  *total = o.n;
  return 0;
}
int oracle_csv_rows(const pie_archive_view* v, int64_t* row_offsets, uint8_t* out_data, uint64_t capacity,
                    uint64_t* total) {
  return rows_single(csv_show_range, v, row_offsets, out_data, capacity, total);
}

typedef struct {
  row_range_fn fn;
  const pie_archive_view* v;
  int64_t s0, s1;
  int64_t* row_offsets;
  csv_out out;
} csv_job;

static void* csv_worker(void* arg) {
  csv_job* j = (csv_job*)arg;
  j->fn(j->v, j->s0, j->s1, j->row_offsets, &j->out);
  return NULL;
}

/* Threaded variant for the CPU baseline: a counting pass per show range, a prefix over the ranges,
 * then the writing pass — the same per-row code as the single-threaded entry point. */
static int rows_mt(row_range_fn fn, const pie_archive_view* v, int64_t* row_offsets, uint8_t* out_data,
                   uint64_t capacity, uint64_t* total, int nthreads) {
  const int64_t S = v->n_shows;
  if (nthreads <= 1 || S < 2 * (int64_t)nthreads) return rows_single(fn, v, row_offsets, out_data, capacity, total);
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
  csv_job* jobs = (csv_job*)malloc(sizeof(csv_job) * (size_t)nthreads);
  const int64_t E = v->entry_offsets[S];
  int64_t s0 = 0;
  for (int t = 0; t < nthreads; ++t) {
    int64_t s1 = s0;
    const int64_t target = (int64_t)((double)E * (t + 1) / nthreads);
    if (t == nthreads - 1) s1 = S;
    else while (s1 < S && v->entry_offsets[s1] < target) s1++;
    jobs[t] = (csv_job){fn, v, s0, s1, NULL, {NULL, 0, 0}};
    s0 = s1;
  }
  for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, csv_worker, &jobs[t]);
  for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  uint64_t base = 0;
  for (int t = 0; t < nthreads; ++t) {
    const uint64_t n = jobs[t].out.n;
    jobs[t].row_offsets = row_offsets;
    jobs[t].out = (csv_out){out_data, base, capacity}; /* p + n addressing: n starts at the range's base */
    base += n;
  }
  *total = base;
  row_offsets[v->n_entries] = (int64_t)base;
  for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, csv_worker, &jobs[t]);
  for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  free(th);
  free(jobs);
  return 0;
}
int oracle_csv_rows_mt(const pie_archive_view* v, int64_t* row_offsets, uint8_t* out_data, uint64_t capacity,
                       uint64_t* total, int nthreads) {
  return rows_mt(csv_show_range, v, row_offsets, out_data, capacity, total, nthreads);
}
/* JSON.stringify(buildArchiveEntryPayload(show, entry)) + '\n' per entry */
int oracle_payload_rows_mt(const pie_archive_view* v, int64_t* row_offsets, uint8_t* out_data, uint64_t capacity,
                           uint64_t* total, int nthreads) {
  return rows_mt(payload_show_range, v, row_offsets, out_data, capacity, total, nthreads);
}

/* ---- computeMetrics(show) (public/app.js:5024-5047) ------------------------------------------------------- */
/* Number.prototype.toFixed(2): the exact value of |x| * 100, rounded to an integer with ties up, by integer
 * arithmetic on the mantissa (ECMA-262 21.1.3.3: "n / 10^f - x as close to zero as possible; two such n: the
 * larger"); "-" iff x < 0; |x| >= 1e21 and non-finite values go through Number::toString. */
static int js_to_fixed2(double x, char* out) {
  if (x != x) { memcpy(out, "NaN", 3); return 3; }
  if (fabs(x) >= 1e21) return js_number_to_string(x, out);
  char* o = out;
  if (x < 0) *o++ = '-';
  int e2;
  const double fr = frexp(fabs(x), &e2);                   /* |x| = fr * 2^e2, fr in [0.5, 1) */
  const unsigned __int128 m = (unsigned __int128)(uint64_t)ldexp(fr, 53); /* 53-bit integer mantissa */
  const int e = e2 - 53;                                   /* |x| = m * 2^e */
  unsigned __int128 n;                                     /* round(|x| * 100) */
  if (e >= 0) n = (m * 100) << e;                          /* < 1e23 < 2^77 */
  else if (-e >= 120) n = 0;
  else n = (m * 100 + ((unsigned __int128)1 << (-e - 1))) >> -e;
  char digits[48];
  int k = 0;
  unsigned __int128 ip = n / 100;
  const int frac = (int)(n % 100);
  do { digits[k++] = (char)('0' + (int)(ip % 10)); ip /= 10; } while (ip);
  while (k) *o++ = digits[--k];
  *o++ = '.';
  *o++ = (char)('0' + frac / 10);
  *o++ = (char)('0' + frac % 10);
  return (int)(o - out);
}

static int str_eq_lit(const pie_strcol* c, int64_t i, const char* lit) {
  const int n = c->offsets[i + 1] - c->offsets[i];
  return n == (int)strlen(lit) && memcmp(c->data + c->offsets[i], lit, (size_t)n) == 0;
}
/* array index key (ECMA-262 6.1.7): canonical decimal string of an integer in 0 .. 2^32-2 */
static int array_index_key(const uint8_t* s, int n, uint64_t* value) {
  if (n < 1 || n > 10 || (n > 1 && s[0] == '0')) return 0;
  uint64_t v = 0;
  for (int i = 0; i < n; ++i) { if (s[i] < '0' || s[i] > '9') return 0; v = v * 10 + (uint64_t)(s[i] - '0'); }
  if (v > 4294967294ull) return 0;
  *value = v;
  return 1;
}
typedef struct { int first; int count; int is_index; uint64_t index; int order; } issue_slot;

int oracle_compute_metrics(const pie_archive_view* v, int32_t* out, uint8_t* text, int64_t stride) {
  for (int64_t s = 0; s < v->n_shows; ++s) {
    const int e0 = v->entry_offsets[s], e1 = v->entry_offsets[s + 1];
    int planned = 0, completed = 0, no_launch = 0, abort_ = 0, dn = 0;
    double sum = 0.0;
    issue_slot* slots = (issue_slot*)malloc(sizeof(issue_slot) * (size_t)(e1 - e0 + 1));
    int ns = 0;
    for (int e = e0; e < e1; ++e) {
      planned += str_eq_lit(&v->planned, e, "Yes");
      const int comp = str_eq_lit(&v->status, e, "Completed");
      completed += comp;
      no_launch += str_eq_lit(&v->status, e, "No-launch");
      abort_ += str_eq_lit(&v->status, e, "Abort");
      if (v->delay_valid[e]) { sum = sum + v->delay_sec[e]; dn++; }
      const uint8_t* p = v->primary_issue.data + v->primary_issue.offsets[e];
      const int n = v->primary_issue.offsets[e + 1] - v->primary_issue.offsets[e];
      if (comp || n == 0) continue;
      int k = 0;
      for (; k < ns; ++k) {
        const int f = slots[k].first;
        const int fn = v->primary_issue.offsets[f + 1] - v->primary_issue.offsets[f];
        if (fn == n && memcmp(v->primary_issue.data + v->primary_issue.offsets[f], p, (size_t)n) == 0) break;
      }
      if (k == ns) {
        slots[ns].first = e; slots[ns].count = 0; slots[ns].order = ns;
        slots[ns].is_index = array_index_key(p, n, &slots[ns].index);
        ns++;
      }
      slots[k].count++;
    }
    /* Object.entries order: array-index keys ascending, then creation order; then a stable sort by count desc:
     * insertion sort on the combined key keeps it simple and stable */
    for (int i = 1; i < ns; ++i) {
      issue_slot x = slots[i];
      int j = i - 1;
      for (; j >= 0; --j) {
        const issue_slot* y = &slots[j];
        int before; /* does x come before y? */
        if (x.count != y->count) before = x.count > y->count;
        else if (x.is_index != y->is_index) before = x.is_index;
        else if (x.is_index) before = x.index < y->index;
        else before = x.order < y->order;
        if (!before) break;
        slots[j + 1] = slots[j];
      }
      slots[j + 1] = x;
    }
    double rate = 0;
    if (planned) {
      const double q = ((double)completed / (double)planned) * 100.0;
      const double r = floor(q);
      rate = (q - r >= 0.5) ? r + 1 : r; /* Math.round: ties toward +inf */
    }
    out[0 * stride + s] = (int32_t)rate;
    out[1 * stride + s] = completed;
    out[2 * stride + s] = no_launch;
    out[3 * stride + s] = abort_;
    for (int k = 0; k < 3; ++k) out[(4 + k) * stride + s] = k < ns ? slots[k].first : -1;
    char buf[64];
    int len;
    if (dn) len = js_to_fixed2(sum / (double)dn, buf);
    else { memcpy(buf, "0.00", 4); len = 4; }
    out[7 * stride + s] = len;
    memset(text + 32 * s, 0, 32);
    memcpy(text + 32 * s, buf, (size_t)len);
    free(slots);
  }
  return 0;
}

/* Number::toString of an array (checker for the product's Ryu): out is n x 32 bytes */
void oracle_number_to_string_batch(const double* x, int64_t n, char* out, int32_t* lens) {
  for (int64_t i = 0; i < n; ++i) lens[i] = js_number_to_string(x[i], out + 32 * i);
}

/* =====================================================================================================
 * Stored documents -> the archive table: JSON.parse(row.data) as _mapArchiveRow applies it (reference
 * server/storage/sqlProvider.js:892-926: null unless the text parses to an object; arrays are objects) followed
 * by the projection on the table's schema (sph_pie_b200/columnar.py pack_shows, i.e. sqlProvider.js:361-409).
 * Written independently of the kernels: recursive descent over ECMA-404, a byte at a time, numbers by strtod
 * (glibc: correctly rounded).  Two passes per document, as the table is laid out from the counts of the first.
 * Also the CPU baseline of the ingest path.
 * ===================================================================================================== */
#define ING_PLANES 26
#define ING_ENTRIES 23
#define ING_CREW_ITEMS 24
#define ING_ACTION_ITEMS 25
#define ING_MAX_DEPTH 64

typedef struct pie_strcol_mut_ { int32_t* offsets; uint8_t* data; } ing_col;

typedef struct {
  /* the 23 heaps in table order (7 show columns, crew items, 14 entry columns, action items) */
  ing_col heap[23];
  int32_t *entry_offsets, *crew_list, *actions_list;
  double *created_at, *archived_at, *delay_sec, *entry_ts;
  uint8_t* delay_valid;
} ing_out;

typedef struct {
  const uint8_t *p, *end;
  int depth;
  int syntax;  /* the text is not JSON: the row is dropped */
  int hard;    /* first of PIE_ERR_SCHEMA / PIE_ERR_UNSUPPORTED_JSON met (reported only if the text is JSON) */
  int write;   /* second pass */
  uint64_t pos[ING_PLANES];
  const ing_out* out;
  int64_t show;
} ing;

static void ing_hard(ing* j, int code) { if (!j->hard) j->hard = code; }
static void ing_ws(ing* j) { while (j->p < j->end && (*j->p == ' ' || *j->p == '\t' || *j->p == '\n' || *j->p == '\r')) ++j->p; }
static int ing_hex(int c) {
  if (c >= '0' && c <= '9') return c - '0';
  if (c >= 'a' && c <= 'f') return c - 'a' + 10;
  if (c >= 'A' && c <= 'F') return c - 'A' + 10;
  return -1;
}

/* a string, the cursor on its opening quote.  dst (may be NULL) receives the unescaped UTF-8; *len its length;
 * key (may be NULL) the first 31 bytes, NUL-terminated; *lone: an escape named an unpaired surrogate */
static void ing_string(ing* j, uint8_t* dst, uint64_t* len, char* key, int* lone) {
  uint64_t n = 0;
  uint32_t high = 0;
  *lone = 0;
#define ING_PUT(b) do { uint8_t b_ = (uint8_t)(b); if (dst) dst[n] = b_; if (key && n < 31) key[n] = (char)b_; ++n; } while (0)
#define ING_PUT3(cp) do { ING_PUT(0xE0 | ((cp) >> 12)); ING_PUT(0x80 | (((cp) >> 6) & 0x3F)); ING_PUT(0x80 | ((cp) & 0x3F)); } while (0)
  ++j->p;
  for (;;) {
    if (j->p >= j->end) { j->syntax = 1; break; }
    uint32_t c = *j->p++;
    if (c == '\\') {
      if (j->p >= j->end) { j->syntax = 1; break; }
      const int e = *j->p++;
      uint32_t cp;
      if (e == 'u') {
        if (j->end - j->p < 4) { j->syntax = 1; break; }
        cp = 0;
        for (int k = 0; k < 4; ++k) {
          const int h = ing_hex(j->p[k]);
          if (h < 0) { j->syntax = 1; break; }
          cp = cp * 16 + (uint32_t)h;
        }
        if (j->syntax) break;
        j->p += 4;
        if (high) {
          if (cp >= 0xDC00 && cp <= 0xDFFF) {
            const uint32_t full = 0x10000 + ((high - 0xD800) << 10) + (cp - 0xDC00);
            high = 0;
            ING_PUT(0xF0 | (full >> 18)); ING_PUT(0x80 | ((full >> 12) & 0x3F)); ING_PUT(0x80 | ((full >> 6) & 0x3F)); ING_PUT(0x80 | (full & 0x3F));
            continue;
          }
          *lone = 1;
          ING_PUT3(high);
          high = 0;
        }
        if (cp >= 0xD800 && cp <= 0xDBFF) { high = cp; continue; }
        if (cp >= 0xDC00 && cp <= 0xDFFF) *lone = 1;
      } else {
        switch (e) {
          case '"': cp = '"'; break; case '\\': cp = '\\'; break; case '/': cp = '/'; break; case 'b': cp = 8; break;
          case 'f': cp = 12; break; case 'n': cp = 10; break; case 'r': cp = 13; break; case 't': cp = 9; break;
          default: j->syntax = 1; cp = 0; break;
        }
        if (j->syntax) break;
        if (high) { *lone = 1; ING_PUT3(high); high = 0; }
      }
      if (cp < 0x80) ING_PUT(cp);
      else if (cp < 0x800) { ING_PUT(0xC0 | (cp >> 6)); ING_PUT(0x80 | (cp & 0x3F)); }
      else ING_PUT3(cp);
      continue;
    }
    if (high) { *lone = 1; ING_PUT3(high); high = 0; }
    if (c == '"') break;
    if (c < 0x20) { j->syntax = 1; break; }
    ING_PUT(c);
    if (c >= 0x80) { /* UTF-8, well-formed sequences only (Unicode table 3-7); otherwise reported, walk goes on */
      int need = 0;
      uint32_t lo = 0x80, hi = 0xBF;
      if (c < 0xC2) ing_hard(j, PIE_ERR_UNSUPPORTED_JSON);
      else if (c < 0xE0) need = 1;
      else if (c < 0xF0) { need = 2; if (c == 0xE0) lo = 0xA0; if (c == 0xED) hi = 0x9F; }
      else if (c < 0xF5) { need = 3; if (c == 0xF0) lo = 0x90; if (c == 0xF4) hi = 0x8F; }
      else ing_hard(j, PIE_ERR_UNSUPPORTED_JSON);
      for (int k = 0; k < need; ++k) {
        if (j->p >= j->end || *j->p < lo || *j->p > hi) { ing_hard(j, PIE_ERR_UNSUPPORTED_JSON); break; }
        ING_PUT(*j->p++);
        lo = 0x80; hi = 0xBF;
      }
    }
  }
#undef ING_PUT3
#undef ING_PUT
  if (key) key[n < 31 ? n : 31] = 0;
  *len = n;
}

/* a number, the cursor on '-' or a digit: grammar by hand, value by strtod */
static double ing_number(ing* j) {
  const uint8_t* s = j->p;
  if (j->p < j->end && *j->p == '-') ++j->p;
  if (j->p >= j->end || *j->p < '0' || *j->p > '9') { j->syntax = 1; return 0; }
  if (*j->p == '0') { ++j->p; if (j->p < j->end && *j->p >= '0' && *j->p <= '9') { j->syntax = 1; return 0; } }
  else while (j->p < j->end && *j->p >= '0' && *j->p <= '9') ++j->p;
  if (j->p < j->end && *j->p == '.') {
    ++j->p;
    if (j->p >= j->end || *j->p < '0' || *j->p > '9') { j->syntax = 1; return 0; }
    while (j->p < j->end && *j->p >= '0' && *j->p <= '9') ++j->p;
  }
  if (j->p < j->end && (*j->p == 'e' || *j->p == 'E')) {
    ++j->p;
    if (j->p < j->end && (*j->p == '+' || *j->p == '-')) ++j->p;
    if (j->p >= j->end || *j->p < '0' || *j->p > '9') { j->syntax = 1; return 0; }
    while (j->p < j->end && *j->p >= '0' && *j->p <= '9') ++j->p;
  }
  const size_t n = (size_t)(j->p - s);
  char stack[128];
  char* buf = n < sizeof(stack) ? stack : (char*)malloc(n + 1);
  memcpy(buf, s, n);
  buf[n] = 0;
  const double v = strtod(buf, NULL);
  if (buf != stack) free(buf);
  return v;
}

static int ing_literal(ing* j, const char* lit) {
  const size_t n = strlen(lit);
  if ((size_t)(j->end - j->p) < n || memcmp(j->p, lit, n) != 0) { j->syntax = 1; return 0; }
  j->p += n;
  return 1;
}

/* any value, validated and thrown away */
static void ing_skip(ing* j) {
  ing_ws(j);
  if (j->p >= j->end) { j->syntax = 1; return; }
  const int c = *j->p;
  if (c == '"') { uint64_t n; int lone; ing_string(j, NULL, &n, NULL, &lone); return; }
  if (c == '-' || (c >= '0' && c <= '9')) { ing_number(j); return; }
  if (c == 't') { ing_literal(j, "true"); return; }
  if (c == 'f') { ing_literal(j, "false"); return; }
  if (c == 'n') { ing_literal(j, "null"); return; }
  if (c != '{' && c != '[') { j->syntax = 1; return; }
  if (j->depth >= ING_MAX_DEPTH) { j->hard = PIE_ERR_UNSUPPORTED_JSON; j->syntax = 2; return; } /* 2: stop, not a drop */
  ++j->depth;
  ++j->p;
  const int close = c == '{' ? '}' : ']';
  ing_ws(j);
  if (j->p < j->end && *j->p == close) { ++j->p; --j->depth; return; }
  for (;;) {
    if (c == '{') {
      ing_ws(j);
      if (j->p >= j->end || *j->p != '"') { j->syntax = 1; return; }
      uint64_t n; int lone;
      ing_string(j, NULL, &n, NULL, &lone);
      if (j->syntax) return;
      ing_ws(j);
      if (j->p >= j->end || *j->p != ':') { j->syntax = 1; return; }
      ++j->p;
    }
    ing_skip(j);
    if (j->syntax) return;
    ing_ws(j);
    if (j->p >= j->end) { j->syntax = 1; return; }
    if (*j->p == ',') { ++j->p; continue; }
    if (*j->p == close) { ++j->p; --j->depth; return; }
    j->syntax = 1;
    return;
  }
}

/* a value that belongs in text heap h: string or null, anything else is a schema error (and still has to be JSON) */
static void ing_text(ing* j, int h) {
  ing_ws(j);
  if (j->p < j->end && *j->p == '"') {
    uint64_t n; int lone;
    ing_string(j, j->write ? j->out->heap[h].data + j->pos[h] : NULL, &n, NULL, &lone);
    j->pos[h] += n;
    if (lone) ing_hard(j, PIE_ERR_SCHEMA);
    return;
  }
  if (j->p < j->end && *j->p == 'n') { ing_literal(j, "null"); return; }
  ing_hard(j, PIE_ERR_SCHEMA);
  ing_skip(j);
}

/* crew / actions: a list of strings when the value is an array, else empty */
static void ing_list(ing* j, int h, int items) {
  ing_ws(j);
  if (j->p >= j->end || *j->p != '[') { ing_skip(j); return; }
  if (j->depth >= ING_MAX_DEPTH) { j->hard = PIE_ERR_UNSUPPORTED_JSON; j->syntax = 2; return; }
  ++j->depth;
  ++j->p;
  ing_ws(j);
  if (j->p < j->end && *j->p == ']') { ++j->p; --j->depth; return; }
  for (;;) {
    if (j->write) j->out->heap[h].offsets[j->pos[items]] = (int32_t)j->pos[h];
    ++j->pos[items];
    ing_text(j, h);
    if (j->syntax) return;
    ing_ws(j);
    if (j->p >= j->end) { j->syntax = 1; return; }
    if (*j->p == ',') { ++j->p; continue; }
    if (*j->p == ']') { ++j->p; --j->depth; return; }
    j->syntax = 1;
    return;
  }
}

/* a finite number, or absent (NaN) */
static double ing_time(ing* j) {
  ing_ws(j);
  if (j->p < j->end && (*j->p == '-' || (*j->p >= '0' && *j->p <= '9'))) {
    const double v = ing_number(j);
    return isfinite(v) ? v : NAN;
  }
  ing_skip(j);
  return NAN;
}

static const char* const kIngShowKeys[11] = {"id", "date", "time", "label", "leadPilot", "monkeyLead", "notes", "crew",
                                             "createdAt", "archivedAt", "entries"};
static const char* const kIngEntryKeys[17] = {"id", "unitId", "planned", "launched", "status", "primaryIssue", "subIssue",
                                              "otherDetail", "severity", "rootCause", "operator", "batteryId", "commandRx",
                                              "notes", "actions", "delaySec", "ts"};
static int ing_key_index(const char* const* keys, int nkeys, const char* key, uint64_t len) {
  if (len > 30) return -1;
  for (int i = 0; i < nkeys; ++i)
    if (strlen(keys[i]) == len && memcmp(keys[i], key, len) == 0) return i;
  return -1;
}

/* the members of an object whose '{' has been taken: calls member(j, key index or -1) with the cursor behind ':' */
static void ing_entry(ing* j) { /* the cursor on the element of `entries`, whatever it is */
  const uint64_t row = j->pos[ING_ENTRIES]++;
  double delay = 0.0, ts = NAN;
  int valid = 0;
  if (j->write) {
    for (int h = 0; h < 14; ++h) j->out->heap[8 + h].offsets[row] = (int32_t)j->pos[8 + h];
    j->out->actions_list[row] = (int32_t)j->pos[ING_ACTION_ITEMS];
  }
  ing_ws(j);
  if (j->p < j->end && *j->p == '{') {
    if (j->depth >= ING_MAX_DEPTH) { j->hard = PIE_ERR_UNSUPPORTED_JSON; j->syntax = 2; return; }
    ++j->depth;
    ++j->p;
    uint32_t seen = 0;
    ing_ws(j);
    if (j->p < j->end && *j->p == '}') { ++j->p; --j->depth; }
    else for (;;) {
      ing_ws(j);
      if (j->p >= j->end || *j->p != '"') { j->syntax = 1; return; }
      char key[32]; uint64_t klen; int lone;
      ing_string(j, NULL, &klen, key, &lone);
      if (j->syntax) return;
      ing_ws(j);
      if (j->p >= j->end || *j->p != ':') { j->syntax = 1; return; }
      ++j->p;
      const int k = ing_key_index(kIngEntryKeys, 17, key, klen);
      if (k >= 0) { if (seen >> k & 1) ing_hard(j, PIE_ERR_UNSUPPORTED_JSON); seen |= 1u << k; }
      if (k >= 0 && k < 14) ing_text(j, 8 + k);
      else if (k == 14) ing_list(j, 22, ING_ACTION_ITEMS);
      else if (k == 15) { /* delaySec: number | null */
        ing_ws(j);
        if (j->p < j->end && (*j->p == '-' || (*j->p >= '0' && *j->p <= '9'))) { delay = ing_number(j); valid = 1; }
        else if (j->p < j->end && *j->p == 'n') ing_literal(j, "null");
        else { ing_hard(j, PIE_ERR_SCHEMA); ing_skip(j); }
      } else if (k == 16) ts = ing_time(j);
      else ing_skip(j);
      if (j->syntax) return;
      ing_ws(j);
      if (j->p >= j->end) { j->syntax = 1; return; }
      if (*j->p == ',') { ++j->p; continue; }
      if (*j->p == '}') { ++j->p; --j->depth; break; }
      j->syntax = 1;
      return;
    }
  } else {
    ing_skip(j); /* not an object: a row without fields */
    if (j->syntax) return;
  }
  if (j->write) {
    j->out->delay_sec[row] = delay;
    j->out->delay_valid[row] = (uint8_t)valid;
    j->out->entry_ts[row] = ts;
  }
}

/* one document; returns 0 kept, 1 dropped, or a negative pie_status */
static int ing_document(ing* j) {
  ing_ws(j);
  int is_object = 0;
  if (j->p < j->end && *j->p == '{') {
    is_object = 1;
    j->depth = 1;
    ++j->p;
    uint32_t seen = 0;
    ing_ws(j);
    if (j->p < j->end && *j->p == '}') { ++j->p; }
    else for (;;) {
      ing_ws(j);
      if (j->p >= j->end || *j->p != '"') { j->syntax = 1; break; }
      char key[32]; uint64_t klen; int lone;
      ing_string(j, NULL, &klen, key, &lone);
      if (j->syntax) break;
      ing_ws(j);
      if (j->p >= j->end || *j->p != ':') { j->syntax = 1; break; }
      ++j->p;
      const int k = ing_key_index(kIngShowKeys, 11, key, klen);
      if (k >= 0) { if (seen >> k & 1) ing_hard(j, PIE_ERR_UNSUPPORTED_JSON); seen |= 1u << k; }
      if (k >= 0 && k < 7) ing_text(j, k);
      else if (k == 7) ing_list(j, 7, ING_CREW_ITEMS);
      else if (k == 8 || k == 9) {
        const double v = ing_time(j);
        if (j->write) (k == 8 ? j->out->created_at : j->out->archived_at)[j->show] = v;
      } else if (k == 10) {
        ing_ws(j);
        if (j->p < j->end && *j->p == '[') {
          ++j->depth;
          ++j->p;
          ing_ws(j);
          if (j->p < j->end && *j->p == ']') { ++j->p; --j->depth; }
          else for (;;) {
            ing_entry(j);
            if (j->syntax) break;
            ing_ws(j);
            if (j->p >= j->end) { j->syntax = 1; break; }
            if (*j->p == ',') { ++j->p; continue; }
            if (*j->p == ']') { ++j->p; --j->depth; break; }
            j->syntax = 1;
            break;
          }
        } else ing_skip(j);
      } else ing_skip(j);
      if (j->syntax) break;
      ing_ws(j);
      if (j->p >= j->end) { j->syntax = 1; break; }
      if (*j->p == ',') { ++j->p; continue; }
      if (*j->p == '}') { ++j->p; break; }
      j->syntax = 1;
      break;
    }
  } else if (j->p < j->end && *j->p == '[') {
    is_object = 1; /* typeof [] === 'object': an empty show */
    ing_skip(j);
  } else {
    ing_skip(j);
  }
  if (j->syntax == 2) return PIE_ERR_UNSUPPORTED_JSON; /* too deep: reported at once */
  if (!j->syntax) { ing_ws(j); if (j->p < j->end) j->syntax = 1; }
  if (j->syntax) return 1;
  if (j->hard) return j->hard;
  return is_object ? 0 : 1;
}

typedef struct {
  const uint8_t* text; const int64_t* offsets; int64_t d0, d1;
  uint32_t* rows; /* [n][26] */
  uint8_t* doc_status;
  const ing_out* out;
  int write;
  int64_t bad_doc; int bad_code;
} ing_job;

static void* ing_worker(void* arg) {
  ing_job* job = (ing_job*)arg;
  job->bad_doc = -1;
  job->bad_code = 0;
  for (int64_t s = job->d0; s < job->d1; ++s) {
    ing j;
    memset(&j, 0, sizeof(j));
    j.p = job->text + job->offsets[s];
    j.end = job->text + job->offsets[s + 1];
    j.write = job->write;
    j.out = job->out;
    j.show = s;
    uint32_t* row = job->rows + s * ING_PLANES;
    if (job->write) {
      for (int p = 0; p < ING_PLANES; ++p) j.pos[p] = row[p];
      for (int h = 0; h < 7; ++h) job->out->heap[h].offsets[s] = (int32_t)j.pos[h];
      job->out->entry_offsets[s] = (int32_t)j.pos[ING_ENTRIES];
      job->out->crew_list[s] = (int32_t)j.pos[ING_CREW_ITEMS];
      job->out->created_at[s] = NAN;
      job->out->archived_at[s] = NAN;
      if (job->doc_status[s] == 0) ing_document(&j);
      continue;
    }
    const int r = ing_document(&j);
    if (r != 0) {
      memset(j.pos, 0, sizeof(j.pos));
      if (r < 0 && job->bad_doc < 0) { job->bad_doc = s; job->bad_code = r; }
    }
    job->doc_status[s] = r == 0 ? 0 : 1;
    for (int p = 0; p < ING_PLANES; ++p) row[p] = (uint32_t)j.pos[p];
  }
  return NULL;
}

static void ing_run(ing_job* proto, int64_t n, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > n) nthreads = n > 0 ? (int)n : 1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
  ing_job* jobs = (ing_job*)malloc(sizeof(ing_job) * (size_t)nthreads);
  for (int t = 0; t < nthreads; ++t) {
    jobs[t] = *proto;
    jobs[t].d0 = n * t / nthreads;
    jobs[t].d1 = n * (t + 1) / nthreads;
    pthread_create(&th[t], NULL, ing_worker, &jobs[t]);
  }
  proto->bad_doc = -1;
  proto->bad_code = 0;
  for (int t = 0; t < nthreads; ++t) {
    pthread_join(th[t], NULL);
    if (jobs[t].bad_doc >= 0 && proto->bad_doc < 0) { proto->bad_doc = jobs[t].bad_doc; proto->bad_code = jobs[t].bad_code; }
  }
  free(jobs);
  free(th);
}

/* first pass: rows[n][26] receive the counts and are turned into exclusive prefixes; totals[26]; status = {code, doc} */
int oracle_ingest_measure(const uint8_t* text, const int64_t* offsets, int64_t n_docs, uint32_t* rows, uint8_t* doc_status,
                          int64_t* totals, int32_t* status, int nthreads) {
  ing_job job;
  memset(&job, 0, sizeof(job));
  job.text = text; job.offsets = offsets; job.rows = rows; job.doc_status = doc_status;
  ing_run(&job, n_docs, nthreads);
  status[0] = job.bad_code;
  status[1] = (int32_t)job.bad_doc;
  for (int p = 0; p < ING_PLANES; ++p) {
    uint64_t run = 0;
    for (int64_t s = 0; s < n_docs; ++s) {
      const uint32_t v = rows[s * ING_PLANES + p];
      rows[s * ING_PLANES + p] = (uint32_t)run;
      run += v;
    }
    totals[p] = (int64_t)run;
    if (run > 0x7fffffffull && status[0] == 0) { status[0] = PIE_ERR_CAPACITY; status[1] = -1; }
  }
  return 0;
}

/* second pass into a table sized from the totals (same field order as pie_archive_view) */
int oracle_ingest_fill(const uint8_t* text, const int64_t* offsets, int64_t n_docs, uint32_t* rows, uint8_t* doc_status,
                       const int64_t* totals, const pie_archive_table* t, int nthreads) {
  ing_out out;
  const pie_strcol_mut* cols[23] = {&t->show_id, &t->show_date, &t->show_time, &t->show_label, &t->lead_pilot, &t->monkey_lead,
                                    &t->show_notes, &t->crew.items, &t->entry_id, &t->unit_id, &t->planned, &t->launched,
                                    &t->status, &t->primary_issue, &t->sub_issue, &t->other_detail, &t->severity, &t->root_cause,
                                    &t->operator_name, &t->battery_id, &t->command_rx, &t->notes, &t->actions.items};
  for (int h = 0; h < 23; ++h) { out.heap[h].offsets = cols[h]->offsets; out.heap[h].data = cols[h]->data; }
  out.entry_offsets = t->entry_offsets; out.crew_list = t->crew.list_offsets; out.actions_list = t->actions.list_offsets;
  out.created_at = t->created_at; out.archived_at = t->archived_at; out.delay_sec = t->delay_sec;
  out.entry_ts = t->entry_ts; out.delay_valid = t->delay_valid;
  ing_job job;
  memset(&job, 0, sizeof(job));
  job.text = text; job.offsets = offsets; job.rows = rows; job.doc_status = doc_status; job.out = &out; job.write = 1;
  ing_run(&job, n_docs, nthreads);
  /* the terminal offsets */
  for (int h = 0; h < 7; ++h) out.heap[h].offsets[n_docs] = (int32_t)totals[h];
  out.entry_offsets[n_docs] = (int32_t)totals[ING_ENTRIES];
  out.crew_list[n_docs] = (int32_t)totals[ING_CREW_ITEMS];
  out.heap[7].offsets[totals[ING_CREW_ITEMS]] = (int32_t)totals[7];
  for (int h = 8; h < 22; ++h) out.heap[h].offsets[totals[ING_ENTRIES]] = (int32_t)totals[h];
  out.actions_list[totals[ING_ENTRIES]] = (int32_t)totals[ING_ACTION_ITEMS];
  out.heap[22].offsets[totals[ING_ACTION_ITEMS]] = (int32_t)totals[22];
  return 0;
}
