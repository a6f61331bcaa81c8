// The walk of one stored show document (ECMA-404 recogniser + projection on the archive table), shared by the two
// passes of the ingest kernels (json_ingest.cu).  Everything is __host__ __device__ so that the SAME code is also built
// for the host by tests/native/ingest_host.cpp and checked against the oracle without a GPU; the product only ever
// runs it on the device.
#pragma once
#include <stdint.h>

#include "../../include/sph_pie_b200.h"
#include "pie_numparse.cuh"

#if defined(__CUDACC__)
#define PIE_JW_HD __host__ __device__ __forceinline__
#define PIE_JW_NOINLINE __host__ __device__ __noinline__
#else
#define PIE_JW_HD inline
#define PIE_JW_NOINLINE
#endif

namespace pie {
namespace jw {

PIE_JW_HD double jw_nan() { return np_bits_to_double(0x7ff8000000000000ull); }
PIE_JW_HD bool jw_is_finite(double x) {
  uint64_t b;
#if defined(__CUDA_ARCH__)
  b = (uint64_t)__double_as_longlong(x);
#else
  __builtin_memcpy(&b, &x, 8);
#endif
  return ((b >> 52) & 0x7ff) != 0x7ff;
}
PIE_JW_HD uint64_t jw_load_word(const uint64_t* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(reinterpret_cast<const unsigned long long*>(p));
#else
  return *p;
#endif
}

constexpr int kHeaps = PIE_INGEST_HEAPS;              // 23 string heaps
constexpr int kPlaneEntries = PIE_IT_ENTRIES;         // 23
constexpr int kPlaneCrewItems = PIE_IT_CREW_ITEMS;    // 24
constexpr int kPlaneActionItems = PIE_IT_ACTION_ITEMS;  // 25
constexpr int kPlanes = PIE_INGEST_TOTALS;            // 26
constexpr int kHeapCrew = 7;
constexpr int kHeapEntry0 = 8;
constexpr int kHeapActions = 22;
constexpr int kMaxDepth = 64;

// ---- the document text, 8 aligned bytes at a time ---------------------------------------------------------------
struct DocCursor {
  const uint64_t* w;  // next word to request
  uint64_t cur, nxt;  // bytes not yet consumed (low byte first); the word after
  int left;           // valid bytes in cur (0 = end of the document)
  int words_left;     // words not yet moved into cur
  int tail_bytes;     // valid bytes of the last word

  PIE_JW_HD void open(const uint8_t* data, int64_t begin, int64_t end) {
    const int64_t n = end - begin;
    left = 0;
    words_left = 0;
    tail_bytes = 0;
    cur = nxt = 0;
    w = nullptr;
    if (n <= 0) return;
    const uintptr_t a = reinterpret_cast<uintptr_t>(data + begin);
    const int skip = (int)(a & 7);
    const uint64_t* first = reinterpret_cast<const uint64_t*>(a - skip);
    const int64_t words = (skip + n + 7) >> 3;
    tail_bytes = (int)((skip + n - 1) & 7) + 1;
    cur = jw_load_word(first) >> (8 * skip);
    left = words == 1 ? (int)n : 8 - skip;
    words_left = (int)(words - 1);
    w = first + 1;
    if (words_left > 0) nxt = jw_load_word(w++);
  }
  PIE_JW_HD int peek() const { return left > 0 ? (int)(cur & 0xFF) : -1; }
  PIE_JW_HD void next() {
    cur >>= 8;
    if (--left == 0 && words_left > 0) {
      cur = nxt;
      --words_left;
      left = words_left == 0 ? tail_bytes : 8;
      if (words_left > 0) nxt = jw_load_word(w++);
    }
  }
};

// ---- keys ---------------------------------------------------------------------------------------------------------
template <int L>
PIE_JW_HD constexpr uint64_t key_word(const char (&s)[L], int from) {
  uint64_t v = 0;
  for (int j = 0; j < 8; ++j)
    if (from + j < L - 1) v |= (uint64_t)(unsigned char)s[from + j] << (8 * j);
  return v;
}
#define PIE_KEY_IS(lit) (len == sizeof(lit) - 1 && k0 == key_word(lit, 0) && k1 == key_word(lit, 8))

enum ShowKey { kSkCrew = 7, kSkCreatedAt = 8, kSkArchivedAt = 9, kSkEntries = 10 };
enum EntryKey { kEkActions = 14, kEkDelaySec = 15, kEkTs = 16 };

// keys of the show document in the order of the table's show columns (columnar.py SHOW_KEY_TO_COL), then the rest
PIE_JW_HD int match_show_key(uint32_t len, uint64_t k0, uint64_t k1) {
  if (PIE_KEY_IS("id")) return 0;
  if (PIE_KEY_IS("date")) return 1;
  if (PIE_KEY_IS("time")) return 2;
  if (PIE_KEY_IS("label")) return 3;
  if (PIE_KEY_IS("leadPilot")) return 4;
  if (PIE_KEY_IS("monkeyLead")) return 5;
  if (PIE_KEY_IS("notes")) return 6;
  if (PIE_KEY_IS("crew")) return kSkCrew;
  if (PIE_KEY_IS("createdAt")) return kSkCreatedAt;
  if (PIE_KEY_IS("archivedAt")) return kSkArchivedAt;
  if (PIE_KEY_IS("entries")) return kSkEntries;
  return -1;
}
// keys of an entry in the order of the table's entry columns (ENTRY_KEY_TO_COL)
PIE_JW_HD int match_entry_key(uint32_t len, uint64_t k0, uint64_t k1) {
  if (PIE_KEY_IS("id")) return 0;
  if (PIE_KEY_IS("unitId")) return 1;
  if (PIE_KEY_IS("planned")) return 2;
  if (PIE_KEY_IS("launched")) return 3;
  if (PIE_KEY_IS("status")) return 4;
  if (PIE_KEY_IS("primaryIssue")) return 5;
  if (PIE_KEY_IS("subIssue")) return 6;
  if (PIE_KEY_IS("otherDetail")) return 7;
  if (PIE_KEY_IS("severity")) return 8;
  if (PIE_KEY_IS("rootCause")) return 9;
  if (PIE_KEY_IS("operator")) return 10;
  if (PIE_KEY_IS("batteryId")) return 11;
  if (PIE_KEY_IS("commandRx")) return 12;
  if (PIE_KEY_IS("notes")) return 13;
  if (PIE_KEY_IS("actions")) return kEkActions;
  if (PIE_KEY_IS("delaySec")) return kEkDelaySec;
  if (PIE_KEY_IS("ts")) return kEkTs;
  return -1;
}

// ---- strings ------------------------------------------------------------------------------------------------------
// where the unescaped bytes of a string go: nowhere (dst == nullptr: only counted), to memory, or — for keys — into
// two words that are compared with the known keys
template <bool kKey>
struct StrSink {
  uint8_t* dst;
  uint32_t len;
  uint64_t k0, k1;
  PIE_JW_HD void put(uint32_t b) {
    if (kKey) {
      if (len < 8) k0 |= (uint64_t)b << (8 * len);
      else if (len < 16) k1 |= (uint64_t)b << (8 * (len - 8));
    } else if (dst) {
      dst[len] = (uint8_t)b;
    }
    ++len;
  }
};

enum StrResult { kStrOk = 0, kStrSyntax = 1, kStrBadUtf8 = 2 };

PIE_JW_HD int hex_value(int c) {
  if (c >= '0' && c <= '9') return c - '0';
  c |= 0x20;
  if (c >= 'a' && c <= 'f') return c - 'a' + 10;
  return -1;
}

// The cursor stands behind the opening quote; on success it stands behind the closing one.  *lone is set when an
// escape names a surrogate code unit without its partner (a JS string that has no UTF-8 form).
template <bool kKey>
PIE_JW_NOINLINE int scan_string(DocCursor& c, StrSink<kKey>& out, bool* lone) {
  uint32_t high = 0;  // pending high surrogate of a \uD8xx escape
  for (;;) {
    const int ch = c.peek();
    if (ch < 0) return kStrSyntax;
    c.next();
    if (ch == '\\') {
      const int e = c.peek();
      if (e < 0) return kStrSyntax;
      c.next();
      uint32_t cp;
      if (e == 'u') {
        cp = 0;
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {
          const int h = hex_value(c.peek());
          if (h < 0) return kStrSyntax;
          c.next();
          cp = cp * 16 + (uint32_t)h;
        }
        if (high) {
          if (cp >= 0xDC00 && cp <= 0xDFFF) {
            cp = 0x10000 + ((high - 0xD800) << 10) + (cp - 0xDC00);
            high = 0;
            out.put(0xF0 | (cp >> 18));
            out.put(0x80 | ((cp >> 12) & 0x3F));
            out.put(0x80 | ((cp >> 6) & 0x3F));
            out.put(0x80 | (cp & 0x3F));
            continue;
          }
          *lone = true;  // the pending one stays alone; this unit starts over
          out.put(0xE0 | (high >> 12)); out.put(0x80 | ((high >> 6) & 0x3F)); out.put(0x80 | (high & 0x3F));
          high = 0;
        }
        if (cp >= 0xD800 && cp <= 0xDBFF) { high = cp; continue; }
        if (cp >= 0xDC00 && cp <= 0xDFFF) *lone = true;
      } else {
        switch (e) {
          case '"': cp = '"'; break;
          case '\\': cp = '\\'; break;
          case '/': cp = '/'; break;
          case 'b': cp = 8; break;
          case 'f': cp = 12; break;
          case 'n': cp = 10; break;
          case 'r': cp = 13; break;
          case 't': cp = 9; break;
          default: return kStrSyntax;
        }
        if (high) {
          *lone = true;
          out.put(0xE0 | (high >> 12)); out.put(0x80 | ((high >> 6) & 0x3F)); out.put(0x80 | (high & 0x3F));
          high = 0;
        }
      }
      if (cp < 0x80) {
        out.put(cp);
      } else if (cp < 0x800) {
        out.put(0xC0 | (cp >> 6));
        out.put(0x80 | (cp & 0x3F));
      } else {
        out.put(0xE0 | (cp >> 12));
        out.put(0x80 | ((cp >> 6) & 0x3F));
        out.put(0x80 | (cp & 0x3F));
      }
      continue;
    }
    if (high) {
      *lone = true;
      out.put(0xE0 | (high >> 12)); out.put(0x80 | ((high >> 6) & 0x3F)); out.put(0x80 | (high & 0x3F));
      high = 0;
    }
    if (ch == '"') return kStrOk;
    if (ch < 0x20) return kStrSyntax;  // control characters must be escaped
    out.put((uint32_t)ch);
    if (ch >= 0x80) {  // UTF-8: well-formed sequences only (Unicode table 3-7)
      int need;
      int lo = 0x80, hi = 0xBF;
      if (ch < 0xC2) return kStrBadUtf8;
      if (ch < 0xE0) need = 1;
      else if (ch < 0xF0) { need = 2; if (ch == 0xE0) lo = 0xA0; if (ch == 0xED) hi = 0x9F; }
      else if (ch < 0xF5) { need = 3; if (ch == 0xF0) lo = 0x90; if (ch == 0xF4) hi = 0x8F; }
      else return kStrBadUtf8;
#pragma unroll 1
      for (int k = 0; k < need; ++k) {
        const int b = c.peek();
        if (b < lo || b > hi) return kStrBadUtf8;
        c.next();
        out.put((uint32_t)b);
        lo = 0x80;
        hi = 0xBF;
      }
    }
  }
}

// ---- the walk -----------------------------------------------------------------------------------------------------
struct IngestOut {
  int32_t* off[kHeaps];
  uint8_t* data[kHeaps];
  int32_t* entry_offsets;
  int32_t* crew_list;
  int32_t* actions_list;
  double* created_at;
  double* archived_at;
  double* delay_sec;
  uint8_t* delay_valid;
  double* entry_ts;
};

enum Sem : int { kSemTop = 0, kSemShow, kSemCrew, kSemEntries, kSemEntry, kSemActions, kSemSkip, kSemDone };
enum Expect : int { kXValue = 0, kXValueOrClose, kXKeyOrClose, kXKey, kXColon, kXCommaOrClose, kXEnd };

// what the walk of one document reports
enum DocResult : int {
  kDocOk = 0,
  kDocDropped = 1,  // not JSON, or JSON that is not an object (sqlProvider.js:897-904: the row maps to null)
  // hard errors (the call fails): values = -pie_status
  kDocSchema = -PIE_ERR_SCHEMA,
  kDocUnsupported = -PIE_ERR_UNSUPPORTED_JSON,
};

// cnt[p]: pass 1 = what the document adds to plane p; pass 2 = the running position in plane p
template <bool kFill>
PIE_JW_HD int walk_document(DocCursor& c, uint32_t (&cnt)[kPlanes], const IngestOut& out, int64_t s,
                            const Pow5Table& pow5) {
  int depth = 0;
  uint64_t kinds = 0;  // bit d: the container open at depth d+1 is an object
  int sem = kSemTop, saved_sem = kSemDone, skip_base = 0;
  int expect = kXValue;
  int key = -1;  // key of the member whose value comes next (show / entry objects)
  uint32_t seen_show = 0, seen_entry = 0;
  int hard = 0;  // first hard error met; reported only if the document is JSON at all
  bool top_is_object = false;
  // the open entry
  double e_delay = 0.0, e_ts = 0.0;
  bool e_valid = false;
  uint32_t e_row = 0;

#define PIE_HARD(code) do { if (!hard) hard = (code); } while (0)

  auto begin_entry = [&]() {
    e_row = cnt[kPlaneEntries]++;
    seen_entry = 0;
    e_delay = 0.0;
    e_valid = false;
    e_ts = jw_nan();
    if (kFill) {
#pragma unroll
      for (int h = kHeapEntry0; h < kHeapEntry0 + 14; ++h) out.off[h][e_row] = (int32_t)cnt[h];
      out.actions_list[e_row] = (int32_t)cnt[kPlaneActionItems];
    }
  };
  auto end_entry = [&]() {
    if (kFill) {
      out.delay_sec[e_row] = e_delay;
      out.delay_valid[e_row] = e_valid ? 1 : 0;
      out.entry_ts[e_row] = e_ts;
    }
  };

  for (;;) {
    int ch = c.peek();
    while (ch == ' ' || ch == '\n' || ch == '\r' || ch == '\t') { c.next(); ch = c.peek(); }
    if (ch < 0) {
      if (expect != kXEnd) return kDocDropped;
      break;
    }
    if (expect == kXEnd) return kDocDropped;  // something after the value
    c.next();

    // ---- punctuation the grammar state asks for
    if (expect == kXColon) {
      if (ch != ':') return kDocDropped;
      expect = kXValue;
      continue;
    }
    bool closing = false;
    if (expect == kXCommaOrClose) {
      const bool in_object = (kinds >> (depth - 1)) & 1;
      if (ch == ',') { expect = in_object ? kXKey : kXValue; continue; }
      if (ch != (in_object ? '}' : ']')) return kDocDropped;
      closing = true;
    } else if (expect == kXKeyOrClose && ch == '}') {
      closing = true;
    } else if (expect == kXValueOrClose && ch == ']') {
      closing = true;
    }
    if (closing) {
      --depth;
      if (sem == kSemSkip) { if (depth == skip_base) sem = saved_sem; }
      else if (sem == kSemEntry) { end_entry(); sem = kSemEntries; }
      else if (sem == kSemActions) sem = kSemEntry;
      else if (sem == kSemShow) sem = kSemDone;
      else sem = kSemShow;  // crew, entries
      expect = depth == 0 ? kXEnd : kXCommaOrClose;
      continue;
    }

    // ---- a key
    if (expect == kXKeyOrClose || expect == kXKey) {
      if (ch != '"') return kDocDropped;
      bool lone = false;
      if (sem == kSemShow || sem == kSemEntry) {
        StrSink<true> sink{nullptr, 0, 0, 0};
        const int r = scan_string<true>(c, sink, &lone);
        if (r == kStrSyntax) return kDocDropped;
        if (r == kStrBadUtf8) return kDocUnsupported;
        if (sem == kSemShow) {
          key = match_show_key(sink.len, sink.k0, sink.k1);
          if (key >= 0) { if (seen_show >> key & 1) PIE_HARD(kDocUnsupported); seen_show |= 1u << key; }
        } else {
          key = match_entry_key(sink.len, sink.k0, sink.k1);
          if (key >= 0) { if (seen_entry >> key & 1) PIE_HARD(kDocUnsupported); seen_entry |= 1u << key; }
        }
      } else {
        StrSink<false> sink{nullptr, 0, 0, 0};
        const int r = scan_string<false>(c, sink, &lone);
        if (r == kStrSyntax) return kDocDropped;
        if (r == kStrBadUtf8) return kDocUnsupported;
      }
      expect = kXColon;
      continue;
    }

    // ---- a value (expect is kXValue, or kXValueOrClose with something that is not ']').  Its role:
    //   heap >= 0   a text field of the table: string or null, anything else is a schema error
    //   is_item     an element of crew / actions: counted even when null
    //   num_role    1 createdAt, 2 archivedAt, 3 ts (a finite number or absent), 4 delaySec (number | null)
    int heap = -1, num_role = 0;
    bool is_item = false;
    int open_sem = kSemSkip;  // what an opening bracket starts
    const int at = sem;
    if (at == kSemShow) {
      if (key >= 0 && key < 7) heap = key;
      else if (key == kSkCreatedAt) num_role = 1;
      else if (key == kSkArchivedAt) num_role = 2;
      else if (key == kSkCrew && ch == '[') open_sem = kSemCrew;
      else if (key == kSkEntries && ch == '[') open_sem = kSemEntries;
    } else if (at == kSemEntry) {
      if (key >= 0 && key < 14) heap = kHeapEntry0 + key;
      else if (key == kEkTs) num_role = 3;
      else if (key == kEkDelaySec) num_role = 4;
      else if (key == kEkActions && ch == '[') open_sem = kSemActions;
    } else if (at == kSemCrew) {
      heap = kHeapCrew;
      is_item = true;
    } else if (at == kSemActions) {
      heap = kHeapActions;
      is_item = true;
    } else if (at == kSemEntries) {
      begin_entry();  // whatever the element is, it is a row: a non-object has no fields (pack_shows: `e = {}`)
      if (ch == '{') open_sem = kSemEntry;
      else end_entry();
    } else if (at == kSemTop) {
      if (ch == '{') { open_sem = kSemShow; top_is_object = true; }
      else if (ch == '[') { top_is_object = true; }  // typeof [] === 'object': an empty show, not a dropped row
    }
    key = -1;
    if (is_item) {
      const int items = at == kSemCrew ? kPlaneCrewItems : kPlaneActionItems;
      if (kFill) out.off[heap][cnt[items]] = (int32_t)cnt[heap];
      ++cnt[items];
    }

    if (ch == '{' || ch == '[') {
      if (depth >= kMaxDepth) return kDocUnsupported;
      if (heap >= 0 || num_role == 4) PIE_HARD(kDocSchema);
      if (ch == '{') kinds |= 1ull << depth;
      else kinds &= ~(1ull << depth);
      if (open_sem == kSemSkip) {
        if (sem != kSemSkip) { saved_sem = sem == kSemTop ? kSemDone : sem; skip_base = depth; sem = kSemSkip; }
      } else {
        sem = open_sem;
      }
      ++depth;
      expect = ch == '{' ? kXKeyOrClose : kXValueOrClose;
      continue;
    }
    if (ch == '"') {
      bool lone = false;
      StrSink<false> sink{nullptr, 0, 0, 0};
      if (heap >= 0 && kFill) sink.dst = out.data[heap] + cnt[heap];
      const int r = scan_string<false>(c, sink, &lone);
      if (r == kStrSyntax) return kDocDropped;
      if (r == kStrBadUtf8) return kDocUnsupported;
      if (heap >= 0) {
        cnt[heap] += sink.len;
        if (lone) PIE_HARD(kDocSchema);
      }
      if (num_role == 4) PIE_HARD(kDocSchema);  // delaySec: number or null
    } else if (ch == '-' || (ch >= '0' && ch <= '9')) {
      // the number parser wants the first byte back: a source that replays it
      struct Replay {
        DocCursor& c;
        int first;
        PIE_JW_HD int peek() const { return first >= 0 ? first : c.peek(); }
        PIE_JW_HD void next() { if (first >= 0) first = -1; else c.next(); }
      } src{c, ch};
      double v = 0.0;
      int r;
      if (num_role) r = parse_json_number_from<true>(src, pow5, &v);
      else r = parse_json_number_from<false>(src, pow5, &v);
      if (r == kNumSyntax) return kDocDropped;
      if (r == kNumUndecided) PIE_HARD(kDocUnsupported);
      if (heap >= 0) PIE_HARD(kDocSchema);
      if (num_role == 4) { e_delay = v; e_valid = true; }
      else if (num_role == 3) e_ts = jw_is_finite(v) ? v : jw_nan();
      else if (num_role && kFill) (num_role == 1 ? out.created_at : out.archived_at)[s] = jw_is_finite(v) ? v : jw_nan();
    } else {
      const char* lit = ch == 't' ? "rue" : ch == 'f' ? "alse" : ch == 'n' ? "ull" : nullptr;
      if (!lit) return kDocDropped;
      for (; *lit; ++lit) {
        if (c.peek() != *lit) return kDocDropped;
        c.next();
      }
      if (ch != 'n' && (heap >= 0 || num_role == 4)) PIE_HARD(kDocSchema);  // true / false where text / a number belongs
    }
    expect = depth == 0 ? kXEnd : kXCommaOrClose;
  }
#undef PIE_HARD
  if (hard) return hard;
  return top_is_object ? kDocOk : kDocDropped;
}

}  // namespace jw
}  // namespace pie
