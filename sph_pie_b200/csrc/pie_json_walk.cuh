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

PIE_JW_HD int jw_ctz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __ffsll((long long)x) - 1;
#else
  return __builtin_ctzll(x);
#endif
}

// ---- the document text, 8 aligned bytes at a time ---------------------------------------------------------------
struct DocCursor {
  const uint64_t* w;  // next word to request
  uint64_t cur, nxt;  // bytes not yet consumed (low byte first, zero above them); the word after
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
  PIE_JW_HD void refill() {
    if (words_left > 0) {
      cur = nxt;
      --words_left;
      left = words_left == 0 ? tail_bytes : 8;
      if (words_left > 0) nxt = jw_load_word(w++);
    } else {
      cur = 0;
    }
  }
  // address of the next byte the cursor will hand out
  PIE_JW_HD const uint8_t* position() const {
    const uint8_t* word = reinterpret_cast<const uint8_t*>(w - 1);  // the last word fetched: nxt, or cur at the end
    return words_left > 0 ? word - left : word + tail_bytes - left;
  }
  PIE_JW_HD int peek() const { return left > 0 ? (int)(cur & 0xFF) : -1; }
  PIE_JW_HD void next() {
    cur >>= 8;
    if (--left == 0) refill();
  }
  // The next up-to-8 bytes of the text wherever the cursor stands in its word: what is left of the current word with
  // the start of the next one behind it.  *avail = how many of them are text (8 except at the end of the document).
  // A string is then scanned in ceil(length / 8) steps whatever its phase against the aligned words — lanes that walk
  // documents of the same make stay in step even after one value was a byte longer in one of them.
  PIE_JW_HD uint64_t window(int* avail) const {
    uint64_t w8 = cur;
    int n = left;
    if (left < 8 && words_left > 0) {
      w8 |= nxt << (8 * left);  // left >= 1 here: a cursor with bytes to come never rests on an empty word
      n = left + (words_left == 1 ? tail_bytes : 8);
      if (n > 8) n = 8;
    }
    *avail = n;
    return w8;
  }
  // steps over n bytes of the window
  PIE_JW_HD void advance(int n) {
    if (n < left) {
      cur >>= 8 * n;
      left -= n;
      return;
    }
    n -= left;
    left = 0;
    refill();
    if (n > 0) {  // into the next word: n < 8 of its bytes
      cur >>= 8 * n;
      left -= n;
      if (left == 0) refill();
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

enum ShowKey { kSkCrew = 7, kSkCreatedAt = 8, kSkArchivedAt = 9, kSkEntries = 10, kSkUpdatedAt = 11, kSkDeletedAt = 12 };
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
  if (PIE_KEY_IS("updatedAt")) return kSkUpdatedAt;
  if (PIE_KEY_IS("deletedAt")) return kSkDeletedAt;
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

// ---- the walk -----------------------------------------------------------------------------------------------------
// What the second walk knows of an entry, written as ONE 96-byte row (three full 32-byte sectors) instead of 18
// words scattered over 18 arrays — every lane keeps a sector open in L2 for each stream it writes, and ~130 k lanes
// x 30 streams do not fit; json_ingest.cu turns the rows into the table's columns with coalesced accesses.
struct alignas(32) EntryRow {
  int32_t off[14];       // where the entry's text starts in each of the 14 entry heaps
  int32_t actions_list;  // first item of its actions
  int32_t pad0;
  double delay;
  double ts;
  uint32_t valid;
  uint32_t pad1[3];
};
static_assert(sizeof(EntryRow) == 96, "three sectors");

struct IngestOut {
  EntryRow* rows;  // nullptr: the entry columns are written directly (the host build of the walk)
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
  // the four time fields of the show (PIE_TF_*): value arrays (created_at / archived_at above are two of them; the
  // other two may be nullptr), what a field holds when it is not a finite number (may be nullptr), and the text the
  // documents are in — a string in a time field is recorded by the place of its first character
  double* time_val[PIE_TF_COUNT];
  uint8_t* time_kind;
  const uint8_t* text;
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
  kDocRunning = 100,  // step(): the document is not through yet
};

PIE_JW_HD int hex_value(int c) {
  if (c >= '0' && c <= '9') return c - '0';
  c |= 0x20;
  if (c >= 'a' && c <= 'f') return c - 'a' + 10;
  return -1;
}

// The walk is a resumable state machine: step() takes one token of the grammar and, when that token opens a string,
// the string to its closing quote (8 bytes at a time); step_member() a key with its scalar value.  The kernels run
// it in a loop in which every lane fetches its next document as soon as its current one ends.
//   cnt[p]: pass 1 = what the document adds to plane p; pass 2 = the running position in plane p
template <bool kFill>
struct DocWalker {
  DocCursor c;
  int64_t s;  // the document (row of the show columns)
  // grammar
  uint64_t kinds;  // bit d: the container open at depth d+1 is an object
  int depth, sem, saved_sem, skip_base, expect;
  int key;  // key of the member whose value comes next (show / entry objects), -1 = not one the table holds
  uint32_t seen_show, seen_entry;
  int hard;  // first hard error met; reported only if the document is JSON at all
  bool top_is_object;
  // the open entry
  double e_delay, e_ts;
  bool e_valid;
  uint32_t e_row;
  // the open string
  int str_mode;  // 0 none, 1 key, 2 value
  int str_heap;  // value: the heap that takes it, -1 = only recognised
  uint32_t str_len;
  uint64_t k0, k1;  // key: its first 16 bytes
  uint8_t* dst;     // value, pass 2: where its bytes go
  uint32_t high;    // pending high surrogate of a \uD8xx escape
  bool lone;        // an escape named a surrogate without its partner (a JS string that has no UTF-8 form)
  bool str_closed;  // the closing quote has been taken, close_string() is due

  PIE_JW_HD void begin(const uint8_t* text, int64_t from, int64_t to, int64_t doc) {
    c.open(text, from, to);
    s = doc;
    kinds = 0;
    depth = 0;
    sem = kSemTop;
    saved_sem = kSemDone;
    skip_base = 0;
    expect = kXValue;
    key = -1;
    seen_show = seen_entry = 0;
    hard = 0;
    top_is_object = false;
    e_delay = e_ts = 0.0;
    e_valid = false;
    e_row = 0;
    str_mode = 0;
    str_heap = -1;
    str_len = 0;
    k0 = k1 = 0;
    dst = nullptr;
    high = 0;
    lone = false;
    str_closed = false;
  }
  PIE_JW_HD void set_hard(int code) { if (!hard) hard = code; }

  PIE_JW_HD void put(uint32_t b) {
    if (str_mode == 1) {
      if (str_len < 8) k0 |= (uint64_t)b << (8 * str_len);
      else if (str_len < 16) k1 |= (uint64_t)b << (8 * (str_len - 8));
    } else if (kFill && dst) {
      store_bytes((uint64_t)b, 1);
    }
    ++str_len;
  }
  // Pass 2, a value's bytes on their way out: n (1..8) bytes, byte by byte at dst.  (Gathering them in a register and
  // storing aligned 32-bit words was measured: 30 % faster when a warp's lanes walk identical documents — byte stores
  // are one L2 transaction per byte and lane — and 30 % slower when they do not, which is the case that matters: the
  // alignment arithmetic is more instructions, and a diverged warp pays for every instruction several times.)
  PIE_JW_HD void store_bytes(uint64_t chunk, int n) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k < n) dst[k] = (uint8_t)(chunk >> (8 * k));
    dst += n;
  }
  PIE_JW_HD void flush_bytes() {}
  // n (1..8) plain bytes, the low bytes of `chunk` (zero above them)
  PIE_JW_HD void put_run(uint64_t chunk, int n) {
    if (str_mode == 1) {
      if (str_len < 8) {
        k0 |= chunk << (8 * str_len);
        if (str_len > 0 && str_len + n > 8) k1 |= chunk >> (8 * (8 - str_len));
      } else if (str_len < 16) {
        k1 |= chunk << (8 * (str_len - 8));
      }
    } else if (kFill && dst) {
      store_bytes(chunk, n);
    }
    str_len += (uint32_t)n;
  }

  PIE_JW_HD void begin_entry(uint32_t (&cnt)[kPlanes], const IngestOut& out) {
    e_row = cnt[kPlaneEntries]++;
    seen_entry = 0;
    e_delay = 0.0;
    e_valid = false;
    e_ts = jw_nan();
    if (kFill) {
      if (out.rows) {
#if defined(__CUDA_ARCH__)
        uint4* row = reinterpret_cast<uint4*>(out.rows + e_row);
        row[0] = make_uint4(cnt[kHeapEntry0 + 0], cnt[kHeapEntry0 + 1], cnt[kHeapEntry0 + 2], cnt[kHeapEntry0 + 3]);
        row[1] = make_uint4(cnt[kHeapEntry0 + 4], cnt[kHeapEntry0 + 5], cnt[kHeapEntry0 + 6], cnt[kHeapEntry0 + 7]);
        row[2] = make_uint4(cnt[kHeapEntry0 + 8], cnt[kHeapEntry0 + 9], cnt[kHeapEntry0 + 10], cnt[kHeapEntry0 + 11]);
        row[3] = make_uint4(cnt[kHeapEntry0 + 12], cnt[kHeapEntry0 + 13], cnt[kPlaneActionItems], 0u);
#endif
      } else {
#pragma unroll
        for (int h = kHeapEntry0; h < kHeapEntry0 + 14; ++h) out.off[h][e_row] = (int32_t)cnt[h];
        out.actions_list[e_row] = (int32_t)cnt[kPlaneActionItems];
      }
    }
  }
  PIE_JW_HD void end_entry(const IngestOut& out) {
    if (kFill) {
      if (out.rows) {
#if defined(__CUDA_ARCH__)
        uint4* row = reinterpret_cast<uint4*>(out.rows + e_row);
        const unsigned long long d = (unsigned long long)__double_as_longlong(e_delay), t = (unsigned long long)__double_as_longlong(e_ts);
        row[4] = make_uint4((uint32_t)d, (uint32_t)(d >> 32), (uint32_t)t, (uint32_t)(t >> 32));
        row[5] = make_uint4(e_valid ? 1u : 0u, 0u, 0u, 0u);
#endif
      } else {
        out.delay_sec[e_row] = e_delay;
        out.delay_valid[e_row] = e_valid ? 1 : 0;
        out.entry_ts[e_row] = e_ts;
      }
    }
  }
  // a value is through: what may follow it (a comma right behind it is taken at once)
  PIE_JW_HD void after_value() {
    if (depth == 0) {
      expect = kXEnd;
      return;
    }
    expect = kXCommaOrClose;
    if (c.peek() == ',') {
      c.next();
      expect = ((kinds >> (depth - 1)) & 1) ? kXKey : kXValue;
    }
  }

  // the string that just closed: a key names the member, a value is counted into its heap
  PIE_JW_HD void close_string(uint32_t (&cnt)[kPlanes]) {
    const int mode = str_mode;
    str_mode = 0;
    str_closed = false;
    if (mode == 1) {
      if (sem == kSemShow) {
        key = match_show_key(str_len, k0, k1);
        if (key >= 0) { if (seen_show >> key & 1) set_hard(kDocUnsupported); seen_show |= 1u << key; }
      } else if (sem == kSemEntry) {
        key = match_entry_key(str_len, k0, k1);
        if (key >= 0) { if (seen_entry >> key & 1) set_hard(kDocUnsupported); seen_entry |= 1u << key; }
      }
      expect = kXColon;
      if (c.peek() == ':') { c.next(); expect = kXValue; }
    } else {
      if (kFill && dst) flush_bytes();
      if (str_heap >= 0) {
        cnt[str_heap] += str_len;
        if (lone) set_hard(kDocSchema);
      }
      after_value();
    }
  }
  PIE_JW_HD bool in_string() const { return str_mode != 0 && !str_closed; }

  // ---- inside a string: up to 8 plain bytes at once, then whatever ends the run.  Everything but the plain run —
  // escapes, UTF-8 sequences, a surrogate left alone — is put together in one word (`extra`, at most 7 bytes a step)
  // and leaves through ONE put_run at the end: the byte-store code exists twice in the kernel, not twenty times.
  PIE_JW_HD int string_step(uint32_t (&cnt)[kPlanes]) {
    if (c.left == 0) return kDocDropped;  // the text ends inside the string
    uint64_t extra = 0;
    int extra_n = 0;
#define PIE_EXTRA(b) do { extra |= (uint64_t)((b) & 0xFFu) << (8 * extra_n); ++extra_n; } while (0)
#define PIE_EXTRA3(cp) do { PIE_EXTRA(0xE0 | ((cp) >> 12)); PIE_EXTRA(0x80 | (((cp) >> 6) & 0x3F)); PIE_EXTRA(0x80 | ((cp) & 0x3F)); } while (0)
    if (high && c.peek() != '\\') {
      // a high surrogate is pending and no escape follows that could complete it: it stays alone (a JS string without
      // a UTF-8 form) — its three bytes now, the rest of the string from the next step on
      lone = true;
      PIE_EXTRA3(high);
      high = 0;
    }
    int avail;
    const uint64_t x = c.window(&avail);
    const uint64_t k80 = 0x8080808080808080ull, k01 = 0x0101010101010101ull;
    const uint64_t q = x ^ (k01 * 0x22), bs = x ^ (k01 * 0x5C);
    // 0x80 in every byte that is >= 0x80, < 0x20, '"' or '\\' — exact up to and including the first such byte
    const uint64_t special = (x & k80) | ((x - k01 * 0x20) & ~x & k80) | ((q - k01) & ~q & k80) | ((bs - k01) & ~bs & k80);
    int n = special ? (jw_ctz64(special) >> 3) : 8;
    if (n > avail) n = avail;
    if (extra_n) n = 0;  // (rare) the lone surrogate's bytes go first; the plain run waits for the next step
    bool at_special = true;
    if (n > 0) {
      put_run(n == 8 ? x : (x & ((1ull << (8 * n)) - 1)), n);
      c.advance(n);
      // the rest of this word may start with the byte that ended the run: take it in the same step
      at_special = false;
      if (c.left != 0 && n != 8) {
        const uint64_t y = c.cur & 0xFF;
        at_special = y == '"' || y == '\\' || y < 0x20 || y >= 0x80;
      }
    } else if (extra_n) {
      at_special = false;
    }
    if (at_special) {
      const int ch = c.peek();
      c.next();
      if (ch == '"') {  // the closing quote: what follows from it (close_string) is left to the caller, after its loop
        str_closed = true;
      } else if (ch == '\\') {
        const int e = c.peek();
        if (e < 0) return kDocDropped;
        c.next();
        uint32_t cp = 0;
        bool emit = true;
        if (e == 'u') {
#pragma unroll 1
          for (int k = 0; k < 4; ++k) {
            const int h = hex_value(c.peek());
            if (h < 0) return kDocDropped;
            c.next();
            cp = cp * 16 + (uint32_t)h;
          }
          if (high) {
            if (cp >= 0xDC00 && cp <= 0xDFFF) {
              cp = 0x10000 + ((high - 0xD800) << 10) + (cp - 0xDC00);
              high = 0;
              PIE_EXTRA(0xF0 | (cp >> 18));
              PIE_EXTRA(0x80 | ((cp >> 12) & 0x3F));
              PIE_EXTRA(0x80 | ((cp >> 6) & 0x3F));
              PIE_EXTRA(0x80 | (cp & 0x3F));
              emit = false;
            } else {  // the pending one stays alone; this unit starts over
              lone = true;
              PIE_EXTRA3(high);
              high = 0;
            }
          }
          if (emit) {
            if (cp >= 0xD800 && cp <= 0xDBFF) { high = cp; emit = false; }
            else if (cp >= 0xDC00 && cp <= 0xDFFF) lone = true;
          }
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
            default: return kDocDropped;
          }
          if (high) {  // \uD8xx followed by a simple escape
            lone = true;
            PIE_EXTRA3(high);
            high = 0;
          }
        }
        if (emit) {
          if (cp < 0x80) {
            PIE_EXTRA(cp);
          } else if (cp < 0x800) {
            PIE_EXTRA(0xC0 | (cp >> 6));
            PIE_EXTRA(0x80 | (cp & 0x3F));
          } else {
            PIE_EXTRA3(cp);
          }
        }
      } else if (ch < 0x20) {
        return kDocDropped;  // control characters must be escaped
      } else {
        // ch >= 0x80.  UTF-8: well-formed sequences only (Unicode table 3-7); anything else is reported, and the
        // walk goes on as if the bytes were text so that "is it JSON at all" is still answered
        PIE_EXTRA(ch);
        int need = 0;
        int lo = 0x80, hi = 0xBF;
        if (ch < 0xC2) set_hard(kDocUnsupported);
        else if (ch < 0xE0) need = 1;
        else if (ch < 0xF0) { need = 2; if (ch == 0xE0) lo = 0xA0; if (ch == 0xED) hi = 0x9F; }
        else if (ch < 0xF5) { need = 3; if (ch == 0xF0) lo = 0x90; if (ch == 0xF4) hi = 0x8F; }
        else set_hard(kDocUnsupported);
#pragma unroll 1
        for (int k = 0; k < need; ++k) {
          const int b = c.peek();
          if (b < lo || b > hi) { set_hard(kDocUnsupported); break; }
          c.next();
          PIE_EXTRA(b);
          lo = 0x80;
          hi = 0xBF;
        }
      }
    }
    if (extra_n) put_run(extra, extra_n);
#undef PIE_EXTRA3
#undef PIE_EXTRA
    return kDocRunning;
  }

  // ---- between strings: one token of the grammar
  PIE_JW_HD int token_step(uint32_t (&cnt)[kPlanes], const IngestOut& out, const Pow5Table& pow5) {
    int ch = c.peek();
    while (ch == ' ' || ch == '\n' || ch == '\r' || ch == '\t') { c.next(); ch = c.peek(); }
    if (ch < 0) {
      if (expect != kXEnd) return kDocDropped;
      if (hard) return hard;
      return top_is_object ? kDocOk : kDocDropped;
    }
    if (expect == kXEnd) return kDocDropped;  // something after the value
    c.next();

    // punctuation the grammar state asks for
    if (expect == kXColon) {
      if (ch != ':') return kDocDropped;
      expect = kXValue;
      return kDocRunning;
    }
    bool closing = false;
    if (expect == kXCommaOrClose) {
      const bool in_object = (kinds >> (depth - 1)) & 1;
      if (ch == ',') { expect = in_object ? kXKey : kXValue; return kDocRunning; }
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
      else if (sem == kSemEntry) { end_entry(out); sem = kSemEntries; }
      else if (sem == kSemActions) sem = kSemEntry;
      else if (sem == kSemShow) sem = kSemDone;
      else sem = kSemShow;  // crew, entries
      after_value();
      return kDocRunning;
    }

    // a key
    if (expect == kXKeyOrClose || expect == kXKey) {
      if (ch != '"') return kDocDropped;
      str_mode = 1;
      str_heap = -1;
      str_len = 0;
      k0 = k1 = 0;
      dst = nullptr;
      high = 0;
      lone = false;
      key = -1;
      return kDocRunning;
    }

    // a value (expect is kXValue, or kXValueOrClose with something that is not ']').  Its role:
    //   heap >= 0   a text field of the table: string or null, anything else is a schema error
    //   is_item     an element of crew / actions: counted even when null
    //   num_role    1 createdAt, 2 archivedAt, 3 ts (a finite number or absent), 4 delaySec (number | null)
    int heap = -1, num_role = 0;
    int time_field = -1;  // the value belongs to one of the show's time fields (PIE_TF_*)
    bool is_item = false;
    int open_sem = kSemSkip;  // what an opening bracket starts
    const int at = sem;
    if (at == kSemShow) {
      if (key >= 0 && key < 7) heap = key;
      else if (key == kSkCreatedAt) { num_role = 1; time_field = PIE_TF_CREATED; }
      else if (key == kSkArchivedAt) { num_role = 2; time_field = PIE_TF_ARCHIVED; }
      else if (key == kSkUpdatedAt) { num_role = 5; time_field = PIE_TF_UPDATED; }
      else if (key == kSkDeletedAt) { num_role = 6; time_field = PIE_TF_DELETED; }
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
      begin_entry(cnt, out);  // whatever the element is, it is a row: a non-object has no fields (pack_shows: `e = {}`)
      if (ch == '{') open_sem = kSemEntry;
      else end_entry(out);
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

    if (kFill && time_field >= 0) {  // what the field holds, for _getTimestamp's coercions (rare keys: four a document)
      const int kind = (ch == '{' || ch == '[') ? PIE_TK_OTHER : ch == '"' ? PIE_TK_STRING : ch == 'n' ? PIE_TK_NULL
                       : ch == 't' ? PIE_TK_TRUE : ch == 'f' ? PIE_TK_FALSE : PIE_TK_NUMBER;  // a number: settled below
      if (out.time_kind) out.time_kind[s * PIE_TF_COUNT + time_field] = (uint8_t)kind;
      if (ch == '"' && out.time_val[time_field])  // NaN whose payload is where the string's text starts
        out.time_val[time_field][s] = np_bits_to_double(0x7ff8000000000000ull | ((uint64_t)(c.position() - out.text) & 0x7ffffffffffffull));
    }
    if (ch == '{' || ch == '[') {
      if (depth >= kMaxDepth) return kDocUnsupported;
      if (heap >= 0 || num_role == 4) set_hard(kDocSchema);
      if (ch == '{') kinds |= 1ull << depth;
      else kinds &= ~(1ull << depth);
      if (open_sem == kSemSkip) {
        if (sem != kSemSkip) { saved_sem = sem == kSemTop ? kSemDone : sem; skip_base = depth; sem = kSemSkip; }
      } else {
        sem = open_sem;
      }
      ++depth;
      expect = ch == '{' ? kXKeyOrClose : kXValueOrClose;
      return kDocRunning;
    }
    if (ch == '"') {
      str_mode = 2;
      str_heap = heap;
      str_len = 0;
      dst = (kFill && heap >= 0) ? out.data[heap] + cnt[heap] : nullptr;
      high = 0;
      lone = false;
      if (num_role == 4) set_hard(kDocSchema);  // delaySec: number or null
      return kDocRunning;
    }
    if (ch == '-' || (ch >= '0' && ch <= '9')) {
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
      if (r == kNumUndecided) set_hard(kDocUnsupported);
      if (heap >= 0) set_hard(kDocSchema);
      if (num_role == 4) { e_delay = v; e_valid = true; }
      else if (num_role == 3) e_ts = jw_is_finite(v) ? v : jw_nan();
      else if (num_role && kFill) {
        if (out.time_val[time_field]) out.time_val[time_field][s] = jw_is_finite(v) ? v : jw_nan();
        if (!jw_is_finite(v) && out.time_kind) out.time_kind[s * PIE_TF_COUNT + time_field] = PIE_TK_NONFINITE;
      }
    } else {
      const char* lit = ch == 't' ? "rue" : ch == 'f' ? "alse" : ch == 'n' ? "ull" : nullptr;
      if (!lit) return kDocDropped;
      for (; *lit; ++lit) {
        if (c.peek() != *lit) return kDocDropped;
        c.next();
      }
      if (ch != 'n' && (heap >= 0 || num_role == 4)) set_hard(kDocSchema);  // true / false where text / a number belongs
    }
    after_value();
    return kDocRunning;
  }

  // One step: the next token, and when that opens a string (or the walk is inside one) the string to its end.
  PIE_JW_HD int step(uint32_t (&cnt)[kPlanes], const IngestOut& out, const Pow5Table& pow5) {
    int r = kDocRunning;
    if (str_mode == 0) {
      r = token_step(cnt, out, pow5);
      if (r != kDocRunning) return r;
    }
#pragma unroll 1
    while (in_string()) {
      r = string_step(cnt);
      if (r != kDocRunning) return r;
    }
    if (str_closed) close_string(cnt);
    return r;
  }
  // One member of an object (key, colon, and the value when it is a scalar) or one element of an array per step:
  // fewer turns of the caller's loop than step().
  PIE_JW_HD int step_member(uint32_t (&cnt)[kPlanes], const IngestOut& out, const Pow5Table& pow5) {
    int r = kDocRunning;
#pragma unroll 1
    for (int part = 0; part < 2; ++part) {  // the key, then its value
      if (str_mode == 0) {
        r = token_step(cnt, out, pow5);
        if (r != kDocRunning || str_mode == 0) return r;
      }
      const bool was_key = str_mode == 1;
#pragma unroll 1
      while (in_string()) {
        r = string_step(cnt);
        if (r != kDocRunning) return r;
      }
      if (str_closed) close_string(cnt);
      if (!was_key || expect != kXValue) break;
    }
    return r;
  }
};

}  // namespace jw
}  // namespace pie
