// JSON ingest, the warp-cooperative path: ONE WARP PER DOCUMENT, no lane ever walks a grammar byte by byte.
//
// The thread-per-document walk (pie_json_walk.cuh) spends its time on divergence: 32 lanes in 32 different tokens of
// 32 different documents execute the union of their paths (10 of 32 lanes active, profiles/ncu_r01_ingest_summary.md).
// Here a warp takes a document in two stages that keep all lanes on the same instruction:
//
//   stage 1  structural index (lane = 32 bytes of text, 1 KB per turn): byte classes by SIMD-in-register compares on
//            a byte-transposed copy of the lane's 32 bytes (so that the per-word flags merge into 32-bit position
//            masks by shifts, without a movemask per word), escaped characters by the carry trick of simdjson's
//            stage 1 with the carry passed between lanes, quote parity by prefix-xor + ballot, bracket depth /
//            entry index / colon and quote counts by warp scans.  It leaves two lists in shared memory, so that stage
//            2 never searches: the quotes in text order (each with "an escape starts between the quote before and this
//            one"), and the members by their colons (with the key's closing quote and the entry they are in).
//   stage 2  member-parallel projection (lane = one `"key":value` member): key by one probe of a perfect hash and a
//            16-byte compare, value by its first byte; lengths of the text values are ranked per heap in document
//            order (match_any + shuffles).  Pass 1 checks, counts and writes down what it learnt as records (where
//            every value is and where it goes; of a number only that its shape is one the parser always decides);
//            pass 2 of a document with records is a scatter that converts the numbers on its way.
//
// The path decides ONLY documents of the shape the provider writes (sqlProvider.js:361-409 _normalizeShow /
// _normalizeEntry through JSON.stringify, :682 / :696):
//     doc     := '{' [ member (',' member)* ] '}'                        no whitespace outside strings
//     member  := string ':' ( scalar | string | '[' strings ']' when the key is crew | '[' entries ']' when entries )
//     entry   := '{' the 17 known keys exactly once each, any order, other keys with scalar / string values '}'
//                actions -> '[' strings ']', text keys -> string | null, delaySec -> number | null
// recognised exactly (adjacency rules on the class masks + depth rules, see build_index / the member loop; every
// string and every scalar is accounted for), at most 16 KB and 63 entries.  EVERYTHING ELSE IS DECLINED, never guessed: the document
// goes on the list of the thread-per-document walk, which stays the complete ECMA-404 recogniser and the only place
// that reports errors.  Declining is always safe; accepting is what the damage tests in tests/test_gpu_ingest.py hold
// to the oracle byte by byte.
#pragma once
#include "pie_device.cuh"
#include "pie_json_walk.cuh"

// build-time experiment switches (scripts/_variants): which helpers are out of line, the number fast path
// Measured on B200 (2^19 documents, pass 1): everything inline, generic number parser 9.6 ms; number fast path 10.0;
// parse_number_at out of line 10.5 / 9.9 without the fast path; every helper out of line 11.9.  Pass 1 waits for
// instructions (8 "no instruction" stalls per issue: > 64 KB of SASS, 28 warps in different places of it), so anything
// that adds code or jumps far costs more than the instructions it saves.
#ifndef PIE_JF_NOINLINE_MASK
#define PIE_JF_NOINLINE_MASK 0
#endif
#ifndef PIE_JF_FASTNUM
#define PIE_JF_FASTNUM 0
#endif
// 1: pass 1 converts only numbers of the everyday form (-?digits(.digits)?, at most 15 digits, within 16 bytes) and declines
// a document with any other number to the walk — no Eisel-Lemire, no cursor in the kernel, 56 registers instead of 72.
// Measured and NOT shipped: a quarter of the bench's documents hold a 17-digit delaySec (54.976914585768995), which
// needs the general parser, and the walk's tail on them costs more than the smaller kernel saves (26.8 + 21.0 ms).
#ifndef PIE_JF_NUMBERS_EVERYDAY_ONLY
#define PIE_JF_NUMBERS_EVERYDAY_ONLY 0
#endif
#define PIE_JF_INL(bit) ((PIE_JF_NOINLINE_MASK >> (bit)) & 1)
#if PIE_JF_INL(0)
#define PIE_JF_ESC_FN __device__ __noinline__
#else
#define PIE_JF_ESC_FN __device__ __forceinline__
#endif
#if PIE_JF_INL(1)
#define PIE_JF_ITEMS_FN __device__ __noinline__
#else
#define PIE_JF_ITEMS_FN __device__ __forceinline__
#endif
#if PIE_JF_INL(2)
#define PIE_JF_NUM_FN __device__ __noinline__
#else
#define PIE_JF_NUM_FN __device__ __forceinline__
#endif

namespace pie {
namespace jf {

using namespace jw;

constexpr int kFastMaxBytes = 16384;   // alignment skip + document length a warp takes (positions are 14 bits)
constexpr int kFastMaxEntries = 63;    // entry + 1 is 6 bits of a member's word
// What a warp's lists hold.  Two configurations of the same kernels: the one nearly every document fits (6 KB of shared
// memory a warp: 7 warps x 5 CTAs per SM), and a roomy one for the long documents the first one declines, so that a show
// of 40 entries does not fall to the thread-per-document walk (which needs ~1 ms per KB of ONE document).
struct CapsSmall {
  static constexpr int kWarps = 7;
  static constexpr int kQuotes = 1472;   // twice the strings: ~0.18 per byte in the provider's documents, so ~8.5 KB of them
  static constexpr int kMembers = 480;
  static constexpr int kNumbers = 64;
  static constexpr int kArrays = 32;
};
struct CapsBig {
  static constexpr int kWarps = 4;
  static constexpr int kQuotes = 3584;   // a quote's index is 12 bits of a member's word
  static constexpr int kMembers = 1152;
  static constexpr int kNumbers = 128;
  static constexpr int kArrays = 64;
};
constexpr int kFastMaxEsc = 32;
constexpr int kFastMaxArrays = 32;
constexpr int kInlineCopy = 16;  // pass 2 without records: longer values are copied by the whole warp
constexpr int kLaneCopy = 16;    // pass 2 with records: the same (lanes copying their own values up to 96 bytes, 8 bytes a turn,
                                 // was measured: 10.2 ms instead of 8.3 — the longest value of a turn decides)
constexpr uint32_t kFull = 0xffffffffu;
constexpr uint32_t kAllEntryKeys = 0x1ffffu;  // the 17 keys of an entry
constexpr uint32_t kPosMask = 0x3fffu;
constexpr uint8_t kRouteSlow = 0, kRouteFast = 1, kRouteRecords = 2, kRouteFastBig = 3,  // Big: pass 2 parses it again with CapsBig
                  kRouteLong = 4;  // too long for the warp path: the walk's, known before pass 1 starts

// Stage 1 leaves two lists, so that stage 2 never searches: the quotes that open or close a string in text order
// (string i is the pair 2 i, 2 i + 1), and the members by their colons with what stage 2 would otherwise look up.
template <class Caps>
struct alignas(16) WarpShared {
  uint16_t qpos[Caps::kQuotes];    // position | (an escape starts between the quote before and this one) << 15
  uint32_t mem[Caps::kMembers];    // colon position | index in qpos of the key's closing quote << 14 | (entry + 1, 0 = the show) << 26
  uint32_t cnt[kPlanes];           // measure: what the document adds to each plane; fill: the running position in it
  uint32_t seen[kFastMaxEntries + 1];
  uint32_t num[Caps::kNumbers];    // position | entry << 14 | role << 21
  union {
    struct {                          // pass 2 without records: values with escapes, unescaped at the end of the document
      uint32_t esc_src[kFastMaxEsc];  // position of the first byte | raw length << 16
      uint32_t esc_dst[kFastMaxEsc];  // where it goes in its heap
      uint8_t esc_heap[kFastMaxEsc];
    };
    uint32_t arr[Caps::kArrays][3];   // pass 1 with records: the crew / actions arrays, their elements' records come last
  };
  uint32_t n_num, n_esc, seen_show, n_arr, n_items, pad[3];
};

// ---- records ------------------------------------------------------------------------------------------------------
// What pass 1 knows of a document when it is through with it is everything pass 2 needs, so it writes it down: one
// 8-byte record per member, then one per number and one per element of crew / actions, in a pool inside the scratch
// area (kPoolUnitsPerDoc 8-byte units per document on average; a document that finds the pool full is parsed again by
// pass 2 as before).  Pass 2 of such a document is a scatter: no index, no keys, no grammar.
//   member  lo = first byte | raw length << 14 | has escapes << 28        hi = heap | entry << 5 | offset in the document's
//           part of the heap << 12 | kind << 29
//   number  lo = role | entry << 3 | position << 10 (pass 2 converts it from the text)
//   element lo as a member's; hi = heap | offset << 5 | index in the document's part of the list << 19
enum RecKind : uint32_t { kRecNone = 0, kRecText, kRecArray, kRecNullDelay, kRecNanTs, kRecTimeKind, kRecTimeString };
constexpr int kPoolUnitsPerDoc = 288;
struct DocRec {
  unsigned long long members_at, extra_at;  // units into the pool
  uint32_t members, numbers, items, pad;
};
struct RecCtx {
  unsigned long long* pool;    // nullptr: no records
  unsigned long long* cursor;  // units handed out
  unsigned long long capacity;
  DocRec* doc_rec;
};

// keys by a perfect hash of (first 8 bytes, length): one probe, one compare, no chain of 17 compares that parts the lanes
struct KeySlot {
  uint64_t k0, k1;
  int32_t len, id;
};
constexpr uint32_t kShowKeyMul = 0x2265b1f5u, kEntryKeyMul = 0xab99254bu;
__host__ __device__ constexpr uint32_t key_slot(uint64_t k0, uint32_t len, uint32_t mul) {
  return ((((uint32_t)k0 ^ (uint32_t)(k0 >> 32)) + len) * mul) >> 27;
}
struct KeyTables {
  KeySlot show[32], entry[32];
};
template <int L>
__host__ __device__ constexpr void key_put(KeySlot (&t)[32], uint32_t mul, const char (&lit)[L], int id, bool* clash) {
  const uint64_t k0 = key_word(lit, 0), k1 = key_word(lit, 8);
  KeySlot& sl = t[key_slot(k0, L - 1, mul)];
  if (sl.len != 0) *clash = true;
  sl.k0 = k0;
  sl.k1 = k1;
  sl.len = L - 1;
  sl.id = id;
}
__host__ __device__ constexpr KeyTables make_key_tables(bool* clash) {
  KeyTables t{};
  for (int i = 0; i < 32; ++i) {
    t.show[i] = KeySlot{0, 0, 0, -1};
    t.entry[i] = KeySlot{0, 0, 0, -1};
  }
  // the ids are those of match_show_key / match_entry_key (pie_json_walk.cuh)
  key_put(t.show, kShowKeyMul, "id", 0, clash);
  key_put(t.show, kShowKeyMul, "date", 1, clash);
  key_put(t.show, kShowKeyMul, "time", 2, clash);
  key_put(t.show, kShowKeyMul, "label", 3, clash);
  key_put(t.show, kShowKeyMul, "leadPilot", 4, clash);
  key_put(t.show, kShowKeyMul, "monkeyLead", 5, clash);
  key_put(t.show, kShowKeyMul, "notes", 6, clash);
  key_put(t.show, kShowKeyMul, "crew", kSkCrew, clash);
  key_put(t.show, kShowKeyMul, "createdAt", kSkCreatedAt, clash);
  key_put(t.show, kShowKeyMul, "archivedAt", kSkArchivedAt, clash);
  key_put(t.show, kShowKeyMul, "entries", kSkEntries, clash);
  key_put(t.show, kShowKeyMul, "updatedAt", kSkUpdatedAt, clash);
  key_put(t.show, kShowKeyMul, "deletedAt", kSkDeletedAt, clash);
  key_put(t.entry, kEntryKeyMul, "id", 0, clash);
  key_put(t.entry, kEntryKeyMul, "unitId", 1, clash);
  key_put(t.entry, kEntryKeyMul, "planned", 2, clash);
  key_put(t.entry, kEntryKeyMul, "launched", 3, clash);
  key_put(t.entry, kEntryKeyMul, "status", 4, clash);
  key_put(t.entry, kEntryKeyMul, "primaryIssue", 5, clash);
  key_put(t.entry, kEntryKeyMul, "subIssue", 6, clash);
  key_put(t.entry, kEntryKeyMul, "otherDetail", 7, clash);
  key_put(t.entry, kEntryKeyMul, "severity", 8, clash);
  key_put(t.entry, kEntryKeyMul, "rootCause", 9, clash);
  key_put(t.entry, kEntryKeyMul, "operator", 10, clash);
  key_put(t.entry, kEntryKeyMul, "batteryId", 11, clash);
  key_put(t.entry, kEntryKeyMul, "commandRx", 12, clash);
  key_put(t.entry, kEntryKeyMul, "notes", 13, clash);
  key_put(t.entry, kEntryKeyMul, "actions", kEkActions, clash);
  key_put(t.entry, kEntryKeyMul, "delaySec", kEkDelaySec, clash);
  key_put(t.entry, kEntryKeyMul, "ts", kEkTs, clash);
  return t;
}
constexpr bool key_tables_clash() {
  bool clash = false;
  make_key_tables(&clash);
  return clash;
}
static_assert(!key_tables_clash(), "the key hashes must be perfect");

// what every warp of a CTA reads per value: the table's pointers and the key tables, once in shared memory (indexing
// the kernel parameter / constant memory with a lane-dependent index would serialise on the constant cache)
struct TablePointers {
  uint8_t* data[kHeaps];
  int32_t* off[kHeaps];
  KeyTables keys;
};

struct Doc {
  const uint8_t* ab;  // 32-byte aligned address at or before the document's first byte
  int skip, span;     // the document is ab[skip .. span)
  int nwords;
};

// the first 16 bytes of a key as the walk collects them (len <= 16)
__device__ __forceinline__ void load_key(const uint8_t* p, int len, const uint8_t* limit, uint64_t* k0, uint64_t* k1) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const unsigned long long* w = reinterpret_cast<const unsigned long long*>(a & ~(uintptr_t)7);
  const int sh = (int)(a & 7) * 8;
  const uint64_t w0 = __ldg(w);
  const uint64_t w1 = reinterpret_cast<const uint8_t*>(w + 1) < limit ? __ldg(w + 1) : 0ull;
  const uint64_t w2 = reinterpret_cast<const uint8_t*>(w + 2) < limit ? __ldg(w + 2) : 0ull;
  uint64_t a0 = w0, a1 = w1;
  if (sh) {
    a0 = (w0 >> sh) | (w1 << (64 - sh));
    a1 = (w1 >> sh) | (w2 << (64 - sh));
  }
  if (len < 8) {
    a0 &= (1ull << (8 * len)) - 1ull;
    a1 = 0;
  } else if (len < 16) {
    a1 &= (1ull << (8 * (len - 8))) - 1ull;
  }
  *k0 = a0;
  *k1 = a1;
}
__device__ __forceinline__ int match_key(const KeySlot (&t)[32], uint32_t mul, int len, uint64_t k0, uint64_t k1) {
  const KeySlot& sl = t[key_slot(k0, (uint32_t)len, mul)];
  return (sl.len == len && sl.k0 == k0 && sl.k1 == k1) ? sl.id : -1;
}

// the 8 bytes at p (any alignment) from the aligned words that hold them; `limit`: the aligned end of the document's
// last 32-byte word (bytes behind it read as zero)
__device__ __forceinline__ uint64_t load8(const uint8_t* p, const uint8_t* limit) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const unsigned long long* w = reinterpret_cast<const unsigned long long*>(a & ~(uintptr_t)7);
  const int sh = (int)(a & 7) * 8;
  const uint64_t w0 = __ldg(w);
  if (!sh) return w0;
  const uint64_t w1 = reinterpret_cast<const uint8_t*>(w + 1) < limit ? __ldg(w + 1) : 0ull;
  return (w0 >> sh) | (w1 << (64 - sh));
}
// n (<= 8) bytes of x to dst
__device__ __forceinline__ void store_upto8(uint8_t* dst, uint64_t x, int n) {
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (k < n) dst[k] = (uint8_t)(x >> (8 * k));
}
// len plain bytes from src to dst, 8 at a time
__device__ __forceinline__ void copy_plain(const uint8_t* src, int len, uint8_t* dst, const uint8_t* limit) {
  for (int i = 0; i < len; i += 8) store_upto8(dst + i, load8(src + i, limit), len - i);
}

// ---- escapes ------------------------------------------------------------------------------------------------------
// what the escapes of the raw string ab[src .. src + raw) save: raw bytes minus unescaped bytes; *bad when an escape
// is not one of ECMA-404's, or names a surrogate (the walk decides those).  Only called for strings stage 1 flagged.
// Backslashes are found 8 bytes at a time.
PIE_JF_ESC_FN int escape_savings(const uint8_t* ab, int src, int raw, const uint8_t* limit, bool* bad) {
  int saved = 0, i = 0;
  while (i < raw) {
    {
      const uint64_t x = load8(ab + src + i, limit) ^ 0x5c5c5c5c5c5c5c5cull;
      const uint64_t hit = (x - 0x0101010101010101ull) & ~x & 0x8080808080808080ull;  // exact up to the first hit
      const int j = hit ? (__ffsll((long long)hit) - 1) >> 3 : 8;
      if (i + j >= raw) break;
      i += j;
      if (j == 8) continue;
    }
    const uint8_t e = i + 1 < raw ? ab[src + i + 1] : 0;
    if (e == 'u') {
      if (i + 5 >= raw) { *bad = true; break; }
      uint32_t cp = 0;
      for (int k = 2; k < 6; ++k) {
        const int h = hex_value(ab[src + i + k]);
        if (h < 0) *bad = true;
        cp = cp * 16 + (uint32_t)(h & 15);
      }
      if (cp >= 0xD800 && cp <= 0xDFFF) *bad = true;
      saved += 6 - (cp < 0x80 ? 1 : cp < 0x800 ? 2 : 3);
      i += 6;
    } else if (e == '"' || e == '\\' || e == '/' || e == 'b' || e == 'f' || e == 'n' || e == 'r' || e == 't') {
      saved += 1;
      i += 2;
    } else {
      *bad = true;
      break;
    }
    if (*bad) break;
  }
  return saved;
}
// the unescaped bytes of ab[src .. src + raw) to dst (escapes already validated); returns how many.  Plain runs go 8
// bytes at a time.
PIE_JF_ESC_FN int unescape_copy(const uint8_t* ab, int src, int raw, uint8_t* dst, const uint8_t* limit) {
  int i = 0, o = 0;
  while (i < raw) {
    const uint64_t w = load8(ab + src + i, limit);
    const uint64_t x = w ^ 0x5c5c5c5c5c5c5c5cull;
    const uint64_t hit = (x - 0x0101010101010101ull) & ~x & 0x8080808080808080ull;
    int run = hit ? (__ffsll((long long)hit) - 1) >> 3 : 8;
    if (run > raw - i) run = raw - i;
    store_upto8(dst + o, w, run);
    o += run;
    i += run;
    if (run == 8 || i >= raw) continue;
    const uint8_t e = ab[src + i + 1];  // the byte at i is a backslash
    if (e == 'u') {
      uint32_t cp = 0;
      for (int k = 2; k < 6; ++k) cp = cp * 16 + (uint32_t)(hex_value(ab[src + i + k]) & 15);
      if (cp < 0x80) {
        dst[o++] = (uint8_t)cp;
      } else if (cp < 0x800) {
        dst[o++] = (uint8_t)(0xC0 | (cp >> 6));
        dst[o++] = (uint8_t)(0x80 | (cp & 0x3F));
      } else {
        dst[o++] = (uint8_t)(0xE0 | (cp >> 12));
        dst[o++] = (uint8_t)(0x80 | ((cp >> 6) & 0x3F));
        dst[o++] = (uint8_t)(0x80 | (cp & 0x3F));
      }
      i += 6;
    } else {
      const uint8_t m = e == 'b' ? 8 : e == 'f' ? 12 : e == 'n' ? 10 : e == 'r' ? 13 : e == 't' ? 9 : e;  // " \ / themselves
      dst[o++] = m;
      i += 2;
    }
  }
  return o;
}

// Well-formed UTF-8 (Unicode table 3-7) among the bytes >= 0x80 of a lane's word (`hib`: their positions, pos0: the
// word's first position): sequences that start here are followed to their end (into the next word if need be);
// continuation bytes owed to a sequence that starts in an earlier word belong to the lane that owns its lead.
__device__ __noinline__ bool utf8_ok(const uint8_t* ab, uint32_t hib, int pos0, int skip, int span) {
  int owed = 0;  // continuation bytes at the start of the word that a lead before it answers for
  for (int k = 1; k <= 3 && pos0 - k >= skip; ++k) {
    const uint8_t c = ab[pos0 - k];
    if ((c & 0xC0) == 0x80) continue;
    if (c >= 0xC2) {
      const int need = c >= 0xF0 ? 3 : c >= 0xE0 ? 2 : 1;
      owed = need - k + 1;
      if (owed < 0) owed = 0;
    }
    break;
  }
  int done = owed;  // bytes of the word below this index are settled
  while (hib) {
    const int bit = __ffs(hib) - 1;
    hib &= hib - 1;
    if (bit < done) continue;
    const int p = pos0 + bit;
    const uint8_t c = ab[p];
    if (c < 0xC2 || c > 0xF4) return false;  // a continuation byte nobody owns, an overlong lead, beyond U+10FFFF
    int need = 1, lo = 0x80, hi = 0xBF;
    if (c >= 0xF0) { need = 3; if (c == 0xF0) lo = 0x90; if (c == 0xF4) hi = 0x8F; }
    else if (c >= 0xE0) { need = 2; if (c == 0xE0) lo = 0xA0; if (c == 0xED) hi = 0x9F; }
    for (int k = 1; k <= need; ++k) {
      if (p + k >= span) return false;
      const uint8_t cc = ab[p + k];
      if (cc < lo || cc > hi) return false;
      lo = 0x80;
      hi = 0xBF;
    }
    done = bit + need + 1;
  }
  return true;
}

// ---- numbers ------------------------------------------------------------------------------------------------------
// eight ASCII digits, the first in the low byte, as a number
__device__ __forceinline__ uint32_t digits8(uint64_t x) {
  x = (x & 0x0f0f0f0f0f0f0f0full) * 2561 >> 8;
  x = (x & 0x00ff00ff00ff00ffull) * 6553601 >> 16;
  return (uint32_t)((x & 0x0000ffff0000ffffull) * 42949672960001ull >> 32);
}
// bit 7 of every byte of x that is an ASCII digit
__device__ __forceinline__ uint64_t digit_flags(uint64_t x) {
  const uint64_t lo7 = x & 0x7f7f7f7f7f7f7f7full;
  const uint64_t ge0 = lo7 + 0x5050505050505050ull;  // bit 7: >= '0'
  const uint64_t gt9 = lo7 + 0x4646464646464646ull;  // bit 7: > '9'
  return ge0 & ~gt9 & ~x & 0x8080808080808080ull;
}
// The JSON number at ab[pos ..) (the document ends at span): its value, and in *term the byte behind it (-1: none).
// Numbers of the everyday form -?digits(.digits)? with at most 15 digits that end within 16 bytes are read 8 bytes at
// a time and converted by ONE exact IEEE operation, as parse_json_number_from does for them (Clinger's case: the
// digits as an integer below 2^53, over an exact power of ten); everything else goes through that parser.
// kGeneric = false (pass 1, PIE_JF_NUMBERS_EVERYDAY_ONLY): a number that is not of the everyday form is kNumUndecided —
// the document is declined and the walk's parser decides it; pass 1 then carries no general number parser at all.
template <bool kGeneric, bool kFastFirst = false>
PIE_JF_NUM_FN int parse_number_at(const uint8_t* ab, int pos, int span, const uint8_t* limit, const Pow5Table& pow5,
                                            double* value, int* term) {
  // kFastFirst: pass 2 with records, which has the room for it (pass 1 does not: profiles/ncu_r02_ingest_summary.md)
  if (PIE_JF_FASTNUM || !kGeneric || kFastFirst) {
    uint64_t lo = load8(ab + pos, limit), hi = load8(ab + pos + 8, limit);
    const int avail = span - pos;  // >= 1
    const bool neg = (lo & 0xff) == '-';
    int at = 0;
    if (neg) {  // drop the sign
      lo = (lo >> 8) | (hi << 56);
      hi >>= 8;
      at = 1;
    }
    const uint64_t d_lo = digit_flags(lo), d_hi = digit_flags(hi);
    // length of the run of digits at the start
    const uint64_t nd_lo = ~d_lo & 0x8080808080808080ull;
    int n1 = nd_lo ? (__ffsll((long long)nd_lo) - 1) >> 3 : 8;
    if (n1 == 8) {
      const uint64_t nd_hi = ~d_hi & 0x8080808080808080ull;
      n1 += nd_hi ? (__ffsll((long long)nd_hi) - 1) >> 3 : 8;
    }
    // the bytes from the first non-digit on, as a 128-bit shift
    if (n1 >= 1 && n1 <= 15 && !((lo & 0xff) == '0' && n1 > 1)) {
      // integer part: n1 digits (no leading zero unless it is the only digit)
      uint64_t a = lo, b = hi;
      uint64_t ip;
      if (n1 <= 8) {
        ip = digits8((a & (n1 == 8 ? ~0ull : ((1ull << (8 * n1)) - 1ull))) << (8 * (8 - n1)));
      } else {
        const int r = n1 - 8;
        ip = (uint64_t)digits8(a) * (r == 1 ? 10ull : r == 2 ? 100ull : r == 3 ? 1000ull : r == 4 ? 10000ull : r == 5 ? 100000ull
                                                         : r == 6 ? 1000000ull : 10000000ull) +
             digits8((b & ((1ull << (8 * r)) - 1ull)) << (8 * (8 - r)));
      }
      // what follows the integer part
      uint64_t rest_lo, rest_hi;
      if (n1 < 8) { rest_lo = (a >> (8 * n1)) | (b << (64 - 8 * n1)); rest_hi = b >> (8 * n1); }
      else if (n1 == 8) { rest_lo = b; rest_hi = 0; }
      else { rest_lo = b >> (8 * (n1 - 8)); rest_hi = 0; }
      int used = at + n1;
      int nf = 0;
      uint64_t fp = 0;
      bool ok = true;
      if ((rest_lo & 0xff) == '.') {
        const uint64_t f = (rest_lo >> 8) | (rest_hi << 56);
        const uint64_t nd = ~digit_flags(f) & 0x8080808080808080ull;
        nf = nd ? (__ffsll((long long)nd) - 1) >> 3 : 8;
        ok = nf >= 1 && nf <= 7 && n1 + nf <= 15 && used + 1 + nf < 16;  // the digit run must end inside what was loaded
        if (ok) fp = digits8((f & ((1ull << (8 * nf)) - 1ull)) << (8 * (8 - nf)));
        used += 1 + nf;
        rest_lo = f >> (8 * nf);
      } else {
        ok = used < 16;
      }
      const uint32_t t = (uint32_t)(rest_lo & 0xff);
      if (ok && used <= avail && t != 'e' && t != 'E' && t != '.' && !(t >= '0' && t <= '9')) {
        static const double kPow10[8] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7};
        const uint64_t scale = nf == 0 ? 1ull : nf == 1 ? 10ull : nf == 2 ? 100ull : nf == 3 ? 1000ull : nf == 4 ? 10000ull
                               : nf == 5 ? 100000ull : nf == 6 ? 1000000ull : 10000000ull;
        const uint64_t w = ip * scale + fp;
        double d = (double)w;
        if (nf) d = d / kPow10[nf];
        *value = neg ? -d : d;
        *term = used < avail ? (int)ab[pos + used] : -1;
        return kNumOk;
      }
    }
  }
  if (!kGeneric) {
    *term = -1;
    return kNumUndecided;
  }
  DocCursor src;
  src.open(ab, pos, span);
  const int rc = parse_json_number_from<true>(src, pow5, value);
  *term = src.peek();
  return rc;
}

// Pass 1 does not convert numbers, it only makes sure that pass 2 can: ECMA-404's grammar without an exponent part and
// at most 19 digits in all.  Such a number is always decided by parse_json_number_from (no digit is dropped, and the
// Eisel-Lemire product can only be undecided for powers of ten outside 10^-27 .. 10^55): pass 2 converts it from the
// text, and pass 1 carries no number parser.  Anything else (1e21, 25 digits) declines the document to the walk.
__device__ __forceinline__ bool number_shape(const uint8_t* ab, int pos, int span, int* term) {
  int i = pos, digits = 0;
  int c = i < span ? (int)ab[i] : -1;
  if (c == '-') c = ++i < span ? (int)ab[i] : -1;
  if (c < '0' || c > '9') return false;
  if (c == '0') {
    c = ++i < span ? (int)ab[i] : -1;
    if (c >= '0' && c <= '9') return false;  // no leading zeros
    digits = 1;
  } else {
    while (c >= '0' && c <= '9') {
      ++digits;
      c = ++i < span ? (int)ab[i] : -1;
    }
  }
  if (c == '.') {
    c = ++i < span ? (int)ab[i] : -1;
    if (c < '0' || c > '9') return false;
    while (c >= '0' && c <= '9') {
      ++digits;
      c = ++i < span ? (int)ab[i] : -1;
    }
  }
  *term = c;
  return digits <= 19 && c != 'e' && c != 'E';
}

// ---- stage 1 ------------------------------------------------------------------------------------------------------
// bit 7 of every byte of y (bytes < 0x80, `hi` = the bytes that are not) that equals the byte replicated in c4
__device__ __forceinline__ uint32_t eq_flags(uint32_t y, uint32_t hi, uint32_t c4) {
  return ~(((y ^ c4) + 0x7f7f7f7fu) | hi) & 0x80808080u;
}

// escaped characters of a word of backslash positions (simdjson's find_escaped on 32 bits); cin: the word's first
// byte is escaped from the word before; *cout: the next word's first byte is
__device__ __forceinline__ uint32_t find_escaped(uint32_t b, uint32_t cin, uint32_t* cout) {
  b &= ~cin;
  const uint32_t follows = (b << 1) | cin;
  const uint32_t odd_starts = b & ~0x55555555u & ~follows;
  const uint32_t sum = odd_starts + b;
  *cout = sum < odd_starts ? 1u : 0u;
  const uint32_t invert = sum << 1;
  return (0x55555555u ^ invert) & follows;
}

struct IndexTotals {
  int members, entries, quotes;
};

// Builds the two lists of the document in ws; false = declined.  kCheck (the measuring pass): every rule that makes
// "accepted" mean "a document of the shape in the header" that does not need a key or a value.
template <bool kCheck, class Caps>
__device__ __forceinline__ bool build_index(WarpShared<Caps>& ws, const Doc& dc, IndexTotals* tot) {
  const int lane = threadIdx.x & 31;
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t c_esc = 0, c_instr = 0, c_tops = 0, c_bsq = 0;
  int c_depth = 0, c_lb = 0, c_co = 0, c_q = 0;
  bool bad = false;
  const int nblk = (dc.nwords + 31) >> 5;
  for (int blk = 0; blk < nblk; ++blk) {
    const int wi = blk * 32 + lane;
    const int pos0 = wi * 32;
    uint32_t L[8];
    if (wi < dc.nwords) {
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(dc.ab + pos0));
      const uint4 b = __ldg(reinterpret_cast<const uint4*>(dc.ab + pos0 + 16));
      L[0] = a.x; L[1] = a.y; L[2] = a.z; L[3] = a.w;
      L[4] = b.x; L[5] = b.y; L[6] = b.z; L[7] = b.w;
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) L[k] = 0;
    }
    uint32_t q = 0, bsl = 0, op = 0, cl = 0, brm = 0, co = 0, cm = 0, hib = 0, ctl_any = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      // word w of the transposed copy holds the bytes at positions w, 8 + w, 16 + w, 24 + w of the lane's 32: a flag
      // in bit 7 of its byte b belongs at bit 8 b + w of the position mask, i.e. one shift by 7 - w merges it
      const int h = w >> 2, k = w & 3;
      const uint32_t sel = (uint32_t)k | ((uint32_t)(4 + k) << 4);
      const uint32_t x = __byte_perm(__byte_perm(L[h], L[2 + h], sel), __byte_perm(L[4 + h], L[6 + h], sel), 0x5410);
      const uint32_t hi = x & 0x80808080u, y = x & 0x7f7f7f7fu;
      const uint32_t z = (y | 0x20202020u) ^ 0x7b7b7b7bu;  // 0 for [ {, 6 for ] }
      q |= eq_flags(y, hi, 0x22222222u) >> (7 - w);
      bsl |= eq_flags(y, hi, 0x5c5c5c5cu) >> (7 - w);
      op |= (~((z + 0x7f7f7f7fu) | hi) & 0x80808080u) >> (7 - w);
      cl |= (~(((z ^ 0x06060606u) + 0x7f7f7f7fu) | hi) & 0x80808080u) >> (7 - w);
      brm |= ((y << 2) & 0x80808080u) >> (7 - w);  // bit 5 tells { } from [ ]
      co |= eq_flags(y, hi, 0x3a3a3a3au) >> (7 - w);
      if (kCheck) {
        cm |= eq_flags(y, hi, 0x2c2c2c2cu) >> (7 - w);
        hib |= hi >> (7 - w);
        ctl_any |= ~((y + 0x60606060u) | hi);  // bit 7: a byte below 0x20
      }
    }
    // bytes of the word that are the document's
    uint32_t vm = 0;
    {
      int lo = dc.skip - pos0, he = dc.span - pos0;
      if (lo < 0) lo = 0;
      if (he > 32) he = 32;
      if (lo < 32 && he > lo) vm = (he >= 32 ? kFull : ((1u << he) - 1u)) & ~((1u << lo) - 1u);
    }
    q &= vm; bsl &= vm; op &= vm; cl &= vm; co &= vm;
    // escaped characters: the carry runs from lane to lane; a lane's carry out depends on its carry in only when all
    // its 32 bytes are backslashes, so this settles in two or three turns
    uint32_t esc = 0;
    if (__any_sync(kFull, bsl != 0) || c_esc) {
      uint32_t cin = lane == 0 ? c_esc : 0u, cout = 0;
      for (;;) {
        esc = find_escaped(bsl, cin, &cout);
        uint32_t nc = __shfl_up_sync(kFull, cout, 1);
        if (lane == 0) nc = c_esc;
        const bool changed = nc != cin;
        cin = nc;
        if (!__any_sync(kFull, changed)) break;
      }
      c_esc = __shfl_sync(kFull, cout, 31);
    }
    const uint32_t qr = q & ~esc;
    const uint32_t bstart = bsl & ~esc;
    // inside a string: from the opening quote up to (not including) the closing one
    uint32_t is = qr;
    is ^= is << 1; is ^= is << 2; is ^= is << 4; is ^= is << 8; is ^= is << 16;
    const uint32_t pm = __ballot_sync(kFull, __popc(qr) & 1);
    if ((__popc(pm & lt) & 1) ^ c_instr) is = ~is;
    c_instr ^= __popc(pm) & 1u;
    is &= vm;
    op &= ~is; cl &= ~is; co &= ~is;
    const uint32_t lb = op & brm;
    // depth, entries, colons and quotes before the word
    uint32_t sa = (uint32_t)__popc(op) | ((uint32_t)__popc(cl) << 16);
    uint32_t sb = (uint32_t)__popc(lb) | ((uint32_t)__popc(co) << 16);
    uint32_t sq = (uint32_t)__popc(qr);
    const uint32_t own_a = sa, own_b = sb, own_q = sq;
#pragma unroll
    for (int dlt = 1; dlt < 32; dlt <<= 1) {
      const uint32_t ta = __shfl_up_sync(kFull, sa, dlt), tb = __shfl_up_sync(kFull, sb, dlt), tq = __shfl_up_sync(kFull, sq, dlt);
      if (lane >= dlt) { sa += ta; sb += tb; sq += tq; }
    }
    const uint32_t ex_a = sa - own_a, ex_b = sb - own_b;
    const int d0 = c_depth + (int)(ex_a & 0xffff) - (int)(ex_a >> 16);
    const int lb0 = c_lb + (int)(ex_b & 0xffff);
    const int co0 = c_co + (int)(ex_b >> 16);
    const int q0 = c_q + (int)(sq - own_q);
    const uint32_t tot_a = __shfl_sync(kFull, sa, 31), tot_b = __shfl_sync(kFull, sb, 31);
    c_depth += (int)(tot_a & 0xffff) - (int)(tot_a >> 16);
    c_lb += (int)(tot_b & 0xffff);
    c_co += (int)(tot_b >> 16);
    c_q += (int)__shfl_sync(kFull, sq, 31);
    // the quotes, each with "an escape starts between the quote before and this one" (asked of closing quotes: the
    // string has escapes).  What lies before a lane's first quote comes from the lanes below, back to the last quote.
    {
      const uint32_t G = __ballot_sync(kFull, qr != 0);
      const uint32_t after_last = qr ? (bstart >> (31 - __clz(qr))) >> 1 : bstart;  // escape starts behind the lane's last quote
      const uint32_t F = __ballot_sync(kFull, after_last != 0);
      uint32_t before;  // an escape start between the last quote of the lanes below and this lane
      if (G & lt) before = ((F & lt) >> (31 - __clz(G & lt))) != 0;
      else before = ((F & lt) != 0) | c_bsq;
      c_bsq = G ? ((F >> (31 - __clz(G))) != 0) : (c_bsq | (F != 0));
      // the flagged quotes of the lane: the first quote behind each escape start, and the lane's first quote when an
      // escape start waits below
      uint32_t flagged = before ? (qr & (0u - qr)) : 0u;
      for (uint32_t m = bstart; m; m &= m - 1) {
        const uint32_t above = qr & ~((2u << (__ffs(m) - 1)) - 1u);
        flagged |= above & (0u - above);
      }
      int idx = q0;
      for (uint32_t m = qr; m; m &= m - 1, ++idx) {
        const int bit = __ffs(m) - 1;
        if (idx < Caps::kQuotes) ws.qpos[idx] = (uint16_t)((uint32_t)(pos0 + bit) | (((flagged >> bit) & 1u) << 15));
      }
    }
    // the members: a colon, the index of the quote before it (the key's closing quote), the entry it is in
    {
      int idx = co0;
      for (uint32_t m = co; m; m &= m - 1, ++idx) {
        const int bit = __ffs(m) - 1;
        const uint32_t below = (1u << bit) - 1u;
        const int d = d0 + __popc(op & below) - __popc(cl & below);
        const int e1 = d == 3 ? lb0 + __popc(lb & below) - 1 : 0;  // entry + 1 (the show's own brace is the first)
        const int kq = q0 + __popc(qr & below) - 1;
        if (!(d == 1 || d == 3) || kq < 1 || e1 > kFastMaxEntries) bad = true;  // a colon in an array, or no key before it
        if (idx < Caps::kMembers)
          ws.mem[idx] = (uint32_t)(pos0 + bit) | ((uint32_t)(kq & 0xfff) << 14) | ((uint32_t)(e1 & 0x3f) << 26);
      }
    }
    if (kCheck) {
      cm &= vm & ~is; hib &= vm;
      const uint32_t rb = cl & brm;
      const uint32_t lk = op & ~lb, rk = cl & ~rb;
      const uint32_t cq = qr & ~is;
      const uint32_t sc = vm & ~is & ~qr & ~(op | cl | co | cm);  // bytes of scalars
      // the class of the byte before each position: the masks moved up by one, the top bits handed to the next lane
      const uint32_t tops = (cq >> 31) | ((lb >> 31) << 1) | ((lk >> 31) << 2) | ((rb >> 31) << 3) | ((rk >> 31) << 4) |
                            ((co >> 31) << 5) | ((cm >> 31) << 6) | ((sc >> 31) << 7);
      uint32_t pt = __shfl_up_sync(kFull, tops, 1);
      if (lane == 0) pt = c_tops;
      c_tops = __shfl_sync(kFull, tops, 31);
      const uint32_t p_cq = (cq << 1) | (pt & 1), p_lb = (lb << 1) | ((pt >> 1) & 1), p_lk = (lk << 1) | ((pt >> 2) & 1),
                     p_rb = (rb << 1) | ((pt >> 3) & 1), p_rk = (rk << 1) | ((pt >> 4) & 1), p_co = (co << 1) | ((pt >> 5) & 1),
                     p_cm = (cm << 1) | ((pt >> 6) & 1), p_sc = (sc << 1) | ((pt >> 7) & 1);
      const uint32_t first = wi == 0 ? (1u << dc.skip) : 0u;
      uint32_t wrong = co & ~p_cq;                       // a colon follows a string
      wrong |= (qr & is) & ~(p_lb | p_cm | p_co | p_lk); // a string follows { , : [
      wrong |= lb & ~(p_lk | p_cm | first);              // { opens the document or an element
      wrong |= lk & ~p_co;                               // [ is a member's value
      wrong |= rb & ~(p_lb | p_cq | p_sc | p_rk);        // } follows { or a value
      wrong |= rk & ~(p_lk | p_cq | p_rb);               // ] follows [ or an element
      wrong |= cm & ~(p_cq | p_sc | p_rk | p_rb);        // , follows a value
      wrong |= (sc & ~p_sc) & ~p_co;                     // a scalar is a member's value
      if (wrong) bad = true;
      // control characters are nowhere in such a document.  The flags of the word's 32 bytes include its neighbours'
      // bytes at the document's ends: look byte by byte there (a lane or two per document)
      if (ctl_any & 0x80808080u) {
        if (vm == kFull) bad = true;
        else
          for (uint32_t m = vm; m; m &= m - 1)
            if (dc.ab[pos0 + __ffs(m) - 1] < 0x20) bad = true;
      }
      // brackets: the kind follows from the depth ({ at 0 and 2, [ at 1 and 3), so matching pairs need no stack
      if (op | cl) {
        int d = d0;
        for (uint32_t m = op | cl; m; m &= m - 1) {
          const int bit = __ffs(m) - 1, pos = pos0 + bit;
          const bool open = (op >> bit) & 1u, brc = (brm >> bit) & 1u;
          bool ok;
          if (open) ok = brc ? ((d == 0 && pos == dc.skip) || d == 2) : (d == 1 || d == 3);
          else ok = brc ? (d == 3 || (d == 1 && pos == dc.span - 1)) : (d == 2 || d == 4);
          if (!ok) bad = true;
          d += open ? 1 : -1;
        }
      }
      // bytes >= 0x80 outside strings are bytes of scalars, which stage 2 reads one by one; inside strings: UTF-8
      if (hib && !utf8_ok(dc.ab, hib, pos0, dc.skip, dc.span)) bad = true;
    }
  }
  tot->members = c_co;
  tot->entries = c_lb - 1;
  tot->quotes = c_q;
  if (c_co > Caps::kMembers || c_lb - 1 > kFastMaxEntries || c_q > Caps::kQuotes) bad = true;
  if (kCheck && (c_depth != 0 || c_instr != 0 || c_lb < 1)) bad = true;
  return !__any_sync(kFull, bad);
}

// exclusive sum of `val` over the lower lanes that hold the same key (a key nobody shares: 64 + lane)
__device__ __forceinline__ uint32_t group_prefix(int key, uint32_t val, uint32_t lt, bool* leader) {
  const uint32_t peers = __match_any_sync(kFull, key);
  uint32_t pending = peers & lt, sum = 0;
  while (__any_sync(kFull, pending != 0)) {
    const int j = pending ? __ffs(pending) - 1 : 0;
    const uint32_t v = __shfl_sync(kFull, val, j);
    if (pending) {
      sum += v;
      pending &= pending - 1;
    }
  }
  *leader = (peers >> (threadIdx.x & 31)) == 1u;  // the highest lane of the group
  return sum;
}

// The elements of a crew / actions array that opens at `at` ('['): strings only; qi = the index in qpos of the first
// quote behind the bracket.  kMode 0: counts them and their unescaped bytes; 1: writes offsets and bytes from
// (item0, dst0) on; 2: writes their records (heap; item0 / dst0 relative to the document).  false = not the shape.
template <int kMode, class Caps>
PIE_JF_ITEMS_FN bool walk_items(const WarpShared<Caps>& ws, const Doc& dc, int at, int qi, int n_quotes, uint32_t* n_items,
                                           uint32_t* n_bytes, int32_t* off, uint8_t* data, uint32_t item0, uint32_t dst0,
                                           unsigned long long* recs = nullptr, uint32_t heap = 0) {
  uint32_t N = 0, B = 0;
  int pos = at + 1;
  const uint8_t* limit = dc.ab + dc.nwords * 32;
  if (pos < dc.span && dc.ab[pos] != ']') {
    for (;;) {
      if (qi + 1 >= n_quotes) return false;
      const uint32_t qo = ws.qpos[qi], qc = ws.qpos[qi + 1];
      if ((int)(qo & kPosMask) != pos) return false;  // the element is not a string
      const int cq = (int)(qc & kPosMask);
      if (cq + 1 >= dc.span) return false;
      const int raw = cq - pos - 1;
      int len = raw;
      const bool esc = (qc >> 15) != 0;
      if (kMode == 1) {
        off[item0 + N] = (int32_t)(dst0 + B);
        if (esc) len = unescape_copy(dc.ab, pos + 1, raw, data + dst0 + B, limit);
        else copy_plain(dc.ab + pos + 1, raw, data + dst0 + B, limit);
      } else if (esc) {
        bool bad = false;
        len = raw - escape_savings(dc.ab, pos + 1, raw, limit, &bad);
        if (bad) return false;
      }
      if (kMode == 2) {
        const uint32_t lo = (uint32_t)(pos + 1) | ((uint32_t)raw << 14) | ((uint32_t)esc << 28);
        const uint32_t hi = heap | ((dst0 + B) << 5) | ((item0 + N) << 19);
        recs[N] = (unsigned long long)lo | ((unsigned long long)hi << 32);
      }
      B += (uint32_t)len;
      ++N;
      const uint8_t nc = dc.ab[cq + 1];
      if (nc == ']') break;
      if (nc != ',') return false;
      pos = cq + 2;
      qi += 2;
    }
  }
  *n_items = N;
  *n_bytes = B;
  return true;
}

// ---- one document -------------------------------------------------------------------------------------------------
// kFill = false: validates, counts (planes_row receives the document's 26 counts) and, when the pool has room, writes
// the document's records; returns kRouteSlow (declined), kRouteFast (accepted; pass 2 parses it again) or kRouteRecords.
// kFill = true: planes_row holds where the document's part of every plane starts; writes its part of the table.
template <bool kFill, class Caps>
__device__ __forceinline__ int fast_doc(WarpShared<Caps>& ws, const TablePointers& tp, const uint8_t* __restrict__ text, int64_t from,
                                        int64_t to, int64_t s, int64_t n_docs, uint32_t* __restrict__ planes_row,
                                        const IngestOut& out, const Pow5Table& pow5, const RecCtx& rc) {
  const int lane = threadIdx.x & 31;
  const uint32_t lt = (1u << lane) - 1u;
  const int64_t n = to - from;
  if (n < 2) return kRouteSlow;
  Doc dc;
  {
    const uintptr_t a0 = reinterpret_cast<uintptr_t>(text + from);
    dc.skip = (int)(a0 & 31);
    dc.ab = reinterpret_cast<const uint8_t*>(a0 - dc.skip);
  }
  if (dc.skip + n > kFastMaxBytes) return kRouteSlow;
  dc.span = dc.skip + (int)n;
  dc.nwords = (dc.span + 31) >> 5;
  const uint8_t* ab = dc.ab;

  __syncwarp();
  if (lane == 0) { ws.n_num = 0; ws.n_esc = 0; ws.seen_show = 0; ws.n_arr = 0; ws.n_items = 0; }
  if (lane < kPlanes) ws.cnt[lane] = kFill ? planes_row[lane] : 0u;
  IndexTotals tot;
  if (!build_index<!kFill, Caps>(ws, dc, &tot)) return kRouteSlow;
  if (!kFill)
    for (int e = lane; e < tot.entries; e += 32) ws.seen[e] = 0;
  // records: room for a record per member now, for the numbers and the elements when they are counted.  (The last
  // document is parsed again by pass 2: it writes the terminal offsets from the counters it ends on.)
  bool rec_on = false;
  unsigned long long members_at = 0;
  if (!kFill && rc.pool && s != n_docs - 1) {
    if (lane == 0) members_at = atomicAdd(rc.cursor, (unsigned long long)tot.members);
    members_at = __shfl_sync(kFull, members_at, 0);
    rec_on = members_at + (unsigned long long)tot.members <= rc.capacity;
  }
  unsigned long long* const slots = rc.pool + members_at;
  __syncwarp();
  const uint32_t row0 = kFill ? ws.cnt[kPlaneEntries] : 0u;
  if (kFill) {
    // the show's row: every text field starts where the previous show's ended (an absent key is '')
    if (lane < 7) tp.off[lane][s] = (int32_t)ws.cnt[lane];
    else if (lane == 7) out.entry_offsets[s] = (int32_t)row0;
    else if (lane == 8) out.crew_list[s] = (int32_t)ws.cnt[kPlaneCrewItems];
    else if (lane < 9 + PIE_TF_COUNT) { if (out.time_val[lane - 9]) out.time_val[lane - 9][s] = jw_nan(); }
    else if (lane == 9 + PIE_TF_COUNT) { if (out.time_kind) *reinterpret_cast<uint32_t*>(out.time_kind + s * PIE_TF_COUNT) = 0u; }
    __syncwarp();
  }

  bool bad = false;
  uint32_t accounted = 0;  // strings the members answer for (keys, string values, elements)
  const int steps = (tot.members + 31) >> 5;
  for (int t = 0; t < steps; ++t) {
    const int k = t * 32 + lane;
    const bool act = k < tot.members;
    // ---- the member: every lane runs the same straight code; a lane without a member runs it on the last member and
    // leaves no trace
    const uint32_t rec = ws.mem[act ? k : tot.members - 1];
    const int c = (int)(rec & kPosMask), kq = (int)((rec >> 14) & 0xfff), e1 = (int)(rec >> 26);
    const bool in_entry = e1 != 0;
    const int e = in_entry ? e1 - 1 : 0;
    const uint32_t row = row0 + (uint32_t)e;
    const uint32_t qk = ws.qpos[kq];
    const int ko = (int)(ws.qpos[kq - 1] & kPosMask), kc = (int)(qk & kPosMask);
    const int klen = kc - ko - 1;
    bool mbad = false;
    if (!kFill) {
      const uint8_t pk = ab[ko - 1];
      mbad = kc != c - 1 || !(pk == '{' || pk == ',') || (qk >> 15) != 0;  // a key right before the colon, without escapes
    }
    int key = -1;
    if (klen >= 1 && klen <= 16) {
      uint64_t k0, k1;
      load_key(ab + ko + 1, klen, ab + dc.nwords * 32, &k0, &k1);
      key = in_entry ? match_key(tp.keys.entry, kEntryKeyMul, klen, k0, k1) : match_key(tp.keys.show, kShowKeyMul, klen, k0, k1);
    }
    if (!kFill && act && key >= 0) {
      const uint32_t bit = 1u << key;
      const uint32_t old = in_entry ? atomicOr(&ws.seen[e], bit) : atomicOr(&ws.seen_show, bit);
      if (old & bit) mbad = true;  // a known key twice
    }
    // what the value is for
    int heap = -1, tf = -1, numrole = 0, arr_heap = -1;
    bool is_entries = false;
    if (in_entry) {
      if (key >= 0 && key < 14) heap = kHeapEntry0 + key;
      numrole = key == kEkTs ? 3 : key == kEkDelaySec ? 4 : 0;
      if (key == kEkActions) arr_heap = kHeapActions;
    } else {
      if (key >= 0 && key < 7) heap = key;
      if (key == kSkCreatedAt) { tf = PIE_TF_CREATED; numrole = 1; }
      else if (key == kSkArchivedAt) { tf = PIE_TF_ARCHIVED; numrole = 2; }
      else if (key == kSkUpdatedAt) { tf = PIE_TF_UPDATED; numrole = 5; }
      else if (key == kSkDeletedAt) { tf = PIE_TF_DELETED; numrole = 6; }
      if (key == kSkCrew) arr_heap = kHeapCrew;
      is_entries = key == kSkEntries;
    }
    const int v = c + 1;
    const uint8_t ch = v < dc.span ? ab[v] : 0;
    uint32_t L = 0, N = 0;
    int vkind = 0;  // 1 plain string (or null) for a heap, 2 string with escapes for a heap, 3 elements of an array
    int vsrc = 0, vraw = 0;
    uint32_t strings = 1;
    uint32_t rkind = kRecNone, rlo = 0;  // pass 1: the member's record
    if (ch == '"') {
      // the string behind the colon is the next pair of quotes
      const bool have = kq + 2 < tot.quotes;
      const uint32_t qc = ws.qpos[have ? kq + 2 : kq];
      const int cq = (int)(qc & kPosMask);
      const bool esc = (qc >> 15) != 0;
      if (!have || cq + 1 >= dc.span) mbad = true;
      strings = 2;
      if (!kFill && !mbad) {
        const uint8_t nc = ab[cq + 1];
        if (nc != ',' && nc != '}') mbad = true;
      }
      vsrc = v + 1;
      vraw = cq - v - 1;
      if (heap >= 0) {
        L = (uint32_t)vraw;
        vkind = 1;
        if (esc && !mbad) {
          bool ebad = false;
          L = (uint32_t)(vraw - escape_savings(ab, vsrc, vraw, ab + dc.nwords * 32, &ebad));
          vkind = 2;
          if (ebad) mbad = true;
        }
      } else {
        if (esc) mbad = true;  // nobody unescapes it here, so nobody validates it: the walk does
        if (numrole == 4 || arr_heap == kHeapActions) mbad = true;  // delaySec / actions that are text
        if (tf >= 0) { rkind = kRecTimeString; rlo = (uint32_t)(v + 1); }
        else if (numrole == 3) rkind = kRecNanTs;
        if (kFill && act) {
          if (tf >= 0) {
            if (out.time_kind) out.time_kind[s * PIE_TF_COUNT + tf] = (uint8_t)PIE_TK_STRING;
            if (out.time_val[tf])
              out.time_val[tf][s] = np_bits_to_double(0x7ff8000000000000ull | ((uint64_t)((ab + v + 1) - out.text) & 0x7ffffffffffffull));
          } else if (numrole == 3) {
            out.entry_ts[row] = jw_nan();
          }
        }
      }
    } else if (ch == '[') {
      if (arr_heap >= 0) {
        vkind = 3;
        if (!walk_items<0, Caps>(ws, dc, v, kq + 1, tot.quotes, &N, &L, nullptr, nullptr, 0, 0)) mbad = true;
        strings += N;
      } else if (!is_entries) {
        mbad = true;  // an array under any other key
      }
      heap = arr_heap;
    } else if (ch == '-' || (ch >= '0' && ch <= '9')) {
      if (heap >= 0 || arr_heap == kHeapActions) mbad = true;  // a number where text / a list belongs
      heap = -1;
      if (act) {
        const uint32_t slot = atomicAdd(&ws.n_num, 1u);
        if (slot < (uint32_t)Caps::kNumbers) ws.num[slot] = (uint32_t)v | ((uint32_t)e << 14) | ((uint32_t)numrole << 21);
        else mbad = true;
      }
    } else {
      // null, true, false
      const uint32_t w4 = (uint32_t)ch | ((uint32_t)(v + 1 < dc.span ? ab[v + 1] : 0) << 8) |
                          ((uint32_t)(v + 2 < dc.span ? ab[v + 2] : 0) << 16) | ((uint32_t)(v + 3 < dc.span ? ab[v + 3] : 0) << 24);
      int ln = 0;
      if (w4 == 0x6c6c756eu) ln = 4;       // "null"
      else if (w4 == 0x65757274u) ln = 4;  // "true"
      else if (w4 == 0x736c6166u && v + 4 < dc.span && ab[v + 4] == 'e') ln = 5;  // "false"
      const uint8_t nc = (ln && v + ln < dc.span) ? ab[v + ln] : 0;
      if (nc != ',' && nc != '}') mbad = true;
      if (ch != 'n' && (heap >= 0 || numrole == 4)) mbad = true;  // true / false where text / a number belongs
      if (arr_heap == kHeapActions) mbad = true;
      if (heap >= 0) vkind = 1;  // null: the empty text (its offset is still written)
      if (tf >= 0) { rkind = kRecTimeKind; rlo = (uint32_t)(ch == 'n' ? PIE_TK_NULL : ch == 't' ? PIE_TK_TRUE : PIE_TK_FALSE); }
      else if (numrole == 3) rkind = kRecNanTs;
      else if (numrole == 4) rkind = kRecNullDelay;
      if (kFill && act && !mbad) {
        if (tf >= 0) {
          if (out.time_kind)
            out.time_kind[s * PIE_TF_COUNT + tf] = (uint8_t)(ch == 'n' ? PIE_TK_NULL : ch == 't' ? PIE_TK_TRUE : PIE_TK_FALSE);
        } else if (numrole == 3) {
          out.entry_ts[row] = jw_nan();
        } else if (numrole == 4) {
          out.delay_sec[row] = 0.0;
          out.delay_valid[row] = 0;
        }
      }
    }
    if (!act) {
      mbad = false;
      strings = 0;
    }
    if (mbad || !act) {
      heap = -1;
      vkind = 0;
      L = N = 0;
    }
    bad |= mbad;
    accounted += strings;
    if (!kFill && !rec_on) {
      if (heap >= 0) {
        if (L) atomicAdd(&ws.cnt[heap], L);
        if (N) atomicAdd(&ws.cnt[heap == kHeapCrew ? kPlaneCrewItems : kPlaneActionItems], N);
      }
      continue;
    }
    // ---- where this member's bytes go — the members of a heap in document order
    bool leader;
    const uint32_t pre = group_prefix(heap >= 0 ? heap : 64 + lane, L | (N << 16), lt, &leader);
    const uint32_t preL = pre & 0xffffu, preN = pre >> 16;
    const int items_plane = heap == kHeapCrew ? kPlaneCrewItems : kPlaneActionItems;
    uint32_t dst = 0, item0 = 0;
    if (heap >= 0) {
      dst = ws.cnt[heap] + preL;
      if (vkind == 3) item0 = ws.cnt[items_plane] + preN;
    }
    __syncwarp();
    if (heap >= 0 && leader) {
      ws.cnt[heap] = dst + L;
      if (vkind == 3) ws.cnt[items_plane] = item0 + N;
    }
    if (!kFill) {
      // pass 1 with records: the counters run from 0, so dst / item0 are relative to the document's part
      uint32_t rhi = 0;
      if (vkind == 1 || vkind == 2) {
        rkind = kRecText;
        rlo = (uint32_t)vsrc | ((uint32_t)vraw << 14) | ((uint32_t)(vkind == 2) << 28);
        if (L == 0) rlo = 0;  // null or '': nothing to copy
        rhi = (uint32_t)heap | ((uint32_t)e << 5) | (dst << 12);
      } else if (vkind == 3) {
        rkind = kRecArray;
        rlo = item0;
        rhi = (uint32_t)heap | ((uint32_t)e << 5);
        if (N) {  // its elements' records are written when all of the document's are counted
          const uint32_t slot = atomicAdd(&ws.n_arr, 1u);
          const uint32_t first = atomicAdd(&ws.n_items, N);
          if (slot < (uint32_t)Caps::kArrays) {
            ws.arr[slot][0] = (uint32_t)v | ((uint32_t)(kq + 1) << 14) | ((uint32_t)(heap == kHeapActions) << 26);
            ws.arr[slot][1] = dst | (item0 << 14);
            ws.arr[slot][2] = first;
          }
        }
      } else {
        rhi = (uint32_t)e << 5;
        if (rkind == kRecTimeKind || rkind == kRecTimeString) rhi = (uint32_t)tf;
      }
      if (act) slots[k] = (unsigned long long)rlo | ((unsigned long long)(rhi | (rkind << 29)) << 32);
      __syncwarp();
      continue;
    }
    if (heap >= 0) {
      if (heap >= kHeapEntry0 && heap < kHeapActions) tp.off[heap][row] = (int32_t)dst;
      if (vkind == 3) {
        if (heap == kHeapActions) out.actions_list[row] = (int32_t)item0;
        uint32_t nn, bb;
        walk_items<1, Caps>(ws, dc, v, kq + 1, tot.quotes, &nn, &bb, tp.off[heap], tp.data[heap], item0, dst);
      } else if (vkind == 2) {
        const uint32_t slot = atomicAdd(&ws.n_esc, 1u);
        if (slot < (uint32_t)kFastMaxEsc) {
          ws.esc_src[slot] = (uint32_t)vsrc | ((uint32_t)vraw << 16);
          ws.esc_dst[slot] = dst;
          ws.esc_heap[slot] = (uint8_t)heap;
        } else {
          unescape_copy(ab, vsrc, vraw, tp.data[heap] + dst, ab + dc.nwords * 32);
        }
      } else if (vkind == 1 && L <= (uint32_t)kInlineCopy) {
        uint8_t* dp = tp.data[heap] + dst;
        store_upto8(dp, load8(ab + vsrc, ab + dc.nwords * 32), (int)L);
        if (L > 8) store_upto8(dp + 8, load8(ab + vsrc + 8, ab + dc.nwords * 32), (int)L - 8);
      }
    }
    // long plain values: the whole warp copies each
    uint32_t longm = __ballot_sync(kFull, vkind == 1 && L > (uint32_t)kInlineCopy);
    while (longm) {
      const int j = __ffs(longm) - 1;
      longm &= longm - 1;
      const int src_j = __shfl_sync(kFull, vsrc, j);
      const uint32_t len_j = __shfl_sync(kFull, L, j), dst_j = __shfl_sync(kFull, dst, j);
      const int heap_j = __shfl_sync(kFull, heap, j);
      uint8_t* dp = tp.data[heap_j] + dst_j;
      for (uint32_t i = lane; i < len_j; i += 32) dp[i] = ab[src_j + i];
    }
    __syncwarp();
  }

  // ---- numbers: a lane each, the text through a register window (DocCursor: 8 aligned bytes at a time)
  __syncwarp();
  unsigned long long extra_at = 0;
  if (!kFill && rec_on) {
    if (ws.n_arr > (uint32_t)Caps::kArrays) rec_on = false;
    const unsigned long long units = (unsigned long long)ws.n_num + ws.n_items;
    if (lane == 0) extra_at = atomicAdd(rc.cursor, units);
    extra_at = __shfl_sync(kFull, extra_at, 0);
    if (extra_at + units > rc.capacity) rec_on = false;
  }
  {
    const uint32_t nn = ws.n_num < (uint32_t)Caps::kNumbers ? ws.n_num : (uint32_t)Caps::kNumbers;
    for (uint32_t i = lane; i < nn; i += 32) {
      const uint32_t rec = ws.num[i];
      const int pos = (int)(rec & kPosMask), role = (int)(rec >> 21);
      const uint32_t row = row0 + ((rec >> 14) & 127u);
      double v = 0.0;
      int term = -1;
      if (!kFill) {
        // the shape only; the record says where the number is and what it is for
        if (!number_shape(ab, pos, dc.span, &term) || (term != ',' && term != '}')) bad = true;
        if (rec_on) rc.pool[extra_at + i] = (unsigned long long)((uint32_t)role | (((rec >> 14) & 127u) << 3) | ((uint32_t)pos << 10));
        continue;
      }
      parse_number_at<true>(ab, pos, dc.span, ab + dc.nwords * 32, pow5, &v, &term);
      if (role == 3) {
        out.entry_ts[row] = jw_is_finite(v) ? v : jw_nan();
      } else if (role == 4) {
        out.delay_sec[row] = v;
        out.delay_valid[row] = 1;
      } else if (role) {
        const int tf = role == 1 ? PIE_TF_CREATED : role == 2 ? PIE_TF_ARCHIVED : role == 5 ? PIE_TF_UPDATED : PIE_TF_DELETED;
        const bool fin = jw_is_finite(v);
        if (out.time_val[tf]) out.time_val[tf][s] = fin ? v : jw_nan();
        if (out.time_kind) out.time_kind[s * PIE_TF_COUNT + tf] = (uint8_t)(fin ? PIE_TK_NUMBER : PIE_TK_NONFINITE);
      }
    }
  }
  if (kFill) {
    // ---- values with escapes: a lane each
    const uint32_t ne = ws.n_esc < (uint32_t)kFastMaxEsc ? ws.n_esc : (uint32_t)kFastMaxEsc;
    for (uint32_t i = lane; i < ne; i += 32) {
      const uint32_t sr = ws.esc_src[i];
      unescape_copy(ab, (int)(sr & 0xffffu), (int)(sr >> 16), tp.data[ws.esc_heap[i]] + ws.esc_dst[i], ab + dc.nwords * 32);
    }
    if (s == n_docs - 1) {  // the terminal offsets: where the last document ended
      __syncwarp();
      const uint32_t rows = row0 + (uint32_t)tot.entries;
      if (lane < 7) tp.off[lane][n_docs] = (int32_t)ws.cnt[lane];
      else if (lane == 7) out.entry_offsets[n_docs] = (int32_t)rows;
      else if (lane == 8) out.crew_list[n_docs] = (int32_t)ws.cnt[kPlaneCrewItems];
      else if (lane == 9) tp.off[kHeapCrew][ws.cnt[kPlaneCrewItems]] = (int32_t)ws.cnt[kHeapCrew];
      else if (lane == 10) out.actions_list[rows] = (int32_t)ws.cnt[kPlaneActionItems];
      else if (lane == 11) tp.off[kHeapActions][ws.cnt[kPlaneActionItems]] = (int32_t)ws.cnt[kHeapActions];
      else if (lane >= 12 && lane < 26) tp.off[kHeapEntry0 + lane - 12][rows] = (int32_t)ws.cnt[kHeapEntry0 + lane - 12];
    }
    return kRouteFast;
  } else {
    // ---- the elements of crew / actions: their records, a lane per array
    if (rec_on) {
      const uint32_t na = ws.n_arr;
      for (uint32_t i = lane; i < na; i += 32) {
        const uint32_t a0 = ws.arr[i][0], a1 = ws.arr[i][1];
        uint32_t nn, bb;
        walk_items<2, Caps>(ws, dc, (int)(a0 & kPosMask), (int)((a0 >> 14) & 0xfff), tot.quotes, &nn, &bb, nullptr, nullptr, a1 >> 14,
                      a1 & 0x3fffu, rc.pool + extra_at + ws.n_num + ws.arr[i][2],
                      (a0 >> 26) ? (uint32_t)kHeapActions : (uint32_t)kHeapCrew);
      }
    }
    // ---- measure: is it the shape, all of it?
    for (int e = lane; e < tot.entries; e += 32)
      if (ws.seen[e] != kAllEntryKeys) bad = true;  // an entry without one of its keys: the walk fills the gaps
#pragma unroll
    for (int dlt = 16; dlt > 0; dlt >>= 1) accounted += __shfl_xor_sync(kFull, accounted, dlt);
    if (2 * accounted != (uint32_t)tot.quotes) bad = true;  // a string that is neither key, value nor element
    if (__any_sync(kFull, bad)) return kRouteSlow;
    __syncwarp();
    if (lane == 0) {
      ws.cnt[kPlaneEntries] = (uint32_t)tot.entries;
      if (rec_on) rc.doc_rec[s] = DocRec{members_at, extra_at, (uint32_t)tot.members, ws.n_num, ws.n_items, 0u};
    }
    __syncwarp();
    if (lane < kPlanes) planes_row[lane] = ws.cnt[lane];
    return rec_on ? kRouteRecords : kRouteFast;
  }
}

// ---- pass 2 of a document with records: a scatter ----------------------------------------------------------------
template <class Caps>
__device__ __forceinline__ void fill_records(WarpShared<Caps>& ws, const TablePointers& tp, const uint8_t* __restrict__ text, int64_t from,
                                             int64_t to, int64_t s, const uint32_t* __restrict__ planes_row, const IngestOut& out,
                                             const RecCtx& rc, const Pow5Table& pow5) {
  const int lane = threadIdx.x & 31;
  const uintptr_t a0 = reinterpret_cast<uintptr_t>(text + from);
  const int skip = (int)(a0 & 31);
  const uint8_t* ab = reinterpret_cast<const uint8_t*>(a0 - skip);
  const uint8_t* limit = ab + (((skip + (int)(to - from)) + 31) & ~31);
  __syncwarp();
  if (lane < kPlanes) ws.cnt[lane] = planes_row[lane];
  const DocRec dr = rc.doc_rec[s];
  __syncwarp();
  const uint32_t row0 = ws.cnt[kPlaneEntries];
  if (lane < 7) tp.off[lane][s] = (int32_t)ws.cnt[lane];
  else if (lane == 7) out.entry_offsets[s] = (int32_t)row0;
  else if (lane == 8) out.crew_list[s] = (int32_t)ws.cnt[kPlaneCrewItems];
  else if (lane < 9 + PIE_TF_COUNT) { if (out.time_val[lane - 9]) out.time_val[lane - 9][s] = jw_nan(); }
  else if (lane == 9 + PIE_TF_COUNT) { if (out.time_kind) *reinterpret_cast<uint32_t*>(out.time_kind + s * PIE_TF_COUNT) = 0u; }
  __syncwarp();
  const unsigned long long* slots = rc.pool + dr.members_at;
  const int steps = ((int)dr.members + 31) >> 5;
  for (int t = 0; t < steps; ++t) {
    const int k = t * 32 + lane;
    unsigned long long r = 0;
    if (k < (int)dr.members) r = slots[k];
    const uint32_t lo = (uint32_t)r, hi = (uint32_t)(r >> 32);
    const uint32_t kind = hi >> 29, heap = hi & 31u, row = row0 + ((hi >> 5) & 127u);
    int src = 0;
    uint32_t len = 0, dst = 0;
    bool plain = false;
    if (kind == kRecText) {
      src = (int)(lo & kPosMask);
      len = (lo >> 14) & kPosMask;
      dst = ws.cnt[heap] + ((hi >> 12) & kPosMask);
      if (heap >= (uint32_t)kHeapEntry0) tp.off[heap][row] = (int32_t)dst;
      if ((lo >> 28) & 1u) {
        unescape_copy(ab, src, (int)len, tp.data[heap] + dst, limit);
      } else {
        plain = true;
        if (len <= (uint32_t)kLaneCopy) copy_plain(ab + src, (int)len, tp.data[heap] + dst, limit);
      }
    } else if (kind == kRecArray) {
      if (heap == (uint32_t)kHeapActions) out.actions_list[row] = (int32_t)(ws.cnt[kPlaneActionItems] + lo);
    } else if (kind == kRecNullDelay) {
      out.delay_sec[row] = 0.0;
      out.delay_valid[row] = 0;
    } else if (kind == kRecNanTs) {
      out.entry_ts[row] = jw_nan();
    } else if (kind == kRecTimeKind) {
      if (out.time_kind) out.time_kind[s * PIE_TF_COUNT + (hi & 3u)] = (uint8_t)lo;
    } else if (kind == kRecTimeString) {
      const uint32_t tf = hi & 3u;
      if (out.time_kind) out.time_kind[s * PIE_TF_COUNT + tf] = (uint8_t)PIE_TK_STRING;
      if (out.time_val[tf])
        out.time_val[tf][s] = np_bits_to_double(0x7ff8000000000000ull | ((uint64_t)((ab + lo) - out.text) & 0x7ffffffffffffull));
    }
    // long plain values: the whole warp copies each
    uint32_t longm = __ballot_sync(kFull, plain && len > (uint32_t)kLaneCopy);
    while (longm) {
      const int j = __ffs(longm) - 1;
      longm &= longm - 1;
      const int src_j = __shfl_sync(kFull, src, j);
      const uint32_t len_j = __shfl_sync(kFull, len, j), dst_j = __shfl_sync(kFull, dst, j), heap_j = __shfl_sync(kFull, heap, j);
      uint8_t* dp = tp.data[heap_j] + dst_j;
      for (uint32_t i = lane; i < len_j; i += 32) dp[i] = ab[src_j + i];
    }
  }
  const unsigned long long* extra = rc.pool + dr.extra_at;
  for (uint32_t i = lane; i < dr.numbers; i += 32) {
    // a number: converted here, from the text, by the parser the walk uses (pass 1 made sure it decides it)
    const uint32_t tag = (uint32_t)extra[i];
    double v = 0.0;
    if (tag & 7u) {  // a number under a key the table does not hold needs no value
      int term;
      parse_number_at<true, true>(ab, (int)((tag >> 10) & kPosMask), skip + (int)(to - from), limit, pow5, &v, &term);
    }
    const int role = (int)(tag & 7u);
    const uint32_t row = row0 + ((tag >> 3) & 127u);
    const bool fin = jw_is_finite(v);
    if (role == 3) {
      out.entry_ts[row] = fin ? v : jw_nan();
    } else if (role == 4) {
      out.delay_sec[row] = v;
      out.delay_valid[row] = 1;
    } else if (role) {
      const int tf = role == 1 ? PIE_TF_CREATED : role == 2 ? PIE_TF_ARCHIVED : role == 5 ? PIE_TF_UPDATED : PIE_TF_DELETED;
      if (out.time_val[tf]) out.time_val[tf][s] = fin ? v : jw_nan();
      if (out.time_kind) out.time_kind[s * PIE_TF_COUNT + tf] = (uint8_t)(fin ? PIE_TK_NUMBER : PIE_TK_NONFINITE);
    }
  }
  const unsigned long long* items = extra + dr.numbers;
  for (uint32_t i = lane; i < dr.items; i += 32) {
    const unsigned long long r = items[i];
    const uint32_t lo = (uint32_t)r, hi = (uint32_t)(r >> 32);
    const uint32_t heap = hi & 31u;
    const uint32_t dst = ws.cnt[heap] + ((hi >> 5) & kPosMask);
    const uint32_t item = ws.cnt[heap == (uint32_t)kHeapCrew ? kPlaneCrewItems : kPlaneActionItems] + (hi >> 19);
    tp.off[heap][item] = (int32_t)dst;
    const int src = (int)(lo & kPosMask), raw = (int)((lo >> 14) & kPosMask);
    if ((lo >> 28) & 1u) unescape_copy(ab, src, raw, tp.data[heap] + dst, limit);
    else copy_plain(ab + src, raw, tp.data[heap] + dst, limit);
  }
}

}  // namespace jf
}  // namespace pie
