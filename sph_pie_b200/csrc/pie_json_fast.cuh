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
//            entry index / colon index by warp scans.  Masks (quotes, brackets, escape starts) and the list of the
//            colons stay in shared memory.
//   stage 2  member-parallel projection (lane = one `"key":value` member, found by its colon): key by two 64-bit
//            compares, value by its first byte; lengths of the text values are ranked per heap in document order
//            (match_any + shuffles), copied lane-parallel (short) or by the whole warp (long); numbers and escaped
//            strings are queued and converted with all lanes busy at the end of the document.
//
// The path decides ONLY documents of the shape the provider writes (sqlProvider.js:361-409 _normalizeShow /
// _normalizeEntry through JSON.stringify, :682 / :696):
//     doc     := '{' [ member (',' member)* ] '}'                        no whitespace outside strings
//     member  := string ':' ( scalar | string | '[' strings ']' when the key is crew | '[' entries ']' when entries )
//     entry   := '{' the 17 known keys exactly once each, any order, other keys with scalar / string values '}'
//                actions -> '[' strings ']', text keys -> string | null, delaySec -> number | null
// recognised exactly (adjacency rules on the class masks + depth rules, see build_index / the member loop; every
// string and every scalar is accounted for), at most 8 KB.  EVERYTHING ELSE IS DECLINED, never guessed: the document
// goes on the list of the thread-per-document walk, which stays the complete ECMA-404 recogniser and the only place
// that reports errors.  Declining is always safe; accepting is what the damage tests in tests/test_gpu_ingest.py hold
// to the oracle byte by byte.
#pragma once
#include "pie_device.cuh"
#include "pie_json_walk.cuh"

namespace pie {
namespace jf {

using namespace jw;

constexpr int kFastWarps = 8;
constexpr int kFastThreads = kFastWarps * 32;
constexpr int kFastMaxBytes = 8192;             // alignment skip + document length a warp takes
constexpr int kFastWords = kFastMaxBytes / 32;  // 32-byte words of text
constexpr int kFastMaxMembers = 768;
constexpr int kFastMaxEntries = 128;
constexpr int kFastMaxNumbers = 192;
constexpr int kFastMaxEsc = 48;
constexpr int kInlineCopy = 8;  // longer values are copied by the whole warp
constexpr uint32_t kFull = 0xffffffffu;
constexpr uint32_t kAllEntryKeys = 0x1ffffu;  // the 17 keys of an entry
constexpr uint8_t kRouteSlow = 0, kRouteFast = 1;

struct alignas(16) WarpShared {
  uint32_t q[kFastWords];   // quotes that open or close a string
  uint32_t op[kFastWords];  // '{' '[' outside strings
  uint32_t cl[kFastWords];  // '}' ']' outside strings
  uint32_t lb[kFastWords];  // '{' outside strings
  uint32_t bs[kFastWords];  // backslashes that start an escape
  uint16_t ent[kFastWords];  // '{' before the word
  uint16_t colon[kFastMaxMembers];
  uint8_t depth[kFastWords];  // bracket depth before the word
  uint32_t cnt[kPlanes];      // measure: what the document adds to each plane; fill: the running position in it
  uint32_t seen[kFastMaxEntries];
  uint32_t num[kFastMaxNumbers];  // position | entry << 13 | role << 21
  uint32_t esc_src[kFastMaxEsc];  // position of the first byte | raw length << 16
  uint32_t esc_dst[kFastMaxEsc];  // where it goes in its heap
  uint8_t esc_heap[kFastMaxEsc];
  uint32_t n_num, n_esc, seen_show, pad;
};

// what every warp of a CTA reads per value: the table's pointers, once in shared memory (indexing the kernel
// parameter with a lane-dependent heap would serialise on the constant cache)
struct TablePointers {
  uint8_t* data[kHeaps];
  int32_t* off[kHeaps];
};

struct Doc {
  const uint8_t* ab;  // 32-byte aligned address at or before the document's first byte
  int skip, span;     // the document is ab[skip .. span)
  int nwords;
};

// ---- searches in the masks ----------------------------------------------------------------------------------------
__device__ __forceinline__ int next_bit(const uint32_t* m, int p, int nwords) {  // first set bit behind position p
  int w = p >> 5;
  uint32_t x = m[w] & (0xfffffffeu << (p & 31));
  while (!x) {
    if (++w >= nwords) return -1;
    x = m[w];
  }
  return w * 32 + __ffs(x) - 1;
}
__device__ __forceinline__ int prev_bit(const uint32_t* m, int p) {  // last set bit before position p
  int w = p >> 5;
  uint32_t x = m[w] & ((1u << (p & 31)) - 1u);
  while (!x) {
    if (--w < 0) return -1;
    x = m[w];
  }
  return w * 32 + 31 - __clz(x);
}
__device__ __forceinline__ bool any_bit_between(const uint32_t* m, int a, int b) {  // any set bit in (a, b), a < b
  const int wa = a >> 5, wb = b >> 5;
  const uint32_t above = 0xfffffffeu << (a & 31), below = (1u << (b & 31)) - 1u;
  if (wa == wb) return (m[wa] & above & below) != 0;
  if (m[wa] & above) return true;
  for (int w = wa + 1; w < wb; ++w)
    if (m[w]) return true;
  return (m[wb] & below) != 0;
}
__device__ __forceinline__ int depth_at(const WarpShared& ws, int p) {
  const int w = p >> 5;
  const uint32_t below = (1u << (p & 31)) - 1u;
  return (int)ws.depth[w] + __popc(ws.op[w] & below) - __popc(ws.cl[w] & below);
}
__device__ __forceinline__ int entry_at(const WarpShared& ws, int p) {  // inside an entry: which one
  const int w = p >> 5;
  return (int)ws.ent[w] + __popc(ws.lb[w] & ((1u << (p & 31)) - 1u)) - 2;  // the show's brace and the entry's own
}

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

// ---- escapes ------------------------------------------------------------------------------------------------------
// what the escapes of the string (v, cq) save: raw bytes minus unescaped bytes; *bad when an escape is not one of
// ECMA-404's, or names a surrogate (the walk decides those)
__device__ __forceinline__ int escape_savings(const WarpShared& ws, const uint8_t* ab, int v, int cq, bool* bad) {
  int saved = 0;
  int p = v;
  for (;;) {
    const int w0 = p >> 5;
    uint32_t x = ws.bs[w0] & (0xfffffffeu << (p & 31));
    int w = w0;
    while (!x && (w + 1) * 32 <= cq) x = ws.bs[++w];
    if (!x) break;
    p = w * 32 + __ffs(x) - 1;
    if (p >= cq) break;
    const uint8_t e = ab[p + 1];
    if (e == 'u') {
      if (p + 5 >= cq) { *bad = true; break; }
      uint32_t cp = 0;
      for (int k = 2; k < 6; ++k) {
        const int h = hex_value(ab[p + k]);
        if (h < 0) *bad = true;
        cp = cp * 16 + (uint32_t)(h & 15);
      }
      if (cp >= 0xD800 && cp <= 0xDFFF) *bad = true;
      saved += 6 - (cp < 0x80 ? 1 : cp < 0x800 ? 2 : 3);
    } else if (e == '"' || e == '\\' || e == '/' || e == 'b' || e == 'f' || e == 'n' || e == 'r' || e == 't') {
      saved += 1;
    } else {
      *bad = true;
      break;
    }
    if (*bad) break;
  }
  return saved;
}
// the unescaped bytes of ab[src .. src + raw) to dst (escapes already validated); returns how many
__device__ __forceinline__ int unescape_copy(const uint8_t* ab, int src, int raw, uint8_t* dst) {
  int i = 0, o = 0;
  while (i < raw) {
    const uint8_t c = ab[src + i];
    if (c != '\\') {
      dst[o++] = c;
      ++i;
      continue;
    }
    const uint8_t e = ab[src + i + 1];
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

// well-formed UTF-8 (Unicode table 3-7) in ab[a .. b): sequences that start here are followed to their end; bytes owed
// to a sequence that starts before a belong to the lane that owns its lead
__device__ __noinline__ bool utf8_ok(const uint8_t* ab, int a, int b, int skip, int span) {
  int owed = 0;
  for (int k = 1; k <= 3 && a - k >= skip; ++k) {
    const uint8_t c = ab[a - k];
    if ((c & 0xC0) == 0x80) continue;
    if (c >= 0xC2) {
      const int need = c >= 0xF0 ? 3 : c >= 0xE0 ? 2 : 1;
      owed = need - k + 1;
      if (owed < 0) owed = 0;
    }
    break;
  }
  int p = a + owed;
  while (p < b) {
    const uint8_t c = ab[p];
    if (c < 0x80) { ++p; continue; }
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
    p += need + 1;
  }
  return true;
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
  int members, entries, strings;
};

// Builds the masks of the document in ws; false = declined.  kCheck (the measuring pass): every rule that makes
// "accepted" mean "a document of the shape in the header" that does not need a key or a value.
template <bool kCheck>
__device__ __forceinline__ bool build_index(WarpShared& ws, const Doc& dc, IndexTotals* tot) {
  const int lane = threadIdx.x & 31;
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t c_esc = 0, c_instr = 0, c_tops = 0;
  int c_depth = 0, c_lb = 0, c_co = 0;
  uint32_t oq_count = 0;
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
    uint32_t q = 0, bsl = 0, op = 0, cl = 0, lb = 0, rb = 0, co = 0, cm = 0, ctl = 0, hib = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      // word w of the transposed copy holds the bytes at positions w, 8 + w, 16 + w, 24 + w of the lane's 32: a flag
      // in bit 7 of its byte b belongs at bit 8 b + w of the position mask, i.e. one shift by 7 - w merges it
      const int h = w >> 2, k = w & 3;
      const uint32_t sel = (uint32_t)k | ((uint32_t)(4 + k) << 4);
      const uint32_t x = __byte_perm(__byte_perm(L[h], L[2 + h], sel), __byte_perm(L[4 + h], L[6 + h], sel), 0x5410);
      const uint32_t hi = x & 0x80808080u, y = x & 0x7f7f7f7fu;
      const uint32_t z = (y | 0x20202020u) ^ 0x7b7b7b7bu;  // 0 for [ {, 6 for ] }
      const uint32_t fop = ~((z + 0x7f7f7f7fu) | hi) & 0x80808080u;
      const uint32_t fcl = ~(((z ^ 0x06060606u) + 0x7f7f7f7fu) | hi) & 0x80808080u;
      const uint32_t brace = y << 2;  // bit 5 tells { } from [ ]
      q |= eq_flags(y, hi, 0x22222222u) >> (7 - w);
      bsl |= eq_flags(y, hi, 0x5c5c5c5cu) >> (7 - w);
      op |= fop >> (7 - w);
      cl |= fcl >> (7 - w);
      lb |= (fop & brace) >> (7 - w);
      co |= eq_flags(y, hi, 0x3a3a3a3au) >> (7 - w);
      if (kCheck) {
        rb |= (fcl & brace) >> (7 - w);
        cm |= eq_flags(y, hi, 0x2c2c2c2cu) >> (7 - w);
        ctl |= (~((y + 0x60606060u) | hi) & 0x80808080u) >> (7 - w);
        hib |= hi >> (7 - w);
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
    q &= vm; bsl &= vm; op &= vm; cl &= vm; lb &= vm; co &= vm;
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
    op &= ~is; cl &= ~is; lb &= ~is; co &= ~is;
    const uint32_t oq = qr & is;
    oq_count += __popc(oq);
    // depth, entries and colons before the word
    uint32_t sa = (uint32_t)__popc(op) | ((uint32_t)__popc(cl) << 16);
    uint32_t sb = (uint32_t)__popc(lb) | ((uint32_t)__popc(co) << 16);
    const uint32_t own_a = sa, own_b = sb;
#pragma unroll
    for (int dlt = 1; dlt < 32; dlt <<= 1) {
      const uint32_t ta = __shfl_up_sync(kFull, sa, dlt), tb = __shfl_up_sync(kFull, sb, dlt);
      if (lane >= dlt) { sa += ta; sb += tb; }
    }
    const uint32_t ex_a = sa - own_a, ex_b = sb - own_b;
    const int d0 = c_depth + (int)(ex_a & 0xffff) - (int)(ex_a >> 16);
    const int lb0 = c_lb + (int)(ex_b & 0xffff);
    const int co0 = c_co + (int)(ex_b >> 16);
    const uint32_t tot_a = __shfl_sync(kFull, sa, 31), tot_b = __shfl_sync(kFull, sb, 31);
    c_depth += (int)(tot_a & 0xffff) - (int)(tot_a >> 16);
    c_lb += (int)(tot_b & 0xffff);
    c_co += (int)(tot_b >> 16);
    if (wi < dc.nwords) {
      ws.q[wi] = qr;
      ws.op[wi] = op;
      ws.cl[wi] = cl;
      ws.lb[wi] = lb;
      ws.bs[wi] = bstart;
      ws.ent[wi] = (uint16_t)lb0;
      ws.depth[wi] = (uint8_t)(d0 < 0 ? 255 : d0 > 255 ? 255 : d0);
      int idx = co0;
      for (uint32_t m = co; m; m &= m - 1, ++idx)
        if (idx < kFastMaxMembers) ws.colon[idx] = (uint16_t)(pos0 + __ffs(m) - 1);
    }
    if (kCheck) {
      rb &= vm & ~is; cm &= vm & ~is; ctl &= vm; hib &= vm;
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
      wrong |= oq & ~(p_lb | p_cm | p_co | p_lk);        // a string follows { , : [
      wrong |= lb & ~(p_lk | p_cm | first);              // { opens the document or an element
      wrong |= lk & ~p_co;                               // [ is a member's value
      wrong |= rb & ~(p_lb | p_cq | p_sc | p_rk);        // } follows { or a value
      wrong |= rk & ~(p_lk | p_cq | p_rb);               // ] follows [ or an element
      wrong |= cm & ~(p_cq | p_sc | p_rk | p_rb);        // , follows a value
      wrong |= (sc & ~p_sc) & ~p_co;                     // a scalar is a member's value
      wrong |= ctl;                                      // control characters are nowhere
      wrong |= hib & ~is;                                // bytes >= 0x80 only inside strings
      if (wrong) bad = true;
      // brackets: the kind follows from the depth ({ at 0 and 2, [ at 1 and 3), so matching pairs need no stack
      if (op | cl) {
        int d = d0;
        for (uint32_t m = op | cl; m; m &= m - 1) {
          const int bit = __ffs(m) - 1, pos = pos0 + bit;
          const bool open = (op >> bit) & 1u, brc = ((lb | rb) >> bit) & 1u;
          bool ok;
          if (open) ok = brc ? ((d == 0 && pos == dc.skip) || d == 2) : (d == 1 || d == 3);
          else ok = brc ? (d == 3 || (d == 1 && pos == dc.span - 1)) : (d == 2 || d == 4);
          if (!ok) bad = true;
          d += open ? 1 : -1;
        }
      }
      if (hib && !utf8_ok(dc.ab, pos0 > dc.skip ? pos0 : dc.skip, pos0 + 32 < dc.span ? pos0 + 32 : dc.span, dc.skip, dc.span))
        bad = true;
    }
  }
#pragma unroll
  for (int dlt = 16; dlt > 0; dlt >>= 1) oq_count += __shfl_xor_sync(kFull, oq_count, dlt);
  tot->members = c_co;
  tot->entries = c_lb - 1;
  tot->strings = (int)oq_count;
  if (c_co > kFastMaxMembers || c_lb - 1 > kFastMaxEntries) bad = true;
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

// The elements of a crew / actions array that opens at `at` ('['): strings only.  kCopy = false: counts them and their
// unescaped bytes; kCopy = true: writes offsets and bytes from (item0, dst0) on.  false = not the shape.
template <bool kCopy>
__device__ __forceinline__ bool walk_items(const WarpShared& ws, const Doc& dc, int at, uint32_t* n_items, uint32_t* n_bytes,
                                           int32_t* off, uint8_t* data, uint32_t item0, uint32_t dst0) {
  uint32_t N = 0, B = 0;
  int pos = at + 1;
  if (pos < dc.span && dc.ab[pos] != ']') {
    for (;;) {
      if (pos >= dc.span || dc.ab[pos] != '"') return false;
      const int cq = next_bit(ws.q, pos, dc.nwords);
      if (cq < 0 || cq + 1 >= dc.span) return false;
      const int raw = cq - pos - 1;
      int len = raw;
      const bool esc = raw > 0 && any_bit_between(ws.bs, pos, cq);
      if (kCopy) {
        off[item0 + N] = (int32_t)(dst0 + B);
        if (esc) len = unescape_copy(dc.ab, pos + 1, raw, data + dst0 + B);
        else
          for (int i = 0; i < raw; ++i) data[dst0 + B + i] = dc.ab[pos + 1 + i];
      } else if (esc) {
        bool bad = false;
        len = raw - escape_savings(ws, dc.ab, pos, cq, &bad);
        if (bad) return false;
      }
      B += (uint32_t)len;
      ++N;
      const uint8_t nc = dc.ab[cq + 1];
      if (nc == ']') break;
      if (nc != ',') return false;
      pos = cq + 2;
    }
  }
  *n_items = N;
  *n_bytes = B;
  return true;
}

// ---- one document -------------------------------------------------------------------------------------------------
// kFill = false: validates, counts (planes_row receives the document's 26 counts); true = accepted.
// kFill = true: planes_row holds where the document's part of every plane starts; writes its part of the table.
template <bool kFill>
__device__ __forceinline__ bool fast_doc(WarpShared& ws, const TablePointers& tp, const uint8_t* __restrict__ text, int64_t from,
                                         int64_t to, int64_t s, int64_t n_docs, uint32_t* __restrict__ planes_row,
                                         const IngestOut& out, const Pow5Table& pow5) {
  const int lane = threadIdx.x & 31;
  const uint32_t lt = (1u << lane) - 1u;
  const int64_t n = to - from;
  if (n < 2) return false;
  Doc dc;
  {
    const uintptr_t a0 = reinterpret_cast<uintptr_t>(text + from);
    dc.skip = (int)(a0 & 31);
    dc.ab = reinterpret_cast<const uint8_t*>(a0 - dc.skip);
  }
  if (dc.skip + n > kFastMaxBytes) return false;
  dc.span = dc.skip + (int)n;
  dc.nwords = (dc.span + 31) >> 5;
  const uint8_t* ab = dc.ab;

  __syncwarp();
  if (lane == 0) { ws.n_num = 0; ws.n_esc = 0; ws.seen_show = 0; }
  if (lane < kPlanes) ws.cnt[lane] = kFill ? planes_row[lane] : 0u;
  IndexTotals tot;
  if (!build_index<!kFill>(ws, dc, &tot)) return false;
  if (!kFill)
    for (int e = lane; e < tot.entries; e += 32) ws.seen[e] = 0;
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
    int heap = -1;       // the heap that receives bytes from this member
    uint32_t L = 0, N = 0;
    int vkind = 0;       // 1 plain string, 2 string with escapes, 3 elements of an array
    int vsrc = 0, vraw = 0, arr_at = 0;
    uint32_t row = 0;
    if (k < tot.members) {
      const int c = ws.colon[k];
      const int kc = c - 1;
      const int ko = prev_bit(ws.q, kc);
      const int d = depth_at(ws, c);
      bool mbad = ko < 0 || !(d == 1 || d == 3);
      if (!kFill && !mbad) {
        const uint8_t pk = ab[ko - 1];
        if (!(pk == '{' || pk == ',')) mbad = true;          // the string before the colon is a key
        if (kc - ko > 1 && any_bit_between(ws.bs, ko, kc)) mbad = true;  // an escape in a key: the walk's business
      }
      if (!mbad) {
        const int klen = kc - ko - 1;
        int key = -1;
        if (klen >= 1 && klen <= 16) {
          uint64_t k0, k1;
          load_key(ab + ko + 1, klen, ab + dc.nwords * 32, &k0, &k1);
          key = d == 1 ? match_show_key((uint32_t)klen, k0, k1) : match_entry_key((uint32_t)klen, k0, k1);
        }
        int e = 0;
        if (d == 3) {
          e = entry_at(ws, c);
          row = row0 + (uint32_t)e;
        }
        if (!kFill && key >= 0) {
          const uint32_t bit = 1u << key;
          const uint32_t old = d == 1 ? atomicOr(&ws.seen_show, bit) : atomicOr(&ws.seen[e], bit);
          if (old & bit) mbad = true;  // a known key twice
        }
        // what the value is for
        int tf = -1, numrole = 0, arr_heap = -1;
        bool is_entries = false;
        if (d == 1) {
          if (key >= 0 && key < 7) heap = key;
          else if (key == kSkCreatedAt) { tf = PIE_TF_CREATED; numrole = 1; }
          else if (key == kSkArchivedAt) { tf = PIE_TF_ARCHIVED; numrole = 2; }
          else if (key == kSkUpdatedAt) { tf = PIE_TF_UPDATED; numrole = 5; }
          else if (key == kSkDeletedAt) { tf = PIE_TF_DELETED; numrole = 6; }
          else if (key == kSkCrew) arr_heap = kHeapCrew;
          else if (key == kSkEntries) is_entries = true;
        } else {
          if (key >= 0 && key < 14) heap = kHeapEntry0 + key;
          else if (key == kEkTs) numrole = 3;
          else if (key == kEkDelaySec) numrole = 4;
          else if (key == kEkActions) arr_heap = kHeapActions;
        }
        const int v = c + 1;
        const uint8_t ch = v < dc.span ? ab[v] : 0;
        accounted += 1;
        if (ch == '"') {
          const int cq = next_bit(ws.q, v, dc.nwords);
          if (cq < 0 || cq + 1 >= dc.span) {
            mbad = true;
          } else {
            accounted += 1;
            if (!kFill) {
              const uint8_t nc = ab[cq + 1];
              if (nc != ',' && nc != '}') mbad = true;
            }
            const int raw = cq - v - 1;
            const bool esc = raw > 0 && any_bit_between(ws.bs, v, cq);
            if (heap >= 0) {
              vsrc = v + 1;
              vraw = raw;
              L = (uint32_t)raw;
              vkind = 1;
              if (esc) {
                bool ebad = false;
                L = (uint32_t)(raw - escape_savings(ws, ab, v, cq, &ebad));
                vkind = 2;
                if (ebad) mbad = true;
              }
            } else {
              if (esc) mbad = true;  // nobody unescapes it here, so nobody validates it: the walk does
              if (numrole == 4 || arr_heap == kHeapActions) mbad = true;  // delaySec / actions that are text
              if (kFill) {
                if (tf >= 0) {
                  if (out.time_kind) out.time_kind[s * PIE_TF_COUNT + tf] = (uint8_t)PIE_TK_STRING;
                  if (out.time_val[tf])
                    out.time_val[tf][s] =
                        np_bits_to_double(0x7ff8000000000000ull | ((uint64_t)((ab + v + 1) - out.text) & 0x7ffffffffffffull));
                } else if (numrole == 3) {
                  out.entry_ts[row] = jw_nan();
                }
              }
            }
          }
        } else if (ch == '[') {
          if (arr_heap >= 0) {
            heap = arr_heap;
            vkind = 3;
            arr_at = v;
            if (!walk_items<false>(ws, dc, v, &N, &L, nullptr, nullptr, 0, 0)) mbad = true;
            accounted += N;
          } else if (!is_entries) {
            mbad = true;  // an array under any other key
          }
          heap = arr_heap;
        } else if (ch == '-' || (ch >= '0' && ch <= '9')) {
          if (heap >= 0 || arr_heap == kHeapActions) mbad = true;  // a number where text / a list belongs
          heap = -1;
          const uint32_t slot = atomicAdd(&ws.n_num, 1u);
          if (slot < (uint32_t)kFastMaxNumbers) ws.num[slot] = (uint32_t)v | ((uint32_t)e << 13) | ((uint32_t)numrole << 21);
          else mbad = true;
        } else {
          const char* lit = ch == 'n' ? "null" : ch == 't' ? "true" : ch == 'f' ? "false" : nullptr;
          int ln = 0;
          if (!lit) {
            mbad = true;
          } else {
            for (; lit[ln]; ++ln)
              if (v + ln >= dc.span || ab[v + ln] != (uint8_t)lit[ln]) mbad = true;
            const uint8_t nc = v + ln < dc.span ? ab[v + ln] : 0;
            if (nc != ',' && nc != '}') mbad = true;
          }
          if (ch != 'n' && (heap >= 0 || numrole == 4)) mbad = true;  // true / false where text / a number belongs
          if (arr_heap == kHeapActions) mbad = true;
          if (heap >= 0) vkind = 1;  // null: the empty text (its offset is still written)
          if (kFill && !mbad) {
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
      }
      if (mbad) {
        bad = true;
        heap = -1;
        vkind = 0;
        L = N = 0;
      }
    }
    if (!kFill) {
      if (heap >= 0) {
        if (L) atomicAdd(&ws.cnt[heap], L);
        if (N) atomicAdd(&ws.cnt[heap == kHeapCrew ? kPlaneCrewItems : kPlaneActionItems], N);
      }
      continue;
    }
    // ---- fill: where this member's bytes go — the members of a heap in document order
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
    if (heap >= 0) {
      if (heap >= kHeapEntry0 && heap < kHeapActions) tp.off[heap][row] = (int32_t)dst;
      if (vkind == 3) {
        if (heap == kHeapActions) out.actions_list[row] = (int32_t)item0;
        uint32_t nn, bb;
        walk_items<true>(ws, dc, arr_at, &nn, &bb, tp.off[heap], tp.data[heap], item0, dst);
      } else if (vkind == 2) {
        const uint32_t slot = atomicAdd(&ws.n_esc, 1u);
        if (slot < (uint32_t)kFastMaxEsc) {
          ws.esc_src[slot] = (uint32_t)vsrc | ((uint32_t)vraw << 16);
          ws.esc_dst[slot] = dst;
          ws.esc_heap[slot] = (uint8_t)heap;
        } else {
          unescape_copy(ab, vsrc, vraw, tp.data[heap] + dst);
        }
      } else if (vkind == 1 && L <= (uint32_t)kInlineCopy) {
        uint8_t* dp = tp.data[heap] + dst;
        for (uint32_t i = 0; i < L; ++i) dp[i] = ab[vsrc + i];
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

  // ---- numbers: a lane each
  __syncwarp();
  {
    const uint32_t nn = ws.n_num < (uint32_t)kFastMaxNumbers ? ws.n_num : (uint32_t)kFastMaxNumbers;
    for (uint32_t i = lane; i < nn; i += 32) {
      const uint32_t rec = ws.num[i];
      const int pos = (int)(rec & 8191u), role = (int)(rec >> 21);
      const uint32_t row = row0 + ((rec >> 13) & 255u);
      MemSource src{ab, (int64_t)dc.span, (int64_t)pos};
      double v = 0.0;
      const int rc = role ? parse_json_number_from<true>(src, pow5, &v) : parse_json_number_from<false>(src, pow5, &v);
      if (!kFill) {
        const uint8_t term = src.i < dc.span ? ab[src.i] : 0;
        if (rc != kNumOk || (term != ',' && term != '}')) bad = true;
      } else if (role == 3) {
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
      unescape_copy(ab, (int)(sr & 0xffffu), (int)(sr >> 16), tp.data[ws.esc_heap[i]] + ws.esc_dst[i]);
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
    return true;
  } else {
    // ---- measure: is it the shape, all of it?
    for (int e = lane; e < tot.entries; e += 32)
      if (ws.seen[e] != kAllEntryKeys) bad = true;  // an entry without one of its keys: the walk fills the gaps
#pragma unroll
    for (int dlt = 16; dlt > 0; dlt >>= 1) accounted += __shfl_xor_sync(kFull, accounted, dlt);
    if (accounted != (uint32_t)tot.strings) bad = true;  // a string that is neither key, value nor element
    if (__any_sync(kFull, bad)) return false;
    __syncwarp();
    if (lane == 0) ws.cnt[kPlaneEntries] = (uint32_t)tot.entries;
    __syncwarp();
    if (lane < kPlanes) planes_row[lane] = ws.cnt[lane];
    return true;
  }
}

}  // namespace jf
}  // namespace pie
