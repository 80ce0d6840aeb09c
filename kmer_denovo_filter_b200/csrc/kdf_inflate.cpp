// kdf_inflate.cpp — raw DEFLATE (RFC 1951) decoder and CRC-32 for whole BGZF blocks.
//
// Why not zlib's inflate(): a BAM scan is inflate-bound on the host (the decoder threads
// spend ~2/3 of their time in it) and zlib's streaming decoder is built for input/output
// that arrives in pieces.  A BGZF block is different: <= 64 KiB, entirely in memory, its
// exact inflated size known up front (ISIZE), never referring to an earlier block.  That
// permits a decoder with a 64-bit bit buffer refilled by one unaligned load, one table
// look-up per symbol (11-bit litlen / 8-bit distance primary tables with sub-tables for
// the longer codes), up to three literals per refill and 16-byte match copies — while the
// last few hundred bytes of each block go through a byte-exact loop so nothing is ever
// written outside [out, out + out_len) (neighbouring blocks are being inflated by other
// threads into the same chunk).
//
// This replaces what samtools/htslib does inside the reference's `samtools fasta` /
// pysam reads (reference: src/kmer_denovo_filter/kmer_utils.py:310-340 pipes
// `samtools fasta`; htslib's bgzf.c inflates each block with zlib or libdeflate).
// zlib stays linked: `KDF_BAM_ZLIB=1` routes every block through it (A/B + fallback) and
// its crc32() covers CPUs without PCLMULQDQ and the < 64-byte tails.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "kdf_inflate.h"

namespace kdf {
namespace {

constexpr int LL_BITS = 11, D_BITS = 8, PC_BITS = 7;
constexpr uint32_t LL_MASK = (1u << LL_BITS) - 1, D_MASK = (1u << D_BITS) - 1;
// table entry: bits 0-3 code bits to consume, 4-5 number of literals it yields (1, or 2 when
// two literal codes fit the primary index together), 8-11 extra bits (for F_SUB: index bits
// of the sub-table), 12-14 flags, 16-31 payload (literal(s) / base length / base distance /
// sub-table start)
constexpr uint32_t F_LIT = 0x10, F_LIT2 = 0x20, F_ANYLIT = 0x30, F_EOB = 0x1000, F_BAD = 0x2000, F_SUB = 0x4000;
constexpr int LL_SIZE = (1 << LL_BITS) + 288 * 16, D_SIZE = (1 << D_BITS) + 32 * 128;

struct Tables {
  uint32_t ll[LL_SIZE];
  uint32_t d[D_SIZE];
};

const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31,
                               35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513,
                                769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10,
                                11, 11, 12, 12, 13, 13};

inline uint32_t ll_entry(int sym) {
  if (sym < 256) return F_LIT | ((uint32_t)sym << 16);
  if (sym == 256) return F_EOB;
  if (sym < 286) return ((uint32_t)LEN_BASE[sym - 257] << 16) | ((uint32_t)LEN_EXTRA[sym - 257] << 8);
  return F_BAD;
}
inline uint32_t d_entry(int sym) {
  if (sym < 30) return ((uint32_t)DIST_BASE[sym] << 16) | ((uint32_t)DIST_EXTRA[sym] << 8);
  return F_BAD;
}
inline uint32_t pc_entry(int sym) { return (uint32_t)sym << 16; }

// Canonical Huffman code of lens[0..n) -> look-up table indexed by the next `tb` stream
// bits (LSB first).  `kind`: 0 litlen, 1 distance, 2 code-length code.  A code set that
// leaves part of the code space unused is accepted only when it is a single 1-bit litlen or
// distance code (what zlib accepts; encoders emit it for a block with one distance), the
// unused half then decodes to F_BAD.  No code at all is legal for distances (a literal-only block).
//
// Symbols are taken in canonical order (by length, then value) while `code` walks the
// codewords in bit-reversed form, which is how the stream presents them: the next codeword
// of the same length is a bit-reversed increment, and stepping to a longer length leaves the
// reversed value as it is.  The primary table grows by doubling (after the codes of length l
// are placed in its first 2^l entries, that prefix is copied once to make 2^(l+1)), so its
// cost is one store per symbol plus 2^tb copied entries; codes longer than tb go to
// sub-tables, one per distinct tb-bit prefix, which are contiguous in canonical order.
bool build_table(uint32_t* table, int tb, const uint8_t* lens, int n, int kind) {
  int count[16] = {0};
  for (int i = 0; i < n; ++i) ++count[lens[i]];
  count[0] = 0;
  int maxlen = 0, total = 0;
  for (int l = 1; l <= 15; ++l)
    if (count[l]) maxlen = l, total += count[l];
  const uint32_t tsize = 1u << tb;
  if (total == 0) {
    if (kind != 1) return false;
    for (uint32_t i = 0; i < tsize; ++i) table[i] = F_BAD | 1;
    return true;
  }
  int left = 1;
  for (int l = 1; l <= 15; ++l) {
    left = (left << 1) - count[l];
    if (left < 0) return false;   // over-subscribed
  }
  auto entry = [kind](int sym) { return kind == 0 ? ll_entry(sym) : kind == 1 ? d_entry(sym) : pc_entry(sym); };
  if (left > 0) {
    if (maxlen != 1 || kind == 2) return false;   // incomplete (zlib: never allowed for the code-length code)
    int sym = 0;
    while (lens[sym] != 1) ++sym;
    for (uint32_t i = 0; i < tsize; i += 2) table[i] = entry(sym) | 1u, table[i + 1] = F_BAD | 1;
    return true;
  }
  // counting sort by code length
  uint16_t sorted[288];
  int first[17];
  first[1] = 0;
  for (int l = 1; l <= 15; ++l) first[l + 1] = first[l] + count[l];
  {
    int at[16];
    for (int l = 1; l <= 15; ++l) at[l] = first[l];
    for (int s = 0; s < n; ++s)
      if (lens[s]) sorted[at[lens[s]]++] = (uint16_t)s;
  }
  uint16_t lit_rev[256];   // reversed codeword of each literal shorter than tb (for the pairs below)
  uint32_t code = 0;       // bit-reversed codeword of the next symbol
  const int top = maxlen < tb ? maxlen : tb;
  for (int l = 1; l <= top; ++l) {
    for (int i = first[l]; i < first[l + 1]; ++i) {
      const int sym = sorted[i];
      table[code] = entry(sym) | (uint32_t)l;
      if (kind == 0 && sym < 256 && l < tb) lit_rev[sym] = (uint16_t)code;
      uint32_t bit = 1u << (l - 1);
      while (code & bit) bit >>= 1;
      code = bit ? ((code & (bit - 1)) | bit) : 0;   // 0: that was the all-ones codeword, the last one
    }
    if (l < tb) memcpy(table + (1u << l), table, sizeof(uint32_t) << l);
  }
  for (int l = top + 1; l < tb; ++l) memcpy(table + (1u << l), table, sizeof(uint32_t) << l);
  if (kind == 0) {
    // two literals in one look-up: every index that starts with the code of `a` followed by
    // the code of `b` yields both (quality strings and packed bases are literal runs, and
    // the look-up -> shift -> look-up dependency is what bounds a literal run)
    for (int la = 1; la < tb; ++la)
      for (int ia = first[la]; ia < first[la + 1]; ++ia) {
        const uint32_t a = sorted[ia];
        if (a >= 256) continue;
        const uint32_t ra = lit_rev[a];
        for (int lb = 1; la + lb <= tb && lb <= maxlen; ++lb)
          for (int ib = first[lb]; ib < first[lb + 1]; ++ib) {
            const uint32_t b = sorted[ib];
            if (b >= 256) continue;
            const uint32_t e = F_LIT2 | (uint32_t)(la + lb) | ((a | (b << 8)) << 16);
            for (uint32_t i = ra | ((uint32_t)lit_rev[b] << la); i < tsize; i += 1u << (la + lb)) table[i] = e;
          }
      }
  }
  // codes longer than tb
  uint32_t next_free = tsize, cur_prefix = ~0u, sub_start = 0, sub_bits = 0;
  for (int l = tb + 1; l <= maxlen; ++l) {
    for (int i = first[l]; i < first[l + 1]; ++i) {
      const uint32_t prefix = code & (tsize - 1);
      if (prefix != cur_prefix) {
        // size the sub-table: it must hold every remaining code that shares this prefix, i.e.
        // grow it until the codes of the lengths it spans fill it (zlib's inftrees.c does the same)
        cur_prefix = prefix;
        sub_bits = (uint32_t)(l - tb);
        int room = 1 << sub_bits, ll_ = l, remaining = first[l + 1] - i;
        while (ll_ < maxlen) {
          room -= remaining;
          if (room <= 0) break;
          ++sub_bits, ++ll_;
          room <<= 1;
          remaining = count[ll_];
        }
        sub_start = next_free;
        next_free += 1u << sub_bits;
        table[prefix] = F_SUB | (uint32_t)tb | (sub_bits << 8) | (sub_start << 16);
      }
      const uint32_t e = entry(sorted[i]) | (uint32_t)(l - tb);
      for (uint32_t j = code >> tb; j < (1u << sub_bits); j += 1u << (l - tb)) table[sub_start + j] = e;
      uint32_t bit = 1u << (l - 1);
      while (code & bit) bit >>= 1;
      code = bit ? ((code & (bit - 1)) | bit) : 0;
    }
  }
  return true;
}

struct St {
  uint64_t buf = 0;
  unsigned cnt = 0;        // valid bits in buf (bits above them are either zero or the true next bits)
  const uint8_t* p;        // next byte to load
  const uint8_t* in_end;
  const uint8_t* in_fast;  // p <= in_fast: 8-byte loads at p .. p + 21 stay inside the input (null: never)
  uint8_t* out;
  uint8_t* out0;
  uint8_t* out_end;
  uint8_t* out_fast;       // out <= out_fast: 6 literals + a 258-byte match + 15 bytes of copy slack fit (null: never)
  unsigned overrun = 0;    // zero bytes supplied past the end of the input
};

#define KDF_REFILL_FAST(buf, cnt, p)  \
  do {                                \
    uint64_t w_;                      \
    memcpy(&w_, (p), 8);              \
    (buf) |= w_ << (cnt);             \
    (p) += (63 - (cnt)) >> 3;         \
    (cnt) |= 56;                      \
  } while (0)

// >= 57 bits afterwards; past the end of the input zero bits are supplied and counted
inline void refill(St& s) {
  if (s.in_fast && s.p <= s.in_fast) {
    KDF_REFILL_FAST(s.buf, s.cnt, s.p);
    return;
  }
  while (s.cnt <= 56) {
    if (s.p < s.in_end)
      s.buf |= (uint64_t)*s.p++ << s.cnt;
    else
      ++s.overrun;
    s.cnt += 8;
  }
}
inline uint32_t take(St& s, unsigned n) {
  uint32_t v = (uint32_t)(s.buf & ((1ull << n) - 1));
  s.buf >>= n;
  s.cnt -= n;
  return v;
}

// The bulk of a Huffman block.  → 0 end of block, 1 left the fast region (continue in
// slow_loop), -1 corrupt.
__attribute__((target_clones("bmi2", "default"))) int fast_loop(St& s, const uint32_t* __restrict ll, const uint32_t* __restrict dt) {
  uint64_t buf = s.buf;
  unsigned cnt = s.cnt;
  const uint8_t* p = s.p;
  uint8_t* out = s.out;
  uint8_t* const out0 = s.out0;
  const uint8_t* const in_fast = s.in_fast;
  uint8_t* const out_fast = s.out_fast;
  int ret = 1;
  if (in_fast && out_fast)
    while (p <= in_fast && out <= out_fast) {
      KDF_REFILL_FAST(buf, cnt, p);
      uint32_t e = ll[buf & LL_MASK];
      if (e & F_ANYLIT) {   // up to three look-ups (six literals) on one refill (3 x 15 bits <= 56)
        uint16_t lit = (uint16_t)(e >> 16);
        buf >>= e & 15, cnt -= e & 15;
        memcpy(out, &lit, 2);
        out += (e >> 4) & 3;
        e = ll[buf & LL_MASK];
        if (e & F_ANYLIT) {
          lit = (uint16_t)(e >> 16);
          buf >>= e & 15, cnt -= e & 15;
          memcpy(out, &lit, 2);
          out += (e >> 4) & 3;
          e = ll[buf & LL_MASK];
          if (e & F_ANYLIT) {
            lit = (uint16_t)(e >> 16);
            buf >>= e & 15, cnt -= e & 15;
            memcpy(out, &lit, 2);
            out += (e >> 4) & 3;
            continue;
          }
        }
        KDF_REFILL_FAST(buf, cnt, p);   // e stays valid: the low bits do not change
      }
      if (e & F_SUB) {
        buf >>= LL_BITS, cnt -= LL_BITS;
        e = ll[(e >> 16) + (uint32_t)(buf & ((1u << ((e >> 8) & 15)) - 1))];
        if (e & F_LIT) {
          buf >>= e & 15, cnt -= e & 15;
          *out++ = (uint8_t)(e >> 16);
          continue;
        }
      }
      if (e & (F_EOB | F_BAD)) {
        if (e & F_BAD) {
          ret = -1;
        } else {
          buf >>= e & 15, cnt -= e & 15;
          ret = 0;
        }
        break;
      }
      buf >>= e & 15, cnt -= e & 15;
      unsigned xb = (e >> 8) & 15;
      const unsigned len = (e >> 16) + (unsigned)(buf & ((1u << xb) - 1));
      buf >>= xb, cnt -= xb;
      uint32_t de = dt[buf & D_MASK];
      if (de & F_SUB) {
        buf >>= D_BITS, cnt -= D_BITS;
        de = dt[(de >> 16) + (uint32_t)(buf & ((1u << ((de >> 8) & 15)) - 1))];
      }
      if (de & F_BAD) {
        ret = -1;
        break;
      }
      buf >>= de & 15, cnt -= de & 15;
      xb = (de >> 8) & 15;
      const unsigned dist = (de >> 16) + (unsigned)(buf & ((1u << xb) - 1));
      buf >>= xb, cnt -= xb;
      if (dist > (size_t)(out - out0)) {
        ret = -1;
        break;
      }
      const uint8_t* src = out - dist;
      uint8_t* const dend = out + len;
      if (dist >= 16) {
        do {
          memcpy(out, src, 16);
          out += 16, src += 16;
        } while (out < dend);
      } else if (dist == 1) {
        const uint64_t v = 0x0101010101010101ull * *src;
        do {
          memcpy(out, &v, 8);
          out += 8;
        } while (out < dend);
      } else if (dist >= 8) {
        do {
          memcpy(out, src, 8);
          out += 8, src += 8;
        } while (out < dend);
      } else {
        do *out++ = *src++;
        while (out < dend);
      }
      out = dend;
    }
  s.buf = buf, s.cnt = cnt, s.p = p, s.out = out;
  return ret;
}

// The edges of a block: bounds-checked input and output, byte copies.  → 0 / -1.
int slow_loop(St& s, const uint32_t* ll, const uint32_t* dt) {
  for (;;) {
    refill(s);
    if (s.overrun > 16) return -1;
    uint32_t e = ll[s.buf & LL_MASK];
    if (e & F_SUB) {
      take(s, LL_BITS);
      e = ll[(e >> 16) + (uint32_t)(s.buf & ((1u << ((e >> 8) & 15)) - 1))];
    }
    if (e & F_BAD) return -1;
    take(s, e & 15);
    if (e & F_ANYLIT) {
      const unsigned nl = (e >> 4) & 3;
      if (nl > (size_t)(s.out_end - s.out)) return -1;
      *s.out++ = (uint8_t)(e >> 16);
      if (nl == 2) *s.out++ = (uint8_t)(e >> 24);
      continue;
    }
    if (e & F_EOB) return 0;
    const unsigned len = (e >> 16) + take(s, (e >> 8) & 15);
    uint32_t de = dt[s.buf & D_MASK];
    if (de & F_SUB) {
      take(s, D_BITS);
      de = dt[(de >> 16) + (uint32_t)(s.buf & ((1u << ((de >> 8) & 15)) - 1))];
    }
    if (de & F_BAD) return -1;
    take(s, de & 15);
    const unsigned dist = (de >> 16) + take(s, (de >> 8) & 15);
    if (dist > (size_t)(s.out - s.out0) || len > (size_t)(s.out_end - s.out)) return -1;
    const uint8_t* src = s.out - dist;
    for (unsigned i = 0; i < len; ++i) s.out[i] = src[i];
    s.out += len;
  }
}

const Tables* fixed_tables() {
  static const Tables* t = [] {
    Tables* f = new Tables;
    uint8_t lens[288];
    for (int i = 0; i < 144; ++i) lens[i] = 8;
    for (int i = 144; i < 256; ++i) lens[i] = 9;
    for (int i = 256; i < 280; ++i) lens[i] = 7;
    for (int i = 280; i < 288; ++i) lens[i] = 8;
    build_table(f->ll, LL_BITS, lens, 288, 0);
    uint8_t dl[32];
    for (int i = 0; i < 32; ++i) dl[i] = 5;
    build_table(f->d, D_BITS, dl, 32, 1);
    return f;
  }();
  return t;
}

bool read_dynamic(St& s, Tables& t) {
  refill(s);
  const unsigned hlit = take(s, 5) + 257, hdist = take(s, 5) + 1, hclen = take(s, 4) + 4;
  if (hlit > 286 || hdist > 30) return false;
  static const uint8_t ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  uint8_t pl[19] = {0};
  for (unsigned i = 0; i < hclen; ++i) {
    if (s.cnt < 3) refill(s);
    pl[ORDER[i]] = (uint8_t)take(s, 3);
  }
  uint32_t pc[1 << PC_BITS];
  if (!build_table(pc, PC_BITS, pl, 19, 2)) return false;
  uint8_t lens[286 + 30 + 138];
  unsigned i = 0;
  const unsigned n = hlit + hdist;
  while (i < n) {
    refill(s);
    if (s.overrun > 16) return false;
    const uint32_t e = pc[s.buf & ((1u << PC_BITS) - 1)];
    if (e & F_BAD) return false;
    take(s, e & 15);
    const unsigned sym = e >> 16;
    if (sym < 16) {
      lens[i++] = (uint8_t)sym;
      continue;
    }
    unsigned rep;
    uint8_t v = 0;
    if (sym == 16) {
      if (i == 0) return false;
      v = lens[i - 1];
      rep = 3 + take(s, 2);
    } else if (sym == 17) {
      rep = 3 + take(s, 3);
    } else {
      rep = 11 + take(s, 7);
    }
    if (i + rep > n) return false;
    memset(lens + i, v, rep);
    i += rep;
  }
  if (lens[256] == 0) return false;   // no end-of-block code
  return build_table(t.ll, LL_BITS, lens, (int)hlit, 0) && build_table(t.d, D_BITS, lens + hlit, (int)hdist, 1);
}

}  // namespace

bool inflate_raw(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len) {
  St s;
  s.p = in;
  s.in_end = in + in_len;
  s.in_fast = in_len >= 32 ? in + in_len - 32 : nullptr;
  s.out = s.out0 = out;
  s.out_end = out + out_len;
  s.out_fast = out_len >= 320 ? out + out_len - 320 : nullptr;
  static thread_local Tables* dyn = nullptr;
  for (;;) {
    refill(s);
    if (s.overrun > 16) return false;
    const unsigned bfinal = take(s, 1), btype = take(s, 2);
    if (btype == 0) {
      // stored: to the byte boundary, LEN, ~LEN, then LEN bytes straight from the input
      take(s, s.cnt & 7);
      if (s.cnt < s.overrun * 8) return false;   // the header itself lay past the input
      s.p -= (s.cnt - s.overrun * 8) >> 3;       // hand the unread whole bytes back (not the padding)
      s.buf = 0, s.cnt = 0, s.overrun = 0;
      if ((size_t)(s.in_end - s.p) < 4) return false;
      const unsigned len = s.p[0] | (s.p[1] << 8), nlen = s.p[2] | (s.p[3] << 8);
      if ((len ^ nlen) != 0xFFFFu) return false;
      s.p += 4;
      if (len > (size_t)(s.in_end - s.p) || len > (size_t)(s.out_end - s.out)) return false;
      memcpy(s.out, s.p, len);
      s.p += len, s.out += len;
    } else if (btype == 3) {
      return false;
    } else {
      const Tables* t;
      if (btype == 1) {
        t = fixed_tables();
      } else {
        if (!dyn) dyn = new Tables;   // one per decoder thread, for the life of the thread
        if (!read_dynamic(s, *dyn)) return false;
        t = dyn;
      }
      int r = fast_loop(s, t->ll, t->d);
      if (r == 1) r = slow_loop(s, t->ll, t->d);
      if (r < 0) return false;
    }
    if (bfinal) break;
  }
  if ((size_t)s.overrun * 8 > s.cnt) return false;   // symbols were decoded from bits past the input
  return s.out == s.out_end;
}

// ---- CRC-32 (IEEE 802.3, the gzip polynomial) -------------------------------------------------
#if defined(__x86_64__)
// Carry-less-multiply folding ("Fast CRC Computation for Generic Polynomials Using PCLMULQDQ
// Instruction", Gopal et al., Intel 2009): 64 bytes per iteration folded onto four 128-bit
// accumulators, then 4 -> 1, 128 -> 64 bits and a Barrett reduction.  `len` >= 64 and a
// multiple of 16; `crc` is the raw (pre-inverted) register.  Checked against zlib's crc32()
// in tests/test_host_inflate.py.
__attribute__((target("pclmul,sse4.1"))) uint32_t crc32_clmul(const uint8_t* buf, size_t len, uint32_t crc) {
  alignas(16) static const uint64_t k1k2[2] = {0x0154442bd4ull, 0x01c6e41596ull};   // x^(4*128+32), x^(4*128-32) mod P
  alignas(16) static const uint64_t k3k4[2] = {0x01751997d0ull, 0x00ccaa009eull};   // x^(128+32), x^(128-32) mod P
  alignas(16) static const uint64_t k5k0[2] = {0x0163cd6124ull, 0x0000000000ull};   // x^64 mod P
  alignas(16) static const uint64_t poly[2] = {0x01db710641ull, 0x01f7011641ull};   // P, floor(x^64 / P)
  __m128i x0, x1, x2, x3, x4, x5, x6, x7, x8, y5, y6, y7, y8;
  x1 = _mm_loadu_si128((const __m128i*)(buf + 0x00));
  x2 = _mm_loadu_si128((const __m128i*)(buf + 0x10));
  x3 = _mm_loadu_si128((const __m128i*)(buf + 0x20));
  x4 = _mm_loadu_si128((const __m128i*)(buf + 0x30));
  x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)crc));
  x0 = _mm_load_si128((const __m128i*)k1k2);
  buf += 64, len -= 64;
  while (len >= 64) {
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x6 = _mm_clmulepi64_si128(x2, x0, 0x00);
    x7 = _mm_clmulepi64_si128(x3, x0, 0x00);
    x8 = _mm_clmulepi64_si128(x4, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x2 = _mm_clmulepi64_si128(x2, x0, 0x11);
    x3 = _mm_clmulepi64_si128(x3, x0, 0x11);
    x4 = _mm_clmulepi64_si128(x4, x0, 0x11);
    y5 = _mm_loadu_si128((const __m128i*)(buf + 0x00));
    y6 = _mm_loadu_si128((const __m128i*)(buf + 0x10));
    y7 = _mm_loadu_si128((const __m128i*)(buf + 0x20));
    y8 = _mm_loadu_si128((const __m128i*)(buf + 0x30));
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x5), y5);
    x2 = _mm_xor_si128(_mm_xor_si128(x2, x6), y6);
    x3 = _mm_xor_si128(_mm_xor_si128(x3, x7), y7);
    x4 = _mm_xor_si128(_mm_xor_si128(x4, x8), y8);
    buf += 64, len -= 64;
  }
  x0 = _mm_load_si128((const __m128i*)k3k4);
  x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
  x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
  x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
  x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
  x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
  x1 = _mm_xor_si128(_mm_xor_si128(x1, x3), x5);
  x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
  x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
  x1 = _mm_xor_si128(_mm_xor_si128(x1, x4), x5);
  while (len >= 16) {
    x2 = _mm_loadu_si128((const __m128i*)buf);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
    buf += 16, len -= 16;
  }
  x2 = _mm_clmulepi64_si128(x1, x0, 0x10);
  x3 = _mm_setr_epi32(~0, 0, ~0, 0);
  x1 = _mm_srli_si128(x1, 8);
  x1 = _mm_xor_si128(x1, x2);
  x0 = _mm_loadl_epi64((const __m128i*)k5k0);
  x2 = _mm_srli_si128(x1, 4);
  x1 = _mm_and_si128(x1, x3);
  x1 = _mm_clmulepi64_si128(x1, x0, 0x00);
  x1 = _mm_xor_si128(x1, x2);
  x0 = _mm_load_si128((const __m128i*)poly);
  x2 = _mm_and_si128(x1, x3);
  x2 = _mm_clmulepi64_si128(x2, x0, 0x10);
  x2 = _mm_and_si128(x2, x3);
  x2 = _mm_clmulepi64_si128(x2, x0, 0x00);
  x1 = _mm_xor_si128(x1, x2);
  return (uint32_t)_mm_extract_epi32(x1, 1);
}
#endif

uint32_t crc32_of(const uint8_t* buf, size_t len) {
  uint32_t crc = (uint32_t)crc32(0L, Z_NULL, 0);
#if defined(__x86_64__)
  static const bool have = __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1") &&
                           getenv("KDF_CRC_ZLIB") == nullptr;
  if (have && len >= 64) {
    const size_t body = len & ~(size_t)15;
    crc = ~crc32_clmul(buf, body, ~crc);
    buf += body, len -= body;
  }
#endif
  return len ? (uint32_t)crc32(crc, buf, (uInt)len) : crc;
}

}  // namespace kdf
