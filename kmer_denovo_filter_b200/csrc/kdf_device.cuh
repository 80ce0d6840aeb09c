// kdf_device.cuh — device-side building blocks of the sm_100a k-mer engine.
//
// Everything here is integer work bound by HBM / L2 sector traffic; there is no
// dense contraction anywhere on this path, so no tensor-core (tcgen05) code.
//
// Canonical k-mer semantics follow the reference (kmer_utils.py:30-38
// canonicalize, :91-121 _extract_read_kmers) in the 2-bit encoding A0 C1 G2 T3
// with the first base most significant, where lexicographic min == integer min.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define KDF_HD __host__ __device__ __forceinline__
#define KDF_D __device__ __forceinline__
#else
#define KDF_HD inline
#define KDF_D inline
#endif

namespace kdf {

typedef unsigned long long u64;
typedef unsigned int u32;

static constexpr u64 EMPTY = ~0ull;

// ---------------------------------------------------------------- keys ----
template <int KW> struct Key;
template <> struct Key<1> {
  u64 lo;
  KDF_HD bool operator==(const Key& o) const { return lo == o.lo; }
};
template <> struct Key<2> {
  u64 lo, hi;
  KDF_HD bool operator==(const Key& o) const { return lo == o.lo && hi == o.hi; }
};

// ------------------------------------------------------------- hashing ----
// One 64-bit multiply per key.  The high 32 bits of the product select the
// table partition (top bits) and the bucket; the low 32 bits select the owner
// rank of a multi-GPU run, so owner and bucket are decorrelated.
static constexpr u64 HASH_MUL = 0x9E3779B97F4A7C15ULL;
static constexpr u64 HASH_MUL_HI = 0xD6E8FEB86659FD93ULL;

KDF_HD u64 hash_key(const Key<1>& k) {
  u64 x = k.lo ^ (k.lo >> 32);
  return x * HASH_MUL;
}
KDF_HD u64 hash_key(const Key<2>& k) {
  u64 y = k.lo ^ (k.hi * HASH_MUL_HI);
  y ^= y >> 32;
  return y * HASH_MUL;
}
KDF_HD u32 mulhi32(u32 a, u32 b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (u32)(((u64)a * (u64)b) >> 32);
#endif
}
KDF_HD u64 mulhi64(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
  return __umul64hi(a, b);
#else
  return (u64)(((unsigned __int128)a * (unsigned __int128)b) >> 64);
#endif
}
// partition of a key among 2^log2_parts hash ranges (0 when log2_parts == 0)
KDF_HD u32 part_of(u64 h, int log2_parts) {
  return log2_parts ? (u32)(h >> 32) >> (32 - log2_parts) : 0u;
}
// bucket inside a table (or table slice) that covers one partition: uses the
// hash bits below the partition bits
KDF_HD u32 bucket_of(u64 h, int log2_parts, u32 n_buckets) {
  return mulhi32((u32)(h >> 32) << log2_parts, n_buckets);
}
// owner rank (multi-GPU): low product bits, re-mixed
KDF_HD u32 owner_of(u64 h, u32 n_ranks) {
  u32 l = (u32)h;
  l ^= l >> 15;
  l *= 0x2C1B3C6Du;
  return mulhi32(l, n_ranks);
}
// 64-bit mixer (random-access microbenchmark addresses only)
KDF_HD u64 mix64(u64 x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}

// -------------------------------------------------- reverse complement ----
KDF_HD u64 brev64(u64 x) {
#if defined(__CUDA_ARCH__)
  return __brevll(x);
#else
  x = ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
  x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
  x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
  x = ((x >> 8) & 0x00FF00FF00FF00FFULL) | ((x & 0x00FF00FF00FF00FFULL) << 8);
  x = ((x >> 16) & 0x0000FFFF0000FFFFULL) | ((x & 0x0000FFFF0000FFFFULL) << 16);
  return (x >> 32) | (x << 32);
#endif
}
// reverse the order of the 32 two-bit groups of a word and complement them
KDF_HD u64 revcomp_word(u64 x) {
  u64 r = brev64(~x);
  return ((r >> 1) & 0x5555555555555555ULL) | ((r & 0x5555555555555555ULL) << 1);
}
// reverse complement of a k-mer key, k <= 32
KDF_HD Key<1> revcomp(const Key<1>& f, int k) {
  Key<1> r;
  r.lo = revcomp_word(f.lo) >> (64 - 2 * k);
  return r;
}
// reverse complement of a k-mer key, 33 <= k <= 64
KDF_HD Key<2> revcomp(const Key<2>& f, int k) {
  u64 a = revcomp_word(f.lo);  // becomes the high word of the reversed 128 bits
  u64 b = revcomp_word(f.hi);
  int sh = 128 - 2 * k;  // 0..62
  Key<2> r;
  if (sh == 0) {
    r.hi = a;
    r.lo = b;
  } else {
    r.hi = a >> sh;
    r.lo = (b >> sh) | (a << (64 - sh));
  }
  return r;
}
KDF_HD Key<1> kmin(const Key<1>& a, const Key<1>& b) {
  Key<1> r;
  r.lo = a.lo < b.lo ? a.lo : b.lo;
  return r;
}
KDF_HD Key<2> kmin(const Key<2>& a, const Key<2>& b) {
  bool a_le = (a.hi < b.hi) || (a.hi == b.hi && a.lo <= b.lo);
  return a_le ? a : b;
}

// ------------------------------------------------------- stream access ----
struct StreamView {
  const u64* codes;
  const u32* valid;
  u64 n_bases;
  u64 n_words;
  u64 w_begin, w_end;  // words whose window starts a kernel processes (default: all)
};

KDF_HD u64 ld_code(const StreamView& s, u64 w) {
#if defined(__CUDA_ARCH__)
  return w < s.n_words ? __ldg(s.codes + w) : 0ull;
#else
  return w < s.n_words ? s.codes[w] : 0ull;
#endif
}
KDF_HD u32 ld_valid(const StreamView& s, u64 w) {
#if defined(__CUDA_ARCH__)
  return w < s.n_words ? __ldg(s.valid + w) : 0u;
#else
  return w < s.n_words ? s.valid[w] : 0u;
#endif
}

// Rolling iterator over the 32 window starts of stream word `w`:
// positions 32w .. 32w+31.  `fwd`/`rc` are kept incrementally; the base and
// validity shift registers move 2 / 1 bits per step.
template <int KW> struct WindowIter;

// bit (63 - j) of the result: bases j .. j+k-1 of the 64-base validity register
// are all valid (runs of k ones, by doubling: k = sum of powers of two)
KDF_HD u64 runs_of_k(u64 vv, int k) {
  u64 p = vv, res = ~0ull;
  int off = 0;
  for (int b = 0; (1 << b) <= k; ++b) {
    if (k & (1 << b)) {
      res &= (p << off);
      off += 1 << b;
    }
    p &= (b < 6) ? (p << (1 << b)) : 0ull;
  }
  return res;
}

template <> struct WindowIter<1> {
  u64 b0, b1;   // base shift register (b0 holds the current window at its top)
  u32 okm;      // window validity of the 32 starts of this word (MSB = current start)
  Key<1> rc;
  int k;
  KDF_HD WindowIter(const StreamView& s, u64 w, int k_) : k(k_) {
    b0 = ld_code(s, w);
    b1 = ld_code(s, w + 1);
    u64 vv = ((u64)ld_valid(s, w) << 32) | (u64)ld_valid(s, w + 1);
    okm = (u32)(runs_of_k(vv, k) >> 32);
    Key<1> f = fwd();
    rc = revcomp(f, k);
  }
  KDF_HD bool any_valid() const { return okm != 0; }
  KDF_HD Key<1> fwd() const {
    Key<1> f;
    f.lo = b0 >> (64 - 2 * k);
    return f;
  }
  KDF_HD bool ok() const { return (okm >> 31) != 0; }
  KDF_HD Key<1> canonical() const { return kmin(fwd(), rc); }
  KDF_HD void advance() {
    b0 = (b0 << 2) | (b1 >> 62);
    b1 <<= 2;
    okm <<= 1;
    u64 nb = (b0 >> (64 - 2 * k)) & 3ull;  // newest base of the new window
    rc.lo = (rc.lo >> 2) | ((3ull - nb) << (2 * (k - 1)));
  }
};

template <> struct WindowIter<2> {
  u64 b0, b1, b2;
  u32 okm;     // window validity of the 32 starts of this word (MSB = current start)
  Key<2> rc;
  int k;
  KDF_HD WindowIter(const StreamView& s, u64 w, int k_) : k(k_) {
    b0 = ld_code(s, w);
    b1 = ld_code(s, w + 1);
    b2 = ld_code(s, w + 2);
    // 96 validity bits: v0 = words w, w+1 ; v1 = word w+2 in its top half.  A window
    // of k <= 64 bases starting in word w needs run(start, 32) & run(start+32, k-32)
    // for k > 32: both runs lie inside the 64-bit views below.
    u64 v0 = ((u64)ld_valid(s, w) << 32) | (u64)ld_valid(s, w + 1);
    u64 v1 = ((u64)ld_valid(s, w + 1) << 32) | (u64)ld_valid(s, w + 2);
    u64 r = runs_of_k(v0, k < 32 ? k : 32);
    if (k > 32) r &= runs_of_k(v1, k - 32);
    okm = (u32)(r >> 32);
    rc = revcomp(fwd(), k);
  }
  KDF_HD bool any_valid() const { return okm != 0; }
  KDF_HD Key<2> fwd() const {
    int sh = 128 - 2 * k;  // 0..62
    Key<2> f;
    if (sh == 0) {
      f.hi = b0;
      f.lo = b1;
    } else {
      f.hi = b0 >> sh;
      f.lo = (b1 >> sh) | (b0 << (64 - sh));
    }
    return f;
  }
  KDF_HD bool ok() const { return (okm >> 31) != 0; }
  KDF_HD Key<2> canonical() const { return kmin(fwd(), rc); }
  KDF_HD void advance() {
    b0 = (b0 << 2) | (b1 >> 62);
    b1 = (b1 << 2) | (b2 >> 62);
    b2 <<= 2;
    okm <<= 1;
    // newest base = last base of the new window = bits just above the cut
    int sh = 128 - 2 * k;
    u64 nb = (sh == 0 ? b1 : (b1 >> sh)) & 3ull;
    rc.lo = (rc.lo >> 2) | (rc.hi << 62);
    rc.hi = (rc.hi >> 2) | ((3ull - nb) << (2 * (k - 1) - 64));
  }
};

// Chunked iterator over the 32 window starts of stream word `w` (the form the
// stream kernels use).  WindowIter above rolls every register by one base per
// window (~14 integer instructions of shifting per step); here the 64 (96) bases
// a thread needs are laid out ONCE so that the window at offset u of the current
// chunk is a fixed-distance funnel shift:
//   forward:  (c0:c1[:c2]) = (b0:b1[:b2]) >> (64*KW - 2k); window u = the 64*KW-bit
//             field at bit offset 2u, low 2k bits
//   reverse:  (r0:r1[:r2]) = reverse complement of all the bases; window u =
//             ((r..) >> 2u), low 2k bits
// With u a compile-time constant (unrolled chunk loop) that is two SHF per 64 bits
// and a mask; next<C>() moves the registers on by C bases once per chunk.
template <int KW> struct WindowChunks;

template <> struct WindowChunks<1> {
  u64 c0, c1, r0, r1, kmask;
  u32 okm;  // window validity of the remaining starts (MSB = offset 0 of the chunk)
  KDF_HD WindowChunks(const StreamView& s, u64 w, int k) {
    const u64 b0 = ld_code(s, w), b1 = ld_code(s, w + 1);
    const u64 vv = ((u64)ld_valid(s, w) << 32) | (u64)ld_valid(s, w + 1);
    okm = (u32)(runs_of_k(vv, k) >> 32);
    const int t = 64 - 2 * k;  // 0..62
    kmask = k == 32 ? ~0ull : ((1ull << (2 * k)) - 1);
    if (t == 0) {
      c0 = b0;
      c1 = b1;
    } else {
      c0 = b0 >> t;
      c1 = (b0 << (64 - t)) | (b1 >> t);
    }
    r0 = revcomp_word(b1);
    r1 = revcomp_word(b0);
  }
  KDF_HD bool any_valid() const { return okm != 0; }
  KDF_HD bool ok(int u) const { return ((okm >> (31 - u)) & 1u) != 0; }
  KDF_HD Key<1> key(int u) const {  // canonical k-mer of the window at chunk offset u (0..31)
    Key<1> f, r;
    f.lo = (u == 0 ? c0 : ((c0 << (2 * u)) | (c1 >> (64 - 2 * u)))) & kmask;
    r.lo = (u == 0 ? r1 : ((r1 >> (2 * u)) | (r0 << (64 - 2 * u)))) & kmask;
    return kmin(f, r);
  }
  template <int C> KDF_HD void next() {  // C in 1..31
    c0 = (c0 << (2 * C)) | (c1 >> (64 - 2 * C));
    c1 <<= 2 * C;
    r1 = (r1 >> (2 * C)) | (r0 << (64 - 2 * C));
    r0 >>= 2 * C;
    okm <<= C;
  }
};

template <> struct WindowChunks<2> {
  u64 c0, c1, c2, r0, r1, r2, kmask;  // kmask: key bits of the most significant word
  u32 okm;
  KDF_HD WindowChunks(const StreamView& s, u64 w, int k) {
    const u64 b0 = ld_code(s, w), b1 = ld_code(s, w + 1), b2 = ld_code(s, w + 2);
    const u64 v0 = ((u64)ld_valid(s, w) << 32) | (u64)ld_valid(s, w + 1);
    const u64 v1 = ((u64)ld_valid(s, w + 1) << 32) | (u64)ld_valid(s, w + 2);
    u64 r = runs_of_k(v0, k < 32 ? k : 32);
    if (k > 32) r &= runs_of_k(v1, k - 32);
    okm = (u32)(r >> 32);
    const int t = 128 - 2 * k;  // 0..62
    kmask = k == 64 ? ~0ull : ((1ull << (2 * k - 64)) - 1);
    if (t == 0) {
      c0 = b0;
      c1 = b1;
      c2 = b2;
    } else {
      c0 = b0 >> t;
      c1 = (b0 << (64 - t)) | (b1 >> t);
      c2 = (b1 << (64 - t)) | (b2 >> t);
    }
    r0 = revcomp_word(b2);
    r1 = revcomp_word(b1);
    r2 = revcomp_word(b0);
  }
  KDF_HD bool any_valid() const { return okm != 0; }
  KDF_HD bool ok(int u) const { return ((okm >> (31 - u)) & 1u) != 0; }
  KDF_HD Key<2> key(int u) const {
    Key<2> f, r;
    if (u == 0) {
      f.hi = c0 & kmask;
      f.lo = c1;
      r.hi = r1 & kmask;
      r.lo = r2;
    } else {
      f.hi = ((c0 << (2 * u)) | (c1 >> (64 - 2 * u))) & kmask;
      f.lo = (c1 << (2 * u)) | (c2 >> (64 - 2 * u));
      r.hi = ((r1 >> (2 * u)) | (r0 << (64 - 2 * u))) & kmask;
      r.lo = (r2 >> (2 * u)) | (r1 << (64 - 2 * u));
    }
    return kmin(f, r);
  }
  template <int C> KDF_HD void next() {
    c0 = (c0 << (2 * C)) | (c1 >> (64 - 2 * C));
    c1 = (c1 << (2 * C)) | (c2 >> (64 - 2 * C));
    c2 <<= 2 * C;
    r2 = (r2 >> (2 * C)) | (r1 << (64 - 2 * C));
    r1 = (r1 >> (2 * C)) | (r0 << (64 - 2 * C));
    r0 >>= 2 * C;
    okm <<= C;
  }
};

// Random-access extraction of the window starting at stream position p
// (used by the per-read scan, where lanes stride through one read).
template <int KW> struct WindowAt;

template <> struct WindowAt<1> {
  KDF_HD static bool get(const StreamView& s, u64 p, int k, Key<1>& out) {
    u64 w = p >> 5;
    int off = (int)(p & 31);
    u64 c0 = ld_code(s, w), c1 = ld_code(s, w + 1);
    u64 vv = ((u64)ld_valid(s, w) << 32) | (u64)ld_valid(s, w + 1);
    u64 x = off ? ((c0 << (2 * off)) | (c1 >> (64 - 2 * off))) : c0;
    vv <<= off;
    Key<1> f;
    f.lo = x >> (64 - 2 * k);
    out = kmin(f, revcomp(f, k));
    return ((~vv) >> (64 - k)) == 0;
  }
};
template <> struct WindowAt<2> {
  KDF_HD static bool get(const StreamView& s, u64 p, int k, Key<2>& out) {
    u64 w = p >> 5;
    int off = (int)(p & 31);
    u64 c0 = ld_code(s, w), c1 = ld_code(s, w + 1), c2 = ld_code(s, w + 2);
    u64 v0 = ((u64)ld_valid(s, w) << 32) | (u64)ld_valid(s, w + 1);
    u64 v1 = ((u64)ld_valid(s, w + 2) << 32);
    u64 x0 = c0, x1 = c1;
    if (off) {
      x0 = (c0 << (2 * off)) | (c1 >> (64 - 2 * off));
      x1 = (c1 << (2 * off)) | (c2 >> (64 - 2 * off));
      v0 = (v0 << off) | (v1 >> (64 - off));
    }
    int sh = 128 - 2 * k;
    Key<2> f;
    if (sh == 0) {
      f.hi = x0;
      f.lo = x1;
    } else {
      f.hi = x0 >> sh;
      f.lo = (x1 >> sh) | (x0 << (64 - sh));
    }
    out = kmin(f, revcomp(f, k));
    return k == 64 ? (~v0 == 0) : (((~v0) >> (64 - k)) == 0);
  }
};

// ---------------------------------------------------- K7: hit coverage ----
// Reference positions under one hit window (core/bam_scanner.py:97-117
// _collect_kmer_ref_positions): query positions off .. off+k-1 that a CIGAR M / = / X
// operation aligns.  One key per covered base, (contig << 40) | (ref_pos << 1) | first,
// where `first` says that no earlier hit of the same read covers the base (hits of a
// read come sorted by offset, so that is "query position >= prev_end"): counting all
// keys of a position gives the k-mer coverage, counting the `first` ones the number of
// reads.  cigar words are BAM's (len << 4 | op).  Returns the number of keys written.
KDF_HD int expand_hit(const u32* cigar, u64 n_ops, long long ref_start, u32 off, int k, u32 prev_end,
                      u64 contig, u64* out) {
  int n = 0;
  u64 q = 0;
  long long r = ref_start;
  const u64 lo = off, hi = (u64)off + (u64)k;
  for (u64 i = 0; i < n_ops && q < hi; ++i) {
    const u32 op = cigar[i] & 15u;
    const u64 len = cigar[i] >> 4;
    if (op == 0 || op == 7 || op == 8) {
      const u64 a = q > lo ? q : lo;
      const u64 b = (q + len) < hi ? (q + len) : hi;
      for (u64 qq = a; qq < b; ++qq) {
        const long long rp = r + (long long)(qq - q);
        out[n++] = (contig << 40) | (((u64)rp & 0x7fffffffffull) << 1) | (qq >= prev_end ? 1ull : 0ull);
      }
      q += len;
      r += (long long)len;
    } else if (op == 1 || op == 4) {
      q += len;
    } else if (op == 2 || op == 3) {
      r += (long long)len;
    }
  }
  return n;
}

}  // namespace kdf
