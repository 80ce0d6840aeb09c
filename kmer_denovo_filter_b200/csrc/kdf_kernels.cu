// kdf_kernels.cu — sm_100a kernels + C ABI (include/kdf.h) of the k-mer engine.
//
// Kernel inventory (SURVEY §2 native table):
//   K1  k_extract           rolling extract + canonicalise
//   K2  k_stream            K1 fused with a bucketed open-addressing table op:
//                           insert+count / insert / count-if-present /
//                           mark-if-present / emit-hits; table in L2/HBM or,
//                           for small read-only sets, copied to shared memory
//       k_update_keys       the same table ops on explicit key arrays
//       k_stream<FILT>      the probing ops behind a two-bit filter in L2
//                           (kdf_table_build_filter): one 32-bit load per window,
//                           the table is probed only from the queue drain
//   K2p k_bin_stream/_keys  hash-range binning of canonical k-mers (shared-memory
//                           staged), feeding the per-bin count of kdf_count_bins
//   K2c k_packed_keys       the per-bin count in packed form (saturating counter in
//                           the key's spare bits, keys-only L2 slice; queued
//                           compare-and-swaps), and the probing ops of
//                           kdf_update_bins; k_emit_packed = emit + clear
//   K3  k_threshold_compact threshold + stream compaction (dump -L, == 0, <= pmc)
//   K4  k_lookup_keys       batched membership / count lookup
//   K5  k_scan_reads        dense per-read scan + distinct reduction
//       k_reduce_hits       sparse per-read reduction of an emitted hit list
//   K6  k_bin_stream<OWNER> owner (x hash range) binning; with peer-mapped
//                           destinations the flush IS the multi-GPU all-to-all
//   K7  k_hit_coverage      reference positions under the hit windows (+ cub sort/RLE)
//       k_valid_fill/_clear validity bitmap from its sparse form (PCIe format)
//
// Nothing here is a dense contraction, so there is no tensor-core code: the
// work is bounded by L2 / HBM sector traffic and instruction issue.
//
// Mapping: stream kernels give each thread one 32-base word = 32 consecutive
// window starts, so code/valid loads are 8-byte coalesced and the canonical
// k-mer is maintained by a 2-bit rolling update.  A table probe reads one
// 32-byte bucket (4 x 64-bit keys or 2 x 128-bit keys) with a single 256-bit
// load; CHUNK probes are issued back to back before the first is resolved.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <math.h>

#include <string>
#include <type_traits>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_run_length_encode.cuh>

#include "../../include/kdf.h"
#include "kdf_device.cuh"

using namespace kdf;

// ------------------------------------------------------------ errors ------
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CUDA_TRY(expr)                                                        \
  do {                                                                        \
    cudaError_t _e = (expr);                                                  \
    if (_e != cudaSuccess)                                                    \
      return fail(KDF_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

struct kdf_table {
  int k;
  int key_words;
  u64 capacity;
  void* base;
  int sm_count;
  int log2_parts;  // hash bits consumed above the bucket bits (table slices)
  u32* filter = nullptr;   // optional two-bit membership filter over the keys (kdf_table_build_filter)
  u32 filter_mask = 0;     // number of 32-bit filter words - 1 (a power of two)
};

// SoA table: keys[capacity] (KW words each, 32-byte buckets), p0[capacity], p1[capacity]
template <int KW> struct TableView {
  u64* keys;
  u32* p0;
  u32* p1;
  u32 n_buckets;
  int log2_parts;
  bool fast_empty;  // k % 32 != 0 (see has_empty)
  const u32* filter;  // two-bit membership filter or NULL
  u32 filter_mask;
};
template <int KW> struct SPB { static constexpr int v = 4; };  // slots per bucket

template <int KW>
static TableView<KW> view_of_table(const kdf_table* t) {
  TableView<KW> v;
  v.keys = (u64*)t->base;
  v.p0 = (u32*)((char*)t->base + t->capacity * 8ull * KW);
  v.p1 = v.p0 + t->capacity;
  v.n_buckets = (u32)(t->capacity / SPB<KW>::v);
  v.log2_parts = t->log2_parts;
  v.fast_empty = (t->k % 32) != 0;
  v.filter = t->filter;
  v.filter_mask = t->filter_mask;
  return v;
}

// ----------------------------------------------------- bucket primitives --
// A bucket is 4 slots for either key width: 32 bytes (one sector, one 256-bit
// load) of 64-bit keys, 64 bytes (two sectors of one line, two 256-bit loads) of
// 128-bit keys.  Four slots keep "home bucket full" rare at load 0.3-0.5 (5-14 %);
// with two 128-bit slots per bucket it was 14-26 % and every such probe pays a
// second, dependent round trip.
template <int KW> struct Bucket {
  u64 q[4 * KW];
};
__device__ __forceinline__ void ld256(const u64* p, u64& a, u64& b, u64& c, u64& d) {
  asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];"
               : "=l"(a), "=l"(b), "=l"(c), "=l"(d)
               : "l"(p)
               : "memory");
}
template <int KW> __device__ __forceinline__ Bucket<KW> ld_bucket(const u64* p) {
  Bucket<KW> b;
  ld256(p, b.q[0], b.q[1], b.q[2], b.q[3]);
  if (KW == 2) ld256(p + 4, b.q[4 * (KW - 1)], b.q[4 * (KW - 1) + 1], b.q[4 * (KW - 1) + 2], b.q[4 * (KW - 1) + 3]);
  return b;
}
template <int KW> __device__ __forceinline__ Bucket<KW> lds_bucket(const u64* p) {
  Bucket<KW> b;
#pragma unroll
  for (int i = 0; i < 2 * KW; ++i) {
    ulonglong2 a = *reinterpret_cast<const ulonglong2*>(p + 2 * i);
    b.q[2 * i] = a.x;
    b.q[2 * i + 1] = a.y;
  }
  return b;
}
// index of the slot holding `key` in the bucket, or -1
__device__ __forceinline__ int match_in(const Bucket<1>& b, const Key<1>& key) {
  int j = -1;
  if (b.q[3] == key.lo) j = 3;
  if (b.q[2] == key.lo) j = 2;
  if (b.q[1] == key.lo) j = 1;
  if (b.q[0] == key.lo) j = 0;
  return j;
}
__device__ __forceinline__ int match_in(const Bucket<2>& b, const Key<2>& key) {
  int j = -1;
  if (b.q[6] == key.lo && b.q[7] == key.hi) j = 3;
  if (b.q[4] == key.lo && b.q[5] == key.hi) j = 2;
  if (b.q[2] == key.lo && b.q[3] == key.hi) j = 1;
  if (b.q[0] == key.lo && b.q[1] == key.hi) j = 0;
  return j;
}
// `fast`: k is not a multiple of 32, so the top 32 bits of a stored key's most
// significant word are never all ones and testing them alone identifies an empty slot
__device__ __forceinline__ bool has_empty(const Bucket<1>& b, Key<1>, bool fast) {
  if (fast) {
    // a slot is empty iff its high half is all ones: max over the four high halves
    u32 h0 = (u32)(b.q[0] >> 32), h1 = (u32)(b.q[1] >> 32), h2 = (u32)(b.q[2] >> 32), h3 = (u32)(b.q[3] >> 32);
    u32 mx = max(max(h0, h1), max(h2, h3));
    return mx == 0xffffffffu;
  }
  return b.q[0] == EMPTY || b.q[1] == EMPTY || b.q[2] == EMPTY || b.q[3] == EMPTY;
}
__device__ __forceinline__ bool has_empty(const Bucket<2>& b, Key<2>, bool fast) {
  if (fast) {
    u32 h0 = (u32)(b.q[1] >> 32), h1 = (u32)(b.q[3] >> 32), h2 = (u32)(b.q[5] >> 32), h3 = (u32)(b.q[7] >> 32);
    return max(max(h0, h1), max(h2, h3)) == 0xffffffffu;
  }
  return (b.q[0] == EMPTY && b.q[1] == EMPTY) || (b.q[2] == EMPTY && b.q[3] == EMPTY) ||
         (b.q[4] == EMPTY && b.q[5] == EMPTY) || (b.q[6] == EMPTY && b.q[7] == EMPTY);
}
// slot j of the bucket may be (a possibly torn view of) an empty slot
__device__ __forceinline__ bool maybe_empty(const Bucket<1>& b, int j, Key<1>) { return b.q[j] == EMPTY; }
__device__ __forceinline__ bool maybe_empty(const Bucket<2>& b, int j, Key<2>) {
  return b.q[2 * j] == EMPTY || b.q[2 * j + 1] == EMPTY;
}

__device__ __forceinline__ Key<1> cas_key(u64* p, const Key<1>& val) {
  Key<1> old;
  old.lo = atomicCAS(p, EMPTY, val.lo);
  return old;
}
__device__ __forceinline__ Key<2> cas_key(u64* p, const Key<2>& val) {
  Key<2> old;
  asm volatile(
      "{\n\t"
      ".reg .b128 c, v, o;\n\t"
      "mov.b128 c, {%2, %3};\n\t"
      "mov.b128 v, {%4, %5};\n\t"
      "atom.relaxed.gpu.global.cas.b128 o, [%6], c, v;\n\t"
      "mov.b128 {%0, %1}, o;\n\t"
      "}\n"
      : "=l"(old.lo), "=l"(old.hi)
      : "l"(EMPTY), "l"(EMPTY), "l"(val.lo), "l"(val.hi), "l"(p)
      : "memory");
  return old;
}
// compare-and-swap of a whole slot with an explicit expected value
__device__ __forceinline__ Key<1> cas_slot(u64* p, const Key<1>& expect, const Key<1>& val) {
  Key<1> old;
  old.lo = atomicCAS(p, expect.lo, val.lo);
  return old;
}
__device__ __forceinline__ Key<2> cas_slot(u64* p, const Key<2>& expect, const Key<2>& val) {
  Key<2> old;
  asm volatile(
      "{\n\t"
      ".reg .b128 c, v, o;\n\t"
      "mov.b128 c, {%2, %3};\n\t"
      "mov.b128 v, {%4, %5};\n\t"
      "atom.relaxed.gpu.global.cas.b128 o, [%6], c, v;\n\t"
      "mov.b128 {%0, %1}, o;\n\t"
      "}\n"
      : "=l"(old.lo), "=l"(old.hi)
      : "l"(expect.lo), "l"(expect.hi), "l"(val.lo), "l"(val.hi), "l"(p)
      : "memory");
  return old;
}
__device__ __forceinline__ bool is_empty_key(const Key<1>& k) { return k.lo == EMPTY; }
__device__ __forceinline__ bool is_empty_key(const Key<2>& k) { return k.lo == EMPTY && k.hi == EMPTY; }

// table operations (also the `mode` values of the C ABI, plus the hit emitter)
constexpr int OP_INSERT_COUNT = KDF_MODE_INSERT_COUNT;
constexpr int OP_INSERT_ONLY = KDF_MODE_INSERT_ONLY;
constexpr int OP_COUNT_IF_PRESENT = KDF_MODE_COUNT_IF_PRESENT;
constexpr int OP_MARK_IF_PRESENT = KDF_MODE_MARK_IF_PRESENT;
constexpr int OP_EMIT_HITS = 4;
constexpr int OP_PACKED_COUNT = 5;  // packed counting (K2c below): plane = state shift, arg = sat
constexpr int OP_PACKED_MARK = 6;

struct HitSink {
  u64* pos;
  u32* slot;
  u64 cap;
  u64* n;
};

struct LocalStats {
  u32 windows, hits, fresh, full;
};

// result codes of a resolved probe
constexpr u32 R_MISS = 0, R_HIT = 1, R_NEW = 2, R_FULL = 3;

template <int OP, int KW>
__device__ __forceinline__ void on_found(const TableView<KW>& t, u64 slot, int plane, u32 arg,
                                         u64 pos, const HitSink& sink) {
  if (OP == OP_INSERT_COUNT || OP == OP_COUNT_IF_PRESENT) {
    atomicAdd((plane ? t.p1 : t.p0) + slot, arg);  // result unused -> RED
  } else if (OP == OP_MARK_IF_PRESENT) {
    atomicOr((plane ? t.p1 : t.p0) + slot, arg);
  } else if (OP == OP_EMIT_HITS) {
    u64 o = atomicAdd(sink.n, 1ull);
    if (o < sink.cap) {
      sink.pos[o] = pos;
      sink.slot[o] = (u32)slot;
    }
  }
}

template <int KW>
__device__ __forceinline__ u32 resolve_packed_count(const TableView<KW>& t, u32 b, const Key<KW>& key,
                                                    int sh, u32 sat);
template <int KW>
__device__ __forceinline__ u32 resolve_packed_mark(const TableView<KW>& t, u32 b, const Key<KW>& key,
                                                   int sh, u32 sat);

// Finish a probe starting at bucket `b` (the rare path: a new key, a full
// bucket, or a probing op that has to look past a full bucket).
template <int KW, int OP>
__device__ __forceinline__ u32 resolve_from(const TableView<KW>& t, u32 b, const Key<KW>& key,
                                            int plane, u32 arg, u64 pos, const HitSink& sink) {
  if constexpr (OP == OP_PACKED_COUNT) return resolve_packed_count<KW>(t, b, key, plane, arg);
  if constexpr (OP == OP_PACKED_MARK) return resolve_packed_mark<KW>(t, b, key, plane, arg);
  constexpr bool kInsert = (OP == OP_INSERT_COUNT || OP == OP_INSERT_ONLY);
  constexpr int S = SPB<KW>::v;
  for (u32 n = 0; n < t.n_buckets; ++n) {
    Bucket<KW> bk = ld_bucket<KW>(t.keys + (u64)b * 4 * KW);
    int j = match_in(bk, key);
    if (j >= 0) {
      on_found<OP, KW>(t, (u64)b * S + j, plane, arg, pos, sink);
      return R_HIT;
    }
    if (kInsert) {
#pragma unroll
      for (int c = 0; c < S; ++c) {
        if (maybe_empty(bk, c, key)) {
          u64 slot = (u64)b * S + c;
          Key<KW> old = cas_key(t.keys + slot * KW, key);
          if (is_empty_key(old)) {
            on_found<OP, KW>(t, slot, plane, arg, pos, sink);
            return R_NEW;
          }
          if (old == key) {
            on_found<OP, KW>(t, slot, plane, arg, pos, sink);
            return R_HIT;
          }
        }
      }
    } else {
      if (has_empty(bk, key, t.fast_empty)) return R_MISS;
    }
    b = (b + 1 == t.n_buckets) ? 0 : b + 1;
  }
  return R_FULL;
}

// Per-warp queue of probes that left the fast path.  Lanes push their rare
// items as they meet them; the warp drains the queue 32 items at a time, so
// the slow path runs with every lane busy instead of a few lanes per call.
#ifndef KDF_STREAM_CHUNK
#define KDF_STREAM_CHUNK 4
#endif
#ifndef KDF_STREAM_BLOCKS
#define KDF_STREAM_BLOCKS 3
#endif
#ifndef KDF_SMEM_CHUNK
#define KDF_SMEM_CHUNK 4
#endif
#ifndef KDF_FILT_CHUNK
#define KDF_FILT_CHUNK 8   // filter loads in flight per thread (k_stream<FILT>)
#endif
#ifndef KDF_SMEM_BLOCKS
#define KDF_SMEM_BLOCKS 1
#endif
#ifndef KDF_KEYS_CHUNK
#define KDF_KEYS_CHUNK 4
#endif
#ifndef KDF_KEYS_BLOCKS
#define KDF_KEYS_BLOCKS 2
#endif
constexpr int SQ_CAP = 96;
template <int KW> struct SlowQueue {
  u64 lo[SQ_CAP];
  u64 hi[KW == 2 ? SQ_CAP : 1];
  u64 pos[SQ_CAP];
  u32 b[SQ_CAP];
  u32 count;
};

template <int KW, int OP>
__device__ __forceinline__ u32 sq_push_or_resolve(SlowQueue<KW>& q, const TableView<KW>& t, u32 b,
                                                  const Key<KW>& key, int plane, u32 arg, u64 pos,
                                                  const HitSink& sink) {
  u32 o = atomicAdd(&q.count, 1u);
  if (o < (u32)SQ_CAP) {
    q.lo[o] = key.lo;
    if (KW == 2) q.hi[o] = ((const u64*)&key)[KW - 1];
    q.pos[o] = pos;
    q.b[o] = b;
    return R_MISS;  // accounted for when drained
  }
  return resolve_from<KW, OP>(t, b, key, plane, arg, pos, sink);  // queue full: do it now
}

// all lanes of the warp: drain whole groups of 32 (or everything when `all`)
template <int KW, int OP>
__device__ __noinline__ u32 sq_drain(SlowQueue<KW>* qp, TableView<KW> t, int plane, u32 arg,
                                     HitSink sink, bool all) {
  SlowQueue<KW>& q = *qp;
  const unsigned lane = threadIdx.x & 31;
  __syncwarp();
  u32 n = q.count;
  if (n > (u32)SQ_CAP) n = SQ_CAP;
  u32 packed = 0;  // hits | fresh << 10 | full << 20 of this lane
  while (n >= 32 || (all && n > 0)) {
    u32 take = n >= 32 ? 32 : n;
    u32 base = n - take;
    if (lane < take) {
      Key<KW> key;
      key.lo = q.lo[base + lane];
      if (KW == 2) ((u64*)&key)[KW - 1] = q.hi[base + lane];
      u32 code = resolve_from<KW, OP>(t, q.b[base + lane], key, plane, arg, q.pos[base + lane], sink);
      packed += (code == R_HIT ? 1u : 0u) + (code == R_NEW ? (1u << 10) : 0u);
      packed |= (code == R_FULL ? (1u << 20) : 0u);
    }
    n = base;
    __syncwarp();
  }
  if (lane == 0) q.count = n;
  __syncwarp();
  return packed;
}

__device__ __forceinline__ void tally_packed(LocalStats& st, u32 packed) {
  st.hits += packed & 1023u;
  st.fresh += (packed >> 10) & 1023u;
  st.full |= packed >> 20;
}

__device__ __forceinline__ void tally(LocalStats& st, u32 code) {
  st.hits += (code == R_HIT) ? 1u : 0u;
  st.fresh += (code == R_NEW) ? 1u : 0u;
  st.full |= (code == R_FULL) ? 1u : 0u;
}

__device__ __forceinline__ void flush_stats(const LocalStats& st, u64* stats) {
  if (!stats) return;
  u32 w = st.windows, h = st.hits, f = st.fresh, fl = st.full;
  for (int o = 16; o; o >>= 1) {
    w += __shfl_xor_sync(0xffffffffu, w, o);
    h += __shfl_xor_sync(0xffffffffu, h, o);
    f += __shfl_xor_sync(0xffffffffu, f, o);
    fl |= __shfl_xor_sync(0xffffffffu, fl, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (w) atomicAdd(stats + KDF_STAT_WINDOWS, (u64)w);
    if (h) atomicAdd(stats + KDF_STAT_HITS, (u64)h);
    if (f) atomicAdd(stats + KDF_STAT_NEW, (u64)f);
    if (fl) atomicOr(stats + KDF_STAT_FULL, 1ull);
  }
}

// ------------------------------------------------------------- K1 ---------
template <int KW>
__global__ void __launch_bounds__(256) k_extract(StreamView s, int k, u64* out_lo, u64* out_hi,
                                                 u32* out_ok) {
  u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x; w < s.n_words; w += stride) {
    WindowIter<KW> it(s, w, k);
    u32 okbits = 0;
    u64 base = w << 5;
#pragma unroll 4
    for (int j = 0; j < 32; ++j) {
      u64 p = base + j;
      bool ok = it.ok();
      Key<KW> c = it.canonical();
      if (p < s.n_bases) {
        out_lo[p] = ok ? c.lo : 0ull;
        if (KW == 2) out_hi[p] = ok ? ((const u64*)&c)[KW - 1] : 0ull;
      }
      okbits |= (ok ? 1u : 0u) << (31 - j);
      it.advance();
    }
    out_ok[w] = okbits;
  }
}

// two filter bits of a key inside one 32-bit word (low product bits: independent of
// the bucket index, which comes from the high ones)
constexpr int PF_WORDS = 4096;
__device__ __forceinline__ void pf_bits(u64 h, u32& word, u32& bits) {
  u32 l = (u32)h;
  l ^= l >> 15;
  l *= 0x2C1B3C6Du;
  word = l & (u32)(PF_WORDS - 1);
  bits = (1u << (l >> 27)) | (1u << ((l >> 22) & 31u));
}

// 32-bit read-only load executed only when `on` (else 0): a predicated instruction, not a branch
__device__ __forceinline__ u32 ldg_if(const u32* p, bool on) {
  u32 v;
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "setp.ne.u32 q, %2, 0;\n\t"
      "mov.u32 %0, 0;\n\t"
      "@q ld.global.nc.u32 %0, [%1];\n\t"
      "}\n"
      : "=r"(v)
      : "l"(p), "r"((u32)on));
  return v;
}

// the same for a filter in global memory (kdf_table_build_filter): any power-of-two
// number of words
__device__ __forceinline__ void gf_bits(u64 h, u32 mask, u32& word, u32& bits) {
  u32 l = (u32)h;
  l ^= l >> 15;
  l *= 0x2C1B3C6Du;
  word = (l ^ (u32)(h >> 32)) & mask;
  bits = (1u << (l >> 27)) | (1u << ((l >> 22) & 31u));
}

template <int KW>
__global__ void __launch_bounds__(256) k_build_filter(TableView<KW> t, u64 capacity, u32* words, u32 mask) {
  u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < capacity; i += stride) {
    Key<KW> key;
    key.lo = __ldcg(t.keys + i * KW);
    if (KW == 2) ((u64*)&key)[KW - 1] = __ldcg(t.keys + i * KW + (KW - 1));
    if (!is_empty_key(key)) {
      u32 word, bits;
      gf_bits(hash_key(key), mask, word, bits);
      atomicOr(words + word, bits);
    }
  }
}

// ------------------------------------------------------------- K2 ---------
// One fast-path decision per probe: for probing ops "bucket has an empty slot
// and no match" (a miss, the common case against a filter set), for inserting
// ops "bucket holds the key" (the common case at sequencing depth).  Anything
// else is queued per warp and resolved 32 items at a time.
template <int KW, int OP, bool SMEM, int CHUNK, bool FILT = false>
__global__ void __launch_bounds__(SMEM ? 512 : 256, SMEM ? KDF_SMEM_BLOCKS : (KW == 2 ? 2 : KDF_STREAM_BLOCKS))
    k_stream(TableView<KW> t, StreamView s, int k, int plane, u32 arg, u64* stats, HitSink sink) {
  extern __shared__ __align__(32) u64 sm_keys[];
  constexpr bool kInsert = (OP == OP_INSERT_COUNT || OP == OP_INSERT_ONLY);
  constexpr int S = SPB<KW>::v;
  static_assert(!(SMEM && kInsert), "shared-memory tables are read-only");
  static_assert(!(FILT && (SMEM || kInsert)), "a filter fronts read-only tables in global memory");
  // Shared-memory tables are small read-only sets (the proband-unique k-mers, VCF-mode
  // filter sets) that almost no window hits, and a bucket probe from shared memory costs
  // two 128-bit loads with bank conflicts plus the compares: the ncu profile of the scan
  // showed l1tex at 78 %.  A 16 KB two-bit filter built next to the table answers
  // "certainly absent" for > 99 % of the windows with one 32-bit load.
  u32* pf = reinterpret_cast<u32*>(sm_keys + (SMEM ? (size_t)t.n_buckets * 4 * KW : 0));
  if (SMEM) {
    u32 n = t.n_buckets * 4 * KW;
    for (u32 i = threadIdx.x; i < n; i += blockDim.x) sm_keys[i] = __ldg(t.keys + i);
    for (u32 i = threadIdx.x; i < (u32)PF_WORDS; i += blockDim.x) pf[i] = 0;
    __syncthreads();
    for (u32 i = threadIdx.x; i < t.n_buckets * S; i += blockDim.x) {
      Key<KW> key;
      key.lo = sm_keys[(size_t)i * KW];
      if (KW == 2) ((u64*)&key)[KW - 1] = sm_keys[(size_t)i * KW + (KW - 1)];
      if (!is_empty_key(key)) {
        u32 word, bits;
        pf_bits(hash_key(key), word, bits);
        atomicOr(pf + word, bits);
      }
    }
    __syncthreads();
  }
  __shared__ SlowQueue<KW> sq[(SMEM ? 512 : 256) / 32];
  SlowQueue<KW>& q = sq[threadIdx.x >> 5];
  if ((threadIdx.x & 31) == 0) q.count = 0;
  __syncwarp();
  LocalStats st = {0, 0, 0, 0};
  // Warp-uniform trip counts (words past the end read as invalid) so that the
  // warp can be re-converged explicitly after every divergent section: without
  // the __syncwarp() calls lanes drift apart for the rest of the kernel.
  u64 stride = (u64)gridDim.x * blockDim.x;
  u64 n_iter = (s.n_words + stride - 1) / stride;
  u64 w0 = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  for (u64 itn = 0; itn < n_iter; ++itn) {
    u64 w = w0 + itn * stride;
    WindowChunks<KW> it(s, w, k);
    if (!__any_sync(0xffffffffu, it.any_valid())) continue;
#pragma unroll 1
    for (int c = 0; c < 32 / CHUNK; ++c) {
      Key<KW> keys[CHUNK];
      Bucket<KW> bk[CHUNK];
      u32 bidx[CHUNK];
      u32 okm = 0, probe = 0, cand = 0;
      if constexpr (FILT || SMEM) {
        // filtered tables: one 32-bit load says "certainly absent" for almost every
        // window.  The CHUNK loads are issued back to back and tested afterwards.  Written
        // as `if (ok) { load; test }` ptxas built one divergent region per window with the
        // load's consumer right behind it — one load in flight per thread, long_scoreboard
        // the top stall (profiles/r2a_ncu_k_stream_filtered.txt).  The global loads are
        // PREDICATED on the window's validity (ldg_if: one PTX instruction, no branch): the
        // kernel is bound by the L1/L2 request rate (l1tex 94 %), and loading for the 20 %
        // of the positions that start no valid window cost exactly those 20 % (6.7 -> 7.3 ms,
        // profiles/r2b_ncu_k_stream_filtered.txt).  Shared-memory filter reads are unconditional.
        u32 fval[CHUNK], fbits[CHUNK];
#pragma unroll
        for (int u = 0; u < CHUNK; ++u) {
          keys[u] = it.key(u);
          const u64 h = hash_key(keys[u]);
          u32 word;
          if (FILT) {
            gf_bits(h, t.filter_mask, word, fbits[u]);
            fval[u] = ldg_if(t.filter + word, it.ok(u));
          } else {
            pf_bits(h, word, fbits[u]);
            fval[u] = pf[word];
          }
          okm |= it.ok(u) ? (1u << u) : 0u;
        }
#pragma unroll
        for (int u = 0; u < CHUNK; ++u)
          if ((fval[u] & fbits[u]) == fbits[u]) cand |= 1u << u;
        cand &= okm;
        if (SMEM && cand) {   // the few candidates probe the shared-memory table
#pragma unroll
          for (int u = 0; u < CHUNK; ++u) {
            if (cand & (1u << u)) {
              bidx[u] = bucket_of(hash_key(keys[u]), t.log2_parts, t.n_buckets);
              bk[u] = lds_bucket<KW>(sm_keys + (u64)bidx[u] * 4 * KW);
            }
          }
          probe = cand;
        }
      } else {
#pragma unroll
        for (int u = 0; u < CHUNK; ++u) {
          bool ok = it.ok(u);
          keys[u] = it.key(u);
          const u64 h = hash_key(keys[u]);
          bidx[u] = bucket_of(h, t.log2_parts, t.n_buckets);
          if (ok) {
            okm |= 1u << u;
            probe |= 1u << u;
            bk[u] = ld_bucket<KW>(t.keys + (u64)bidx[u] * 4 * KW);
          }
        }
      }
      it.template next<CHUNK>();
      st.windows += __popc(okm);
      u64 pos0 = (w << 5) + c * CHUNK;
      if (FILT) {
#pragma unroll
        for (int u = 0; u < CHUNK; ++u)
          if (cand & (1u << u))
            tally(st, sq_push_or_resolve<KW, OP>(q, t, bucket_of(hash_key(keys[u]), t.log2_parts, t.n_buckets),
                                                 keys[u], plane, arg, pos0 + u, sink));
      }
#pragma unroll
      for (int u = 0; u < CHUNK; ++u) {
        if (probe & (1u << u)) {
          int j = match_in(bk[u], keys[u]);
          if (j >= 0) {  // found in the home bucket
            st.hits++;
            on_found<OP, KW>(t, (u64)bidx[u] * S + j, plane, arg, pos0 + u, sink);
          } else if (kInsert) {  // new key (or full bucket): CAS path, queued
            tally(st, sq_push_or_resolve<KW, OP>(q, t, bidx[u], keys[u], plane, arg, pos0 + u, sink));
          } else if (!has_empty(bk[u], keys[u], t.fast_empty)) {  // full bucket: look further, queued
            u32 nb = (bidx[u] + 1 == t.n_buckets) ? 0 : bidx[u] + 1;
            tally(st, sq_push_or_resolve<KW, OP>(q, t, nb, keys[u], plane, arg, pos0 + u, sink));
          }
        }
      }
      __syncwarp();
      if (q.count >= 32) tally_packed(st, sq_drain<KW, OP>(&q, t, plane, arg, sink, false));
    }
  }
  tally_packed(st, sq_drain<KW, OP>(&q, t, plane, arg, sink, true));
  flush_stats(st, stats);
}

// the same table ops on an explicit key array; n may live on the device
// (n_dev != NULL: n = min(*n_dev, n_max)) so that bins filled by k_bin_* can be
// consumed without a host round trip.  Keys are read once: evict-first loads.
template <int KW> __device__ __forceinline__ Key<KW> ld_key_stream(const u64* lo, const u64* hi, u64 i);
template <> __device__ __forceinline__ Key<1> ld_key_stream<1>(const u64* lo, const u64*, u64 i) {
  Key<1> k;
  k.lo = __ldcs(lo + i);
  return k;
}
template <> __device__ __forceinline__ Key<2> ld_key_stream<2>(const u64* lo, const u64* hi, u64 i) {
  Key<2> k;
  if (hi) {
    k.lo = __ldcs(lo + i);
    k.hi = __ldcs(hi + i);
  } else {  // interleaved {lo, hi} pairs (bins)
    ulonglong2 v = __ldcs(reinterpret_cast<const ulonglong2*>(lo) + i);
    k.lo = v.x;
    k.hi = v.y;
  }
  return k;
}

template <int KW, int OP, bool FILT>
__global__ void __launch_bounds__(256, KDF_KEYS_BLOCKS) k_update_keys(TableView<KW> t, const u64* lo, const u64* hi,
                                                     u64 n_max, const u64* n_dev, int plane, u32 arg,
                                                     u64* stats, int filt_log2, u32 filt_val) {
  // filt_log2 > 0: only keys of hash range filt_val (of 2^filt_log2) are applied — a
  // bin that spans several table slices is streamed once per slice
  constexpr int CHUNK = KDF_KEYS_CHUNK;
  constexpr bool kInsert = (OP == OP_INSERT_COUNT || OP == OP_INSERT_ONLY);
  constexpr int S = SPB<KW>::v;
  u64 n = n_max;
  if (n_dev) {
    u64 nd = *n_dev;
    n = nd < n_max ? nd : n_max;
  }
  __shared__ SlowQueue<KW> sq[256 / 32];
  SlowQueue<KW>& q = sq[threadIdx.x >> 5];
  if ((threadIdx.x & 31) == 0) q.count = 0;
  __syncwarp();
  LocalStats st = {0, 0, 0, 0};
  HitSink sink = {nullptr, nullptr, 0, nullptr};
  u64 stride = (u64)gridDim.x * blockDim.x;
  u64 n_iter = (n + stride * CHUNK - 1) / (stride * CHUNK);   // warp-uniform (see k_stream)
  u64 first = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  for (u64 itn = 0; itn < n_iter; ++itn) {
    u64 i0 = first + itn * stride * CHUNK;
    Key<KW> keys[CHUNK];
    Bucket<KW> bk[CHUNK];
    u32 bidx[CHUNK];
    u32 okm = 0;
#pragma unroll
    for (int u = 0; u < CHUNK; ++u) {
      u64 i = i0 + (u64)u * stride;
      if (i < n) {
        okm |= 1u << u;
        keys[u] = ld_key_stream<KW>(lo, hi, i);
      }
    }
#pragma unroll
    for (int u = 0; u < CHUNK; ++u) {
      if (okm & (1u << u)) {
        u64 h = hash_key(keys[u]);
        if (FILT && part_of(h, filt_log2) != filt_val) {
          okm &= ~(1u << u);
        } else {
          bidx[u] = bucket_of(h, t.log2_parts, t.n_buckets);
          bk[u] = ld_bucket<KW>(t.keys + (u64)bidx[u] * 4 * KW);
        }
      }
    }
    st.windows += __popc(okm);
#pragma unroll
    for (int u = 0; u < CHUNK; ++u) {
      if (okm & (1u << u)) {
        int j = match_in(bk[u], keys[u]);
        if (j >= 0) {
          st.hits++;
          on_found<OP, KW>(t, (u64)bidx[u] * S + j, plane, arg, 0, sink);
        } else if (kInsert) {
          tally(st, sq_push_or_resolve<KW, OP>(q, t, bidx[u], keys[u], plane, arg, 0, sink));
        } else if (!has_empty(bk[u], keys[u], t.fast_empty)) {
          u32 nb = (bidx[u] + 1 == t.n_buckets) ? 0 : bidx[u] + 1;
          tally(st, sq_push_or_resolve<KW, OP>(q, t, nb, keys[u], plane, arg, 0, sink));
        }
      }
    }
    __syncwarp();
    if (q.count >= 32) tally_packed(st, sq_drain<KW, OP>(&q, t, plane, arg, sink, false));
  }
  tally_packed(st, sq_drain<KW, OP>(&q, t, plane, arg, sink, true));
  flush_stats(st, stats);
}

// ------------------------------------------------ K2c: packed counting -----
// kdf_count_bins only has to answer "count >= min_child_count ?" and "in the
// reference ?" (the counts themselves are discarded, discovery/pipeline.py:
// 207-226), and k is odd, so a key leaves >= 2 spare bits at the top of its most
// significant word.  The packed form keeps a SATURATING counter there and drops
// the value planes: a slot is one word (pair), and once a k-mer has been seen
// `sat` times every further copy is a plain bucket read — no atomic, no plane
// sector.  At 30x that is ~21 of every 24 copies.  States of the field:
// 1..sat = copies seen (saturating), 0 = "reached sat, then found in the
// reference" (an occupied slot always has count >= 1, so 0 is free).  The all-ones
// word stays the empty marker: it would be the all-T key, which is never canonical.
// slot of the bucket holding `key` under `mask` (key bits of the most significant
// word), or -1; `ms` receives that slot's most significant word
__device__ __forceinline__ int match_packed(const Bucket<1>& b, const Key<1>& key, u64 mask, u64& ms) {
  int j = -1;
#pragma unroll
  for (int c = 3; c >= 0; --c)
    if ((b.q[c] & mask) == key.lo) {
      j = c;
      ms = b.q[c];
    }
  return j;
}
__device__ __forceinline__ int match_packed(const Bucket<2>& b, const Key<2>& key, u64 mask, u64& ms) {
  int j = -1;
#pragma unroll
  for (int c = 3; c >= 0; --c)
    if (b.q[2 * c] == key.lo && (b.q[2 * c + 1] & mask) == key.hi) {
      j = c;
      ms = b.q[2 * c + 1];
    }
  return j;
}
__device__ __forceinline__ bool same_packed(const Key<1>& stored, const Key<1>& key, u64 mask) {
  return (stored.lo & mask) == key.lo;
}
__device__ __forceinline__ bool same_packed(const Key<2>& stored, const Key<2>& key, u64 mask) {
  return stored.lo == key.lo && (stored.hi & mask) == key.hi;
}
// saturating +1 on the state field; `cur` is a (possibly stale) view of the word
__device__ __forceinline__ void packed_bump(u64* pms, u64 cur, int sh, u32 sat) {
  for (;;) {
    if ((u32)(cur >> sh) >= sat) return;
    u64 old = atomicCAS(pms, cur, cur + (1ull << sh));
    if (old == cur) return;
    cur = old;
  }
}

template <int KW>
__device__ __forceinline__ u32 resolve_packed_count(const TableView<KW>& t, u32 b, const Key<KW>& key,
                                                    int sh, u32 sat) {
  constexpr int S = SPB<KW>::v;
  const u64 mask = (1ull << sh) - 1;
  Key<KW> val = key;
  ((u64*)&val)[KW - 1] |= 1ull << sh;  // state 1
  for (u32 n = 0; n < t.n_buckets; ++n) {
    Bucket<KW> bk = ld_bucket<KW>(t.keys + (u64)b * 4 * KW);
    u64 ms = 0;
    int j = match_packed(bk, key, mask, ms);
    if (j >= 0) {
      packed_bump(t.keys + ((u64)b * S + j) * KW + (KW - 1), ms, sh, sat);
      return R_HIT;
    }
#pragma unroll
    for (int c = 0; c < S; ++c) {
      if (maybe_empty(bk, c, key)) {
        u64 slot = (u64)b * S + c;
        Key<KW> old = cas_key(t.keys + slot * KW, val);
        if (is_empty_key(old)) return R_NEW;
        if (same_packed(old, key, mask)) {  // inserted by another thread since the load
          packed_bump(t.keys + slot * KW + (KW - 1), ((const u64*)&old)[KW - 1], sh, sat);
          return R_HIT;
        }
      }
    }
    b = (b + 1 == t.n_buckets) ? 0 : b + 1;
  }
  return R_FULL;
}

template <int KW>
__device__ __forceinline__ u32 resolve_packed_mark(const TableView<KW>& t, u32 b, const Key<KW>& key,
                                                   int sh, u32 sat) {
  constexpr int S = SPB<KW>::v;
  const u64 mask = (1ull << sh) - 1;
  for (u32 n = 0; n < t.n_buckets; ++n) {
    Bucket<KW> bk = ld_bucket<KW>(t.keys + (u64)b * 4 * KW);
    u64 ms = 0;
    int j = match_packed(bk, key, mask, ms);
    if (j >= 0) {
      if ((u32)(ms >> sh) == sat) atomicAnd(t.keys + ((u64)b * S + j) * KW + (KW - 1), mask);
      return R_HIT;
    }
    if (has_empty(bk, key, false)) return R_MISS;
    b = (b + 1 == t.n_buckets) ? 0 : b + 1;
  }
  return R_FULL;
}

// What bounds this kernel (round 2, profiles/r2c_count_variants.md): a lean variant
// without the software pipeline (80 registers, 3 blocks per SM, a third fewer
// instructions) and a split-phase variant (compare-and-swaps issued in one drain,
// checked in the next) were built and measured: 18.8 / 20.9 ms against 18.8 ms for this
// kernel.  Occupancy, instruction count and exposed atomic latency are not what limits it;
// the L2's rate for the mix "one 32-byte read per key + a returning compare-and-swap for one
// key in six" is (microbenchmark: 95 G ops/s = 16.1 ms for the 1.53 G keys of the bench).
// k_update_keys for the packed form: OP_PACKED_COUNT inserts / bumps the child's
// keys, OP_PACKED_MARK turns "saturated" into "saturated, in the reference".
//
// The grid fills the machine once, so a thread walks ~80 chunks one after the other
// and the latencies of a chunk would add up: key load (DRAM) -> bucket load (L2) ->
// CAS.  The loop is therefore software-pipelined three deep: while chunk A is
// resolved, the bucket loads of chunk B and the key loads of chunk C are in flight.
// up to KEY_SEGS key arrays consumed by one k_packed_keys launch (see there)
constexpr int KEY_SEGS = 16;
struct KeySegs {
  const u64* lo[KEY_SEGS];
  const u64* n_dev[KEY_SEGS];   // device count of the segment, or NULL (= n_max)
  int n;
};

template <int KW> __device__ __forceinline__ Key<KW> ld_key_pinned(const u64* lo, u64 i);
template <> __device__ __forceinline__ Key<1> ld_key_pinned<1>(const u64* lo, u64 i) {
  Key<1> k;
  asm volatile("ld.global.cs.u64 %0, [%1];" : "=l"(k.lo) : "l"(lo + i) : "memory");
  return k;
}
template <> __device__ __forceinline__ Key<2> ld_key_pinned<2>(const u64* lo, u64 i) {
  Key<2> k;  // interleaved {lo, hi} pairs (bins)
  asm volatile("ld.global.cs.v2.u64 {%0,%1}, [%2];" : "=l"(k.lo), "=l"(k.hi) : "l"(lo + 2 * i) : "memory");
  return k;
}

// Per-warp queue of the atomics a packed count still owes.  In the stream of a bin
// only ~1 key in 6 needs one (a new key, or a copy that has not saturated yet), so
// done in place they would run with a handful of lanes and the warp would sit out a
// full L2 round trip for each of them.  Lanes queue them instead — with everything
// the first attempt needs, so no bucket is read again — and the warp drains the queue
// 64 at a time (measured: 64 -> 18.3 ms, 128 -> 18.9, 256 -> 21.9 per 1.53 G keys): two
// compare-and-swaps per lane in flight together (one instruction
// site for bumps and inserts alike, so no result is waited for before the last one is
// issued), results checked afterwards.  What the first attempt cannot settle (the slot
// was taken meanwhile, the home bucket was full) goes to the SlowQueue, whose drain
// probes from scratch with every lane busy.
#ifndef KDF_PQ_DRAIN
#define KDF_PQ_DRAIN 64
#endif
constexpr int PQ_DRAIN = KDF_PQ_DRAIN;
constexpr int PQ_CAP = PQ_DRAIN + 128;   // a chunk adds at most 128 items per warp
template <int KW> struct PackedQueue {
  u64 lo[PQ_CAP];
  u64 hi[KW == 2 ? PQ_CAP : 1];
  u64 expect[PQ_CAP];  // bump: the state word as read; insert: unused
  u32 b[PQ_CAP];       // home bucket
  u32 info[PQ_CAP];    // kind (1 bump, 2 insert) | slot << 2
  u32 count;
};
template <int KW> struct PackedQueues {
  PackedQueue<KW> fast;
  SlowQueue<KW> slow;
};

template <int KW>
__device__ __forceinline__ bool pq_push(PackedQueue<KW>& q, const Key<KW>& key, u32 b, u32 info, u64 expect) {
  u32 o = atomicAdd(&q.count, 1u);
  if (o >= (u32)PQ_CAP) return false;  // full: the caller resolves in place
  q.lo[o] = key.lo;
  if (KW == 2) q.hi[o] = ((const u64*)&key)[KW - 1];
  q.expect[o] = expect;
  q.b[o] = b;
  q.info[o] = info;
  return true;
}

// all lanes of the warp: drain batches of PQ_DRAIN (or everything when `all`);
// returns fresh << 10 | full << 20 | hits (the tallies of the resolved items)
template <int KW>
__device__ __noinline__ u32 pq_drain(PackedQueues<KW>* qp, TableView<KW> t, int sh, u32 sat, bool all) {
  constexpr int S = SPB<KW>::v;
  constexpr int R = PQ_DRAIN / 32;
  PackedQueue<KW>& q = qp->fast;
  const unsigned lane = threadIdx.x & 31;
  const u64 one = 1ull << sh;
  const HitSink sink = {nullptr, nullptr, 0, nullptr};
  __syncwarp();
  u32 n = q.count;
  if (n > (u32)PQ_CAP) n = PQ_CAP;
  u32 packed = 0;
  while (n >= (u32)PQ_DRAIN || (all && n > 0)) {
    const u32 take = n >= (u32)PQ_DRAIN ? (u32)PQ_DRAIN : n;
    const u32 base = n - take;
    Key<KW> want[R], got[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {  // issue: nothing below looks at a result
      const u32 idx = base + r * 32 + lane;
      want[r].lo = EMPTY;
      if (KW == 2) ((u64*)&want[r])[KW - 1] = EMPTY;
      got[r] = want[r];
      if (r * 32 + lane < take) {
        const u32 info = q.info[idx];
        Key<KW> key;
        key.lo = q.lo[idx];
        if (KW == 2) ((u64*)&key)[KW - 1] = q.hi[idx];
        Key<KW> val = key;
        if ((info & 3u) == 1) {  // bump: expect the slot as read, write state + 1
          ((u64*)&want[r])[KW - 1] = q.expect[idx];
          if (KW == 2) want[r].lo = key.lo;
          ((u64*)&val)[KW - 1] = q.expect[idx] + one;
        } else {                 // insert: expect the empty slot, write the key in state 1
          ((u64*)&val)[KW - 1] |= one;
        }
        got[r] = cas_slot(t.keys + ((u64)q.b[idx] * S + (info >> 2)) * KW, want[r], val);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {  // check
      const u32 idx = base + r * 32 + lane;
      if (r * 32 + lane < take) {
        const u32 info = q.info[idx];
        const bool won = got[r] == want[r];
        if ((info & 3u) == 1) {
          if (!won)  // bumped by someone else in between: go on from the new value
            packed_bump(t.keys + ((u64)q.b[idx] * S + (info >> 2)) * KW + (KW - 1),
                        ((const u64*)&got[r])[KW - 1], sh, sat);
        } else if (won) {
          packed += 1u << 10;
        } else {     // the slot was taken meanwhile: probe from scratch, with the other such cases
          Key<KW> key;
          key.lo = q.lo[idx];
          if (KW == 2) ((u64*)&key)[KW - 1] = q.hi[idx];
          u32 code = sq_push_or_resolve<KW, OP_PACKED_COUNT>(qp->slow, t, q.b[idx], key, sh, sat, 0, sink);
          packed += (code == R_HIT ? 1u : 0u) + (code == R_NEW ? (1u << 10) : 0u);
          packed |= (code == R_FULL ? (1u << 20) : 0u);
        }
      }
    }
    n = base;
    __syncwarp();
  }
  if (lane == 0) q.count = n;
  __syncwarp();
  return packed;
}

template <int KW, int OP>
using PackedKeysQueue = typename std::conditional<OP == OP_PACKED_COUNT, PackedQueues<KW>, SlowQueue<KW>>::type;

template <int KW, int OP, bool FILT>
__global__ void __launch_bounds__(256, KDF_KEYS_BLOCKS) k_packed_keys(TableView<KW> t, const __grid_constant__ KeySegs segs, u64 n_max,
                                                     int sh, u32 sat, u64* stats,
                                                     int filt_log2, u32 filt_val) {
  // The keys come as segs.n segments (one per sending rank of a multi-GPU run; one
  // otherwise): a single launch walks them all — a launch per ~6 M-key segment cost ~8 us
  // of ramp each, 544 launches per count at 8 GPUs.  The segment table stays in the
  // kernel's parameter space (constant bank), so it costs the 128-register kernel nothing.
  constexpr int CHUNK = KW == 2 ? 2 : KDF_KEYS_CHUNK;   // two bucket sets live in registers
  constexpr int S = SPB<KW>::v;
  const u64 mask = (1ull << sh) - 1;
  const u64* lo = nullptr;
  u64 n = 0;
  using Queue = PackedKeysQueue<KW, OP>;
  extern __shared__ __align__(16) unsigned char pk_smem[];   // one queue per warp
  Queue& q = reinterpret_cast<Queue*>(pk_smem)[threadIdx.x >> 5];
  if ((threadIdx.x & 31) == 0) {
    if constexpr (OP == OP_PACKED_COUNT) {
      q.fast.count = 0;
      q.slow.count = 0;
    } else {
      q.count = 0;
    }
  }
  __syncwarp();
  LocalStats st = {0, 0, 0, 0};
  HitSink sink = {nullptr, nullptr, 0, nullptr};
  const u64 stride = (u64)gridDim.x * blockDim.x;
  u64 n_iter = 0;   // warp-uniform (see k_stream); set per segment below
  const u64 first = (u64)blockIdx.x * blockDim.x + threadIdx.x;

  Key<KW> kA[CHUNK], kB[CHUNK], kC[CHUNK];
  Bucket<KW> bA[CHUNK], bB[CHUNK];
  u32 iA[CHUNK], iB[CHUNK];
  u32 mA = 0, mB = 0, mC = 0;

  // stage 1: the chunk's keys (mask of the in-range ones)
  auto load_keys = [&](u64 itn, Key<KW>* keys) -> u32 {
    u32 okm = 0;
    u64 i0 = first + itn * stride * CHUNK;
#pragma unroll
    for (int u = 0; u < CHUNK; ++u) {
      u64 i = i0 + (u64)u * stride;
      if (itn < n_iter && i < n) {
        okm |= 1u << u;
        keys[u] = ld_key_pinned<KW>(lo, i);
      }
    }
    return okm;
  };
  // stage 2: hash, home bucket loads
  auto load_buckets = [&](const Key<KW>* keys, u32& okm, u32* bidx, Bucket<KW>* bk) {
#pragma unroll
    for (int u = 0; u < CHUNK; ++u) {
      if (okm & (1u << u)) {
        u64 h = hash_key(keys[u]);
        if (FILT && part_of(h, filt_log2) != filt_val) {
          okm &= ~(1u << u);
        } else {
          bidx[u] = bucket_of(h, t.log2_parts, t.n_buckets);
          bk[u] = ld_bucket<KW>(t.keys + (u64)bidx[u] * 4 * KW);
        }
      }
    }
  };
  // stage 3: resolve.  A saturated copy is done once its bucket has been read; a
  // copy that needs an atomic is queued (see PackedQueue) — through ONE push site for
  // bumps and inserts alike: what a key owes is first reduced to (kind, slot, expected
  // word), so the queue code is executed once per chunk position, not once per kind.
  auto resolve = [&](const Key<KW>* keys, u32 okm, const u32* bidx, const Bucket<KW>* bk) {
    st.windows += __popc(okm);
    if constexpr (OP == OP_PACKED_COUNT) {
#pragma unroll
      for (int u = 0; u < CHUNK; ++u) {
        if (okm & (1u << u)) {
          u64 ms = 0;
          const int j = match_packed(bk[u], keys[u], mask, ms);
          u32 info = 0;   // kind (1 bump, 2 insert) | slot << 2; 0 = nothing owed
          if (j >= 0) {
            st.hits++;
#ifndef KDF_DBG_NO_BUMP
            if ((u32)(ms >> sh) < sat) info = 1u | ((u32)j << 2);
#endif
          } else {
#ifndef KDF_DBG_NO_INSERT
            int c = -1;
#pragma unroll
            for (int cc = S - 1; cc >= 0; --cc)
              if (maybe_empty(bk[u], cc, keys[u])) c = cc;
            if (c < 0) {  // home bucket full: probe on from the next one
              u32 nb = (bidx[u] + 1 == t.n_buckets) ? 0 : bidx[u] + 1;
              tally(st, sq_push_or_resolve<KW, OP>(q.slow, t, nb, keys[u], sh, sat, 0, sink));
            } else {
              info = 2u | ((u32)c << 2);
              ms = 0;
            }
#endif
          }
          if (info && !pq_push<KW>(q.fast, keys[u], bidx[u], info, ms)) {   // queue full: in place
            if ((info & 3u) == 1) packed_bump(t.keys + ((u64)bidx[u] * S + (info >> 2)) * KW + (KW - 1), ms, sh, sat);
            else tally(st, resolve_packed_count<KW>(t, bidx[u], keys[u], sh, sat));
          }
        }
      }
      __syncwarp();
      if (q.fast.count >= (u32)PQ_DRAIN) tally_packed(st, pq_drain<KW>(&q, t, sh, sat, false));
      if (q.slow.count >= 32) tally_packed(st, sq_drain<KW, OP>(&q.slow, t, sh, sat, sink, false));
    } else {
      // probing ops: OP_PACKED_MARK on a packed slice, or COUNT_IF_PRESENT /
      // MARK_IF_PRESENT on a table with value planes (then sh = plane, sat = arg)
#pragma unroll
      for (int u = 0; u < CHUNK; ++u) {
        if (okm & (1u << u)) {
          bool full;
          if constexpr (OP == OP_PACKED_MARK) {
            u64 ms = 0;
            int j = match_packed(bk[u], keys[u], mask, ms);
            if (j >= 0) {
              st.hits++;
              if ((u32)(ms >> sh) == sat) atomicAnd(t.keys + ((u64)bidx[u] * S + j) * KW + (KW - 1), mask);
            }
            full = j < 0 && !has_empty(bk[u], keys[u], false);
          } else {
            int j = match_in(bk[u], keys[u]);
            if (j >= 0) {
              st.hits++;
              on_found<OP, KW>(t, (u64)bidx[u] * S + j, sh, sat, 0, sink);
            }
            full = j < 0 && !has_empty(bk[u], keys[u], t.fast_empty);
          }
          if (full) {  // full bucket: look further, queued
            u32 nb = (bidx[u] + 1 == t.n_buckets) ? 0 : bidx[u] + 1;
            tally(st, sq_push_or_resolve<KW, OP>(q, t, nb, keys[u], sh, sat, 0, sink));
          }
        }
      }
      __syncwarp();
      if (q.count >= 32) tally_packed(st, sq_drain<KW, OP>(&q, t, sh, sat, sink, false));
    }
  };

#pragma unroll 1
  for (int seg = 0; seg < segs.n; ++seg) {
    lo = segs.lo[seg];
    n = n_max;
    if (segs.n_dev[seg]) {
      const u64 nd = *segs.n_dev[seg];
      n = nd < n_max ? nd : n_max;
    }
    n_iter = (n + stride * CHUNK - 1) / (stride * CHUNK);
    mA = load_keys(0, kA);
    mB = load_keys(1, kB);
    load_buckets(kA, mA, iA, bA);
    for (u64 itn = 0; itn < n_iter; ++itn) {
      mC = load_keys(itn + 2, kC);
      load_buckets(kB, mB, iB, bB);
      resolve(kA, mA, iA, bA);
#pragma unroll
      for (int u = 0; u < CHUNK; ++u) {
        kA[u] = kB[u];
        bA[u] = bB[u];
        iA[u] = iB[u];
        kB[u] = kC[u];
      }
      mA = mB;
      mB = mC;
    }
  }
  if constexpr (OP == OP_PACKED_COUNT) {
    tally_packed(st, pq_drain<KW>(&q, t, sh, sat, true));
    tally_packed(st, sq_drain<KW, OP>(&q.slow, t, sh, sat, sink, true));
  } else {
    tally_packed(st, sq_drain<KW, OP>(&q, t, sh, sat, sink, true));
  }
  flush_stats(st, stats);
}

// Emit + clear of a packed slice.  keep = "reached sat" (and, unless ignore_ref,
// "not found in the reference"); n_count counts the slots that reached sat whether
// or not the reference holds them (count_all: every occupied slot).  Emitted keys
// have the state field cleared.  A thread takes U buckets per round with all loads
// in flight, and a warp reserves its output range with ONE atomic per round (lane
// counts -> shuffle scan), so the pass is not a chain of same-address atomics.
template <int KW>
__global__ void __launch_bounds__(256) k_emit_packed(TableView<KW> t, int sh, u32 sat, int ignore_ref,
                                                     int count_all, u64* out_lo, u64* out_hi, u64 cap,
                                                     u64* n_out, u64* n_count, u64* n_occupied) {
  constexpr int S = SPB<KW>::v;
  constexpr int U = KW == 2 ? 2 : 4;
  const u64 mask = (1ull << sh) - 1;
  const u64 stride = (u64)gridDim.x * blockDim.x;
  const u64 first = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  const u64 rounds = ((u64)t.n_buckets + stride * U - 1) / (stride * U);
  const unsigned lane = threadIdx.x & 31;
  u32 cnt_ge = 0, cnt_occ = 0;
  for (u64 r = 0; r < rounds; ++r) {
    Bucket<KW> bk[U];
    u32 keepm = 0;   // bit u * S + j: slot j of bucket u is emitted
#pragma unroll
    for (int u = 0; u < U; ++u) {
      u64 b = first + (r * U + u) * stride;
#pragma unroll
      for (int i = 0; i < 4 * KW; ++i) bk[u].q[i] = EMPTY;
      if (b < t.n_buckets) bk[u] = ld_bucket<KW>(t.keys + b * 4 * KW);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      u64 b = first + (r * U + u) * stride;
      bool any = false;
#pragma unroll
      for (int i = 0; i < 4 * KW; ++i) any = any || bk[u].q[i] != EMPTY;
      if (any) {
        asm volatile("st.global.cg.v4.u64 [%0], {%1,%1,%1,%1};" ::"l"(t.keys + b * 4 * KW), "l"(EMPTY) : "memory");
        if (KW == 2)
          asm volatile("st.global.cg.v4.u64 [%0], {%1,%1,%1,%1};" ::"l"(t.keys + b * 4 * KW + 4), "l"(EMPTY) : "memory");
#pragma unroll
        for (int j = 0; j < S; ++j) {
          u64 ms = bk[u].q[j * KW + (KW - 1)];
          bool occ = !(ms == EMPTY && bk[u].q[j * KW] == EMPTY);
          u32 state = (u32)(ms >> sh);
          bool reached = state >= sat || state == 0;
          bool keep = occ && (ignore_ref ? reached : state >= sat);
          cnt_occ += occ ? 1u : 0u;
          cnt_ge += (occ && (count_all || reached)) ? 1u : 0u;
          keepm |= (keep ? 1u : 0u) << (u * S + j);
        }
      }
    }
    // warp-wide exclusive scan of the per-lane counts, one reservation per warp
    u32 mine = __popc(keepm), incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      u32 v = __shfl_up_sync(0xffffffffu, incl, o);
      if ((int)lane >= o) incl += v;
    }
    u32 total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) continue;
    u64 base = 0;
    if (lane == 0) base = atomicAdd(n_out, (u64)total);
    base = __shfl_sync(0xffffffffu, base, 0);
    u64 o = base + (incl - mine);
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int j = 0; j < S; ++j) {
        if (keepm & (1u << (u * S + j))) {
          if (o < cap) {
            if (KW == 1) {
              if (out_lo) out_lo[o] = bk[u].q[j] & mask;
            } else {
              if (out_lo) out_lo[o] = bk[u].q[j * KW];
              if (out_hi) out_hi[o] = bk[u].q[j * KW + (KW - 1)] & mask;
            }
          }
          ++o;
        }
      }
    }
  }
  for (int o = 16; o; o >>= 1) {
    cnt_ge += __shfl_xor_sync(0xffffffffu, cnt_ge, o);
    cnt_occ += __shfl_xor_sync(0xffffffffu, cnt_occ, o);
  }
  if (lane == 0) {
    if (n_count && cnt_ge) atomicAdd(n_count, (u64)cnt_ge);
    if (n_occupied && cnt_occ) atomicAdd(n_occupied, (u64)cnt_occ);
  }
}

// ------------------------------------------------------------- K3 ---------
template <int KW>
__global__ void __launch_bounds__(256) k_threshold_compact(TableView<KW> t, u64 capacity, u32 min0,
                                                           u32 max0, u32 min1, u32 max1, u64* out_lo,
                                                           u64* out_hi, u32* out_p0, u32* out_p1,
                                                           u64 cap, u64* n_out, u32 count_min0,
                                                           u64* n_count, u64* n_occupied) {
  u64 stride = (u64)gridDim.x * blockDim.x;
  u64 first = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  u64 rounds = (capacity + stride - 1) / stride;
  unsigned lane = threadIdx.x & 31;
  u32 cnt_ge = 0, cnt_occ = 0;
  for (u64 r = 0; r < rounds; ++r) {
    u64 i = first + r * stride;
    bool keep = false;
    Key<KW> key;
    key.lo = 0;
    u32 p0 = 0, p1 = 0;
    if (i < capacity) {
      key.lo = __ldcg(t.keys + i * KW);
      if (KW == 2) ((u64*)&key)[KW - 1] = __ldcg(t.keys + i * KW + 1);
      if (!is_empty_key(key)) {
        p0 = __ldcg(t.p0 + i);
        p1 = __ldcg(t.p1 + i);
        keep = p0 >= min0 && p0 <= max0 && p1 >= min1 && p1 <= max1;
        cnt_occ++;
        cnt_ge += (p0 >= count_min0) ? 1u : 0u;
      }
    }
    unsigned m = __ballot_sync(0xffffffffu, keep);
    if (m) {
      u64 base = 0;
      int leader = __ffs(m) - 1;
      if ((int)lane == leader) base = atomicAdd(n_out, (u64)__popc(m));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (keep) {
        u64 o = base + __popc(m & ((1u << lane) - 1));
        if (o < cap) {
          if (out_lo) out_lo[o] = key.lo;
          if (KW == 2 && out_hi) out_hi[o] = ((const u64*)&key)[KW - 1];
          if (out_p0) out_p0[o] = p0;
          if (out_p1) out_p1[o] = p1;
        }
      }
    }
  }
  if (n_count || n_occupied) {
    for (int o = 16; o; o >>= 1) {
      cnt_ge += __shfl_xor_sync(0xffffffffu, cnt_ge, o);
      cnt_occ += __shfl_xor_sync(0xffffffffu, cnt_occ, o);
    }
    if (lane == 0) {
      if (n_count && cnt_ge) atomicAdd(n_count, (u64)cnt_ge);
      if (n_occupied && cnt_occ) atomicAdd(n_occupied, (u64)cnt_occ);
    }
  }
}

// Bucket-wise form used by kdf_count_bins: one thread per 32-byte bucket (4 or 2
// slots, three wide loads), same predicate and compaction as above, and the slice
// is cleared on the way out (CLEAR) so that the next bin needs no fill pass.
template <int KW, bool CLEAR>
__global__ void __launch_bounds__(256) k_emit_buckets(TableView<KW> t, u32 min0, u32 max0, u32 min1,
                                                      u32 max1, u64* out_lo, u64* out_hi,
                                                      u32* out_p0, u32* out_p1, u64 cap, u64* n_out,
                                                      u32 count_min0, u64* n_count, u64* n_occupied) {
  constexpr int S = SPB<KW>::v;
  u64 stride = (u64)gridDim.x * blockDim.x;
  u64 first = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  u64 rounds = ((u64)t.n_buckets + stride - 1) / stride;
  unsigned lane = threadIdx.x & 31;
  u32 cnt_ge = 0, cnt_occ = 0;
  for (u64 r = 0; r < rounds; ++r) {
    u64 b = first + r * stride;
    bool inb = b < t.n_buckets;
    Bucket<KW> bk;
#pragma unroll
    for (int i = 0; i < 4 * KW; ++i) bk.q[i] = EMPTY;
    u32 p0[S], p1[S];
#pragma unroll
    for (int j = 0; j < S; ++j) p0[j] = p1[j] = 0;
    bool any = false;
    if (inb) {
      bk = ld_bucket<KW>(t.keys + b * 4 * KW);
#pragma unroll
      for (int i = 0; i < 4 * KW; ++i) any = any || bk.q[i] != EMPTY;
      if (any) {
        uint4 a = __ldcg(reinterpret_cast<const uint4*>(t.p0 + b * 4));
        uint4 c = __ldcg(reinterpret_cast<const uint4*>(t.p1 + b * 4));
        p0[0] = a.x; p0[1] = a.y; p0[2] = a.z; p0[3] = a.w;
        p1[0] = c.x; p1[1] = c.y; p1[2] = c.z; p1[3] = c.w;
        if (CLEAR) {
          asm volatile("st.global.cg.v4.u64 [%0], {%1,%1,%1,%1};" ::"l"(t.keys + b * 4 * KW), "l"(EMPTY) : "memory");
          if (KW == 2)
            asm volatile("st.global.cg.v4.u64 [%0], {%1,%1,%1,%1};" ::"l"(t.keys + b * 4 * KW + 4), "l"(EMPTY) : "memory");
          *reinterpret_cast<uint4*>(t.p0 + b * 4) = make_uint4(0, 0, 0, 0);
          *reinterpret_cast<uint4*>(t.p1 + b * 4) = make_uint4(0, 0, 0, 0);
        }
      }
    }
    if (!__any_sync(0xffffffffu, any)) continue;
#pragma unroll
    for (int j = 0; j < S; ++j) {
      Key<KW> key;
      key.lo = bk.q[j * KW];
      if (KW == 2) ((u64*)&key)[KW - 1] = bk.q[j * KW + (KW - 1)];
      bool occ = any && !is_empty_key(key);
      bool keep = occ && p0[j] >= min0 && p0[j] <= max0 && p1[j] >= min1 && p1[j] <= max1;
      cnt_occ += occ ? 1u : 0u;
      cnt_ge += (occ && p0[j] >= count_min0) ? 1u : 0u;
      unsigned m = __ballot_sync(0xffffffffu, keep);
      if (m) {
        u64 base = 0;
        int leader = __ffs(m) - 1;
        if ((int)lane == leader) base = atomicAdd(n_out, (u64)__popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (keep) {
          u64 o = base + __popc(m & ((1u << lane) - 1));
          if (o < cap) {
            if (out_lo) out_lo[o] = key.lo;
            if (KW == 2 && out_hi) out_hi[o] = ((const u64*)&key)[KW - 1];
            if (out_p0) out_p0[o] = p0[j];
            if (out_p1) out_p1[o] = p1[j];
          }
        }
      }
    }
  }
  for (int o = 16; o; o >>= 1) {
    cnt_ge += __shfl_xor_sync(0xffffffffu, cnt_ge, o);
    cnt_occ += __shfl_xor_sync(0xffffffffu, cnt_occ, o);
  }
  if (lane == 0) {
    if (n_count && cnt_ge) atomicAdd(n_count, (u64)cnt_ge);
    if (n_occupied && cnt_occ) atomicAdd(n_occupied, (u64)cnt_occ);
  }
}

// ------------------------------------------------------------- K4 ---------
template <int KW>
__device__ __forceinline__ bool find_slot(const TableView<KW>& t, const Key<KW>& key, u64& idx_out) {
  constexpr int S = SPB<KW>::v;
  u32 b = bucket_of(hash_key(key), t.log2_parts, t.n_buckets);
  for (u32 n = 0; n < t.n_buckets; ++n) {
    Bucket<KW> bk = ld_bucket<KW>(t.keys + (u64)b * 4 * KW);
    int j = match_in(bk, key);
    if (j >= 0) {
      idx_out = (u64)b * S + j;
      return true;
    }
    if (has_empty(bk, key, t.fast_empty)) return false;
    b = (b + 1 == t.n_buckets) ? 0 : b + 1;
  }
  return false;
}

template <int KW>
__global__ void __launch_bounds__(256) k_lookup_keys(TableView<KW> t, const u64* lo, const u64* hi,
                                                     u64 n, uint8_t* out_found, u32* out_p0,
                                                     u32* out_p1) {
  u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    Key<KW> key;
    key.lo = lo[i];
    if (KW == 2) ((u64*)&key)[KW - 1] = hi[i];
    u64 idx;
    bool f = find_slot<KW>(t, key, idx);
    if (out_found) out_found[i] = f ? 1 : 0;
    if (out_p0) out_p0[i] = f ? __ldcg(t.p0 + idx) : 0u;
    if (out_p1) out_p1[i] = f ? __ldcg(t.p1 + idx) : 0u;
  }
}

// per-key plane accumulation for existing keys (table growth / count merging)
template <int KW>
__global__ void __launch_bounds__(256) k_add_planes(TableView<KW> t, const u64* lo, const u64* hi,
                                                    u64 n, const u32* add0, const u32* add1,
                                                    u64* n_missing) {
  u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    Key<KW> key;
    key.lo = lo[i];
    if (KW == 2) ((u64*)&key)[KW - 1] = hi[i];
    u64 idx;
    if (find_slot<KW>(t, key, idx)) {
      if (add0 && add0[i]) atomicAdd(t.p0 + idx, add0[i]);
      if (add1 && add1[i]) atomicAdd(t.p1 + idx, add1[i]);
    } else if (n_missing) {
      atomicAdd(n_missing, 1ull);
    }
  }
}

// ------------------------------------------------------------- K5 ---------
// Dense form: one warp per read.  Lanes stride through the read's window
// starts; hits are appended to a per-warp shared list in position order
// (ballot prefix), then the warp counts distinct slot indices among them.
constexpr int SCAN_WARPS = 4;
constexpr int SCAN_HCAP = 1024;

template <int KW>
__global__ void __launch_bounds__(SCAN_WARPS * 32) k_scan_reads(
    TableView<KW> t, StreamView s, int k, const u64* read_starts, const u32* read_lens, u64 n_reads,
    u32 min_distinct,
    u32* out_ndistinct, u32* out_nhits, u64* hit_pos, u32* hit_slot, u64 hit_cap, u64* n_hits,
    u64* stats) {
  __shared__ u32 sh_slot[SCAN_WARPS][SCAN_HCAP];
  __shared__ u32 sh_win[SCAN_WARPS][SCAN_HCAP];
  const unsigned lane = threadIdx.x & 31;
  const unsigned wib = threadIdx.x >> 5;
  u32* my_slot = sh_slot[wib];
  u32* my_win = sh_win[wib];
  u64 warp_global = (u64)blockIdx.x * SCAN_WARPS + wib;
  u64 warp_stride = (u64)gridDim.x * SCAN_WARPS;
  u32 windows_done = 0;
  for (u64 r = warp_global; r < n_reads; r += warp_stride) {
    u64 start = read_starts[r];
    u64 len = read_lens[r];
    u32 nwin = len >= (u64)k ? (u32)(len - k + 1) : 0;
    u32 nh = 0;  // hits so far (warp-uniform)
    for (u32 base = 0; base < nwin; base += 32) {
      u32 i = base + lane;
      bool hit = false;
      u64 idx = 0;
      if (i < nwin) {
        Key<KW> key;
        if (WindowAt<KW>::get(s, start + i, k, key)) {
          windows_done++;
          hit = find_slot<KW>(t, key, idx);
        }
      }
      unsigned m = __ballot_sync(0xffffffffu, hit);
      if (hit) {
        u32 o = nh + __popc(m & ((1u << lane) - 1));
        if (o < SCAN_HCAP) {
          my_slot[o] = (u32)idx;
          my_win[o] = i;
        }
      }
      nh += __popc(m);
    }
    __syncwarp();
    u32 nd;
    bool overflow = nh > SCAN_HCAP;
    if (overflow) {
      nd = KDF_NDISTINCT_OVERFLOW;
    } else {
      u32 mine = 0;
      for (u32 a = lane; a < nh; a += 32) {
        u32 sa = my_slot[a];
        bool dup = false;
        for (u32 b = 0; b < a; ++b) {
          if (my_slot[b] == sa) {
            dup = true;
            break;
          }
        }
        mine += dup ? 0u : 1u;
      }
      for (int o = 16; o; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
      nd = mine;
    }
    if (lane == 0) {
      out_ndistinct[r] = nd;
      out_nhits[r] = nh;
    }
    bool emit = nh > 0 && (overflow || (nd >= min_distinct && nd >= 1));
    if (emit && n_hits) {
      u64 gbase = 0;
      if (lane == 0) gbase = atomicAdd(n_hits, (u64)nh);
      gbase = __shfl_sync(0xffffffffu, gbase, 0);
      if (hit_pos && hit_slot) {
        if (!overflow) {
          for (u32 a = lane; a < nh; a += 32) {
            u64 o = gbase + a;
            if (o < hit_cap) {
              hit_pos[o] = start + my_win[a];
              hit_slot[o] = my_slot[a];
            }
          }
        } else {
          // list did not fit in shared memory: recompute and stream hits out
          u32 done = 0;
          for (u32 base = 0; base < nwin; base += 32) {
            u32 i = base + lane;
            bool hit = false;
            u64 idx = 0;
            if (i < nwin) {
              Key<KW> key;
              if (WindowAt<KW>::get(s, start + i, k, key)) hit = find_slot<KW>(t, key, idx);
            }
            unsigned m = __ballot_sync(0xffffffffu, hit);
            if (hit) {
              u64 o = gbase + done + __popc(m & ((1u << lane) - 1));
              if (o < hit_cap) {
                hit_pos[o] = start + i;
                hit_slot[o] = (u32)idx;
              }
            }
            done += __popc(m);
          }
        }
      }
    }
    __syncwarp();
  }
  if (stats) {
    for (int o = 16; o; o >>= 1) windows_done += __shfl_xor_sync(0xffffffffu, windows_done, o);
    if (lane == 0 && windows_done) atomicAdd(stats + KDF_STAT_WINDOWS, (u64)windows_done);
  }
}

// Sparse form: hits (stream position, slot) sorted by position -> one record per read
// that has hits.  Distinct slots per read are counted by SORTING: k_hit_read_keys turns
// every hit into (read << 32 | slot), a second radix sort groups equal (read, slot)
// pairs, and the thread that owns the first entry of a read walks its run once,
// counting the entries that differ from their predecessor — linear in the run length
// (the first version compared every hit with all earlier ones of its read: quadratic
// on a long read full of hits).
__global__ void __launch_bounds__(256) k_hit_read_keys(const u64* pos, const u32* slot, u64 n_hits,
                                                       const u64* read_starts, u64 n_reads, u64* rkey) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_hits) return;
  u64 p = pos[i];
  u64 lo = 0, hi = n_reads;  // r = last read with read_starts[r] <= p
  while (hi - lo > 1) {
    u64 mid = (lo + hi) >> 1;
    if (read_starts[mid] <= p) lo = mid; else hi = mid;
  }
  rkey[i] = (lo << 32) | (u64)slot[i];
}

__global__ void __launch_bounds__(128) k_reduce_hits(const u64* rkey, u64 n_hits, const u64* pos,
                                                     const u64* read_starts, u64* rec_read,
                                                     u32* rec_ndistinct, u32* rec_nhits, u64* rec_first,
                                                     u64* n_recs) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_hits) return;
  const u64 key = rkey[i];
  const u64 r = key >> 32;
  if (i > 0 && (rkey[i - 1] >> 32) == r) return;  // not the first entry of this read
  u32 nh = 0, nd = 0;
  u64 prev = ~key;
  for (u64 j = i; j < n_hits; ++j) {
    const u64 kj = rkey[j];
    if ((kj >> 32) != r) break;
    nd += kj != prev ? 1u : 0u;
    prev = kj;
    ++nh;
  }
  // index of the read's first hit in the position-sorted list
  const u64 rs = read_starts[r];
  u64 lo = 0, hi = n_hits;
  while (lo < hi) {
    u64 mid = (lo + hi) >> 1;
    if (pos[mid] < rs) lo = mid + 1; else hi = mid;
  }
  u64 o = atomicAdd(n_recs, 1ull);
  rec_read[o] = r;
  rec_ndistinct[o] = nd;
  rec_nhits[o] = nh;
  rec_first[o] = lo;
}

// ----------------------------------------------------------- K2p / K6 -----
// Bin canonical k-mers by hash range (BY_OWNER = false: partition bits of the
// bucket hash, for the L2-sliced count) or by owner rank (BY_OWNER = true, in
// front of the all-to-all).  Every bin is a fixed-capacity region of `bins`;
// keys are staged in shared-memory queues and flushed a queue at a time so
// that global writes are contiguous runs.  A bin that overflows sets
// *overflow and drops the key: the caller must size bins or retry.
constexpr int BIN_THREADS = 256;
constexpr int BIN_MAX_PARTS = 512;

template <int KW>
__device__ __forceinline__ void st_key(u64* bins, u64 idx, const Key<KW>& key);
template <> __device__ __forceinline__ void st_key<1>(u64* bins, u64 idx, const Key<1>& key) {
  bins[idx] = key.lo;
}
template <> __device__ __forceinline__ void st_key<2>(u64* bins, u64 idx, const Key<2>& key) {
  reinterpret_cast<ulonglong2*>(bins)[idx] = make_ulonglong2(key.lo, key.hi);
}

// Where bin p lives: a region of one local array (ptrs == NULL), or its own base
// pointer — which may be PEER memory mapped over NVLink: then the queue flush below
// IS the transfer of the multi-GPU all-to-all (kdf_bin_stream_to).
struct BinDest {
  u64* base;
  u64* const* ptrs;
  u64 cap;
  __device__ __forceinline__ u64* of(u32 p, int kw) const {
    return ptrs ? ptrs[p] : base + (u64)p * cap * kw;
  }
};

template <int KW>
struct BinStage {
  // dynamic shared memory layout: cnt[P] u32 | gbase[P] u64 | queue[P][QCAP] keys
  u32* cnt;
  u64* gbase;
  u64* queue;
  int n_parts;
  int qcap;
  __device__ void init(unsigned char* smem, int P, int Q) {
    n_parts = P;
    qcap = Q;
    gbase = reinterpret_cast<u64*>(smem);
    queue = gbase + P;
    cnt = reinterpret_cast<u32*>(queue + (size_t)P * Q * KW);
    for (int i = threadIdx.x; i < P; i += blockDim.x) cnt[i] = 0;
  }
  static size_t bytes(int P, int Q) { return (size_t)P * 8 + (size_t)P * Q * KW * 8 + (size_t)P * 4; }
  // Append a key (all lanes of the warp call this together; `ok` says whether
  // the lane has one).  Queue overflow falls back to a direct global append.
  // With few bins (owner binning at 2..8 ranks) every lane would hit the same
  // few shared counters, so the warp aggregates: one atomic per bin per warp.
  __device__ __forceinline__ void push(bool ok, u32 p, const Key<KW>& key, const BinDest& dst,
                                       u64* cursors, u64* overflow) {
    u32 o = 0;
    if (n_parts <= 8) {
      const unsigned lane = threadIdx.x & 31;
      for (int b = 0; b < n_parts; ++b) {
        unsigned m = __ballot_sync(0xffffffffu, ok && p == (u32)b);
        if (m) {
          int leader = __ffs(m) - 1;
          u32 base = 0;
          if ((int)lane == leader) base = atomicAdd(&cnt[b], (u32)__popc(m));
          base = __shfl_sync(0xffffffffu, base, leader);
          if (ok && p == (u32)b) o = base + __popc(m & ((1u << lane) - 1));
        }
      }
    } else if (ok) {
      o = atomicAdd(&cnt[p], 1u);
    }
    if (!ok) return;
    if (o < (u32)qcap) {
      u64* q = queue + ((size_t)p * qcap + o) * KW;
      q[0] = key.lo;
      if (KW == 2) q[1] = ((const u64*)&key)[KW - 1];
    } else {
      u64 g = atomicAdd(cursors + p, 1ull);
      if (g < dst.cap) st_key<KW>(dst.of(p, KW), g, key);
      else atomicOr(overflow, 1ull);
    }
  }
  // all threads: write every queue to its bin and reset the counters
  __device__ void flush(const BinDest& dst, u64* cursors, u64* overflow) {
    __syncthreads();
    for (int p = threadIdx.x; p < n_parts; p += blockDim.x) {
      u32 c = cnt[p];
      if (c > (u32)qcap) c = qcap;
      gbase[p] = c ? atomicAdd(cursors + p, (u64)c) : 0ull;
    }
    __syncthreads();
    // many bins: a warp per bin; few bins: the whole block per bin
    const bool wide = n_parts <= 8;
    const unsigned lane = wide ? threadIdx.x : (threadIdx.x & 31);
    const unsigned step = wide ? blockDim.x : 32;
    const unsigned first = wide ? 0 : (threadIdx.x >> 5), n_warps = wide ? 1 : (blockDim.x >> 5);
    for (int p = first; p < n_parts; p += n_warps) {
      u32 c = cnt[p];
      if (c > (u32)qcap) c = qcap;
      u64 g0 = gbase[p];
      u64* out = dst.of(p, KW);
      for (u32 i = lane; i < c; i += step) {
        u64 g = g0 + i;
        const u64* q = queue + ((size_t)p * qcap + i) * KW;
        if (g < dst.cap) {
          Key<KW> key;
          key.lo = q[0];
          if (KW == 2) ((u64*)&key)[KW - 1] = q[1];
          st_key<KW>(out, g, key);
        } else {
          atomicOr(overflow, 1ull);
        }
      }
    }
    __syncthreads();
    for (int p = threadIdx.x; p < n_parts; p += blockDim.x) cnt[p] = 0;
    __syncthreads();
  }
};

// BMODE 0: hash range (part_of); 1: owner rank; 2: composite owner x hash range —
// bin = owner * n_local + part, n_parts = n_owners * n_local, log2_parts = log2(n_local).
// pass_log2 > 0 (BMODE 0 / 2): the hash ranges are split into 2^pass_log2 groups by their
// TOP bits and only the keys of group pass_val are binned — the multi-pass child count,
// which bounds the bins' memory by re-extracting the stream once per group.  Returns
// false for a key of another group.
template <int KW, int BMODE>
__device__ __forceinline__ bool bin_of(const Key<KW>& key, int log2_parts, u32 n_parts, u32 n_owners,
                                       int pass_log2, u32 pass_val, u32& bin) {
  const u64 h = hash_key(key);
  if (BMODE == 1) {
    if (pass_log2 && part_of(h, pass_log2) != pass_val) return false;
    bin = owner_of(h, n_parts);
    return true;
  }
  const u32 tot = part_of(h, pass_log2 + log2_parts);
  if (pass_log2 && (tot >> log2_parts) != pass_val) return false;
  const u32 local = tot & ((1u << log2_parts) - 1u);
  bin = BMODE == 0 ? local : owner_of(h, n_owners) * (n_parts / n_owners) + local;
  return true;
}

// windows handled per thread between two flushes of BinStage (k_bin_keys)
constexpr int BIN_WPR = 16;

// K2p / K6, stream form.  A CTA stages the keys of a whole ROUND (threads x wpr
// windows: 8192 keys) in per-bin shared-memory queues and writes them out as one
// contiguous run per bin.  Against the first version (4096-key rounds, one warp per
// bin in the flush): runs twice as long — 64 keys = 512 bytes at 128 bins, which is
// what the NVLink peer route needs — and a flush whose cost does not grow with the
// number of bins: with short queues the copy-out is FLAT (thread i takes queue slot i
// of all bins, ~8 instructions per slot) instead of a loop per bin whose fixed cost was
// paid for 16 keys.  Few bins (< 32) are spread over `rep` virtual bins each so that
// the lanes of a warp do not serialise on a handful of shared counters.  The window
// loop uses WindowChunks (fixed-distance funnel shifts).
struct BinPlan {
  int threads, wpr, qcap, rep_log2, flat;
  size_t smem;
};

// shared-memory accesses by 32-bit shared-window address (the generic-pointer forms made
// ptxas rebuild the window base — S2R SR_CgaCtaId, MOV, LEA — in front of every store)
__device__ __forceinline__ u32 smem_inc(u32 addr) {
  u32 old;
  asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(addr) : "memory");
  return old;
}
__device__ __forceinline__ void smem_st_key(u32 addr, const Key<1>& key) {
  asm volatile("st.shared.u64 [%0], %1;" ::"r"(addr), "l"(key.lo) : "memory");
}
__device__ __forceinline__ void smem_st_key(u32 addr, const Key<2>& key) {
  asm volatile("st.shared.v2.u64 [%0], {%1, %2};" ::"r"(addr), "l"(key.lo), "l"(key.hi) : "memory");
}

// PASS: pass_log2 > 0 (multi-pass count); REP: fewer than 32 bins, spread over virtual bins
template <int KW, int BMODE, bool PASS, bool REP>
__global__ void __launch_bounds__(1024, 1) k_bin_stream(StreamView s, int k, int log2_parts, u32 n_parts,
                                                        u32 n_owners, int pass_log2_arg, u32 pass_val, int wpr,
                                                        int qcap, u32 qinv, int rep_log2_arg, int flat,
                                                        BinDest dst, u64* cursors, u64* overflow, u64* stats) {
  extern __shared__ __align__(16) unsigned char bin_smem[];
  const int pass_log2 = PASS ? pass_log2_arg : 0;
  const int rep_log2 = REP ? rep_log2_arg : 0;
  const u32 nv = n_parts << rep_log2;   // virtual bins
  // layout: queue[nv][qcap] keys | gptr[nv] u64* | ceff[nv] u32 | cnt[2][nv] u32
  u64* queue = reinterpret_cast<u64*>(bin_smem);
  u64** gptr = reinterpret_cast<u64**>(queue + (size_t)nv * qcap * KW);
  u32* ceff = reinterpret_cast<u32*>(gptr + nv);
  u32* cnt0 = ceff + nv;
  const u32 tid = threadIdx.x, nthr = blockDim.x;
  const u32 lane_rep = REP ? (tid & ((1u << rep_log2) - 1u)) : 0u;
  const u32 queue_s = (u32)__cvta_generic_to_shared(queue);
  const u32 cnt0_s = (u32)__cvta_generic_to_shared(cnt0);
  for (u32 i = tid; i < 2 * nv; i += nthr) cnt0[i] = 0;
  __syncthreads();
  u32 round = 0;
  u32 windows = 0;
  const u64 stride = (u64)gridDim.x * nthr;
  const u64 n_iter = (s.w_end - s.w_begin + stride - 1) / stride;
  const u64 w0 = s.w_begin + (u64)blockIdx.x * nthr + tid;
  for (u64 itn = 0; itn < n_iter; ++itn) {
    const u64 w = w0 + itn * stride;
    WindowChunks<KW> it(s, w, k);  // loads beyond n_words read as invalid
    const bool in_range = w < s.w_end;  // window starts of later words belong to another launch
#pragma unroll 1
    for (int c = 0; c < 8; ++c) {
      u32* cnt = cnt0 + (round & 1u) * nv;
      const u32 cnt_s = cnt0_s + (round & 1u) * nv * 4u;
      u32 pushed = 0;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const Key<KW> key = it.key(u);
        u32 p;
        if (in_range && it.ok(u) && bin_of<KW, BMODE>(key, log2_parts, n_parts, n_owners, pass_log2, pass_val, p)) {
          pushed |= 1u << u;
          const u32 v = REP ? ((p << rep_log2) | lane_rep) : p;
          const u32 o = smem_inc(cnt_s + v * 4u);
          if (o < (u32)qcap) {
            smem_st_key(queue_s + (v * (u32)qcap + o) * (8u * KW), key);
          } else {  // queue full (a skewed round): append directly
            const u64 g = atomicAdd(cursors + p, 1ull);
            if (g < dst.cap) st_key<KW>(dst.of(p, KW), g, key);
            else atomicOr(overflow, 1ull);
          }
        }
      }
      windows += __popc(pushed);
      it.template next<4>();
      if ((((c + 1) * 4) % wpr) != 0) continue;
      // ---- flush: reserve one run per (virtual) bin, copy out, next round
      __syncthreads();
      u32* cnt_next = cnt0 + ((round + 1) & 1u) * nv;
      for (u32 v = tid; v < nv; v += nthr) {
        u32 cc = cnt[v];
        if (cc > (u32)qcap) cc = qcap;
        const u32 p = v >> rep_log2;
        u64 g0 = 0;
        if (cc) {
          g0 = atomicAdd(cursors + p, (u64)cc);
          if (g0 + cc > dst.cap) {
            atomicOr(overflow, 1ull);
            cc = g0 < dst.cap ? (u32)(dst.cap - g0) : 0u;
          }
        }
        ceff[v] = cc;
        gptr[v] = dst.of(p, KW) + g0 * KW;
        cnt_next[v] = 0;
      }
      __syncthreads();
      if (flat) {
        const u32 n_slots = nv * (u32)qcap;
        for (u32 i = tid; i < n_slots; i += nthr) {
          const u32 v = __umulhi(i, qinv);
          const u32 o = i - v * (u32)qcap;
          if (o < ceff[v]) {
            const u64* q = queue + (size_t)i * KW;
            u64* out = gptr[v] + (size_t)o * KW;
            if (KW == 1) out[0] = q[0];
            else *reinterpret_cast<ulonglong2*>(out) = *reinterpret_cast<const ulonglong2*>(q);
          }
        }
      } else {
        const u32 lane = tid & 31u, n_warps = nthr >> 5;
        for (u32 v = tid >> 5; v < nv; v += n_warps) {
          const u32 cc = ceff[v];
          u64* out = gptr[v];
          const u64* q = queue + (size_t)v * qcap * KW;
          for (u32 o = lane; o < cc; o += 32) {
            if (KW == 1) out[o] = q[o];
            else reinterpret_cast<ulonglong2*>(out)[o] = reinterpret_cast<const ulonglong2*>(q)[o];
          }
        }
      }
      __syncthreads();
      ++round;
    }
  }
  if (stats) {
    for (int o = 16; o; o >>= 1) windows += __shfl_xor_sync(0xffffffffu, windows, o);
    if ((threadIdx.x & 31) == 0 && windows) atomicAdd(stats + KDF_STAT_WINDOWS, (u64)windows);
  }
}

template <int KW, int BMODE>
__global__ void __launch_bounds__(BIN_THREADS) k_bin_keys(const u64* lo, const u64* hi, u64 n,
                                                          int log2_parts, u32 n_parts, int pass_log2,
                                                          u32 pass_val, int qcap, BinDest dst, u64* cursors,
                                                          u64* overflow) {
  extern __shared__ __align__(16) unsigned char bin_smem[];
  BinStage<KW> stage;
  stage.init(bin_smem, (int)n_parts, qcap);
  __syncthreads();
  u64 per_round = (u64)gridDim.x * blockDim.x * BIN_WPR;
  u64 n_rounds = (n + per_round - 1) / per_round;
  for (u64 r = 0; r < n_rounds; ++r) {
    u64 base = r * per_round + (u64)blockIdx.x * blockDim.x * BIN_WPR + threadIdx.x;
#pragma unroll 4
    for (int j = 0; j < BIN_WPR; ++j) {
      u64 i = base + (u64)j * blockDim.x;
      bool ok = i < n;
      Key<KW> key = ld_key_stream<KW>(lo, hi, ok ? i : 0);
      u32 p = 0;
      // log2_parts = ALL hash-range bits (pass groups on top), n_parts local bins
      ok = ok && bin_of<KW, BMODE>(key, log2_parts - pass_log2, n_parts, 1, pass_log2, pass_val, p);
      stage.push(ok, p, key, dst, cursors, overflow);
    }
    stage.flush(dst, cursors, overflow);
  }
}

// ------------------------------------------------------------- K7 ---------
// One thread per hit: the keys of its covered reference bases into keys[h * k ..),
// unused slots padded with all ones (they sort last).  Hits are sorted by (read, offset).
__global__ void __launch_bounds__(128) k_hit_coverage(const u32* hit_read, const u32* hit_off, u64 n_hits,
                                                      int k, const int* read_contig,
                                                      const long long* read_ref_start,
                                                      const u64* read_cig_off, const u32* cigar, u64* keys) {
  u64 h = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= n_hits) return;
  const u32 r = hit_read[h];
  const u32 off = hit_off[h];
  const u32 prev_end = (h > 0 && hit_read[h - 1] == r) ? hit_off[h - 1] + (u32)k : 0u;
  u64* out = keys + h * (u64)k;
  const u64 c0 = read_cig_off[r], c1 = read_cig_off[r + 1];
  int n = expand_hit(cigar + c0, c1 - c0, read_ref_start[r], off, k, prev_end, (u64)(u32)read_contig[r], out);
  for (int j = n; j < k; ++j) out[j] = ~0ull;
}

// ------------------------------------------------ table maintenance -------
// validity bitmap from its sparse form: all ones inside n_bases, listed positions cleared
__global__ void __launch_bounds__(256) k_valid_fill(u32* valid, u64 n_words, u64 n_bases) {
  u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += stride) {
    u32 v = 0xffffffffu;
    if (w == n_words - 1 && (n_bases & 31)) v <<= 32 - (u32)(n_bases & 31);
    valid[w] = v;
  }
}
__global__ void __launch_bounds__(256) k_valid_clear(u32* valid, u64 n_bases, const u32* pos, u64 n) {
  u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    u32 p = pos[i];
    if (p < n_bases) atomicAnd(valid + (p >> 5), ~(0x80000000u >> (p & 31)));
  }
}

__global__ void k_or_flag(const u64* src, u64* flags, u64 bit) {
  if (*src) atomicOr(flags, bit);
}
__global__ void __launch_bounds__(256) k_fill_u64x2(ulonglong2* p, u64 n_units, u64 value) {
  u64 stride = (u64)gridDim.x * blockDim.x;
  ulonglong2 v = make_ulonglong2(value, value);
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_units; i += stride) p[i] = v;
}

// ---------------------------------------------- random-access roofline ----
// mode 0: 32-byte-sector reads; 1: read + RED.ADD.32 on the sector (the two figures
// bench.py reports).  Modes >= 2 time other read-modify-write flavours on the same
// access pattern (profiles/r1d_atomics.txt); every CAS below succeeds unless raced:
// 2 ATOM.ADD.32 (result used), 3 CAS.32, 4 CAS.64, 5 RED.ADD.64, 6 RED.AND.64,
// 7 ATOM.EXCH.64, 8 RED.ADD.32 without the read, 10 one 256-bit load per op,
// 11 256-bit load + CAS.64 on one op in six (the mix of a packed count).
__global__ void __launch_bounds__(256) k_bench_random(u32* buf, u64 n_sectors, u64 n_ops, int atomic,
                                                      u64* sink) {
  constexpr int CHUNK = 8;
  u64 stride = (u64)gridDim.x * blockDim.x;
  u64 acc = 0;
  for (u64 i0 = (u64)blockIdx.x * blockDim.x + threadIdx.x; i0 < n_ops; i0 += stride * CHUNK) {
    u64 v[CHUNK];
    u64 sidx[CHUNK];
#pragma unroll
    for (int u = 0; u < CHUNK; ++u) {
      u64 i = i0 + (u64)u * stride;
      sidx[u] = mulhi64(mix64(i + 0x1234567ull), n_sectors);
      v[u] = 0;
      if (i < n_ops && atomic != 8) {
        if (atomic == 10 || atomic == 11) {
          u64 a, b, c, d;
          ld256(reinterpret_cast<const u64*>(buf + sidx[u] * 8), a, b, c, d);
          v[u] = b;
          acc += a ^ c ^ d;
        } else if (atomic == 4 || atomic == 5 || atomic == 6 || atomic == 7) {
          v[u] = __ldcg(reinterpret_cast<const u64*>(buf + sidx[u] * 8 + 2));
        } else {
          v[u] = __ldcg(buf + sidx[u] * 8 + 2);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < CHUNK; ++u) {
      u64 i = i0 + (u64)u * stride;
      acc += v[u];
      if (i >= n_ops) continue;
      u32* p32 = buf + sidx[u] * 8 + 2;
      u64* p64 = reinterpret_cast<u64*>(buf + sidx[u] * 8 + 2);
      switch (atomic) {
        case 1: case 8: atomicAdd(p32, 1u); break;
        case 2: acc += atomicAdd(p32, 1u); break;
        case 3: acc += atomicCAS(p32, (u32)v[u], (u32)v[u] + 1u); break;
        case 4: acc += atomicCAS(p64, v[u], v[u] + 1ull); break;
        case 5: atomicAdd(p64, 1ull); break;
        case 6: atomicAnd(p64, ~(u64)(i & 1)); break;
        case 7: acc += atomicExch(p64, i); break;
        case 11: if ((sidx[u] % 6) == 0) acc += atomicCAS(p64, v[u], v[u] + 1ull); break;
        default: break;
      }
    }
  }
  if (acc == 0xdeadbeefdeadbeefull) *sink = acc;
}

// ------------------------------------------------------------ host side ---
static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}
static int grid_for(const void* func, int block, size_t smem, u64 work_items, int sm_count) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, func, block, smem) != cudaSuccess ||
      per_sm < 1)
    per_sm = 1;
  u64 full = (u64)sm_count * per_sm;
  u64 need = (work_items + block - 1) / block;
  if (need < 1) need = 1;
  return (int)(need < full ? need : full);
}

static int current_sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
    return 148;
  return n;
}

static StreamView view_of(const kdf_stream* s) {
  StreamView v;
  v.codes = (const u64*)s->codes;
  v.valid = s->valid;
  v.n_bases = s->n_bases;
  v.n_words = (s->n_bases + 31) / 32;
  v.w_begin = 0;
  v.w_end = v.n_words;
  return v;
}

// shared-memory budget for a read-only table copy (keys only)
static const size_t SMEM_TABLE_MAX = 160 * 1024;

template <int KW, int OP, bool SMEM, int CHUNK, bool FILT = false>
static int launch_stream(const kdf_table* t, const StreamView& v, int plane, u32 arg, u64* stats,
                         const HitSink& sink, cudaStream_t st) {
  TableView<KW> tv = view_of_table<KW>(t);
  const void* fn = (const void*)k_stream<KW, OP, SMEM, CHUNK, FILT>;
  size_t smem = SMEM ? (size_t)tv.n_buckets * 32 * KW + (size_t)PF_WORDS * 4 : 0;
  int block = SMEM ? 512 : 256;
  if (SMEM) CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int g = grid_for(fn, block, smem, v.n_words, t->sm_count);
  k_stream<KW, OP, SMEM, CHUNK, FILT><<<g, block, smem, st>>>(tv, v, t->k, plane, arg, stats, sink);
  CUDA_TRY(cudaGetLastError());
  return KDF_OK;
}

template <int KW>
static int dispatch_stream(const kdf_table* t, const StreamView& v, int op, int plane, u32 arg,
                           u64* stats, const HitSink& sink, cudaStream_t st) {
  bool small = (size_t)(t->capacity / SPB<KW>::v) * 32 * KW <= SMEM_TABLE_MAX && t->log2_parts == 0;
  bool filt = t->filter != nullptr;
  switch (op) {
    case OP_INSERT_COUNT:
      return launch_stream<KW, OP_INSERT_COUNT, false, KDF_STREAM_CHUNK>(t, v, plane, arg, stats, sink, st);
    case OP_INSERT_ONLY:
      return launch_stream<KW, OP_INSERT_ONLY, false, KDF_STREAM_CHUNK>(t, v, plane, arg, stats, sink, st);
    case OP_COUNT_IF_PRESENT:
      if (small) return launch_stream<KW, OP_COUNT_IF_PRESENT, true, KDF_SMEM_CHUNK>(t, v, plane, arg, stats, sink, st);
      if (filt) return launch_stream<KW, OP_COUNT_IF_PRESENT, false, KDF_FILT_CHUNK, true>(t, v, plane, arg, stats, sink, st);
      return launch_stream<KW, OP_COUNT_IF_PRESENT, false, KDF_STREAM_CHUNK>(t, v, plane, arg, stats, sink, st);
    case OP_MARK_IF_PRESENT:
      if (small) return launch_stream<KW, OP_MARK_IF_PRESENT, true, KDF_SMEM_CHUNK>(t, v, plane, arg, stats, sink, st);
      if (filt) return launch_stream<KW, OP_MARK_IF_PRESENT, false, KDF_FILT_CHUNK, true>(t, v, plane, arg, stats, sink, st);
      return launch_stream<KW, OP_MARK_IF_PRESENT, false, KDF_STREAM_CHUNK>(t, v, plane, arg, stats, sink, st);
    case OP_EMIT_HITS:
      if (small) return launch_stream<KW, OP_EMIT_HITS, true, KDF_SMEM_CHUNK>(t, v, plane, arg, stats, sink, st);
      if (filt) return launch_stream<KW, OP_EMIT_HITS, false, KDF_FILT_CHUNK, true>(t, v, plane, arg, stats, sink, st);
      return launch_stream<KW, OP_EMIT_HITS, false, KDF_STREAM_CHUNK>(t, v, plane, arg, stats, sink, st);
    default:
      return fail(KDF_ERR_ARG, "unknown table operation");
  }
}

template <int KW, int OP>
static int launch_update_keys(const kdf_table* t, const u64* lo, const u64* hi, u64 n_max,
                              const u64* n_dev, int plane, u32 arg, u64* stats, cudaStream_t st,
                              int filt_log2 = 0, u32 filt_val = 0) {
  TableView<KW> tv = view_of_table<KW>(t);
  if (filt_log2 > 0) {
    int g = grid_for((const void*)k_update_keys<KW, OP, true>, 256, 0, (n_max + KDF_KEYS_CHUNK - 1) / KDF_KEYS_CHUNK, t->sm_count);
    k_update_keys<KW, OP, true><<<g, 256, 0, st>>>(tv, lo, hi, n_max, n_dev, plane, arg, stats, filt_log2,
                                                   filt_val);
  } else {
    int g = grid_for((const void*)k_update_keys<KW, OP, false>, 256, 0, (n_max + KDF_KEYS_CHUNK - 1) / KDF_KEYS_CHUNK, t->sm_count);
    k_update_keys<KW, OP, false><<<g, 256, 0, st>>>(tv, lo, hi, n_max, n_dev, plane, arg, stats, 0, 0);
  }
  CUDA_TRY(cudaGetLastError());
  return KDF_OK;
}

template <int KW>
static int dispatch_update_keys(const kdf_table* t, const u64* lo, const u64* hi, u64 n_max,
                                const u64* n_dev, int op, int plane, u32 arg, u64* stats,
                                cudaStream_t st) {
  switch (op) {
    case OP_INSERT_COUNT:
      return launch_update_keys<KW, OP_INSERT_COUNT>(t, lo, hi, n_max, n_dev, plane, arg, stats, st);
    case OP_INSERT_ONLY:
      return launch_update_keys<KW, OP_INSERT_ONLY>(t, lo, hi, n_max, n_dev, plane, arg, stats, st);
    case OP_COUNT_IF_PRESENT:
      return launch_update_keys<KW, OP_COUNT_IF_PRESENT>(t, lo, hi, n_max, n_dev, plane, arg, stats, st);
    case OP_MARK_IF_PRESENT:
      return launch_update_keys<KW, OP_MARK_IF_PRESENT>(t, lo, hi, n_max, n_dev, plane, arg, stats, st);
    default:
      return fail(KDF_ERR_ARG, "unknown table operation");
  }
}

static int clear_table_async(const kdf_table* t, cudaStream_t st) {
  // keys -> all ones, planes -> 0; both regions are multiples of 16 bytes
  u64 key_units = t->capacity * 8ull * t->key_words / 16;
  u64 plane_units = t->capacity * 8ull / 16;
  ulonglong2* kp = (ulonglong2*)t->base;
  ulonglong2* pp = kp + key_units;
  int g = grid_for((const void*)k_fill_u64x2, 256, 0, key_units, t->sm_count);
  k_fill_u64x2<<<g, 256, 0, st>>>(kp, key_units, EMPTY);
  g = grid_for((const void*)k_fill_u64x2, 256, 0, plane_units, t->sm_count);
  k_fill_u64x2<<<g, 256, 0, st>>>(pp, plane_units, 0ull);
  CUDA_TRY(cudaGetLastError());
  return KDF_OK;
}

template <int KW>
static int launch_threshold(const kdf_table* t, u32 min0, u32 max0, u32 min1, u32 max1, u64* out_lo,
                            u64* out_hi, u32* out_p0, u32* out_p1, u64 cap, u64* n_out,
                            u32 count_min0, u64* n_count, u64* n_occ, cudaStream_t st) {
  TableView<KW> tv = view_of_table<KW>(t);
  int g = grid_for((const void*)k_threshold_compact<KW>, 256, 0, t->capacity, t->sm_count);
  k_threshold_compact<KW><<<g, 256, 0, st>>>(tv, t->capacity, min0, max0, min1, max1, out_lo, out_hi,
                                             out_p0, out_p1, cap, n_out, count_min0, n_count, n_occ);
  CUDA_TRY(cudaGetLastError());
  return KDF_OK;
}

template <int KW>
static int launch_emit_buckets(const kdf_table* t, u32 min0, u32 max0, u32 min1, u32 max1, u64* out_lo,
                               u64* out_hi, u32* out_p0, u32* out_p1, u64 cap, u64* n_out,
                               u32 count_min0, u64* n_count, u64* n_occ, cudaStream_t st) {
  TableView<KW> tv = view_of_table<KW>(t);
  int g = grid_for((const void*)k_emit_buckets<KW, true>, 256, 0, tv.n_buckets, t->sm_count);
  k_emit_buckets<KW, true><<<g, 256, 0, st>>>(tv, min0, max0, min1, max1, out_lo, out_hi, out_p0, out_p1,
                                              cap, n_out, count_min0, n_count, n_occ);
  CUDA_TRY(cudaGetLastError());
  return KDF_OK;
}

template <int KW, int OP>
static int launch_packed_keys(const kdf_table* t, const u64* lo, u64 n_max, const u64* n_dev, int sh,
                              u32 sat, u64* stats, cudaStream_t st, int filt_log2, u32 filt_val,
                              int n_src = 1, u64 src_stride = 0, u64 cur_stride = 0) {
  TableView<KW> tv = view_of_table<KW>(t);
  const size_t smem = sizeof(PackedKeysQueue<KW, OP>) * (256 / 32);
  const u64 items = (n_max + KDF_KEYS_CHUNK - 1) / KDF_KEYS_CHUNK;
  for (int s0 = 0; s0 < n_src; s0 += KEY_SEGS) {
    KeySegs segs;
    segs.n = n_src - s0 < KEY_SEGS ? n_src - s0 : KEY_SEGS;
    for (int i = 0; i < KEY_SEGS; ++i) {
      segs.lo[i] = i < segs.n ? lo + (u64)(s0 + i) * src_stride : nullptr;
      segs.n_dev[i] = (i < segs.n && n_dev) ? n_dev + (u64)(s0 + i) * cur_stride : nullptr;
    }
    if (filt_log2 > 0) {
      const void* fn = (const void*)k_packed_keys<KW, OP, true>;
      CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int g = grid_for(fn, 256, smem, items, t->sm_count);
      k_packed_keys<KW, OP, true><<<g, 256, smem, st>>>(tv, segs, n_max, sh, sat, stats, filt_log2, filt_val);
    } else {
      const void* fn = (const void*)k_packed_keys<KW, OP, false>;
      CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int g = grid_for(fn, 256, smem, items, t->sm_count);
      k_packed_keys<KW, OP, false><<<g, 256, smem, st>>>(tv, segs, n_max, sh, sat, stats, 0, 0);
    }
    CUDA_TRY(cudaGetLastError());
  }
  return KDF_OK;
}

template <int KW>
static int launch_emit_packed(const kdf_table* t, int sh, u32 sat, int ignore_ref, int count_all,
                              u64* out_lo, u64* out_hi, u64 cap, u64* n_out, u64* n_count, u64* n_occ,
                              cudaStream_t st) {
  TableView<KW> tv = view_of_table<KW>(t);
  int g = grid_for((const void*)k_emit_packed<KW>, 256, 0, (tv.n_buckets + 3) / 4, t->sm_count);
  k_emit_packed<KW><<<g, 256, 0, st>>>(tv, sh, sat, ignore_ref, count_all, out_lo, out_hi, cap, n_out,
                                       n_count, n_occ);
  CUDA_TRY(cudaGetLastError());
  return KDF_OK;
}

static int check_table_args(int k, uint64_t capacity, const void* slots, const char* who) {
  int kw = kdf_key_words(k);
  if (!kw) return fail(KDF_ERR_ARG, std::string(who) + ": k must be in 1..64");
  if (capacity < 4 || (capacity & 3))
    return fail(KDF_ERR_ARG, std::string(who) + ": capacity must be a multiple of 4, >= 4");
  if (capacity / 4 > 0xffffffffull)
    return fail(KDF_ERR_ARG, std::string(who) + ": capacity exceeds 2^32 buckets");
  if (((uintptr_t)slots & 31) != 0)
    return fail(KDF_ERR_ARG, std::string(who) + ": table memory must be 32-byte aligned");
  return KDF_OK;
}

extern "C" {

int kdf_version(void) { return KDF_VERSION; }
const char* kdf_last_error(void) { return g_err.c_str(); }

int kdf_device_info(int device, kdf_device_props* out) {
  if (!out) return fail(KDF_ERR_ARG, "kdf_device_info: out is NULL");
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0)
    return fail(KDF_ERR_NO_DEVICE, "no CUDA device visible");
  if (device < 0 || device >= n) return fail(KDF_ERR_ARG, "kdf_device_info: bad device index");
  cudaDeviceProp p;
  CUDA_TRY(cudaGetDeviceProperties(&p, device));
  out->sm_count = p.multiProcessorCount;
  out->cc_major = p.major;
  out->cc_minor = p.minor;
  out->l2_bytes = p.l2CacheSize;
  out->hbm_bytes = p.totalGlobalMem;
  strncpy(out->name, p.name, sizeof(out->name) - 1);
  out->name[sizeof(out->name) - 1] = 0;
  return KDF_OK;
}

int kdf_key_words(int k) {
  if (k >= 1 && k <= 32) return 1;
  if (k >= 33 && k <= 64) return 2;
  return 0;
}

size_t kdf_table_bytes(uint64_t capacity, int key_words) {
  if (key_words != 1 && key_words != 2) return 0;
  return (size_t)capacity * (8 * (size_t)key_words + 8);
}

uint64_t kdf_table_capacity_for(uint64_t n_keys) {
  uint64_t c = n_keys * 2;
  if (c < 1024) c = 1024;
  return (c + 7) & ~7ull;  // whole buckets for either key width
}

int kdf_table_clear(kdf_table* t, void* stream) {
  if (!t) return fail(KDF_ERR_ARG, "kdf_table_clear: table is NULL");
  t->filter = nullptr;   // a filter describes the keys it was built from
  t->filter_mask = 0;
  return clear_table_async(t, (cudaStream_t)stream);
}

int kdf_table_create(kdf_table** out, int k, uint64_t capacity, void* slots, void* stream) {
  if (!out || !slots) return fail(KDF_ERR_ARG, "kdf_table_create: NULL argument");
  int rc = check_table_args(k, capacity, slots, "kdf_table_create");
  if (rc != KDF_OK) return rc;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0)
    return fail(KDF_ERR_NO_DEVICE, "no CUDA device visible");
  kdf_table* t = new kdf_table;
  t->k = k;
  t->key_words = kdf_key_words(k);
  t->capacity = capacity;
  t->base = slots;
  t->sm_count = current_sm_count();
  t->log2_parts = 0;
  rc = kdf_table_clear(t, stream);
  if (rc != KDF_OK) {
    delete t;
    return rc;
  }
  *out = t;
  return KDF_OK;
}

int kdf_table_build_filter(kdf_table* t, uint32_t* words, uint64_t n_words, void* stream) {
  if (!t) return fail(KDF_ERR_ARG, "kdf_table_build_filter: table is NULL");
  if (!words || n_words == 0) {   // detach
    t->filter = nullptr;
    t->filter_mask = 0;
    return KDF_OK;
  }
  if ((n_words & (n_words - 1)) || n_words > (1ull << 30))
    return fail(KDF_ERR_ARG, "kdf_table_build_filter: n_words must be a power of two <= 2^30");
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaMemsetAsync(words, 0, n_words * 4, st));
  if (t->key_words == 1) {
    int g = grid_for((const void*)k_build_filter<1>, 256, 0, t->capacity, t->sm_count);
    k_build_filter<1><<<g, 256, 0, st>>>(view_of_table<1>(t), t->capacity, words, (u32)(n_words - 1));
  } else {
    int g = grid_for((const void*)k_build_filter<2>, 256, 0, t->capacity, t->sm_count);
    k_build_filter<2><<<g, 256, 0, st>>>(view_of_table<2>(t), t->capacity, words, (u32)(n_words - 1));
  }
  CUDA_TRY(cudaGetLastError());
  t->filter = words;
  t->filter_mask = (u32)(n_words - 1);
  return KDF_OK;
}

int kdf_table_destroy(kdf_table* t) {
  delete t;
  return KDF_OK;
}

int kdf_table_info(const kdf_table* t, int* k, int* key_words, uint64_t* capacity) {
  if (!t) return fail(KDF_ERR_ARG, "kdf_table_info: table is NULL");
  if (k) *k = t->k;
  if (key_words) *key_words = t->key_words;
  if (capacity) *capacity = t->capacity;
  return KDF_OK;
}

int kdf_table_clear_plane(kdf_table* t, int plane, void* stream) {
  if (!t || (plane != 0 && plane != 1)) return fail(KDF_ERR_ARG, "kdf_table_clear_plane: bad argument");
  char* p = (char*)t->base + t->capacity * 8ull * t->key_words + (plane ? t->capacity * 4ull : 0);
  CUDA_TRY(cudaMemsetAsync(p, 0, t->capacity * 4ull, (cudaStream_t)stream));
  return KDF_OK;
}

int kdf_extract_canonical(const kdf_stream* s, int k, uint64_t* out_lo, uint64_t* out_hi,
                          uint32_t* out_ok, void* stream) {
  if (!s || !out_lo || !out_ok) return fail(KDF_ERR_ARG, "kdf_extract_canonical: NULL argument");
  int kw = kdf_key_words(k);
  if (!kw) return fail(KDF_ERR_ARG, "kdf_extract_canonical: k must be in 1..64");
  if (kw == 2 && !out_hi) return fail(KDF_ERR_ARG, "kdf_extract_canonical: out_hi required for k > 32");
  StreamView v = view_of(s);
  if (v.n_words == 0) return KDF_OK;
  int sm = current_sm_count();
  cudaStream_t st = (cudaStream_t)stream;
  if (kw == 1) {
    int g = grid_for((const void*)k_extract<1>, 256, 0, v.n_words, sm);
    k_extract<1><<<g, 256, 0, st>>>(v, k, (u64*)out_lo, (u64*)out_hi, out_ok);
  } else {
    int g = grid_for((const void*)k_extract<2>, 256, 0, v.n_words, sm);
    k_extract<2><<<g, 256, 0, st>>>(v, k, (u64*)out_lo, (u64*)out_hi, out_ok);
  }
  CUDA_TRY(cudaGetLastError());
  return KDF_OK;
}

int kdf_count_stream(kdf_table* t, const kdf_stream* s, int mode, int plane, uint32_t arg,
                     uint64_t* stats, void* stream) {
  if (!t || !s) return fail(KDF_ERR_ARG, "kdf_count_stream: NULL argument");
  if (plane != 0 && plane != 1) return fail(KDF_ERR_ARG, "kdf_count_stream: plane must be 0 or 1");
  if (mode < 0 || mode > KDF_MODE_MARK_IF_PRESENT) return fail(KDF_ERR_ARG, "kdf_count_stream: unknown mode");
  StreamView v = view_of(s);
  if (v.n_words == 0) return KDF_OK;
  if (mode == KDF_MODE_INSERT_COUNT || mode == KDF_MODE_INSERT_ONLY) {
    t->filter = nullptr;   // new keys: a filter built earlier no longer covers the table
    t->filter_mask = 0;
  }
  HitSink sink = {nullptr, nullptr, 0, nullptr};
  if (t->key_words == 1)
    return dispatch_stream<1>(t, v, mode, plane, arg, (u64*)stats, sink, (cudaStream_t)stream);
  return dispatch_stream<2>(t, v, mode, plane, arg, (u64*)stats, sink, (cudaStream_t)stream);
}

int kdf_scan_stream_hits(const kdf_table* t, const kdf_stream* s, uint64_t* hit_pos,
                         uint32_t* hit_slot, uint64_t hit_cap, uint64_t* n_hits, uint64_t* stats,
                         void* stream) {
  if (!t || !s || !n_hits) return fail(KDF_ERR_ARG, "kdf_scan_stream_hits: NULL argument");
  if (hit_cap && (!hit_pos || !hit_slot)) return fail(KDF_ERR_ARG, "kdf_scan_stream_hits: NULL hit arrays");
  if (t->capacity > 0xffffffffull)
    return fail(KDF_ERR_ARG, "kdf_scan_stream_hits: table capacity must fit 32-bit slot indices");
  StreamView v = view_of(s);
  if (v.n_words == 0) return KDF_OK;
  HitSink sink = {(u64*)hit_pos, hit_slot, hit_cap, (u64*)n_hits};
  if (t->key_words == 1)
    return dispatch_stream<1>(t, v, OP_EMIT_HITS, 0, 0, (u64*)stats, sink, (cudaStream_t)stream);
  return dispatch_stream<2>(t, v, OP_EMIT_HITS, 0, 0, (u64*)stats, sink, (cudaStream_t)stream);
}

int kdf_update_keys(kdf_table* t, const uint64_t* lo, const uint64_t* hi, uint64_t n, int mode,
                    int plane, uint32_t arg, uint64_t* stats, void* stream) {
  if (!t) return fail(KDF_ERR_ARG, "kdf_update_keys: table is NULL");
  if (n == 0) return KDF_OK;
  if (!lo || (t->key_words == 2 && !hi)) return fail(KDF_ERR_ARG, "kdf_update_keys: NULL key array");
  if (plane != 0 && plane != 1) return fail(KDF_ERR_ARG, "kdf_update_keys: plane must be 0 or 1");
  if (mode < 0 || mode > KDF_MODE_MARK_IF_PRESENT) return fail(KDF_ERR_ARG, "kdf_update_keys: unknown mode");
  if (mode == KDF_MODE_INSERT_COUNT || mode == KDF_MODE_INSERT_ONLY) {
    t->filter = nullptr;
    t->filter_mask = 0;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (t->key_words == 1)
    return dispatch_update_keys<1>(t, (const u64*)lo, (const u64*)hi, n, nullptr, mode, plane, arg,
                                   (u64*)stats, st);
  return dispatch_update_keys<2>(t, (const u64*)lo, (const u64*)hi, n, nullptr, mode, plane, arg,
                                 (u64*)stats, st);
}

int kdf_update_bins(kdf_table* t, int n_parts, const uint64_t* bins, uint64_t bin_cap,
                    const uint64_t* cursors, int mode, int plane, uint32_t arg, uint64_t* stats,
                    void* stream) {
  if (!t || !bins || !cursors) return fail(KDF_ERR_ARG, "kdf_update_bins: NULL argument");
  if (n_parts < 1 || n_parts > BIN_MAX_PARTS) return fail(KDF_ERR_ARG, "kdf_update_bins: n_parts must be 1..512");
  if (plane != 0 && plane != 1) return fail(KDF_ERR_ARG, "kdf_update_bins: plane must be 0 or 1");
  if (mode < 0 || mode > KDF_MODE_MARK_IF_PRESENT) return fail(KDF_ERR_ARG, "kdf_update_bins: unknown mode");
  if (bin_cap == 0) return KDF_OK;
  if (mode == KDF_MODE_INSERT_COUNT || mode == KDF_MODE_INSERT_ONLY) {
    t->filter = nullptr;
    t->filter_mask = 0;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int kw = t->key_words;
  // one launch per bin, in hash order: the bucket index grows with the hash, so the
  // keys of bin p only touch the p-th n_parts-th of the table, which stays in L2
  for (int p = 0; p < n_parts; ++p) {
    const u64* b = (const u64*)bins + (u64)p * bin_cap * kw;
    const u64* n_dev = (const u64*)cursors + p;
    int rc;
    if (mode == KDF_MODE_COUNT_IF_PRESENT)   // the probing ops take the software-pipelined kernel
      rc = kw == 1 ? launch_packed_keys<1, OP_COUNT_IF_PRESENT>(t, b, bin_cap, n_dev, plane, arg, (u64*)stats, st, 0, 0)
                   : launch_packed_keys<2, OP_COUNT_IF_PRESENT>(t, b, bin_cap, n_dev, plane, arg, (u64*)stats, st, 0, 0);
    else if (mode == KDF_MODE_MARK_IF_PRESENT)
      rc = kw == 1 ? launch_packed_keys<1, OP_MARK_IF_PRESENT>(t, b, bin_cap, n_dev, plane, arg, (u64*)stats, st, 0, 0)
                   : launch_packed_keys<2, OP_MARK_IF_PRESENT>(t, b, bin_cap, n_dev, plane, arg, (u64*)stats, st, 0, 0);
    else
      rc = kw == 1 ? dispatch_update_keys<1>(t, b, nullptr, bin_cap, n_dev, mode, plane, arg, (u64*)stats, st)
                   : dispatch_update_keys<2>(t, b, nullptr, bin_cap, n_dev, mode, plane, arg, (u64*)stats, st);
    if (rc != KDF_OK) return rc;
  }
  return KDF_OK;
}

int kdf_threshold_compact(const kdf_table* t, uint32_t min0, uint32_t max0, uint32_t min1,
                          uint32_t max1, uint64_t* out_lo, uint64_t* out_hi, uint32_t* out_p0,
                          uint32_t* out_p1, uint64_t cap, uint64_t* n_out, void* stream) {
  if (!t || !n_out) return fail(KDF_ERR_ARG, "kdf_threshold_compact: NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (t->key_words == 1)
    return launch_threshold<1>(t, min0, max0, min1, max1, (u64*)out_lo, (u64*)out_hi, out_p0, out_p1,
                               cap, (u64*)n_out, 0, nullptr, nullptr, st);
  return launch_threshold<2>(t, min0, max0, min1, max1, (u64*)out_lo, (u64*)out_hi, out_p0, out_p1, cap,
                             (u64*)n_out, 0, nullptr, nullptr, st);
}

int kdf_lookup_keys(const kdf_table* t, const uint64_t* lo, const uint64_t* hi, uint64_t n,
                    uint8_t* out_found, uint32_t* out_p0, uint32_t* out_p1, void* stream) {
  if (!t) return fail(KDF_ERR_ARG, "kdf_lookup_keys: table is NULL");
  if (n == 0) return KDF_OK;
  if (!lo || (t->key_words == 2 && !hi)) return fail(KDF_ERR_ARG, "kdf_lookup_keys: NULL key array");
  cudaStream_t st = (cudaStream_t)stream;
  if (t->key_words == 1) {
    TableView<1> tv = view_of_table<1>(t);
    int g = grid_for((const void*)k_lookup_keys<1>, 256, 0, n, t->sm_count);
    k_lookup_keys<1><<<g, 256, 0, st>>>(tv, (const u64*)lo, (const u64*)hi, n, out_found, out_p0, out_p1);
  } else {
    TableView<2> tv = view_of_table<2>(t);
    int g = grid_for((const void*)k_lookup_keys<2>, 256, 0, n, t->sm_count);
    k_lookup_keys<2><<<g, 256, 0, st>>>(tv, (const u64*)lo, (const u64*)hi, n, out_found, out_p0, out_p1);
  }
  CUDA_TRY(cudaGetLastError());
  return KDF_OK;
}

int kdf_add_planes(kdf_table* t, const uint64_t* lo, const uint64_t* hi, uint64_t n,
                   const uint32_t* add0, const uint32_t* add1, uint64_t* n_missing, void* stream) {
  if (!t) return fail(KDF_ERR_ARG, "kdf_add_planes: table is NULL");
  if (n == 0) return KDF_OK;
  if (!lo || (t->key_words == 2 && !hi)) return fail(KDF_ERR_ARG, "kdf_add_planes: NULL key array");
  cudaStream_t st = (cudaStream_t)stream;
  if (t->key_words == 1) {
    TableView<1> tv = view_of_table<1>(t);
    int g = grid_for((const void*)k_add_planes<1>, 256, 0, n, t->sm_count);
    k_add_planes<1><<<g, 256, 0, st>>>(tv, (const u64*)lo, (const u64*)hi, n, add0, add1, (u64*)n_missing);
  } else {
    TableView<2> tv = view_of_table<2>(t);
    int g = grid_for((const void*)k_add_planes<2>, 256, 0, n, t->sm_count);
    k_add_planes<2><<<g, 256, 0, st>>>(tv, (const u64*)lo, (const u64*)hi, n, add0, add1, (u64*)n_missing);
  }
  CUDA_TRY(cudaGetLastError());
  return KDF_OK;
}

int kdf_scan_reads(const kdf_table* t, const kdf_stream* s, const uint64_t* read_starts,
                   const uint32_t* read_lens, uint64_t n_reads, uint32_t min_distinct,
                   uint32_t* out_ndistinct,
                   uint32_t* out_nhits, uint64_t* hit_pos, uint32_t* hit_slot, uint64_t hit_cap,
                   uint64_t* n_hits, uint64_t* stats, void* stream) {
  if (!t || !s) return fail(KDF_ERR_ARG, "kdf_scan_reads: NULL argument");
  if (n_reads == 0) return KDF_OK;
  if (!read_starts || !read_lens || !out_ndistinct || !out_nhits)
    return fail(KDF_ERR_ARG, "kdf_scan_reads: NULL array");
  if (t->capacity > 0xffffffffull)
    return fail(KDF_ERR_ARG, "kdf_scan_reads: table capacity must fit 32-bit slot indices");
  StreamView v = view_of(s);
  cudaStream_t st = (cudaStream_t)stream;
  const int block = SCAN_WARPS * 32;
  if (t->key_words == 1) {
    TableView<1> tv = view_of_table<1>(t);
    int g = grid_for((const void*)k_scan_reads<1>, block, 0, n_reads * 32, t->sm_count);
    k_scan_reads<1><<<g, block, 0, st>>>(tv, v, t->k, (const u64*)read_starts, read_lens, n_reads, min_distinct,
                                         out_ndistinct, out_nhits, (u64*)hit_pos, hit_slot, hit_cap,
                                         (u64*)n_hits, (u64*)stats);
  } else {
    TableView<2> tv = view_of_table<2>(t);
    int g = grid_for((const void*)k_scan_reads<2>, block, 0, n_reads * 32, t->sm_count);
    k_scan_reads<2><<<g, block, 0, st>>>(tv, v, t->k, (const u64*)read_starts, read_lens, n_reads, min_distinct,
                                         out_ndistinct, out_nhits, (u64*)hit_pos, hit_slot, hit_cap,
                                         (u64*)n_hits, (u64*)stats);
  }
  CUDA_TRY(cudaGetLastError());
  return KDF_OK;
}

size_t kdf_reduce_hits_scratch_bytes(uint64_t n_hits) {
  size_t tmp = 0, tmp2 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, (const u64*)nullptr, (u64*)nullptr,
                                  (const u32*)nullptr, (u32*)nullptr, (int)n_hits);
  cub::DeviceRadixSort::SortKeys(nullptr, tmp2, (const u64*)nullptr, (u64*)nullptr, (int)n_hits);
  if (tmp2 > tmp) tmp = tmp2;
  // sorted pos + sorted slot + two (read, slot) key arrays + cub temp, each 256-byte aligned
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  return 3 * al(n_hits * 8) + al(n_hits * 4) + al(tmp) + 256;
}

int kdf_reduce_hits(const uint64_t* hit_pos, const uint32_t* hit_slot, uint64_t n_hits,
                    const uint64_t* read_starts, uint64_t n_reads, void* scratch,
                    size_t scratch_bytes, uint64_t* sorted_pos_out, uint32_t* sorted_slot_out,
                    uint64_t* rec_read, uint32_t* rec_ndistinct, uint32_t* rec_nhits,
                    uint64_t* rec_first, uint64_t* n_recs, void* stream) {
  if (n_hits == 0) return KDF_OK;
  if (!hit_pos || !hit_slot || !read_starts || !n_reads || !scratch || !rec_read || !rec_ndistinct ||
      !rec_nhits || !rec_first || !n_recs)
    return fail(KDF_ERR_ARG, "kdf_reduce_hits: NULL argument");
  if (n_hits > 0x7fffffffull) return fail(KDF_ERR_ARG, "kdf_reduce_hits: too many hits for one call");
  if (n_reads > 0xffffffffull) return fail(KDF_ERR_ARG, "kdf_reduce_hits: more than 2^32 reads in one stream");
  if (scratch_bytes < kdf_reduce_hits_scratch_bytes(n_hits))
    return fail(KDF_ERR_CAPACITY, "kdf_reduce_hits: scratch too small");
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  char* p = (char*)(((uintptr_t)scratch + 255) & ~(uintptr_t)255);
  u64* spos = sorted_pos_out ? (u64*)sorted_pos_out : (u64*)p;
  p += al(n_hits * 8);
  u32* sslot = sorted_slot_out ? sorted_slot_out : (u32*)p;
  p += al(n_hits * 4);
  u64* rkey = (u64*)p;
  p += al(n_hits * 8);
  u64* rkey_sorted = (u64*)p;
  p += al(n_hits * 8);
  size_t tmp = 0, tmp2 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, (const u64*)hit_pos, spos, hit_slot, sslot, (int)n_hits);
  cub::DeviceRadixSort::SortKeys(nullptr, tmp2, (const u64*)rkey, rkey_sorted, (int)n_hits);
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cub::DeviceRadixSort::SortPairs(p, tmp, (const u64*)hit_pos, spos, hit_slot, sslot,
                                           (int)n_hits, 0, 64, st));
  k_hit_read_keys<<<(int)((n_hits + 255) / 256), 256, 0, st>>>(spos, sslot, n_hits, (const u64*)read_starts,
                                                              n_reads, rkey);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cub::DeviceRadixSort::SortKeys(p, tmp2, (const u64*)rkey, rkey_sorted, (int)n_hits, 0, 64, st));
  int g = (int)((n_hits + 127) / 128);
  k_reduce_hits<<<g, 128, 0, st>>>(rkey_sorted, n_hits, spos, (const u64*)read_starts, (u64*)rec_read,
                                   rec_ndistinct, rec_nhits, (u64*)rec_first, (u64*)n_recs);
  CUDA_TRY(cudaGetLastError());
  return KDF_OK;
}

// ---- K7: coverage of the reference by hit k-mers ---------------------------
size_t kdf_hit_coverage_scratch_bytes(uint64_t n_hits, int k) {
  uint64_t n = n_hits * (uint64_t)(k > 0 ? k : 0);
  size_t t1 = 0, t2 = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, t1, (const u64*)nullptr, (u64*)nullptr, (int)n);
  cub::DeviceRunLengthEncode::Encode(nullptr, t2, (const u64*)nullptr, (u64*)nullptr, (u32*)nullptr,
                                     (u64*)nullptr, (int)n);
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  return 2 * al(n * 8) + al(t1 > t2 ? t1 : t2) + 256;
}

int kdf_hit_coverage(const uint32_t* hit_read, const uint32_t* hit_off, uint64_t n_hits, int k,
                     const int32_t* read_contig, const int64_t* read_ref_start,
                     const uint64_t* read_cig_off, const uint32_t* cigar, void* scratch,
                     size_t scratch_bytes, uint64_t* out_keys, uint32_t* out_counts, uint64_t* n_out,
                     void* stream) {
  if (n_hits == 0) return KDF_OK;
  if (!hit_read || !hit_off || !read_contig || !read_ref_start || !read_cig_off || !cigar || !scratch ||
      !out_keys || !out_counts || !n_out)
    return fail(KDF_ERR_ARG, "kdf_hit_coverage: NULL argument");
  if (k < 1 || k > 64) return fail(KDF_ERR_ARG, "kdf_hit_coverage: k must be in 1..64");
  const u64 n = n_hits * (u64)k;
  if (n > 0x7fffffffull) return fail(KDF_ERR_ARG, "kdf_hit_coverage: too many hits for one call");
  if (scratch_bytes < kdf_hit_coverage_scratch_bytes(n_hits, k))
    return fail(KDF_ERR_CAPACITY, "kdf_hit_coverage: scratch too small");
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  char* p = (char*)(((uintptr_t)scratch + 255) & ~(uintptr_t)255);
  u64* keys = (u64*)p;
  p += al(n * 8);
  u64* sorted = (u64*)p;
  p += al(n * 8);
  cudaStream_t st = (cudaStream_t)stream;
  int g = (int)((n_hits + 127) / 128);
  k_hit_coverage<<<g, 128, 0, st>>>(hit_read, hit_off, n_hits, k, (const int*)read_contig,
                                    (const long long*)read_ref_start, (const u64*)read_cig_off, cigar, keys);
  CUDA_TRY(cudaGetLastError());
  size_t t1 = 0, t2 = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, t1, (const u64*)keys, sorted, (int)n);
  CUDA_TRY(cub::DeviceRadixSort::SortKeys(p, t1, (const u64*)keys, sorted, (int)n, 0, 64, st));
  cub::DeviceRunLengthEncode::Encode(nullptr, t2, (const u64*)sorted, (u64*)out_keys, out_counts, (u64*)n_out, (int)n);
  CUDA_TRY(cub::DeviceRunLengthEncode::Encode(p, t2, (const u64*)sorted, (u64*)out_keys, out_counts,
                                              (u64*)n_out, (int)n, st));
  return KDF_OK;
}

int kdf_debug_hit_coverage_host(const uint32_t* hit_read, const uint32_t* hit_off, uint64_t n_hits, int k,
                                const int32_t* read_contig, const int64_t* read_ref_start,
                                const uint64_t* read_cig_off, const uint32_t* cigar, uint64_t* keys) {
  if (k < 1 || k > 64) return fail(KDF_ERR_ARG, "kdf_debug_hit_coverage_host: k must be in 1..64");
  for (uint64_t h = 0; h < n_hits; ++h) {
    const uint32_t r = hit_read[h];
    const uint32_t prev_end = (h > 0 && hit_read[h - 1] == r) ? hit_off[h - 1] + (uint32_t)k : 0u;
    u64* out = (u64*)keys + h * (u64)k;
    int n = expand_hit(cigar + read_cig_off[r], read_cig_off[r + 1] - read_cig_off[r],
                       (long long)read_ref_start[r], hit_off[h], k, prev_end, (u64)(uint32_t)read_contig[r], out);
    for (int j = n; j < k; ++j) out[j] = ~0ull;
  }
  return KDF_OK;
}

// ---- binning -------------------------------------------------------------
static int bin_qcap(int n_parts, int kw) {
  // ~64 KB of queues per CTA -> 3 CTAs per SM
  // a round offers 256 threads x 16 windows = 4096 keys, 4096 / n_parts per queue
  int q = (64 * 1024) / (n_parts * 8 * kw);
  if (q < 16) q = (128 * 1024) / (n_parts * 8 * kw);  // many wide bins: fewer CTAs per SM, deeper queues
  int want = 2 * (BIN_THREADS * BIN_WPR / n_parts) + 32;
  if (q > want) q = want;
  if (q < 8) q = 8;
  return q;
}

// Launch shape of k_bin_stream for n_parts bins (see the kernel's comment).  A round is
// 8192 / kw keys.  Two CTAs of 512 threads per SM while the queues fit ~108 KB; with
// many bins the slack a short queue needs (mean + 5 sigma) makes them larger, and one
// CTA of 1024 threads takes the same round.  KDF_BIN_THREADS / KDF_BIN_WPR override.
static BinPlan bin_plan(int n_parts, int kw) {
  BinPlan pl;
  pl.rep_log2 = 0;
  while ((n_parts << pl.rep_log2) < 32) ++pl.rep_log2;
  const int nv = n_parts << pl.rep_log2;
  const int round_keys = 8192 / kw;
  auto qcap_for = [&](int keys) {
    double mean = (double)keys / nv;
    return (int)(mean + 5.0 * sqrt(mean) + 8.0);
  };
  auto smem_for = [&](int qcap) { return (size_t)nv * qcap * 8 * kw + (size_t)nv * (8 + 4 + 8); };
  pl.threads = env_int("KDF_BIN_THREADS", 0);
  pl.wpr = env_int("KDF_BIN_WPR", 0);
  if (pl.threads != 512 && pl.threads != 1024 && pl.threads != 256) pl.threads = 0;
  if (pl.wpr != 4 && pl.wpr != 8 && pl.wpr != 16 && pl.wpr != 32) pl.wpr = 0;
  if (!pl.threads) pl.threads = smem_for(qcap_for(round_keys)) <= 108 * 1024 ? 512 : 1024;
  if (!pl.wpr) {
    pl.wpr = round_keys / pl.threads;
    if (pl.wpr < 4) pl.wpr = 4;
    if (pl.wpr > 32) pl.wpr = 32;
  }
  const int keys = pl.threads * pl.wpr;
  pl.qcap = qcap_for(keys);
  const size_t budget = 220 * 1024;
  if (smem_for(pl.qcap) > budget) pl.qcap = (int)((budget - (size_t)nv * 20) / ((size_t)nv * 8 * kw));
  if (pl.qcap < 4) pl.qcap = 4;
  pl.smem = smem_for(pl.qcap);
  pl.flat = (keys / nv) < 128 ? 1 : 0;
  return pl;
}

static int bin_stream_impl(const kdf_stream* s, int k, int by_owner, int n_parts, BinDest dst,
                           uint64_t* cursors, uint64_t* overflow, uint64_t* stats, void* stream,
                           uint64_t first_word, uint64_t n_range_words, int pass_log2, uint32_t pass_val) {
  int kw = kdf_key_words(k);
  if (!kw) return fail(KDF_ERR_ARG, "kdf_bin_stream: k must be in 1..64");
  if (n_parts < 1 || n_parts > BIN_MAX_PARTS) return fail(KDF_ERR_ARG, "kdf_bin_stream: n_parts must be 1..512");
  if (pass_log2 < 0 || pass_log2 > 16 || (pass_val >> pass_log2) != 0)
    return fail(KDF_ERR_ARG, "kdf_bin_stream: pass_val must be < 2^pass_log2, pass_log2 <= 16");
  // by_owner: 0 = hash ranges, 1 = owner ranks, R >= 2 = composite (R owners x n_parts/R ranges)
  int log2p = 0;
  int n_owners = 1;
  int bmode = by_owner == 0 ? 0 : (by_owner == 1 ? 1 : 2);
  if (bmode != 1) {
    n_owners = bmode == 2 ? by_owner : 1;
    if (n_parts % n_owners) return fail(KDF_ERR_ARG, "kdf_bin_stream: n_parts must be a multiple of the owner count");
    int n_local = n_parts / n_owners;
    while ((1 << log2p) < n_local) ++log2p;
    if ((1 << log2p) != n_local) return fail(KDF_ERR_ARG, "kdf_bin_stream: hash-range bins need a power-of-two count");
    if (log2p + pass_log2 > 28) return fail(KDF_ERR_ARG, "kdf_bin_stream: too many hash ranges");
  }
  StreamView v = view_of(s);
  if (first_word > v.n_words) first_word = v.n_words;
  v.w_begin = first_word;
  v.w_end = (n_range_words > v.n_words - first_word) ? v.n_words : first_word + n_range_words;
  if (v.w_end == v.w_begin) return KDF_OK;
  int sm = current_sm_count();
  cudaStream_t st = (cudaStream_t)stream;
  const BinPlan pl = bin_plan(n_parts, kw);
  const u32 qinv = (u32)((0x100000000ull + (u64)pl.qcap - 1) / (u64)pl.qcap);
#define KDF_BIN2(KW, BM, PS, RP)                                                                  \
  {                                                                                               \
    const void* fn = (const void*)k_bin_stream<KW, BM, PS, RP>;                                   \
    CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem)); \
    int g = grid_for(fn, pl.threads, pl.smem, v.w_end - v.w_begin, sm);                           \
    k_bin_stream<KW, BM, PS, RP><<<g, pl.threads, pl.smem, st>>>(                                 \
        v, k, log2p, (u32)n_parts, (u32)n_owners, pass_log2, pass_val, pl.wpr, pl.qcap, qinv,     \
        pl.rep_log2, pl.flat, dst, (u64*)cursors, (u64*)overflow, (u64*)stats);                   \
  }
#define KDF_BIN(KW, BM)                                                                           \
  {                                                                                               \
    if (pass_log2 > 0) {                                                                          \
      if (pl.rep_log2 > 0) KDF_BIN2(KW, BM, true, true) else KDF_BIN2(KW, BM, true, false)        \
    } else {                                                                                      \
      if (pl.rep_log2 > 0) KDF_BIN2(KW, BM, false, true) else KDF_BIN2(KW, BM, false, false)      \
    }                                                                                             \
  }
  if (kw == 1) {
    if (bmode == 0) KDF_BIN(1, 0) else if (bmode == 1) KDF_BIN(1, 1) else KDF_BIN(1, 2)
  } else {
    if (bmode == 0) KDF_BIN(2, 0) else if (bmode == 1) KDF_BIN(2, 1) else KDF_BIN(2, 2)
  }
#undef KDF_BIN
#undef KDF_BIN2
  CUDA_TRY(cudaGetLastError());
  return KDF_OK;
}

int kdf_bin_stream_pass(const kdf_stream* s, uint64_t first_word, uint64_t n_words, int k, int by_owner,
                        int n_parts, uint64_t* bins, uint64_t* const* bin_ptrs, uint64_t bin_cap,
                        uint64_t* cursors, uint64_t* overflow, uint64_t* stats, int pass_log2,
                        uint32_t pass_val, void* stream) {
  if (!s || (!bins && !bin_ptrs) || !cursors || !overflow)
    return fail(KDF_ERR_ARG, "kdf_bin_stream_pass: NULL argument");
  BinDest dst = {(u64*)bins, bins ? nullptr : (u64* const*)bin_ptrs, bin_cap};
  return bin_stream_impl(s, k, by_owner, n_parts, dst, cursors, overflow, stats, stream, first_word,
                         n_words, pass_log2, pass_val);
}

int kdf_bin_stream(const kdf_stream* s, int k, int by_owner, int n_parts, uint64_t* bins,
                   uint64_t bin_cap, uint64_t* cursors, uint64_t* overflow, uint64_t* stats,
                   void* stream) {
  if (!bins) return fail(KDF_ERR_ARG, "kdf_bin_stream: NULL argument");
  return kdf_bin_stream_pass(s, 0, ~0ull, k, by_owner, n_parts, bins, nullptr, bin_cap, cursors, overflow,
                             stats, 0, 0, stream);
}

int kdf_bin_stream_to(const kdf_stream* s, int k, int by_owner, int n_parts,
                      uint64_t* const* bin_ptrs, uint64_t bin_cap, uint64_t* cursors,
                      uint64_t* overflow, uint64_t* stats, void* stream) {
  return kdf_bin_stream_to_range(s, 0, ~0ull, k, by_owner, n_parts, bin_ptrs, bin_cap, cursors,
                                 overflow, stats, stream);
}

int kdf_bin_stream_to_range(const kdf_stream* s, uint64_t first_word, uint64_t n_words, int k,
                            int by_owner, int n_parts, uint64_t* const* bin_ptrs, uint64_t bin_cap,
                            uint64_t* cursors, uint64_t* overflow, uint64_t* stats, void* stream) {
  if (!bin_ptrs) return fail(KDF_ERR_ARG, "kdf_bin_stream_to: NULL argument");
  return kdf_bin_stream_pass(s, first_word, n_words, k, by_owner, n_parts, nullptr, bin_ptrs, bin_cap,
                             cursors, overflow, stats, 0, 0, stream);
}

int kdf_bin_stream_range(const kdf_stream* s, uint64_t first_word, uint64_t n_words, int k,
                         int by_owner, int n_parts, uint64_t* bins, uint64_t bin_cap,
                         uint64_t* cursors, uint64_t* overflow, uint64_t* stats, void* stream) {
  if (!bins) return fail(KDF_ERR_ARG, "kdf_bin_stream_range: NULL argument");
  return kdf_bin_stream_pass(s, first_word, n_words, k, by_owner, n_parts, bins, nullptr, bin_cap, cursors,
                             overflow, stats, 0, 0, stream);
}

int kdf_bin_keys_pass(const uint64_t* lo, const uint64_t* hi, uint64_t n, int k, int by_owner,
                      int n_parts, int pass_log2, uint32_t pass_val, uint64_t* bins, uint64_t bin_cap,
                      uint64_t* cursors, uint64_t* overflow, void* stream) {
  if (!bins || !cursors || !overflow) return fail(KDF_ERR_ARG, "kdf_bin_keys: NULL argument");
  int kw = kdf_key_words(k);
  if (!kw) return fail(KDF_ERR_ARG, "kdf_bin_keys: k must be in 1..64");
  if (n == 0) return KDF_OK;
  if (!lo) return fail(KDF_ERR_ARG, "kdf_bin_keys: NULL key array");
  if (n_parts < 1 || n_parts > BIN_MAX_PARTS) return fail(KDF_ERR_ARG, "kdf_bin_keys: n_parts must be 1..512");
  if (pass_log2 < 0 || pass_log2 > 16 || (pass_val >> pass_log2) != 0)
    return fail(KDF_ERR_ARG, "kdf_bin_keys: pass_val must be < 2^pass_log2, pass_log2 <= 16");
  BinDest dst = {(u64*)bins, nullptr, bin_cap};
  int log2p = 0;
  if (!by_owner) {
    while ((1 << log2p) < n_parts) ++log2p;
    if ((1 << log2p) != n_parts) return fail(KDF_ERR_ARG, "kdf_bin_keys: hash-range bins need a power-of-two count");
  }
  int sm = current_sm_count();
  cudaStream_t st = (cudaStream_t)stream;
  int qcap = bin_qcap(n_parts, kw);
#define KDF_BINK(KW, OWN)                                                                         \
  {                                                                                               \
    size_t smem = BinStage<KW>::bytes(n_parts, qcap);                                             \
    const void* fn = (const void*)k_bin_keys<KW, OWN>;                                            \
    CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    int g = grid_for(fn, BIN_THREADS, smem, (n + BIN_WPR - 1) / BIN_WPR, sm);                     \
    k_bin_keys<KW, OWN><<<g, BIN_THREADS, smem, st>>>((const u64*)lo, (const u64*)hi, n,          \
                                                      log2p + pass_log2, (u32)n_parts, pass_log2, \
                                                      pass_val, qcap, dst, (u64*)cursors,         \
                                                      (u64*)overflow);                            \
  }
  if (kw == 1) {
    if (by_owner) KDF_BINK(1, 1) else KDF_BINK(1, 0)
  } else {
    if (by_owner) KDF_BINK(2, 1) else KDF_BINK(2, 0)
  }
#undef KDF_BINK
  CUDA_TRY(cudaGetLastError());
  return KDF_OK;
}

int kdf_bin_keys(const uint64_t* lo, const uint64_t* hi, uint64_t n, int k, int by_owner,
                 int n_parts, uint64_t* bins, uint64_t bin_cap, uint64_t* cursors,
                 uint64_t* overflow, void* stream) {
  return kdf_bin_keys_pass(lo, hi, n, k, by_owner, n_parts, 0, 0, bins, bin_cap, cursors, overflow, stream);
}

int kdf_count_bins_packed(int k, uint32_t min0, uint32_t max0, uint32_t min1, uint32_t max1,
                          uint32_t count_min0, int want_planes) {
  int kw = kdf_key_words(k);
  if (!kw || want_planes) return 0;
  const char* off = getenv("KDF_COUNT_BINS_PLANES");   // A/B switch: force the plane form
  if (off && off[0] == '1') return 0;
  int free_bits = 64 * kw - 2 * k;   // spare bits above the key in its most significant word
  if (free_bits < 2) return 0;
  if (min0 < 1 || max0 != 0xffffffffu || min1 != 0 || (max1 != 0 && max1 != 0xffffffffu)) return 0;
  if (!(count_min0 <= 1 || count_min0 == min0)) return 0;
  if (free_bits < 31 && min0 > (1u << free_bits) - 1) return 0;
  if (min0 > 0x7fffffffu) return 0;
  return 1;
}

int kdf_count_bins(int k, int n_parts, const uint64_t* child_bins, uint64_t child_bin_cap,
                   const uint64_t* child_cursors, const uint64_t* ref_bins, uint64_t ref_bin_cap,
                   const uint64_t* ref_cursors, void* slice, uint64_t slice_capacity,
                   uint32_t min0, uint32_t max0, uint32_t min1, uint32_t max1, uint64_t* out_lo,
                   uint64_t* out_hi, uint32_t* out_p0, uint32_t* out_p1, uint64_t out_cap,
                   uint64_t* n_out, uint32_t count_min0, uint64_t* counters, void* stream) {
  return kdf_count_bins_multi(k, n_parts, 1, 1, child_bins, child_bin_cap, child_cursors, ref_bins,
                              ref_bin_cap, ref_cursors, slice, slice_capacity, min0, max0, min1, max1,
                              out_lo, out_hi, out_p0, out_p1, out_cap, n_out, count_min0, counters,
                              stream);
}

int kdf_count_bins_multi(int k, int n_parts, int n_src, int sub_split, const uint64_t* child_bins,
                         uint64_t child_bin_cap, const uint64_t* child_cursors,
                         const uint64_t* ref_bins, uint64_t ref_bin_cap, const uint64_t* ref_cursors,
                         void* slice, uint64_t slice_capacity, uint32_t min0, uint32_t max0,
                         uint32_t min1, uint32_t max1, uint64_t* out_lo, uint64_t* out_hi,
                         uint32_t* out_p0, uint32_t* out_p1, uint64_t out_cap, uint64_t* n_out,
                         uint32_t count_min0, uint64_t* counters, void* stream) {
  return kdf_count_bins_pass(k, n_parts, n_src, sub_split, 0, 0, child_bins, child_bin_cap, child_cursors,
                             ref_bins, ref_bin_cap, ref_cursors, slice, slice_capacity, min0, max0, min1,
                             max1, out_lo, out_hi, out_p0, out_p1, out_cap, n_out, count_min0, counters,
                             stream);
}

int kdf_count_bins_pass(int k, int n_parts, int n_src, int sub_split, int pass_log2, uint32_t pass_val,
                        const uint64_t* child_bins, uint64_t child_bin_cap, const uint64_t* child_cursors,
                        const uint64_t* ref_bins, uint64_t ref_bin_cap, const uint64_t* ref_cursors,
                        void* slice, uint64_t slice_capacity, uint32_t min0, uint32_t max0,
                        uint32_t min1, uint32_t max1, uint64_t* out_lo, uint64_t* out_hi,
                        uint32_t* out_p0, uint32_t* out_p1, uint64_t out_cap, uint64_t* n_out,
                        uint32_t count_min0, uint64_t* counters, void* stream) {
  if (pass_log2 < 0 || pass_log2 > 16 || (pass_val >> pass_log2) != 0)
    return fail(KDF_ERR_ARG, "kdf_count_bins: pass_val must be < 2^pass_log2, pass_log2 <= 16");
  if (!child_bins || !child_cursors || !slice || !n_out || !counters)
    return fail(KDF_ERR_ARG, "kdf_count_bins: NULL argument");
  if (n_src < 1) return fail(KDF_ERR_ARG, "kdf_count_bins: n_src must be >= 1");
  int log2s = 0;
  while ((1 << log2s) < sub_split) ++log2s;
  if (sub_split < 1 || sub_split > 64 || (1 << log2s) != sub_split)
    return fail(KDF_ERR_ARG, "kdf_count_bins: sub_split must be a power of two <= 64");
  int rc = check_table_args(k, slice_capacity, slice, "kdf_count_bins");
  if (rc != KDF_OK) return rc;
  int log2p = 0;
  while ((1 << log2p) < n_parts) ++log2p;
  if (n_parts < 1 || n_parts > BIN_MAX_PARTS || (1 << log2p) != n_parts)
    return fail(KDF_ERR_ARG, "kdf_count_bins: n_parts must be a power of two <= 512");
  kdf_table t;
  t.k = k;
  t.key_words = kdf_key_words(k);
  t.capacity = slice_capacity;
  t.base = slice;
  t.sm_count = current_sm_count();
  // a slice covers one of 2^pass_log2 * n_parts * sub_split hash ranges: the bins hold the
  // ranges of group pass_val (kdf_bin_stream_pass), slice pf of this call is range pass_base + pf
  t.log2_parts = pass_log2 + log2p + log2s;
  if (t.log2_parts > 28) return fail(KDF_ERR_ARG, "kdf_count_bins: too many hash ranges");
  const int filt_log2 = sub_split > 1 ? t.log2_parts : 0;
  const u32 pass_base = pass_val << (log2p + log2s);
  cudaStream_t st = (cudaStream_t)stream;
  u64* ctr = (u64*)counters;  // [0..3] = stats block (windows = keys applied, full, hits, new), [4] = #(p0 >= count_min0), [5] = #occupied
  const int kw = t.key_words;
  if (kdf_count_bins_packed(k, min0, max0, min1, max1, count_min0, out_p0 || out_p1)) {
    // packed form (K2c): the slice is keys only, the counter saturates at min0
    const int sh = 2 * k - 64 * (kw - 1);
    const u32 sat = min0;
    const int ignore_ref = (max1 != 0) ? 1 : 0;
    const int count_all = (count_min0 <= 1) ? 1 : 0;
    {
      u64 key_units = t.capacity * 8ull * kw / 16;
      int g = grid_for((const void*)k_fill_u64x2, 256, 0, key_units, t.sm_count);
      k_fill_u64x2<<<g, 256, 0, st>>>((ulonglong2*)t.base, key_units, EMPTY);
      CUDA_TRY(cudaGetLastError());
    }
    for (int pf = 0; pf < n_parts * sub_split; ++pf) {
      const int p = pf / sub_split;
      {   // bins are laid out [source][hash range][bin_cap]: one launch walks bin p of every source
        const u64* cb = (const u64*)child_bins + (u64)p * child_bin_cap * kw;
        const u64 sstride = (u64)n_parts * child_bin_cap * kw;
        if (kw == 1)
          rc = launch_packed_keys<1, OP_PACKED_COUNT>(&t, cb, child_bin_cap, (const u64*)child_cursors + p, sh, sat, ctr, st, filt_log2, pass_base + (u32)pf, n_src, sstride, (u64)n_parts);
        else
          rc = launch_packed_keys<2, OP_PACKED_COUNT>(&t, cb, child_bin_cap, (const u64*)child_cursors + p, sh, sat, ctr, st, filt_log2, pass_base + (u32)pf, n_src, sstride, (u64)n_parts);
        if (rc != KDF_OK) return rc;
      }
      if (ref_bins && ref_cursors && !ignore_ref) {
        const u64* rb = (const u64*)ref_bins + (u64)p * ref_bin_cap * kw;
        const u64 sstride = (u64)n_parts * ref_bin_cap * kw;
        if (kw == 1)
          rc = launch_packed_keys<1, OP_PACKED_MARK>(&t, rb, ref_bin_cap, (const u64*)ref_cursors + p, sh, sat, nullptr, st, filt_log2, pass_base + (u32)pf, n_src, sstride, (u64)n_parts);
        else
          rc = launch_packed_keys<2, OP_PACKED_MARK>(&t, rb, ref_bin_cap, (const u64*)ref_cursors + p, sh, sat, nullptr, st, filt_log2, pass_base + (u32)pf, n_src, sstride, (u64)n_parts);
        if (rc != KDF_OK) return rc;
      }
      if (kw == 1)
        rc = launch_emit_packed<1>(&t, sh, sat, ignore_ref, count_all, (u64*)out_lo, (u64*)out_hi, out_cap, (u64*)n_out, ctr + 4, ctr + 5, st);
      else
        rc = launch_emit_packed<2>(&t, sh, sat, ignore_ref, count_all, (u64*)out_lo, (u64*)out_hi, out_cap, (u64*)n_out, ctr + 4, ctr + 5, st);
      if (rc != KDF_OK) return rc;
    }
    return KDF_OK;
  }
  rc = clear_table_async(&t, st);  // once: every emit pass leaves the slice clean
  if (rc != KDF_OK) return rc;
  for (int pf = 0; pf < n_parts * sub_split; ++pf) {
    const int p = pf / sub_split;   // the bin; pf = the slice's hash range
    // bins are laid out [source][hash range][bin_cap]: one insert pass per source
    for (int sidx = 0; sidx < n_src; ++sidx) {
      u64 b = (u64)sidx * n_parts + p;
      const u64* cb = (const u64*)child_bins + b * child_bin_cap * kw;
      if (kw == 1)
        rc = launch_update_keys<1, OP_INSERT_COUNT>(&t, cb, nullptr, child_bin_cap, (const u64*)child_cursors + b, 0, 1, ctr, st, filt_log2, pass_base + (u32)pf);
      else
        rc = launch_update_keys<2, OP_INSERT_COUNT>(&t, cb, nullptr, child_bin_cap, (const u64*)child_cursors + b, 0, 1, ctr, st, filt_log2, pass_base + (u32)pf);
      if (rc != KDF_OK) return rc;
    }
    if (ref_bins && ref_cursors) {
      for (int sidx = 0; sidx < n_src; ++sidx) {
        u64 b = (u64)sidx * n_parts + p;
        const u64* rb = (const u64*)ref_bins + b * ref_bin_cap * kw;
        if (kw == 1)
          rc = launch_update_keys<1, OP_MARK_IF_PRESENT>(&t, rb, nullptr, ref_bin_cap, (const u64*)ref_cursors + b, 1, 1, nullptr, st, filt_log2, pass_base + (u32)pf);
        else
          rc = launch_update_keys<2, OP_MARK_IF_PRESENT>(&t, rb, nullptr, ref_bin_cap, (const u64*)ref_cursors + b, 1, 1, nullptr, st, filt_log2, pass_base + (u32)pf);
        if (rc != KDF_OK) return rc;
      }
    }
    if (kw == 1)
      rc = launch_emit_buckets<1>(&t, min0, max0, min1, max1, (u64*)out_lo, (u64*)out_hi, out_p0, out_p1,
                                  out_cap, (u64*)n_out, count_min0, ctr + 4, ctr + 5, st);
    else
      rc = launch_emit_buckets<2>(&t, min0, max0, min1, max1, (u64*)out_lo, (u64*)out_hi, out_p0, out_p1,
                                  out_cap, (u64*)n_out, count_min0, ctr + 4, ctr + 5, st);
    if (rc != KDF_OK) return rc;
  }
  return KDF_OK;
}

int kdf_valid_from_invalid(uint32_t* valid, uint64_t n_bases, const uint32_t* invalid_pos,
                           uint64_t n_invalid, void* stream) {
  if (n_bases == 0) return KDF_OK;
  if (!valid || (n_invalid && !invalid_pos)) return fail(KDF_ERR_ARG, "kdf_valid_from_invalid: NULL argument");
  if (n_bases > 0xffffffffull) return fail(KDF_ERR_ARG, "kdf_valid_from_invalid: stream longer than 2^32 bases");
  cudaStream_t st = (cudaStream_t)stream;
  const int sm = current_sm_count();
  const u64 n_words = (n_bases + 31) / 32;
  int g = grid_for((const void*)k_valid_fill, 256, 0, n_words, sm);
  k_valid_fill<<<g, 256, 0, st>>>(valid, n_words, n_bases);
  if (n_invalid) {
    g = grid_for((const void*)k_valid_clear, 256, 0, n_invalid, sm);
    k_valid_clear<<<g, 256, 0, st>>>(valid, n_bases, invalid_pos, n_invalid);
  }
  CUDA_TRY(cudaGetLastError());
  return KDF_OK;
}

int kdf_bench_random_access(void* buf, uint64_t buf_bytes, uint64_t n_ops, int atomic,
                            uint64_t* sink, void* stream) {
  if (!buf || buf_bytes < 32 || !sink) return fail(KDF_ERR_ARG, "kdf_bench_random_access: bad argument");
  int sm = current_sm_count();
  int g = grid_for((const void*)k_bench_random, 256, 0, n_ops / 8 + 1, sm);
  k_bench_random<<<g, 256, 0, (cudaStream_t)stream>>>((u32*)buf, buf_bytes / 32, n_ops, atomic, (u64*)sink);
  CUDA_TRY(cudaGetLastError());
  return KDF_OK;
}

// ------------------------------------------------------- host helpers -----
uint64_t kdf_pack_sequences(const char* seqs, const uint64_t* offsets, uint64_t n_seqs,
                            uint64_t* codes, uint32_t* valid, uint64_t* read_offsets) {
  static const struct Lut {
    uint8_t v[256];
    Lut() {
      memset(v, 4, sizeof(v));
      v[(int)'A'] = v[(int)'a'] = 0;
      v[(int)'C'] = v[(int)'c'] = 1;
      v[(int)'G'] = v[(int)'g'] = 2;
      v[(int)'T'] = v[(int)'t'] = 3;
    }
  } lut;
  uint64_t total = 0;
  for (uint64_t i = 0; i < n_seqs; ++i) total += offsets[i + 1] - offsets[i];
  if (n_seqs > 1) total += n_seqs - 1;
  if (read_offsets) {
    uint64_t p = 0;
    for (uint64_t i = 0; i < n_seqs; ++i) {
      read_offsets[i] = p;
      p += offsets[i + 1] - offsets[i] + 1;
    }
    read_offsets[n_seqs] = total;
    if (n_seqs == 0) read_offsets[0] = 0;
  }
  if (!codes || !valid) return total;
  uint64_t n_words = (total + 31) / 32;
  memset(codes, 0, n_words * sizeof(uint64_t));
  memset(valid, 0, n_words * sizeof(uint32_t));
  uint64_t p = 0;
  for (uint64_t i = 0; i < n_seqs; ++i) {
    const unsigned char* b = (const unsigned char*)seqs + offsets[i];
    uint64_t len = offsets[i + 1] - offsets[i];
    for (uint64_t j = 0; j < len; ++j, ++p) {
      uint8_t c = lut.v[b[j]];
      if (c < 4) {
        codes[p >> 5] |= (uint64_t)c << (62 - 2 * (p & 31));
        valid[p >> 5] |= 1u << (31 - (p & 31));
      }
    }
    ++p;  // separator (invalid)
  }
  return total;
}

// Test hook: run the device iterator code path on the CPU (same templates,
// host instantiation) so the bit manipulation can be checked without a GPU.
int kdf_debug_extract_host(const uint64_t* codes, const uint32_t* valid, uint64_t n_bases, int k,
                           int use_random_access, uint64_t* out_lo, uint64_t* out_hi,
                           uint8_t* out_ok) {
  int kw = kdf_key_words(k);
  if (!kw) return fail(KDF_ERR_ARG, "kdf_debug_extract_host: k must be in 1..64");
  StreamView v;
  v.codes = (const u64*)codes;
  v.valid = valid;
  v.n_bases = n_bases;
  v.n_words = (n_bases + 31) / 32;
  v.w_begin = 0;
  v.w_end = v.n_words;
  if (use_random_access >= 2) {
    // the chunked iterator of the stream kernels: 2 = chunks of 4 windows, 3 = chunks of 16
    for (u64 w = 0; w < v.n_words; ++w) {
      auto run = [&](auto& it) {
        for (int j0 = 0; j0 < 32;) {
          const int cn = use_random_access == 2 ? 4 : 16;
          for (int u = 0; u < cn; ++u) {
            u64 p = (w << 5) + j0 + u;
            auto c = it.key(u);
            bool ok = it.ok(u);
            if (p < n_bases) {
              out_lo[p] = ok ? c.lo : 0;
              if (out_hi) out_hi[p] = ok ? ((const u64*)&c)[sizeof(c) / 8 - 1] : 0;
              out_ok[p] = ok;
            }
          }
          if (cn == 4) it.template next<4>(); else it.template next<16>();
          j0 += cn;
        }
      };
      if (kw == 1) {
        WindowChunks<1> it(v, w, k);
        run(it);
        if (out_hi) for (u64 p = w << 5; p < n_bases && p < (w << 5) + 32; ++p) out_hi[p] = 0;
      } else {
        WindowChunks<2> it(v, w, k);
        run(it);
      }
    }
    return KDF_OK;
  }
  for (u64 w = 0; w < v.n_words; ++w) {
    if (kw == 1) {
      WindowIter<1> it(v, w, k);
      for (int j = 0; j < 32; ++j) {
        u64 p = (w << 5) + j;
        Key<1> c = it.canonical();
        bool ok = it.ok();
        if (use_random_access) ok = WindowAt<1>::get(v, p, k, c);
        if (p < n_bases) {
          out_lo[p] = ok ? c.lo : 0;
          if (out_hi) out_hi[p] = 0;
          out_ok[p] = ok;
        }
        it.advance();
      }
    } else {
      WindowIter<2> it(v, w, k);
      for (int j = 0; j < 32; ++j) {
        u64 p = (w << 5) + j;
        Key<2> c = it.canonical();
        bool ok = it.ok();
        if (use_random_access) ok = WindowAt<2>::get(v, p, k, c);
        if (p < n_bases) {
          out_lo[p] = ok ? c.lo : 0;
          out_hi[p] = ok ? c.hi : 0;
          out_ok[p] = ok;
        }
        it.advance();
      }
    }
  }
  return KDF_OK;
}

// Test hook: the hash functions of the device code, on the host.
int kdf_debug_hash_host(const uint64_t* lo, const uint64_t* hi, uint64_t n, int key_words,
                        int log2_parts, uint32_t n_buckets, uint32_t n_ranks, uint32_t* out_part,
                        uint32_t* out_bucket, uint32_t* out_owner) {
  if (!lo || (key_words == 2 && !hi)) return fail(KDF_ERR_ARG, "kdf_debug_hash_host: NULL keys");
  for (u64 i = 0; i < n; ++i) {
    u64 h;
    if (key_words == 1) {
      Key<1> k;
      k.lo = lo[i];
      h = hash_key(k);
    } else {
      Key<2> k;
      k.lo = lo[i];
      k.hi = hi[i];
      h = hash_key(k);
    }
    if (out_part) out_part[i] = part_of(h, log2_parts);
    if (out_bucket) out_bucket[i] = bucket_of(h, log2_parts, n_buckets);
    if (out_owner) out_owner[i] = owner_of(h, n_ranks ? n_ranks : 1);
  }
  return KDF_OK;
}

}  // extern "C"
