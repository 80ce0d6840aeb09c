// kdf_kernels.cu — sm_100a kernels + C ABI (include/kdf.h) of the k-mer engine.
//
// Kernel inventory (SURVEY §2 native table):
//   K1  k_extract          rolling extract + canonicalise
//   K2  k_count_stream     K1 fused with open-addressing insert / count /
//                          update-if-present / mark-if-present
//       k_update_keys      K2 on explicit key arrays
//   K3  k_threshold_compact threshold + stream compaction (dump -L, == 0, <= pmc)
//   K4  k_lookup_keys      batched membership / count lookup
//   K5  k_scan_reads       per-read membership scan + distinct reduction
//   K6  k_partition_stream owner binning in front of the NCCL all-to-all
//
// Mapping: stream kernels give each thread one 32-base word = 32 consecutive
// window starts, so code/valid loads are 8-byte coalesced and the canonical
// k-mer is maintained by a 2-bit rolling update.  Table traffic is the bound:
// one 32-byte sector per probe; probes are issued eight at a time per thread
// (loads first, resolution after) so that every thread keeps eight independent
// sector reads in flight.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <string>

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
  void* slots;
  int sm_count;
};

template <int KW> struct TableView {
  typename SlotOf<KW>::type* slots;
  u64 capacity;
};

// ----------------------------------------------------- slot primitives ----
__device__ __forceinline__ Key<1> ld_key(const Slot1* p) {
  Key<1> k;
  k.lo = __ldcg(&p->key);
  return k;
}
__device__ __forceinline__ Key<2> ld_key(const Slot2* p) {
  ulonglong2 v = __ldcg(reinterpret_cast<const ulonglong2*>(p));
  Key<2> k;
  k.lo = v.x;
  k.hi = v.y;
  return k;
}
__device__ __forceinline__ bool is_empty(const Key<1>& k) { return k.lo == EMPTY; }
__device__ __forceinline__ bool is_empty(const Key<2>& k) { return k.lo == EMPTY && k.hi == EMPTY; }
// may the loaded value be (a possibly torn view of) an empty slot?
__device__ __forceinline__ bool maybe_empty(const Key<1>& k) { return k.lo == EMPTY; }
__device__ __forceinline__ bool maybe_empty(const Key<2>& k) { return k.lo == EMPTY || k.hi == EMPTY; }

__device__ __forceinline__ Key<1> cas_key(Slot1* p, const Key<1>& val) {
  Key<1> old;
  old.lo = atomicCAS(&p->key, EMPTY, val.lo);
  return old;
}
__device__ __forceinline__ Key<2> cas_key(Slot2* p, const Key<2>& val) {
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

template <int MODE, typename SlotT>
__device__ __forceinline__ void apply_plane(SlotT* p, int plane, u32 arg) {
  u32* addr = plane ? &p->p1 : &p->p0;
  if (MODE == KDF_MODE_INSERT_COUNT || MODE == KDF_MODE_COUNT_IF_PRESENT) {
    atomicAdd(addr, arg);  // result unused -> RED
  } else if (MODE == KDF_MODE_MARK_IF_PRESENT) {
    atomicOr(addr, arg);
  }
}

struct LocalStats {
  u32 windows, hits, fresh, full;
};

// Finish one probe whose home slot (idx) has already been loaded into `cur`.
template <int KW, int MODE>
__device__ __forceinline__ void resolve(const TableView<KW>& t, u64 idx, const Key<KW>& key,
                                        Key<KW> cur, int plane, u32 arg, LocalStats& st) {
  typedef typename SlotOf<KW>::type SlotT;
  constexpr bool kInsert = (MODE == KDF_MODE_INSERT_COUNT || MODE == KDF_MODE_INSERT_ONLY);
  for (u64 n = 0; n < t.capacity; ++n) {
    SlotT* p = t.slots + idx;
    if (cur == key) {
      apply_plane<MODE>(p, plane, arg);
      st.hits++;
      return;
    }
    if (kInsert) {
      if (maybe_empty(cur)) {
        Key<KW> old = cas_key(p, key);
        if (is_empty(old)) {
          apply_plane<MODE>(p, plane, arg);
          st.fresh++;
          return;
        }
        if (old == key) {
          apply_plane<MODE>(p, plane, arg);
          st.hits++;
          return;
        }
      }
    } else {
      if (is_empty(cur)) return;  // miss
    }
    idx = (idx + 1 == t.capacity) ? 0 : idx + 1;
    cur = ld_key(t.slots + idx);
  }
  st.full = 1;
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

// ------------------------------------------------------------- K2 ---------
constexpr int CHUNK = 8;

template <int KW, int MODE>
__global__ void __launch_bounds__(256) k_count_stream(TableView<KW> t, StreamView s, int k,
                                                      int plane, u32 arg, u64* stats) {
  LocalStats st = {0, 0, 0, 0};
  u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x; w < s.n_words; w += stride) {
    WindowIter<KW> it(s, w, k);
    if (!it.any_valid()) continue;
#pragma unroll 1
    for (int c = 0; c < 32 / CHUNK; ++c) {
      Key<KW> keys[CHUNK];
      Key<KW> cur[CHUNK];
      u64 idx[CHUNK];
      u32 okm = 0;
#pragma unroll
      for (int u = 0; u < CHUNK; ++u) {
        bool ok = it.ok();
        keys[u] = it.canonical();
        it.advance();
        if (ok) {
          okm |= 1u << u;
          idx[u] = slot_of(hash_key(keys[u]), t.capacity);
          cur[u] = ld_key(t.slots + idx[u]);
        }
      }
      st.windows += __popc(okm);
#pragma unroll
      for (int u = 0; u < CHUNK; ++u) {
        if (okm & (1u << u)) resolve<KW, MODE>(t, idx[u], keys[u], cur[u], plane, arg, st);
      }
    }
  }
  flush_stats(st, stats);
}

template <int KW, int MODE>
__global__ void __launch_bounds__(256) k_update_keys(TableView<KW> t, const u64* lo, const u64* hi,
                                                     u64 n, int plane, u32 arg, u64* stats) {
  LocalStats st = {0, 0, 0, 0};
  u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    Key<KW> key;
    key.lo = lo[i];
    if (KW == 2) ((u64*)&key)[KW - 1] = hi[i];
    u64 idx = slot_of(hash_key(key), t.capacity);
    Key<KW> cur = ld_key(t.slots + idx);
    st.windows++;
    resolve<KW, MODE>(t, idx, key, cur, plane, arg, st);
  }
  flush_stats(st, stats);
}

// ------------------------------------------------------------- K3 ---------
template <int KW>
__global__ void __launch_bounds__(256) k_threshold_compact(TableView<KW> t, u32 min0, u32 max0,
                                                           u32 min1, u32 max1, u64* out_lo,
                                                           u64* out_hi, u32* out_p0, u32* out_p1,
                                                           u64 cap, u64* n_out) {
  typedef typename SlotOf<KW>::type SlotT;
  u64 stride = (u64)gridDim.x * blockDim.x;
  u64 first = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  u64 rounds = (t.capacity + stride - 1) / stride;
  unsigned lane = threadIdx.x & 31;
  for (u64 r = 0; r < rounds; ++r) {
    u64 i = first + r * stride;
    bool keep = false;
    Key<KW> key;
    key.lo = 0;
    u32 p0 = 0, p1 = 0;
    if (i < t.capacity) {
      const SlotT* p = t.slots + i;
      key = ld_key(p);
      if (!is_empty(key)) {
        p0 = __ldcg(&p->p0);
        p1 = __ldcg(&p->p1);
        keep = p0 >= min0 && p0 <= max0 && p1 >= min1 && p1 <= max1;
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
}

// ------------------------------------------------------------- K4 ---------
// static table (no concurrent inserts): keys may use the read-only path
template <int KW>
__device__ __forceinline__ bool find_slot(const TableView<KW>& t, const Key<KW>& key, u64& idx_out) {
  u64 idx = slot_of(hash_key(key), t.capacity);
  for (u64 n = 0; n < t.capacity; ++n) {
    Key<KW> cur = ld_key(t.slots + idx);
    if (cur == key) {
      idx_out = idx;
      return true;
    }
    if (is_empty(cur)) return false;
    idx = (idx + 1 == t.capacity) ? 0 : idx + 1;
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
    if (out_p0) out_p0[i] = f ? __ldcg(&t.slots[idx].p0) : 0u;
    if (out_p1) out_p1[i] = f ? __ldcg(&t.slots[idx].p1) : 0u;
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
      if (add0 && add0[i]) atomicAdd(&t.slots[idx].p0, add0[i]);
      if (add1 && add1[i]) atomicAdd(&t.slots[idx].p1, add1[i]);
    } else if (n_missing) {
      atomicAdd(n_missing, 1ull);
    }
  }
}

// ------------------------------------------------------------- K5 ---------
// One warp per read.  Lanes stride through the read's window starts; hits are
// appended to a per-warp shared list in position order (ballot prefix), then
// the warp counts distinct slot indices among them.
constexpr int SCAN_WARPS = 4;
constexpr int SCAN_HCAP = 1024;
#define KDF_NDISTINCT_OVERFLOW 0xFFFFFFFFu

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

// ------------------------------------------------------------- K6 ---------
constexpr int MAX_RANKS = 64;

template <int KW, bool SCATTER>
__global__ void __launch_bounds__(256) k_partition_stream(StreamView s, int k, u32 n_ranks,
                                                          u64* counts, const u64* bin_offsets,
                                                          u64* cursors, u64* out_lo, u64* out_hi) {
  __shared__ u32 sh_cnt[MAX_RANKS];
  if (!SCATTER) {
    for (int i = threadIdx.x; i < MAX_RANKS; i += blockDim.x) sh_cnt[i] = 0;
    __syncthreads();
  }
  const unsigned lane = threadIdx.x & 31;
  u64 stride = (u64)gridDim.x * blockDim.x;
  u64 n_iter = (s.n_words + stride - 1) / stride;
  u64 w0 = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  for (u64 itn = 0; itn < n_iter; ++itn) {
    u64 w = w0 + itn * stride;
    WindowIter<KW> it(s, w, k);  // loads beyond n_words read as invalid
#pragma unroll 1
    for (int j = 0; j < 32; ++j) {
      bool ok = it.ok();
      Key<KW> key = it.canonical();
      it.advance();
      u32 owner = ok ? owner_of(hash_key(key), n_ranks) : 0xffffffffu;
      if (!SCATTER) {
        if (ok) atomicAdd(&sh_cnt[owner], 1u);
      } else {
        unsigned peers = __match_any_sync(0xffffffffu, owner);
        if (ok) {
          int leader = __ffs(peers) - 1;
          u64 base = 0;
          if ((int)lane == leader) base = atomicAdd(cursors + owner, (u64)__popc(peers));
          base = __shfl_sync(peers, base, leader);
          u64 o = bin_offsets[owner] + base + __popc(peers & ((1u << lane) - 1));
          out_lo[o] = key.lo;
          if (KW == 2) out_hi[o] = ((const u64*)&key)[KW - 1];
        }
      }
    }
  }
  if (!SCATTER) {
    __syncthreads();
    for (u32 i = threadIdx.x; i < n_ranks; i += blockDim.x)
      if (sh_cnt[i]) atomicAdd(counts + i, (u64)sh_cnt[i]);
  }
}

// ------------------------------------------------ table maintenance -------
__global__ void __launch_bounds__(256) k_clear_table(ulonglong2* p, u64 n_units, int key_words) {
  // 16-byte units: KW=1 slot = {key, planes} ; KW=2 slot = {lo, hi} {planes, pad}
  u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_units; i += stride) {
    ulonglong2 v;
    if (key_words == 1)
      v = make_ulonglong2(EMPTY, 0ull);
    else
      v = (i & 1) ? make_ulonglong2(0ull, 0ull) : make_ulonglong2(EMPTY, EMPTY);
    p[i] = v;
  }
}

template <int KW>
__global__ void __launch_bounds__(256) k_clear_plane(TableView<KW> t, int plane) {
  u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < t.capacity; i += stride) {
    if (plane)
      t.slots[i].p1 = 0;
    else
      t.slots[i].p0 = 0;
  }
}

// ---------------------------------------------- random-access roofline ----
__global__ void __launch_bounds__(256) k_bench_random(u32* buf, u64 n_sectors, u64 n_ops, int atomic,
                                                      u64* sink) {
  u64 stride = (u64)gridDim.x * blockDim.x;
  u32 acc = 0;
  for (u64 i0 = (u64)blockIdx.x * blockDim.x + threadIdx.x; i0 < n_ops; i0 += stride * CHUNK) {
    u32 v[CHUNK];
    u64 sidx[CHUNK];
#pragma unroll
    for (int u = 0; u < CHUNK; ++u) {
      u64 i = i0 + (u64)u * stride;
      sidx[u] = mulhi64(mix64(i + 0x1234567ull), n_sectors);
      v[u] = (i < n_ops) ? __ldcg(buf + sidx[u] * 8) : 0u;
    }
#pragma unroll
    for (int u = 0; u < CHUNK; ++u) {
      u64 i = i0 + (u64)u * stride;
      acc += v[u];
      if (atomic && i < n_ops) atomicAdd(buf + sidx[u] * 8 + 2, 1u);
    }
  }
  if (acc == 0xdeadbeefu) *sink = acc;
}

// ------------------------------------------------------------ host side ---
static int grid_for(const void* func, int block, u64 work_items, int sm_count) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, func, block, 0) != cudaSuccess ||
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
  return v;
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
  if (key_words == 1) return (size_t)capacity * sizeof(Slot1);
  if (key_words == 2) return (size_t)capacity * sizeof(Slot2);
  return 0;
}

uint64_t kdf_table_capacity_for(uint64_t n_keys) {
  uint64_t c = n_keys * 2;
  if (c < 1024) c = 1024;
  return (c + 7) & ~7ull;  // whole 128-byte lines for either slot size
}

int kdf_table_clear(kdf_table* t, void* stream) {
  if (!t) return fail(KDF_ERR_ARG, "kdf_table_clear: table is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  u64 units = kdf_table_bytes(t->capacity, t->key_words) / 16;
  int g = grid_for((const void*)k_clear_table, 256, units, t->sm_count);
  k_clear_table<<<g, 256, 0, st>>>((ulonglong2*)t->slots, units, t->key_words);
  CUDA_TRY(cudaGetLastError());
  return KDF_OK;
}

int kdf_table_create(kdf_table** out, int k, uint64_t capacity, void* slots, void* stream) {
  if (!out || !slots) return fail(KDF_ERR_ARG, "kdf_table_create: NULL argument");
  int kw = kdf_key_words(k);
  if (!kw) return fail(KDF_ERR_ARG, "kdf_table_create: k must be in 1..64");
  if (capacity < 2) return fail(KDF_ERR_ARG, "kdf_table_create: capacity must be >= 2");
  if (((uintptr_t)slots & 31) != 0)
    return fail(KDF_ERR_ARG, "kdf_table_create: slots must be 32-byte aligned");
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0)
    return fail(KDF_ERR_NO_DEVICE, "no CUDA device visible");
  kdf_table* t = new kdf_table;
  t->k = k;
  t->key_words = kw;
  t->capacity = capacity;
  t->slots = slots;
  t->sm_count = current_sm_count();
  int rc = kdf_table_clear(t, stream);
  if (rc != KDF_OK) {
    delete t;
    return rc;
  }
  *out = t;
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
  cudaStream_t st = (cudaStream_t)stream;
  if (t->key_words == 1) {
    TableView<1> v{(Slot1*)t->slots, t->capacity};
    int g = grid_for((const void*)k_clear_plane<1>, 256, t->capacity, t->sm_count);
    k_clear_plane<1><<<g, 256, 0, st>>>(v, plane);
  } else {
    TableView<2> v{(Slot2*)t->slots, t->capacity};
    int g = grid_for((const void*)k_clear_plane<2>, 256, t->capacity, t->sm_count);
    k_clear_plane<2><<<g, 256, 0, st>>>(v, plane);
  }
  CUDA_TRY(cudaGetLastError());
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
    int g = grid_for((const void*)k_extract<1>, 256, v.n_words, sm);
    k_extract<1><<<g, 256, 0, st>>>(v, k, (u64*)out_lo, (u64*)out_hi, out_ok);
  } else {
    int g = grid_for((const void*)k_extract<2>, 256, v.n_words, sm);
    k_extract<2><<<g, 256, 0, st>>>(v, k, (u64*)out_lo, (u64*)out_hi, out_ok);
  }
  CUDA_TRY(cudaGetLastError());
  return KDF_OK;
}

}  // extern "C"

template <int KW, int MODE>
static int launch_count_stream(kdf_table* t, const StreamView& v, int plane, u32 arg, u64* stats,
                               cudaStream_t st) {
  TableView<KW> tv{(typename SlotOf<KW>::type*)t->slots, t->capacity};
  int g = grid_for((const void*)k_count_stream<KW, MODE>, 256, v.n_words, t->sm_count);
  k_count_stream<KW, MODE><<<g, 256, 0, st>>>(tv, v, t->k, plane, arg, stats);
  CUDA_TRY(cudaGetLastError());
  return KDF_OK;
}

extern "C" int kdf_count_stream(kdf_table* t, const kdf_stream* s, int mode, int plane, uint32_t arg,
                     uint64_t* stats, void* stream) {
  if (!t || !s) return fail(KDF_ERR_ARG, "kdf_count_stream: NULL argument");
  if (plane != 0 && plane != 1) return fail(KDF_ERR_ARG, "kdf_count_stream: plane must be 0 or 1");
  StreamView v = view_of(s);
  if (v.n_words == 0) return KDF_OK;
  cudaStream_t st = (cudaStream_t)stream;
  u64* sp = (u64*)stats;
#define KDF_DISPATCH(KW)                                                                    \
  switch (mode) {                                                                           \
    case KDF_MODE_INSERT_COUNT:                                                             \
      return launch_count_stream<KW, KDF_MODE_INSERT_COUNT>(t, v, plane, arg, sp, st);      \
    case KDF_MODE_INSERT_ONLY:                                                              \
      return launch_count_stream<KW, KDF_MODE_INSERT_ONLY>(t, v, plane, arg, sp, st);       \
    case KDF_MODE_COUNT_IF_PRESENT:                                                         \
      return launch_count_stream<KW, KDF_MODE_COUNT_IF_PRESENT>(t, v, plane, arg, sp, st);  \
    case KDF_MODE_MARK_IF_PRESENT:                                                          \
      return launch_count_stream<KW, KDF_MODE_MARK_IF_PRESENT>(t, v, plane, arg, sp, st);   \
    default:                                                                                \
      return fail(KDF_ERR_ARG, "kdf_count_stream: unknown mode");                           \
  }
  if (t->key_words == 1) {
    KDF_DISPATCH(1)
  } else {
    KDF_DISPATCH(2)
  }
#undef KDF_DISPATCH
}

template <int KW, int MODE>
static int launch_update_keys(kdf_table* t, const u64* lo, const u64* hi, u64 n, int plane, u32 arg,
                              u64* stats, cudaStream_t st) {
  TableView<KW> tv{(typename SlotOf<KW>::type*)t->slots, t->capacity};
  int g = grid_for((const void*)k_update_keys<KW, MODE>, 256, n, t->sm_count);
  k_update_keys<KW, MODE><<<g, 256, 0, st>>>(tv, lo, hi, n, plane, arg, stats);
  CUDA_TRY(cudaGetLastError());
  return KDF_OK;
}

extern "C" {

int kdf_update_keys(kdf_table* t, const uint64_t* lo, const uint64_t* hi, uint64_t n, int mode,
                    int plane, uint32_t arg, uint64_t* stats, void* stream) {
  if (!t) return fail(KDF_ERR_ARG, "kdf_update_keys: table is NULL");
  if (n == 0) return KDF_OK;
  if (!lo || (t->key_words == 2 && !hi)) return fail(KDF_ERR_ARG, "kdf_update_keys: NULL key array");
  if (plane != 0 && plane != 1) return fail(KDF_ERR_ARG, "kdf_update_keys: plane must be 0 or 1");
  cudaStream_t st = (cudaStream_t)stream;
  const u64* l = (const u64*)lo;
  const u64* h = (const u64*)hi;
  u64* sp = (u64*)stats;
#define KDF_DISPATCH(KW)                                                                       \
  switch (mode) {                                                                              \
    case KDF_MODE_INSERT_COUNT:                                                                \
      return launch_update_keys<KW, KDF_MODE_INSERT_COUNT>(t, l, h, n, plane, arg, sp, st);    \
    case KDF_MODE_INSERT_ONLY:                                                                 \
      return launch_update_keys<KW, KDF_MODE_INSERT_ONLY>(t, l, h, n, plane, arg, sp, st);     \
    case KDF_MODE_COUNT_IF_PRESENT:                                                            \
      return launch_update_keys<KW, KDF_MODE_COUNT_IF_PRESENT>(t, l, h, n, plane, arg, sp, st); \
    case KDF_MODE_MARK_IF_PRESENT:                                                             \
      return launch_update_keys<KW, KDF_MODE_MARK_IF_PRESENT>(t, l, h, n, plane, arg, sp, st); \
    default:                                                                                   \
      return fail(KDF_ERR_ARG, "kdf_update_keys: unknown mode");                               \
  }
  if (t->key_words == 1) {
    KDF_DISPATCH(1)
  } else {
    KDF_DISPATCH(2)
  }
#undef KDF_DISPATCH
}

int kdf_threshold_compact(const kdf_table* t, uint32_t min0, uint32_t max0, uint32_t min1,
                          uint32_t max1, uint64_t* out_lo, uint64_t* out_hi, uint32_t* out_p0,
                          uint32_t* out_p1, uint64_t cap, uint64_t* n_out, void* stream) {
  if (!t || !n_out) return fail(KDF_ERR_ARG, "kdf_threshold_compact: NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (t->key_words == 1) {
    TableView<1> tv{(Slot1*)t->slots, t->capacity};
    int g = grid_for((const void*)k_threshold_compact<1>, 256, t->capacity, t->sm_count);
    k_threshold_compact<1><<<g, 256, 0, st>>>(tv, min0, max0, min1, max1, (u64*)out_lo, (u64*)out_hi,
                                              out_p0, out_p1, cap, (u64*)n_out);
  } else {
    TableView<2> tv{(Slot2*)t->slots, t->capacity};
    int g = grid_for((const void*)k_threshold_compact<2>, 256, t->capacity, t->sm_count);
    k_threshold_compact<2><<<g, 256, 0, st>>>(tv, min0, max0, min1, max1, (u64*)out_lo, (u64*)out_hi,
                                              out_p0, out_p1, cap, (u64*)n_out);
  }
  CUDA_TRY(cudaGetLastError());
  return KDF_OK;
}

int kdf_lookup_keys(const kdf_table* t, const uint64_t* lo, const uint64_t* hi, uint64_t n,
                    uint8_t* out_found, uint32_t* out_p0, uint32_t* out_p1, void* stream) {
  if (!t) return fail(KDF_ERR_ARG, "kdf_lookup_keys: table is NULL");
  if (n == 0) return KDF_OK;
  if (!lo || (t->key_words == 2 && !hi)) return fail(KDF_ERR_ARG, "kdf_lookup_keys: NULL key array");
  cudaStream_t st = (cudaStream_t)stream;
  if (t->key_words == 1) {
    TableView<1> tv{(Slot1*)t->slots, t->capacity};
    int g = grid_for((const void*)k_lookup_keys<1>, 256, n, t->sm_count);
    k_lookup_keys<1><<<g, 256, 0, st>>>(tv, (const u64*)lo, (const u64*)hi, n, out_found, out_p0, out_p1);
  } else {
    TableView<2> tv{(Slot2*)t->slots, t->capacity};
    int g = grid_for((const void*)k_lookup_keys<2>, 256, n, t->sm_count);
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
    TableView<1> tv{(Slot1*)t->slots, t->capacity};
    int g = grid_for((const void*)k_add_planes<1>, 256, n, t->sm_count);
    k_add_planes<1><<<g, 256, 0, st>>>(tv, (const u64*)lo, (const u64*)hi, n, add0, add1, (u64*)n_missing);
  } else {
    TableView<2> tv{(Slot2*)t->slots, t->capacity};
    int g = grid_for((const void*)k_add_planes<2>, 256, n, t->sm_count);
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
    TableView<1> tv{(Slot1*)t->slots, t->capacity};
    int g = grid_for((const void*)k_scan_reads<1>, block, n_reads * 32, t->sm_count);
    k_scan_reads<1><<<g, block, 0, st>>>(tv, v, t->k, (const u64*)read_starts, read_lens, n_reads, min_distinct,
                                         out_ndistinct, out_nhits, (u64*)hit_pos, hit_slot, hit_cap,
                                         (u64*)n_hits, (u64*)stats);
  } else {
    TableView<2> tv{(Slot2*)t->slots, t->capacity};
    int g = grid_for((const void*)k_scan_reads<2>, block, n_reads * 32, t->sm_count);
    k_scan_reads<2><<<g, block, 0, st>>>(tv, v, t->k, (const u64*)read_starts, read_lens, n_reads, min_distinct,
                                         out_ndistinct, out_nhits, (u64*)hit_pos, hit_slot, hit_cap,
                                         (u64*)n_hits, (u64*)stats);
  }
  CUDA_TRY(cudaGetLastError());
  return KDF_OK;
}

int kdf_partition_stream(const kdf_stream* s, int k, int n_ranks, uint64_t* counts,
                         const uint64_t* bin_offsets, uint64_t* cursors, uint64_t* out_lo,
                         uint64_t* out_hi, void* stream) {
  if (!s) return fail(KDF_ERR_ARG, "kdf_partition_stream: NULL stream");
  int kw = kdf_key_words(k);
  if (!kw) return fail(KDF_ERR_ARG, "kdf_partition_stream: k must be in 1..64");
  if (n_ranks < 1 || n_ranks > MAX_RANKS) return fail(KDF_ERR_ARG, "kdf_partition_stream: n_ranks must be 1..64");
  bool scatter = out_lo != nullptr;
  if (!scatter && !counts) return fail(KDF_ERR_ARG, "kdf_partition_stream: counts required for the histogram pass");
  if (scatter && (!bin_offsets || !cursors || (kw == 2 && !out_hi)))
    return fail(KDF_ERR_ARG, "kdf_partition_stream: scatter pass needs bin_offsets, cursors and outputs");
  StreamView v = view_of(s);
  if (v.n_words == 0) return KDF_OK;
  int sm = current_sm_count();
  cudaStream_t st = (cudaStream_t)stream;
#define KDF_PART(KW, SC)                                                                   \
  {                                                                                        \
    int g = grid_for((const void*)k_partition_stream<KW, SC>, 256, v.n_words, sm);         \
    k_partition_stream<KW, SC><<<g, 256, 0, st>>>(v, k, (u32)n_ranks, (u64*)counts,        \
                                                  (const u64*)bin_offsets, (u64*)cursors,  \
                                                  (u64*)out_lo, (u64*)out_hi);             \
  }
  if (kw == 1) {
    if (scatter) KDF_PART(1, true) else KDF_PART(1, false)
  } else {
    if (scatter) KDF_PART(2, true) else KDF_PART(2, false)
  }
#undef KDF_PART
  CUDA_TRY(cudaGetLastError());
  return KDF_OK;
}

int kdf_bench_random_access(void* buf, uint64_t buf_bytes, uint64_t n_ops, int atomic,
                            uint64_t* sink, void* stream) {
  if (!buf || buf_bytes < 32 || !sink) return fail(KDF_ERR_ARG, "kdf_bench_random_access: bad argument");
  int sm = current_sm_count();
  int g = grid_for((const void*)k_bench_random, 256, n_ops / CHUNK + 1, sm);
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

}  // extern "C"
