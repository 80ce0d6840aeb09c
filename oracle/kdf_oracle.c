/*
 * kdf_oracle.c — C/OpenMP twin of the CPU oracle (TEST INFRASTRUCTURE ONLY).
 *
 * Restates, on the packed stream layout, what the reference delegates to
 * Jellyfish: `count -m k -C` (discovery/pipeline.py:114-122), `count --if`
 * (core/jellyfish_wrappers.py:167-176), the reference `query` subtraction
 * (discovery/pipeline.py:286-304), `dump -L` (:207-211) and the per-read scan
 * of core/bam_scanner.py:434-443.  Multi-threaded lock-free-style hash like
 * Jellyfish's (-t threads).  Used by tests (checked equal to the numpy oracle)
 * and by bench.py's cpu_baseline / --impl reference legs.  Never linked into
 * or called by the product.
 *
 * Written independently of the CUDA code: byte-at-a-time rolling k-mers in
 * unsigned __int128, run-length validity, a slot-state claim protocol.
 */
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

typedef struct {
  uint64_t lo, hi;
  uint32_t p0, p1;
  uint32_t state; /* 0 empty, 1 being written, 2 full */
  uint32_t pad;
} oslot;

typedef struct {
  oslot* slots;
  uint64_t cap;
  int full;
} okdf_table;

static inline uint64_t omix(uint64_t x) {
  x ^= x >> 31;
  x *= 0x7fb5d329728ea185ULL;
  x ^= x >> 27;
  x *= 0x81dadef4bc2dd44dULL;
  x ^= x >> 33;
  return x;
}
static inline uint64_t ohash(uint64_t lo, uint64_t hi) { return omix(lo ^ omix(hi ^ 0x5851f42d4c957f2dULL)); }

okdf_table* okdf_table_new(uint64_t cap) {
  okdf_table* t = (okdf_table*)malloc(sizeof(okdf_table));
  if (!t) return NULL;
  t->cap = cap < 2 ? 2 : cap;
  t->full = 0;
  t->slots = (oslot*)calloc(t->cap, sizeof(oslot));
  if (!t->slots) {
    free(t);
    return NULL;
  }
  return t;
}
void okdf_table_free(okdf_table* t) {
  if (!t) return;
  free(t->slots);
  free(t);
}
void okdf_clear_plane(okdf_table* t, int plane) {
#pragma omp parallel for schedule(static)
  for (long i = 0; i < (long)t->cap; ++i) {
    if (plane) t->slots[i].p1 = 0; else t->slots[i].p0 = 0;
  }
}
int okdf_is_full(const okdf_table* t) { return t->full; }

/* mode: 0 insert+count, 1 insert only, 2 count if present, 3 mark (OR) if present */
static inline void oprobe(okdf_table* t, uint64_t lo, uint64_t hi, int mode, int plane, uint32_t arg) {
  uint64_t i = ohash(lo, hi) % t->cap;
  for (uint64_t n = 0; n < t->cap; ++n) {
    oslot* s = &t->slots[i];
    uint32_t st = __atomic_load_n(&s->state, __ATOMIC_ACQUIRE);
    if (st == 0) {
      if (mode >= 2) return; /* miss */
      uint32_t exp = 0;
      if (__atomic_compare_exchange_n(&s->state, &exp, 1, 0, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE)) {
        s->lo = lo;
        s->hi = hi;
        __atomic_store_n(&s->state, 2, __ATOMIC_RELEASE);
        st = 2;
      } else {
        st = exp;
      }
    }
    while (st == 1) st = __atomic_load_n(&s->state, __ATOMIC_ACQUIRE);
    if (s->lo == lo && s->hi == hi) {
      uint32_t* p = plane ? &s->p1 : &s->p0;
      if (mode == 0 || mode == 2) __atomic_fetch_add(p, arg, __ATOMIC_RELAXED);
      else if (mode == 3) __atomic_fetch_or(p, arg, __ATOMIC_RELAXED);
      return;
    }
    if (++i == t->cap) i = 0;
  }
  t->full = 1;
}

static inline int olookup(const okdf_table* t, uint64_t lo, uint64_t hi, uint64_t* idx) {
  uint64_t i = ohash(lo, hi) % t->cap;
  for (uint64_t n = 0; n < t->cap; ++n) {
    const oslot* s = &t->slots[i];
    if (s->state == 0) return 0;
    if (s->lo == lo && s->hi == hi) {
      *idx = i;
      return 1;
    }
    if (++i == t->cap) i = 0;
  }
  return 0;
}

static inline unsigned code_at(const uint64_t* codes, uint64_t p) {
  return (unsigned)((codes[p >> 5] >> (62 - 2 * (p & 31))) & 3);
}
static inline unsigned valid_at(const uint32_t* valid, uint64_t p) {
  return (valid[p >> 5] >> (31 - (p & 31))) & 1u;
}

/* canonical k-mers of windows starting in [w0, w1): calls body(lo, hi, start) */
#define FOR_EACH_WINDOW(codes, valid, n_bases, k, w0, w1, BODY)                    \
  do {                                                                            \
    u128 mask_ = (k) == 64 ? ~(u128)0 : (((u128)1 << (2 * (k))) - 1);             \
    u128 fwd_ = 0, rc_ = 0;                                                       \
    uint64_t run_ = 0;                                                            \
    uint64_t pend_ = (w1) + (uint64_t)(k) - 1;                                    \
    if (pend_ > (n_bases)) pend_ = (n_bases);                                     \
    for (uint64_t p_ = (w0); p_ < pend_; ++p_) {                                  \
      if (!valid_at(valid, p_)) {                                                 \
        run_ = 0;                                                                 \
        continue;                                                                 \
      }                                                                           \
      unsigned c_ = code_at(codes, p_);                                           \
      fwd_ = ((fwd_ << 2) | c_) & mask_;                                          \
      rc_ = (rc_ >> 2) | ((u128)(3 - c_) << (2 * ((k)-1)));                       \
      if (++run_ >= (uint64_t)(k)) {                                              \
        u128 can_ = fwd_ < rc_ ? fwd_ : rc_;                                      \
        uint64_t lo_ = (uint64_t)can_, hi_ = (uint64_t)(can_ >> 64);              \
        uint64_t start_ = p_ + 1 - (uint64_t)(k);                                 \
        BODY                                                                      \
      }                                                                           \
    }                                                                             \
  } while (0)

uint64_t okdf_count_stream(okdf_table* t, const uint64_t* codes, const uint32_t* valid,
                           uint64_t n_bases, int k, int mode, int plane, uint32_t arg, int threads) {
  if (n_bases < (uint64_t)k) return 0;
  uint64_t n_win = n_bases - k + 1;
  uint64_t total = 0;
  if (threads < 1) threads = 1;
  uint64_t chunk = 1 << 16;
  long n_chunks = (long)((n_win + chunk - 1) / chunk);
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads) reduction(+ : total)
  for (long c = 0; c < n_chunks; ++c) {
    uint64_t w0 = (uint64_t)c * chunk, w1 = w0 + chunk;
    if (w1 > n_win) w1 = n_win;
    uint64_t cnt = 0;
    FOR_EACH_WINDOW(codes, valid, n_bases, k, w0, w1, {
      (void)start_;
      oprobe(t, lo_, hi_, mode, plane, arg);
      cnt++;
    });
    total += cnt;
  }
  return total;
}

void okdf_update_keys(okdf_table* t, const uint64_t* lo, const uint64_t* hi, uint64_t n, int mode,
                      int plane, uint32_t arg, int threads) {
#pragma omp parallel for schedule(static) num_threads(threads < 1 ? 1 : threads)
  for (long i = 0; i < (long)n; ++i) oprobe(t, lo[i], hi ? hi[i] : 0, mode, plane, arg);
}

uint64_t okdf_threshold(const okdf_table* t, uint32_t min0, uint32_t max0, uint32_t min1,
                        uint32_t max1, uint64_t* out_lo, uint64_t* out_hi, uint32_t* out_p0,
                        uint32_t* out_p1, uint64_t cap) {
  uint64_t n = 0;
  for (uint64_t i = 0; i < t->cap; ++i) {
    const oslot* s = &t->slots[i];
    if (s->state != 2) continue;
    if (s->p0 < min0 || s->p0 > max0 || s->p1 < min1 || s->p1 > max1) continue;
    if (n < cap) {
      if (out_lo) out_lo[n] = s->lo;
      if (out_hi) out_hi[n] = s->hi;
      if (out_p0) out_p0[n] = s->p0;
      if (out_p1) out_p1[n] = s->p1;
    }
    n++;
  }
  return n;
}

/* per-read distinct / hit counts (core/bam_scanner.py:434-443) */
uint64_t okdf_scan_reads(const okdf_table* t, const uint64_t* codes, const uint32_t* valid,
                         uint64_t n_bases, const uint64_t* read_starts, const uint32_t* read_lens,
                         uint64_t n_reads, int k, uint32_t* out_nd, uint32_t* out_nh, int threads) {
  uint64_t total = 0;
#pragma omp parallel num_threads(threads < 1 ? 1 : threads) reduction(+ : total)
  {
    uint64_t* seen = NULL;
    size_t seen_cap = 0;
#pragma omp for schedule(dynamic, 256)
    for (long r = 0; r < (long)n_reads; ++r) {
      uint64_t s0 = read_starts[r];
      uint64_t len = read_lens[r];
      uint32_t nh = 0, nd = 0;
      if (len >= (uint64_t)k) {
        uint64_t end = s0 + len;
        if (end > n_bases) end = n_bases;
        size_t nseen = 0;
        uint64_t cnt = 0;
        FOR_EACH_WINDOW(codes, valid, end, k, s0, end - k + 1, {
          (void)start_;
          cnt++;
          uint64_t idx;
          if (olookup(t, lo_, hi_, &idx)) {
            nh++;
            int dup = 0;
            for (size_t q = 0; q < nseen; ++q)
              if (seen[q] == idx) { dup = 1; break; }
            if (!dup) {
              if (nseen == seen_cap) {
                seen_cap = seen_cap ? seen_cap * 2 : 256;
                seen = (uint64_t*)realloc(seen, seen_cap * sizeof(uint64_t));
              }
              seen[nseen++] = idx;
              nd++;
            }
          }
        });
        total += cnt;
      }
      out_nd[r] = nd;
      out_nh[r] = nh;
    }
    free(seen);
  }
  return total;
}

int okdf_max_threads(void) { return omp_get_max_threads(); }
