// kdf_host.cpp — host-side (CPU) half of libkdf_sm100.so: a multi-threaded
// BGZF/BAM decoder that emits 2-bit packed read batches (the stream layout of
// include/kdf.h) plus per-record alignment metadata.
//
// Replaces, for BAM input, the two record streams the reference obtains from
// third-party tools:
//   KDF_BAM_FASTA : `samtools fasta -F 0xD00 X.bam`
//                   (core/jellyfish_wrappers.py:159-165,
//                    discovery/pipeline.py:106-112, 369-375): drop records with
//                   flag & 0xD00, then within a run of consecutive records with
//                   the same QNAME keep the first record of each read-part
//                   (READ1 / READ2 / other).
//   KDF_BAM_SCAN  : pysam iteration of the anchoring scan
//                   (core/bam_scanner.py:405-414): skip secondary and duplicate
//                   records, keep supplementary and unmapped ones.
//   KDF_BAM_ALL   : every record.
// CRAM is not supported (needs htslib codecs); see DESIGN.md.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/types.h>
#include <omp.h>
#include <sched.h>
#include <zlib.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include <omp.h>

#include <atomic>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include "../../include/kdf.h"
#include "kdf_inflate.h"

namespace {

thread_local std::string g_host_err;

// std::vector whose resize() leaves new elements uninitialised (the packer zero-fills
// codes / valid itself, in parallel)
template <class T> struct NoInitAlloc : std::allocator<T> {
  template <class U> struct rebind { using other = NoInitAlloc<U>; };
  template <class U> void construct(U* p) noexcept { ::new ((void*)p) U; }
  template <class U, class... A> void construct(U* p, A&&... a) { ::new ((void*)p) U(std::forward<A>(a)...); }
};

// A growable array of plain values for the batch outputs: resize() leaves new elements
// uninitialised (they are written by the parallel fill loops, which is also what first touches
// their pages).  From 256 KB up the storage is its own anonymous mapping grown with mremap(),
// which moves pages instead of copying them — malloc/realloc copies (and faults in the copy,
// on the one thread that lays a round out) until a block passes glibc's moving mmap
// threshold of up to 32 MB.  (std::vector value-initialises and copies on every growth.)
// Test hook: KDF_BAM_FAIL_ALLOC=n makes the n-th growth of a batch buffer fail, so that the
// out-of-memory paths (also the ones inside the OpenMP region) can be exercised.
inline void maybe_fail_alloc() {
  static const long fail_at = getenv("KDF_BAM_FAIL_ALLOC") ? atol(getenv("KDF_BAM_FAIL_ALLOC")) : 0;
  if (fail_at > 0) {
    static std::atomic<long> calls{0};
    if (calls.fetch_add(1, std::memory_order_relaxed) + 1 == fail_at) throw std::bad_alloc();
  }
}

inline bool thp_wanted() {   // KDF_BAM_THP=1: ask for transparent huge pages for the large buffers
  static const bool on = getenv("KDF_BAM_THP") != nullptr && atoi(getenv("KDF_BAM_THP")) != 0;
  return on;
}

template <class T> struct PodVec {
  T* p = nullptr;
  size_t n = 0, cap = 0;
  bool mapped = false;
  static constexpr size_t MAP_FROM = 256u << 10;
  PodVec() = default;
  PodVec(const PodVec&) = delete;
  PodVec& operator=(const PodVec&) = delete;
  ~PodVec() { release(); }
  void release() {
    if (mapped)
      munmap(p, cap * sizeof(T));
    else
      free(p);
    p = nullptr;
    n = cap = 0;
    mapped = false;
  }
  T* data() { return p; }
  const T* data() const { return p; }
  size_t size() const { return n; }
  size_t capacity() const { return cap; }
  bool empty() const { return n == 0; }
  T& operator[](size_t i) { return p[i]; }
  const T& operator[](size_t i) const { return p[i]; }
  void reserve(size_t want) {
    if (want <= cap) return;
    maybe_fail_alloc();
    if (want > SIZE_MAX / sizeof(T) - 4096) throw std::bad_alloc();
    size_t bytes = want * sizeof(T);
    if (bytes >= MAP_FROM) {
      bytes = (bytes + 4095) & ~(size_t)4095;
      void* q;
      if (mapped) {
        q = mremap(p, cap * sizeof(T), bytes, MREMAP_MAYMOVE);
      } else {
        q = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (q != MAP_FAILED && n) memcpy(q, p, n * sizeof(T));
      }
      if (q == MAP_FAILED) throw std::bad_alloc();
#ifdef MADV_HUGEPAGE
      if (thp_wanted() && bytes >= (4u << 20)) madvise(q, bytes, MADV_HUGEPAGE);
#endif
      if (!mapped) free(p);
      p = (T*)q;
      cap = bytes / sizeof(T);
      mapped = true;
      return;
    }
    void* q = realloc(p, bytes);
    if (!q) throw std::bad_alloc();
    p = (T*)q;
    cap = want;
  }
  void resize(size_t m) {
    if (m > cap) reserve(m + m / 2 + 1024);
    n = m;
  }
  void assign(size_t m, T v) {
    resize(m);
    for (size_t i = 0; i < m; ++i) p[i] = v;
  }
  void clear() { n = 0; }
};

struct BlockRef {
  uint64_t coff;   // offset of the block in the compressed bytes
  uint32_t csize;  // total block size
  uint32_t usize;  // uncompressed size (ISIZE)
  uint64_t uoff;   // offset of its data in the inflated chunk
};

// compressed BGZF blocks of one chunk, read but not yet inflated
struct CompBuf {
  std::vector<uint8_t, NoInitAlloc<uint8_t>> bytes;
  std::vector<BlockRef> blocks;
  uint64_t total_u = 0;
  uint64_t ustart = 0;   // offset of its first byte in the file's uncompressed stream
  bool valid = false;
};

// An inflated chunk of the record stream.  data[own .. own + total) is what its BGZF blocks
// inflate to; data[head .. own) holds the bytes of the previous chunk that belong to a record
// completed here.  The record walk (filled while the chunk is inflated) gives the compact
// per-record arrays the consumer works from.
struct Chunk {
  uint8_t* data = nullptr;
  size_t cap = 0;
  size_t own = 0, head = 0, total = 0;
  uint64_t ustart = 0;   // uncompressed-stream offset of data[own]
  bool valid = false;    // inflated, walked and classified
  std::vector<int64_t> w_off;   // n_rec + 1: record i is data[own + w_off[i] .. own + w_off[i + 1]) (may start in the head: negative)
  std::vector<uint16_t> w_flag;
  std::vector<uint32_t> w_lseq, w_prevp;
  std::vector<uint8_t> w_cls;
  size_t n_rec = 0;        // records complete in this chunk
  size_t sel = 0;          // records [0, sel) are consumed
  int64_t exit_off = 0;    // start of the first record that is not complete here (== total: none)
  uint64_t rec_base = 0;   // file-order index of record 0
  std::string walk_err;    // the walk stopped at exit_off on a corrupt block_size
  std::string prev_qname;  // QNAME of the last primary record before this chunk
  std::string last_qname;  // ... and of the last one up to its end
};

struct Bam {
  FILE* fh = nullptr;
  int threads = 1;
  std::vector<std::string> ref_names;
  std::vector<int32_t> ref_lens;
  std::string header_text;
  std::vector<uint8_t> carry;  // undecoded tail of the previous chunk
  bool eof = false;
  bool verify_crc = true;     // check every block's CRC32 (KDF_BAM_CRC=0 switches it off)
  uint64_t record_index = 0;  // file-order index of the next record
  // uncompressed / compressed bytes of all blocks taken so far, and the index of the
  // blocks (uncompressed start -> file offset) that kdf_bam_fetch_records seeks with
  uint64_t u_total = 0, c_total = 0;
  std::vector<std::pair<uint64_t, uint64_t>> blk_index;
  uint64_t first_rec_uoff = 0;   // uncompressed offset of the first record (= header size)
  // (atomics: the read-ahead thread sets them when it reaches the block, while the consumer
  // of an earlier chunk is looking at them)
  std::atomic<uint64_t> u_limit{~0ull};   // deliver no record at or past this uncompressed offset (kdf_bam_set_end)
  std::atomic<uint64_t> u_begin{0};       // records before this offset only update the QNAME-run state (kdf_bam_set_begin)
  uint64_t begin_coff = ~0ull;
  uint32_t begin_in = 0;
  size_t skip_bytes = 0;         // kdf_bam_seek: bytes of the first block that precede the target record
  uint64_t end_coff = ~0ull;     // kdf_bam_set_end: virtual offset (block, offset in block) to stop at
  uint32_t end_in = 0;
  std::string path;
  // collapse state of the FASTA stream (persists across batches): the read parts seen in the
  // current run of same-QNAME primary records (the QNAMEs themselves travel with the chunks)
  unsigned seen_parts = 0;
  // inflate buffers, recycled across chunks and batches: plain malloc memory (no
  // zero-fill, no growth copies) whose pages stay mapped once touched
  std::vector<std::pair<uint8_t*, size_t>> pool;
  std::vector<uint8_t> pending;   // read from the file, not yet taken as blocks
  bool file_eof = false;          // nothing left to read from the file
  uint64_t chunk_bytes = 64ull << 20;   // data per pipeline chunk (KDF_BAM_CHUNK_KB: tests)
  size_t gap = 1u << 20;                // headroom in front of a chunk (KDF_BAM_GAP: tests)
  // the decode pipeline (kdf_bam_next_batch): the chunk being parsed, the inflated
  // chunk after it, and the compressed bytes of the one after that
  Chunk cur, next;
  CompBuf ahead, ahead2;
  bool range_done = false;   // the end of the range (kdf_bam_set_end) has been delivered
  // scratch of a round (kdf_bam_next_batch): the hand-over of the record walk from block to
  // block, each thread's record offsets and where a block's records sit in them, the
  // selection of the consumer
  std::vector<std::atomic<int64_t>> chain;
  std::vector<std::vector<int64_t>> tl_off;
  std::vector<std::vector<uint16_t>> tl_flag;   // (parallel to tl_off: what the walker could classify while hot)
  std::vector<std::vector<uint32_t>> tl_lseq;
  std::vector<std::vector<uint8_t>> tl_cls;
  std::vector<int64_t> fin_off;
  std::vector<uint32_t> blk_first, blk_count;
  std::vector<uint16_t> blk_owner;
  std::vector<uint32_t> s_idx;
  std::vector<uint64_t> s_start;
  std::vector<uint8_t> s_fk;
  std::vector<const uint8_t*> s_sa;
  // where the decode time goes (seconds of the calling thread; printed by kdf_bam_close when
  // KDF_BAM_TIMING is set)
  struct Timing {
    double rounds = 0, wait = 0, select = 0, fill = 0, index = 0, invalid = 0;
    uint64_t n_rounds = 0, n_batches = 0, n_reads = 0;
  } tm;
  uint8_t* get_buf(size_t n, size_t* cap) {
    size_t best = pool.size();
    for (size_t i = 0; i < pool.size(); ++i)
      if (pool[i].second >= n && (best == pool.size() || pool[i].second < pool[best].second)) best = i;
    if (best < pool.size()) {
      uint8_t* p = pool[best].first;
      *cap = pool[best].second;
      pool.erase(pool.begin() + (long)best);
      return p;
    }
    // 2 MB-aligned.  Fresh pages are touched by all inflate threads at once, which is slow
    // (page faults serialise), but the pool makes that a one-off per reader.  Advising
    // transparent huge pages (KDF_BAM_THP=1) removes most of those faults and measured ~5 %
    // faster in steady state, at the price of an occasional compaction stall of a second
    // on the first allocation: off by default.
    const size_t huge = 2u << 20;
    *cap = ((n < huge ? huge : n) + huge - 1) & ~(huge - 1);
    void* p = aligned_alloc(huge, *cap);
#ifdef MADV_HUGEPAGE
    if (p && thp_wanted()) madvise(p, *cap, MADV_HUGEPAGE);
#endif
    return (uint8_t*)p;
  }
  void put_buf(uint8_t* p, size_t cap) {
    if (!p) return;
    if (pool.size() >= 12) {
      free(p);
      return;
    }
    pool.emplace_back(p, cap);
  }
  ~Bam() {
    free(cur.data);
    free(next.data);
    for (auto& e : pool) free(e.first);
  }
};

// impl: 0 the library's own decoder (kdf_inflate.cpp), 1 zlib, -1 whichever KDF_BAM_ZLIB says
bool inflate_block(const uint8_t* src, uint32_t csize, uint8_t* dst, uint32_t usize, bool verify_crc = true,
                   int impl = -1) {
  // BGZF: 18-byte header (with BC subfield), deflate payload, crc32 + isize
  if (csize < 26) return false;
  uint16_t xlen = (uint16_t)(src[10] | (src[11] << 8));
  uint32_t hdr = 12 + xlen;
  if (hdr + 8 > csize) return false;
  static const bool env_zlib = getenv("KDF_BAM_ZLIB") != nullptr && atoi(getenv("KDF_BAM_ZLIB")) != 0;
  if (impl < 0) impl = env_zlib ? 1 : 0;
  if (impl == 0) {
    if (!kdf::inflate_raw(src + hdr, csize - hdr - 8, dst, usize)) return false;
  } else {
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, -15) != Z_OK) return false;
    zs.next_in = const_cast<uint8_t*>(src + hdr);
    zs.avail_in = csize - hdr - 8;
    zs.next_out = dst;
    zs.avail_out = usize;
    int rc = inflate(&zs, Z_FINISH);
    inflateEnd(&zs);
    if (!(rc == Z_STREAM_END && zs.total_out == usize)) return false;
  }
  if (verify_crc) {
    const uint8_t* tl = src + csize - 8;
    uint32_t want = tl[0] | (tl[1] << 8) | (tl[2] << 16) | ((uint32_t)tl[3] << 24);
    if (kdf::crc32_of(dst, usize) != want) return false;
  }
  return true;
}

// Stage 1 (serial): the compressed bytes of up to `want_bytes` of data.  The file is read
// in large slabs and the BGZF block headers walked in memory (one fread per block cost a
// fifth of the decode time); what follows the last block taken stays in b->pending.
// Errors go to `err` (this may run on a worker thread; g_host_err is thread-local).
bool read_comp(Bam* b, uint64_t want_bytes, CompBuf& cb, std::string& err) {
  auto& comp = cb.bytes;
  comp.clear();
  cb.blocks.clear();
  cb.total_u = 0;
  cb.ustart = b->u_total;
  cb.valid = false;
  comp.insert(comp.end(), b->pending.begin(), b->pending.end());
  b->pending.clear();
  // file reads: 8 MB at a time for a sequential pass, about one chunk's worth for region
  // fetches (kdf_bam_set_chunk_bytes of a few hundred KB) — what is read past the chunk is
  // copied to `pending` and back, which for an 8 MB slab cost more than the fetch itself
  const size_t SLAB = want_bytes / 2 > (8u << 20) ? (8u << 20) : (want_bytes / 2 < (128u << 10) ? (128u << 10) : (size_t)(want_bytes / 2));
  auto refill = [&]() -> bool {
    size_t at = comp.size();
    comp.resize(at + SLAB);
    size_t got = fread(comp.data() + at, 1, SLAB, b->fh);
    comp.resize(at + got);
    return got > 0;
  };
  size_t pos = 0;
  while (cb.total_u < want_bytes) {
    if (comp.size() - pos < 18) {
      if (refill()) continue;
      if (comp.size() == pos) {
        b->file_eof = true;
        break;
      }
      err = "not a BGZF block (is this a BAM file?)";
      return false;
    }
    const uint8_t* hdr = comp.data() + pos;
    if (hdr[0] != 31 || hdr[1] != 139 || hdr[2] != 8 || !(hdr[3] & 4)) {
      err = "not a BGZF block (is this a BAM file?)";
      return false;
    }
    uint16_t xlen = (uint16_t)(hdr[10] | (hdr[11] << 8));
    if (comp.size() - pos < (size_t)12 + xlen) {
      if (refill()) continue;
      err = "truncated BGZF header";
      return false;
    }
    // locate the BC subfield (normally the only one)
    const uint8_t* extra = hdr + 12;
    int bsize = -1;
    for (uint32_t p = 0; p + 4 <= xlen;) {
      uint16_t slen = (uint16_t)(extra[p + 2] | (extra[p + 3] << 8));
      if (extra[p] == 'B' && extra[p + 1] == 'C' && slen == 2 && p + 6 <= xlen)
        bsize = (extra[p + 4] | (extra[p + 5] << 8)) + 1;
      p += 4 + slen;
    }
    if (bsize < 0) {
      err = "BGZF block without BC subfield";
      return false;
    }
    if (bsize < 12 + (int)xlen + 8) {
      err = "corrupt BGZF block size";
      return false;
    }
    if (comp.size() - pos < (size_t)bsize) {
      if (refill()) continue;
      err = "truncated BGZF block";
      return false;
    }
    const uint8_t* tl = comp.data() + pos + bsize - 4;
    uint32_t isize = tl[0] | (tl[1] << 8) | (tl[2] << 16) | ((uint32_t)tl[3] << 24);
    if (isize > 65536) {
      err = "corrupt BGZF block (ISIZE above 64 KiB)";
      return false;
    }
    cb.blocks.push_back({(uint64_t)pos, (uint32_t)bsize, isize, cb.total_u});
    if (isize) b->blk_index.emplace_back(b->u_total, b->c_total);
    if (b->c_total == b->end_coff) b->u_limit = b->u_total + b->end_in;
    if (b->c_total == b->begin_coff) b->u_begin = b->u_total + b->begin_in;
    b->u_total += isize;
    b->c_total += (uint64_t)bsize;
    cb.total_u += isize;
    pos += (size_t)bsize;
  }
  b->pending.assign(comp.begin() + (long)pos, comp.end());
  cb.valid = !cb.blocks.empty();
  return true;
}

// a pooled buffer for the data of `cb` behind `gap` bytes of headroom
bool alloc_chunk(Bam* b, const CompBuf& cb, size_t gap, Chunk& c) {
  size_t cap = 0;
  uint8_t* p = b->get_buf(gap + cb.total_u + 1, &cap);
  if (!p) return false;
  c.data = p;
  c.own = c.head = gap;
  c.total = cb.total_u;
  c.ustart = cb.ustart;
  c.cap = cap;
  c.valid = false;   // until inflated
  return true;
}

// one block of `cb` into its place in `c` (any thread)
inline bool inflate_one(const CompBuf& cb, size_t i, Chunk& c, bool verify_crc) {
  const BlockRef& br = cb.blocks[i];
  if (br.usize == 0) return true;
  return inflate_block(cb.bytes.data() + br.coff, br.csize, c.data + c.own + br.uoff, br.usize, verify_crc);
}

// read + inflate `want_bytes` more, appended to a vector (header parsing only)
bool read_chunk(Bam* b, uint64_t want_bytes, std::vector<uint8_t>& out) {
  std::string err;
  CompBuf& cb = b->ahead;
  if (!read_comp(b, want_bytes, cb, err)) {
    g_host_err = err;
    return false;
  }
  if (!cb.valid) {
    b->eof = true;
    return true;
  }
  Chunk c;
  if (!alloc_chunk(b, cb, 0, c)) {
    g_host_err = "out of memory";
    return false;
  }
  int bad = 0;
#pragma omp parallel for schedule(dynamic, 8) num_threads(b->threads) reduction(| : bad)
  for (long i = 0; i < (long)cb.blocks.size(); ++i)
    if (!inflate_one(cb, (size_t)i, c, b->verify_crc)) bad |= 1;
  cb.valid = false;
  if (bad) {
    b->put_buf(c.data, c.cap);
    g_host_err = "BGZF inflate failed (corrupt block or CRC mismatch)";
    return false;
  }
  out.insert(out.end(), c.data + c.own, c.data + c.own + c.total);
  b->put_buf(c.data, c.cap);
  return true;
}

inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
  __builtin_ia32_pause();
#endif
}

inline int32_t rd_i32(const uint8_t* p) {
  int32_t v;
  memcpy(&v, p, 4);
  return v;
}
inline uint16_t rd_u16(const uint8_t* p) {
  uint16_t v;
  memcpy(&v, p, 2);
  return v;
}

bool parse_header(Bam* b) {
  // the header may span several blocks: keep reading until complete
  std::vector<uint8_t>& buf = b->carry;
  auto need = [&](size_t n) -> bool {
    while (buf.size() < n && !b->eof) {
      if (!read_chunk(b, 1 << 20, buf)) return false;
    }
    return buf.size() >= n;
  };
  if (!need(12) || memcmp(buf.data(), "BAM\1", 4) != 0) {
    if (g_host_err.empty()) g_host_err = "missing BAM magic";
    return false;
  }
  int32_t l_text = rd_i32(buf.data() + 4);
  if (l_text < 0 || l_text > (1 << 30)) {
    g_host_err = "corrupt BAM header (l_text)";
    return false;
  }
  if (!need(12 + (size_t)l_text)) return false;
  b->header_text.assign((const char*)buf.data() + 8, (size_t)l_text);
  size_t off = 8 + (size_t)l_text;
  int32_t n_ref = rd_i32(buf.data() + off);
  if (n_ref < 0 || n_ref > (1 << 24)) {
    g_host_err = "corrupt BAM header (n_ref)";
    return false;
  }
  off += 4;
  for (int32_t i = 0; i < n_ref; ++i) {
    if (!need(off + 4)) return false;
    int32_t l_name = rd_i32(buf.data() + off);
    if (l_name < 1 || l_name > (1 << 16)) {
      g_host_err = "corrupt BAM header (reference name length)";
      return false;
    }
    off += 4;
    if (!need(off + (size_t)l_name + 4)) return false;
    b->ref_names.emplace_back((const char*)buf.data() + off, (size_t)(l_name > 0 ? l_name - 1 : 0));
    off += (size_t)l_name;
    b->ref_lens.push_back(rd_i32(buf.data() + off));
    off += 4;
  }
  buf.erase(buf.begin(), buf.begin() + (long)off);
  b->first_rec_uoff = off;
  return true;
}

// nibble -> 2-bit code / validity; index = BAM 4-bit base (=ACMGRSVTWYHKDBN)
const uint8_t NIB_CODE[16] = {0, 0, 1, 0, 2, 0, 0, 0, 3, 0, 0, 0, 0, 0, 0, 0};
const uint8_t NIB_OK[16] = {0, 1, 1, 0, 1, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0};

inline void or_word64(uint64_t* w, uint64_t v, bool shared) {
  if (!v) return;
  if (shared)
    __atomic_fetch_or(w, v, __ATOMIC_RELAXED);
  else
    *w |= v;
}
inline void or_word32(uint32_t* w, uint32_t v, bool shared) {
  if (!v) return;
  if (shared)
    __atomic_fetch_or(w, v, __ATOMIC_RELAXED);
  else
    *w |= v;
}

// two nibble bytes (4 bases) -> 8 code bits (first base most significant) | 4 validity
// bits << 8; indexed by the two bytes as a little-endian u16
struct NibLut {
  uint16_t v[65536];
  NibLut() {
    for (unsigned i = 0; i < 65536; ++i) {
      unsigned b0 = i & 255, b1 = i >> 8;
      unsigned n[4] = {b0 >> 4, b0 & 15, b1 >> 4, b1 & 15};
      unsigned c = 0, ok = 0;
      for (int j = 0; j < 4; ++j) {
        c = (c << 2) | NIB_CODE[n[j]];
        ok = (ok << 1) | NIB_OK[n[j]];
      }
      v[i] = (uint16_t)(c | (ok << 8));
    }
  }
};
const NibLut NIB4;

// 32 bases (16 nibble bytes at `src`, `nb` of them meaningful) -> the aligned code word (first
// base most significant) and its validity word: eight look-ups of four bases each
inline void pack32_lut(const uint8_t* src, uint32_t nb, uint64_t* lw_out, uint32_t* lv_out) {
  uint64_t lw = 0;
  uint32_t lv = 0;
  const unsigned groups = (nb + 3) >> 2;
  for (unsigned g = 0; g < groups; ++g) {
    uint16_t two;
    memcpy(&two, src + 2 * g, 2);   // past an odd end this reads into the qualities: masked by the caller
    const uint16_t e = NIB4.v[two];
    lw |= (uint64_t)(e & 255) << (56 - 8 * g);
    lv |= (uint32_t)(e >> 8) << (28 - 4 * g);
  }
  *lw_out = lw;
  *lv_out = lv;
}

#if defined(__x86_64__)
// The same with byte shuffles as 16-entry tables (nibble -> code, nibble -> valid) and PEXT
// to squeeze the sixteen 4-bit code pairs / 2-bit validity pairs into their words: ~25
// instructions per 32 bases and no table in the data cache (the 128 KB look-up table of the
// scalar form does not fit L1).  All 16 bytes are read.
__attribute__((target("ssse3,bmi2"))) inline void pack32_simd(const uint8_t* src, uint64_t* lw_out, uint32_t* lv_out) {
  const __m128i code_lut = _mm_setr_epi8(0, 0, 1, 0, 2, 0, 0, 0, 3, 0, 0, 0, 0, 0, 0, 0);
  const __m128i ok_lut = _mm_setr_epi8(0, 1, 1, 0, 1, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0);
  const __m128i low4 = _mm_set1_epi8(0x0F);
  const __m128i v = _mm_loadu_si128((const __m128i*)src);
  const __m128i hi = _mm_and_si128(_mm_srli_epi16(v, 4), low4), lo = _mm_and_si128(v, low4);
  // per byte: bits 0-3 = code(first base) << 2 | code(second), bits 4-5 = ok(first) << 1 | ok(second)
  const __m128i c = _mm_or_si128(_mm_slli_epi16(_mm_shuffle_epi8(code_lut, hi), 2), _mm_shuffle_epi8(code_lut, lo));
  const __m128i k = _mm_or_si128(_mm_slli_epi16(_mm_shuffle_epi8(ok_lut, hi), 5), _mm_slli_epi16(_mm_shuffle_epi8(ok_lut, lo), 4));
  const __m128i t = _mm_or_si128(c, k);
  const uint64_t a = __builtin_bswap64((uint64_t)_mm_cvtsi128_si64(t));                       // bytes 0..7, byte 0 on top
  const uint64_t b = __builtin_bswap64((uint64_t)_mm_cvtsi128_si64(_mm_unpackhi_epi64(t, t)));   // bytes 8..15
  *lw_out = (_pext_u64(a, 0x0F0F0F0F0F0F0F0Full) << 32) | _pext_u64(b, 0x0F0F0F0F0F0F0F0Full);
  *lv_out = (uint32_t)((_pext_u64(a, 0x3030303030303030ull) << 16) | _pext_u64(b, 0x3030303030303030ull));
}
const bool g_pack_simd = __builtin_cpu_supports("ssse3") && __builtin_cpu_supports("bmi2") &&
                         getenv("KDF_PACK_SCALAR") == nullptr;
#else
const bool g_pack_simd = false;
#endif

// pack l_seq bases (BAM nibbles) at stream position `start`: 32 bases at a time into an
// aligned word, shifted into place; the first and the last stream word of a read are shared
// with its neighbours (atomic OR).  `readable`: bytes that may be read from `nib` on (the
// vector form reads 16 at a time; the end of a record's sequence is followed by its
// qualities, the end of a chunk by nothing)
void pack_record(const uint8_t* nib, uint32_t l_seq, uint64_t start, uint64_t* codes, uint32_t* valid,
                 size_t readable) {
  if (!l_seq) return;
  const uint64_t first_w = start >> 5, last_w = (start + l_seq - 1) >> 5;
  const unsigned s = (unsigned)(start & 31);
  uint64_t carry_c = 0;
  uint32_t carry_v = 0;
  uint64_t w = first_w;
  for (uint32_t done = 0; done < l_seq; done += 32, ++w) {
    const uint32_t nb = l_seq - done < 32 ? l_seq - done : 32;
    const uint8_t* src = nib + (done >> 1);
    uint64_t lw;
    uint32_t lv;
#if defined(__x86_64__)
    if (g_pack_simd && readable >= (size_t)(done >> 1) + 16)
      pack32_simd(src, &lw, &lv);
    else
#endif
      pack32_lut(src, nb, &lw, &lv);
    if (nb < 32) {
      lw &= ~0ull << (64 - 2 * nb);
      lv &= ~0u << (32 - nb);
    }
    uint64_t oc = carry_c | (lw >> (2 * s));
    uint32_t ov = carry_v | (lv >> s);
    carry_c = s ? lw << (64 - 2 * s) : 0;
    carry_v = s ? lv << (32 - s) : 0;
    const bool shared = (w == first_w) || (w == last_w);
    or_word64(codes + w, oc, shared);
    or_word32(valid + w, ov, shared);
  }
  if (w <= last_w) {   // the shifted remainder spills into one more word
    or_word64(codes + w, carry_c, true);
    or_word32(valid + w, carry_v, true);
  }
}

}  // namespace

// positions of the invalid bases (0 bits) of a validity bitmap, ascending
static uint64_t invalid_positions(const uint32_t* valid, uint64_t n_bases, uint32_t* out, uint64_t cap) {
  uint64_t n = 0;
  const uint64_t n_words = (n_bases + 31) / 32;
  for (uint64_t w = 0; w < n_words; ++w) {
    uint32_t inv = ~valid[w];
    if (w == n_words - 1 && (n_bases & 31)) inv &= ~0u << (32 - (n_bases & 31));   // bits past the end are not bases
    while (inv) {
      int b = __builtin_clz(inv);          // base i of the word sits at bit 31 - i
      if (out && n < cap) out[n] = (uint32_t)(w * 32 + (uint64_t)b);
      ++n;
      inv &= ~(0x80000000u >> b);
    }
  }
  return n;
}

struct kdf_bam_batch_impl {
  PodVec<uint64_t> codes;
  PodVec<uint32_t> valid;
  PodVec<uint32_t> invalid;   // positions of the invalid bases (sparse form of `valid`)
  PodVec<uint64_t> read_starts;
  PodVec<uint32_t> read_lens;
  PodVec<uint64_t> rec_index;
  PodVec<uint64_t> rec_uoff;     // offset of every kept record in the uncompressed stream
  PodVec<uint8_t> fasta_keep;    // the record is part of the KDF_BAM_FASTA stream
  PodVec<int32_t> ref_id, pos, next_ref_id, next_pos;
  PodVec<uint16_t> flag;
  PodVec<uint8_t> mapq;
  PodVec<uint64_t> qname_off, cigar_off, sa_off;  // n+1 each
  PodVec<char> qname_blob, sa_blob;
  PodVec<uint32_t> cigar_blob;
  PodVec<uint64_t> raw_off;    // n+1 (want_meta >= 3): the BAM records themselves
  PodVec<uint8_t> raw_blob;    //   (bytes after block_size), for BAM output
  PodVec<uint64_t> qual_off;   // n+1 (want_meta >= 2)
  PodVec<uint8_t> qual_blob;   // Phred base qualities, l_seq bytes per read
  uint64_t n_bases = 0;
  void reset() {
    codes.clear(), valid.clear(), invalid.clear(), read_starts.clear(), read_lens.clear(), rec_index.clear();
    rec_uoff.clear(), fasta_keep.clear(), ref_id.clear(), pos.clear(), next_ref_id.clear(), next_pos.clear();
    flag.clear(), mapq.clear(), qname_off.clear(), cigar_off.clear(), sa_off.clear(), qname_blob.clear();
    sa_blob.clear(), cigar_blob.clear(), raw_off.clear(), raw_blob.clear(), qual_off.clear(), qual_blob.clear();
    n_bases = 0;
  }
  size_t bytes() const {
    return codes.cap * 8 + valid.cap * 4 + invalid.cap * 4 + read_starts.cap * 8 + read_lens.cap * 4 +
           rec_index.cap * 8 + rec_uoff.cap * 8 + fasta_keep.cap + (ref_id.cap + pos.cap + next_ref_id.cap + next_pos.cap) * 4 +
           flag.cap * 2 + mapq.cap + (qname_off.cap + cigar_off.cap + sa_off.cap + raw_off.cap + qual_off.cap) * 8 +
           qname_blob.cap + sa_blob.cap + cigar_blob.cap * 4 + raw_blob.cap + qual_blob.cap;
  }
};

// Batches handed back (kdf_bam_batch_free) keep their buffers for the next kdf_bam_next_batch
// of any reader: a fresh page costs a fault (microseconds under virtualisation), a recycled
// one nothing.  KDF_BAM_POOL_MB bounds what is held (default 4096; 0: no pooling).
static std::mutex g_impl_mu;
static std::vector<kdf_bam_batch_impl*> g_impl_pool;
static size_t g_impl_pool_bytes = 0;

static kdf_bam_batch_impl* impl_get(size_t hint_bytes, bool want_meta) {
  {
    std::lock_guard<std::mutex> lk(g_impl_mu);
    if (!g_impl_pool.empty()) {
      // one of the same kind (with / without the metadata arrays) if there is one; among those
      // the smallest that is likely to hold the batch (a region fetch should not take the
      // buffers of a whole-file batch away from the decoder running beside it), else the largest
      size_t best = SIZE_MAX;
      for (int any_kind = 0; any_kind < 2 && best == SIZE_MAX; ++any_kind)
        for (size_t i = 0; i < g_impl_pool.size(); ++i) {
          if (!any_kind && (g_impl_pool[i]->ref_id.cap > 0) != want_meta) continue;
          if (best == SIZE_MAX) {
            best = i;
            continue;
          }
          const size_t a = g_impl_pool[i]->bytes(), c = g_impl_pool[best]->bytes();
          if (c >= hint_bytes ? (a >= hint_bytes && a < c) : a > c) best = i;
        }
      kdf_bam_batch_impl* im = g_impl_pool[best];
      g_impl_pool.erase(g_impl_pool.begin() + (long)best);
      g_impl_pool_bytes -= im->bytes();
      im->reset();
      return im;
    }
  }
  return new kdf_bam_batch_impl;
}

static void impl_put(kdf_bam_batch_impl* im) {
  if (!im) return;
  static const size_t limit = [] {
    const char* e = getenv("KDF_BAM_POOL_MB");
    return (size_t)(e ? atof(e) : 4096.0) << 20;
  }();
  const size_t sz = im->bytes();
  {
    std::lock_guard<std::mutex> lk(g_impl_mu);
    if (g_impl_pool_bytes + sz <= limit && g_impl_pool.size() < 16) {
      g_impl_pool.push_back(im);
      g_impl_pool_bytes += sz;
      return;
    }
  }
  delete im;
}

extern "C" {

int kdf_bam_open(const char* path, int n_threads, kdf_bam** out) {
  if (!path || !out) {
    g_host_err = "kdf_bam_open: NULL argument";
    return KDF_ERR_ARG;
  }
  FILE* fh = fopen(path, "rb");
  if (!fh) {
    g_host_err = std::string("cannot open ") + path;
    return KDF_ERR_ARG;
  }
  Bam* b = new (std::nothrow) Bam;
  if (!b) {
    fclose(fh);
    g_host_err = "out of memory";
    return KDF_ERR_ARG;
  }
  b->fh = fh;
  b->path = path;
  b->threads = n_threads > 0 ? n_threads : 1;
  if (const char* e = getenv("KDF_BAM_CRC")) b->verify_crc = e[0] != '0';
  if (const char* e = getenv("KDF_BAM_CHUNK_KB")) {
    long v = atol(e);
    if (v > 0) b->chunk_bytes = (uint64_t)v << 10;
  }
  if (const char* e = getenv("KDF_BAM_GAP")) {
    long v = atol(e);
    if (v >= 0) b->gap = (size_t)v;
  }
  g_host_err.clear();
  bool ok = false;
  try {
    ok = parse_header(b);
  } catch (const std::exception& e) {
    g_host_err = std::string("cannot parse the BAM header: ") + e.what();
  }
  if (!ok) {
    fclose(fh);
    delete b;
    return KDF_ERR_ARG;
  }
  *out = reinterpret_cast<kdf_bam*>(b);
  return KDF_OK;
}

void kdf_bam_close(kdf_bam* h) {
  Bam* b = reinterpret_cast<Bam*>(h);
  if (!b) return;
  if (b->fh) fclose(b->fh);
  if (getenv("KDF_BAM_TIMING") && b->tm.n_batches) {
    const Bam::Timing& t = b->tm;
    fprintf(stderr,
            "[kdf_bam] %s: %llu reads, %llu batches, %d threads; %llu rounds %.3f s (inflate+walk phase %.3f, "
            "select+layout %.3f, classify+pack %.3f) finish %.3f invalid-list %.3f\n",
            b->path.c_str(), (unsigned long long)t.n_reads, (unsigned long long)t.n_batches, b->threads,
            (unsigned long long)t.n_rounds, t.rounds, t.wait, t.select, t.fill, t.index, t.invalid);
  }
  delete b;
}

const char* kdf_host_last_error(void) { return g_host_err.c_str(); }

const char* kdf_bam_header_text(const kdf_bam* h, uint64_t* len) {
  const Bam* b = reinterpret_cast<const Bam*>(h);
  if (len) *len = b->header_text.size();
  return b->header_text.data();
}
int kdf_bam_n_refs(const kdf_bam* h) { return (int)reinterpret_cast<const Bam*>(h)->ref_names.size(); }
const char* kdf_bam_ref_name(const kdf_bam* h, int i) {
  const Bam* b = reinterpret_cast<const Bam*>(h);
  return (i >= 0 && i < (int)b->ref_names.size()) ? b->ref_names[i].c_str() : "";
}
int64_t kdf_bam_ref_len(const kdf_bam* h, int i) {
  const Bam* b = reinterpret_cast<const Bam*>(h);
  return (i >= 0 && i < (int)b->ref_lens.size()) ? b->ref_lens[i] : -1;
}

static int next_batch_impl(Bam* b, int mode, uint64_t max_bases, int want_meta, kdf_bam_batch* out,
                           kdf_bam_batch_impl* im);

int kdf_bam_next_batch(kdf_bam* h, int mode, uint64_t max_bases, int want_meta,
                       kdf_bam_batch* out) {
  Bam* b = reinterpret_cast<Bam*>(h);
  if (!b || !out) {
    g_host_err = "kdf_bam_next_batch: NULL argument";
    return KDF_ERR_ARG;
  }
  if (mode < KDF_BAM_FASTA || mode > KDF_BAM_ALL) {
    g_host_err = "kdf_bam_next_batch: bad mode";
    return KDF_ERR_ARG;
  }
  // no C++ exception may cross the C ABI: an allocation failure (a huge or corrupt
  // input) comes back as an error code
  try {
    std::unique_ptr<kdf_bam_batch_impl, void (*)(kdf_bam_batch_impl*)> im(impl_get(max_bases ? (size_t)(max_bases / 32) * 12 : SIZE_MAX, want_meta != 0), impl_put);
    int rc = next_batch_impl(b, mode, max_bases, want_meta, out, im.get());
    if (rc == KDF_OK) im.release();   // owned by *out until kdf_bam_batch_free
    return rc;
  } catch (const std::bad_alloc&) {
    g_host_err = "out of memory while decoding the BAM";
  } catch (const std::exception& e) {
    g_host_err = std::string("BAM decode failed: ") + e.what();
  } catch (...) {
    g_host_err = "BAM decode failed";
  }
  memset(out, 0, sizeof(*out));
  return KDF_ERR_ARG;
}

// ---- the decode pipeline -------------------------------------------------------------
//
// One call of kdf_bam_next_batch runs ROUNDS; a round is one OpenMP parallel region over two
// chunks of ~64 MB of records:
//
//   consume(cur)   select (serial, over compact per-record arrays only): the `samtools fasta`
//                  collapse state, the keep decision, the batch limit; then, by all threads, the
//                  kept records are packed straight into the batch (2-bit codes + validity,
//                  fixed fields, QNAME / CIGAR / SA / quality / raw blobs).  A chunk is recycled
//                  as soon as its records are consumed: nothing is held back until the batch
//                  is complete.
//   produce(next)  every thread inflates BGZF blocks (kdf_inflate.cpp) and, while a block is
//                  still in its cache, continues the RECORD WALK through it: the chain of
//                  block_size fields is the only way to find BAM records and is inherently
//                  serial, so it is handed from block to block (chain[i] = where the walk
//                  enters block i; a thread waits for the walk of the block before — a
//                  microsecond against the ~50 us of an inflate).  Then, in parallel, each
//                  record's header is validated and classified (flag, l_seq, "same QNAME as
//                  the primary record before it").
//   read(ahead)    one thread reads the compressed bytes of the chunk after that.
//
// A record that is not complete in its chunk (it straddles the boundary, or only its
// block_size field does) is copied into the headroom in front of the next chunk and walked
// there; what is not consumed (the rest of cur, the produced next, the read-ahead) stays in the
// reader for the next call.

constexpr uint8_t C_PRIMARY = 1, C_SAME = 2, C_BAD = 4;
// (produce only) the header / the QNAME comparison could not be done while the block was hot
constexpr uint8_t C_PENDING = 0x40, C_PENDING_SAME = 0x20;
constexpr uint32_t NO_PREV = 0xffffffffu;
constexpr int64_t CHAIN_WAIT = INT64_MIN, CHAIN_ABORT = INT64_MIN + 1;

static inline const uint8_t* find_sa_tag(const uint8_t* t, const uint8_t* end, size_t* len) {
  *len = 0;
  while (t + 3 <= end) {   // walk the aux tags up to the SA:Z tag (a record has one at most)
    char t0 = (char)t[0], t1 = (char)t[1], ty = (char)t[2];
    t += 3;
    size_t adv = 0;
    switch (ty) {
      case 'A': case 'c': case 'C': adv = 1; break;
      case 's': case 'S': adv = 2; break;
      case 'i': case 'I': case 'f': adv = 4; break;
      case 'Z': case 'H': {
        const uint8_t* z = (const uint8_t*)memchr(t, 0, (size_t)(end - t));
        if (!z) return nullptr;
        if (t0 == 'S' && t1 == 'A' && ty == 'Z') {
          *len = (size_t)(z - t);
          return t;
        }
        adv = (size_t)(z - t) + 1;
        break;
      }
      case 'B': {
        if (t + 5 > end) return nullptr;
        char sub = (char)t[0];
        uint32_t cnt;
        memcpy(&cnt, t + 1, 4);
        size_t sz = (sub == 'c' || sub == 'C') ? 1 : ((sub == 's' || sub == 'S') ? 2 : 4);
        adv = 5 + sz * cnt;
        break;
      }
      default: return nullptr;
    }
    if (t >= end) break;
    t += adv;
  }
  return nullptr;
}

static void recycle_chunk(Bam* b, Chunk& c) {
  if (c.data) b->put_buf(c.data, c.cap);
  c.data = nullptr;
  c.cap = 0;
  c.valid = false;
  c.n_rec = c.sel = 0;
}

static std::atomic<int> g_active_decoders{0};   // readers inside kdf_bam_next_batch right now

// A record's header: validate the variable-length fields against the record size `bs` (a
// corrupt l_seq / n_cigar / l_read_name would otherwise send the packer and the metadata pass
// out of bounds) and classify it.  → C_BAD, C_PRIMARY or 0
static inline uint8_t classify_header(const uint8_t* r, uint64_t bs, uint16_t* flag_out, uint32_t* lseq_out) {
  const uint16_t flag = rd_u16(r + 14);
  const uint8_t l_name = r[8];
  const uint16_t n_cig_v = rd_u16(r + 12);
  const int32_t l_seq_s = rd_i32(r + 16);
  *flag_out = flag;
  *lseq_out = (uint32_t)l_seq_s;
  if (l_seq_s < 0 || 32ull + l_name + 4ull * n_cig_v + ((uint64_t)l_seq_s + 1) / 2 + (uint64_t)l_seq_s > bs)
    return C_BAD;
  return (flag & 0xD00) ? 0 : C_PRIMARY;
}

static int next_batch_impl(Bam* b, int mode, uint64_t max_bases, int want_meta, kdf_bam_batch* out,
                           kdf_bam_batch_impl* im) {
  const uint64_t CHUNK_BYTES = b->chunk_bytes;
  const size_t GAP = b->gap;
  const int nthr = b->threads > 0 ? b->threads : 1;   // the most a round may use
  // Several readers decode at once (the discovery pipeline runs the child and both parents
  // together): each round then takes its share of the cores instead of every reader
  // starting a full team (three teams on one set of cores spend their time in barriers
  // and in the hand-over of the record walk).
  struct ActiveGuard {
    ActiveGuard() { g_active_decoders.fetch_add(1, std::memory_order_relaxed); }
    ~ActiveGuard() { g_active_decoders.fetch_sub(1, std::memory_order_relaxed); }
  } active_guard;
  static const int n_procs = omp_get_num_procs() > 0 ? omp_get_num_procs() : 1;
  auto team_size = [&](long blocks, size_t records) {
    const int act = g_active_decoders.load(std::memory_order_relaxed);
    int t = act > 1 ? (n_procs + act - 1) / act : nthr;
    if (t < 2) t = 2;
    if (t > nthr) t = nthr;
    // a region fetch inflates a handful of blocks per round: a full team would spend the
    // round in its barriers
    const long work = blocks / 16 + (long)(records / 16384);   // (~1 ms of work per thread at least)
    if (work < t) t = work < 1 ? 1 : (int)work;
    return t;
  };
  auto bail = [&](const std::string& msg) {
    g_host_err = msg;
    return KDF_ERR_ARG;
  };
  // ---- the batch under construction ----
  uint64_t n_bases = 0;      // stream bases so far (reads are separated by one invalid base)
  size_t n_kept = 0;         // reads so far
  uint64_t words_zeroed = 0; // stream words [0, words_zeroed) are initialised
  if (max_bases) {           // address space only: pages are touched as the batch grows
    im->codes.reserve((max_bases + 31) / 32 + 2);
    im->valid.reserve((max_bases + 31) / 32 + 2);
  }
  auto grow = [](auto& v, size_t n) { v.resize(n); };   // (amortised, new elements untouched)
  im->qname_off.assign(1, 0);
  im->cigar_off.assign(1, 0);
  im->sa_off.assign(1, 0);
  im->qual_off.assign(1, 0);
  im->raw_off.assign(1, 0);
  // selection of the current round (indices into cur's record arrays)
  std::vector<uint32_t>& s_idx = b->s_idx;
  std::vector<uint64_t>& s_start = b->s_start;
  std::vector<uint8_t>& s_fk = b->s_fk;
  std::vector<const uint8_t*>& s_sa = b->s_sa;
  bool done = false, hit_limit = false;
  std::string perr, rerr;

  Chunk& cur = b->cur;
  Chunk& next = b->next;
  if (b->range_done) {
    done = true;   // the end of the range was delivered by an earlier call
    hit_limit = true;
  }

  // select: the serial part of consume(cur)
  auto select = [&]() {
    s_idx.clear();
    s_start.clear();
    s_fk.clear();
    if (!cur.valid) return;
    size_t i = cur.sel;
    for (; i < cur.n_rec; ++i) {
      const uint64_t uoff = cur.ustart + (uint64_t)cur.w_off[i];   // (two's complement: w_off may be negative)
      if (uoff >= b->u_limit.load(std::memory_order_relaxed)) {   // end of this reader's range
        hit_limit = true;
        break;
      }
      const uint8_t cls = cur.w_cls[i];
      if (cls & C_BAD) {
        perr = "corrupt BAM record (field sizes exceed the record)";
        break;
      }
      const uint16_t flag = cur.w_flag[i];
      const uint32_t l_seq = cur.w_lseq[i];
      // membership of the `samtools fasta -F 0xD00` stream, tracked in every mode (the
      // discovery pipeline decodes the child ONCE in scan mode and masks the counting
      // stream with this flag)
      bool fasta_keep = false;
      unsigned seen = b->seen_parts;
      if (cls & C_PRIMARY) {
        if (!(cls & C_SAME)) seen = 0;
        bool r1 = flag & 0x40, r2 = flag & 0x80;
        unsigned part = (r1 && !r2) ? 1u : ((r2 && !r1) ? 2u : 0u);
        if (!(seen & (1u << part))) {
          seen |= 1u << part;
          fasta_keep = true;
        }
      }
      bool keep = true;
      if (mode == KDF_BAM_FASTA) {
        keep = fasta_keep;
      } else if (mode == KDF_BAM_SCAN) {
        if (flag & 0x500) keep = false;
      }
      // warm-up records of a range (kdf_bam_set_begin): parsed for the QNAME-run state only
      if (uoff < b->u_begin.load(std::memory_order_relaxed)) keep = false;
      if (keep) {
        if (max_bases && n_kept + s_idx.size() > 0 && n_bases + l_seq + 1 > max_bases) {
          done = true;
          break;  // this record waits for the next batch: the collapse state is as before it
        }
        const uint64_t start = (n_kept + s_idx.size()) ? n_bases + 1 : 0;
        s_idx.push_back((uint32_t)i);
        s_start.push_back(start);
        s_fk.push_back((uint8_t)(fasta_keep ? 1 : 0));
        n_bases = start + l_seq;
      }
      if (cls & C_PRIMARY) b->seen_parts = seen;
    }
    cur.sel = i;
    if (i == cur.n_rec && !hit_limit && !done && perr.empty() && !cur.walk_err.empty()) perr = cur.walk_err;
  };

  std::vector<std::atomic<int64_t>>& chain = b->chain;
  std::vector<std::vector<int64_t>>& tl_off = b->tl_off;
  if ((int)tl_off.size() < nthr) {
    tl_off.resize((size_t)nthr);
    b->tl_flag.resize((size_t)nthr);
    b->tl_lseq.resize((size_t)nthr);
    b->tl_cls.resize((size_t)nthr);
  }
  static const bool hot_classify = !(getenv("KDF_BAM_WALK_CLASSIFY") && atoi(getenv("KDF_BAM_WALK_CLASSIFY")) == 0);

  while (!done) {
    // ---- what this round does ----
    const bool cur_live = cur.valid && cur.sel < cur.n_rec;
    if (!b->ahead.valid && !next.valid && !b->file_eof) {   // cold pipeline: nothing read ahead yet
      if (!read_comp(b, CHUNK_BYTES, b->ahead, rerr)) return bail(rerr);
    }
    // (a file whose records all followed the header in its first blocks has no chunk to
    // inflate: what parse_header left in `carry` is then walked as a chunk of its own)
    const bool carry_only = !cur.valid && !next.valid && !b->ahead.valid && !b->carry.empty();
    if (carry_only) {
      b->ahead.blocks.clear();
      b->ahead.total_u = 0;
      b->ahead.ustart = b->u_total;
    }
    const bool do_produce = !next.valid && (b->ahead.valid || carry_only);
    if (!cur_live && !do_produce && !next.valid) {
      // nothing left anywhere: the file (or what was read of it) ends here
      if (cur.valid) {
        if (!cur.walk_err.empty()) return bail(cur.walk_err);
        if (cur.exit_off < (int64_t)cur.total) return bail("truncated BAM file (incomplete record at the end)");
      }
      break;
    }
    if (!cur_live && !do_produce) {   // next is ready and cur is spent: step
      if (cur.valid && !cur.walk_err.empty()) return bail(cur.walk_err);
      recycle_chunk(b, cur);
      std::swap(cur, next);
      continue;
    }
    const double t_round = omp_get_wtime();
    // ---- produce(next): buffer, the straddling bytes in front of it, the chain entry ----
    long n_blk = 0;
    if (do_produce) {
      const uint8_t* tail = nullptr;
      size_t tail_len = 0;
      int64_t entry;
      if (cur.valid) {
        tail = cur.data + cur.own + cur.exit_off;   // (exit_off may be negative: still in cur's own head)
        tail_len = (size_t)((int64_t)cur.total - cur.exit_off);
        entry = -(int64_t)tail_len;
      } else if (!b->carry.empty()) {   // first chunk of the file: what followed the header
        tail = b->carry.data();
        tail_len = b->carry.size();
        entry = -(int64_t)tail_len;
      } else {
        entry = (int64_t)b->skip_bytes;   // after kdf_bam_seek: the record starts inside the first block
      }
      const size_t head = tail_len > GAP ? tail_len : GAP;
      size_t cap = 0;
      uint8_t* p = b->get_buf(head + b->ahead.total_u + 1, &cap);
      if (!p) return bail("out of memory");
      next.data = p;
      next.cap = cap;
      next.own = head;
      next.head = head - tail_len;
      next.total = b->ahead.total_u;
      next.ustart = b->ahead.ustart;
      if (tail_len) memcpy(p + next.head, tail, tail_len);
      if (!cur.valid) {
        if (b->skip_bytes) {
          if (b->ahead.blocks.empty() || b->skip_bytes > b->ahead.blocks[0].usize) {
            recycle_chunk(b, next);
            return bail("kdf_bam_seek: offset beyond the end of its block");
          }
          b->skip_bytes = 0;
        }
        b->carry.clear();
        next.prev_qname.clear();   // nothing precedes the first chunk of a file / of a seek
        next.rec_base = b->record_index;
      } else {
        next.prev_qname = cur.last_qname;
        next.rec_base = cur.rec_base + cur.n_rec;
      }
      n_blk = (long)b->ahead.blocks.size();
      if (chain.size() < (size_t)n_blk + 1) {
        std::vector<std::atomic<int64_t>> bigger((size_t)n_blk + 1 + 256);
        chain.swap(bigger);
      }
      for (long i = 0; i <= n_blk; ++i) chain[(size_t)i].store(CHAIN_WAIT, std::memory_order_relaxed);
      chain[0].store(entry, std::memory_order_relaxed);
      b->blk_first.assign((size_t)n_blk, 0);
      b->blk_count.assign((size_t)n_blk, 0);
      b->blk_owner.assign((size_t)n_blk, 0);
      for (int t = 0; t < nthr; ++t) {
        tl_off[(size_t)t].clear();
        b->tl_flag[(size_t)t].clear();
        b->tl_lseq[(size_t)t].clear();
        b->tl_cls[(size_t)t].clear();
      }
      next.walk_err.clear();
    }
    b->ahead2.valid = false;
    const bool do_read = !b->file_eof && (do_produce || !b->ahead.valid);
    bool read_ok = true;
    // bit 0: a block failed to inflate; bit 1: an allocation failed inside the region (no C++
    // exception may leave an OpenMP construct) — seen by every thread after the barrier
    std::atomic<int> bad{0};
    std::atomic<int> walk_corrupt{0};   // a walker met a corrupt block_size
    std::string werr;
    size_t round_kept = 0;
    uint64_t w_first = 0, w_last = 0;   // stream words this round initialises: [w_first, w_last)
    const CompBuf& cb = b->ahead;
    uint8_t* const nbase = do_produce ? next.data + next.own : nullptr;
    const int64_t ntotal = (int64_t)next.total;
    std::vector<uint64_t> part_sum;   // per-thread partial sums of the variable-length sizes
    size_t blocks_rec = 0;            // records the block walkers found (the rest: the finishing walk)
#pragma omp parallel num_threads(team_size(n_blk, cur_live ? cur.n_rec - cur.sel : 0))
    {
      const int tid = omp_get_thread_num();
      const int team = omp_get_num_threads();
#pragma omp single nowait
      {
        try {
          if (do_read) read_ok = read_comp(b, CHUNK_BYTES, b->ahead2, rerr);
        } catch (...) {
          read_ok = false;
          rerr = "out of memory while reading the BAM";
        }
      }
#pragma omp single nowait
      {
        const double t0 = omp_get_wtime();
        try {
          select();
        } catch (...) {
          bad.fetch_or(2, std::memory_order_relaxed);
        }
        b->tm.select += omp_get_wtime() - t0;
      }
      // inflate + record walk, block by block
#pragma omp for schedule(dynamic, 1) nowait
      for (long i = 0; i < n_blk; ++i) {
        const BlockRef& br = cb.blocks[(size_t)i];
        bool ok = true;
        if (br.usize)
          ok = inflate_block(cb.bytes.data() + br.coff, br.csize, nbase + br.uoff, br.usize, b->verify_crc);
        if (!ok) bad.fetch_or(1, std::memory_order_relaxed);
        int64_t o;
        for (unsigned spins = 0;; ++spins) {
          o = chain[(size_t)i].load(std::memory_order_acquire);
          if (o != CHAIN_WAIT) break;
          if (spins < 2000)
            cpu_relax();
          else
            sched_yield();
        }
        if (!ok || o == CHAIN_ABORT) {
          chain[(size_t)i + 1].store(CHAIN_ABORT, std::memory_order_release);
          continue;
        }
        std::vector<int64_t>& mine = tl_off[(size_t)tid];
        const size_t first = mine.size();
        const int64_t u_end = (int64_t)(br.uoff + br.usize);
        bool corrupt = false;
        try {
          while (o + 4 <= u_end) {
            const int32_t bs = rd_i32(nbase + o);
            if (bs < 32 || bs > (1 << 28)) {
              corrupt = true;
              break;   // the chain stops here: `o` is handed on unchanged and nobody can advance it
            }
            if (o + 4 + (int64_t)bs > ntotal) break;   // not complete in this chunk: it goes in front of the next
            mine.push_back(o);
            o += 4 + (int64_t)bs;
          }
        } catch (...) {
          bad.fetch_or(2, std::memory_order_relaxed);
          chain[(size_t)i + 1].store(CHAIN_ABORT, std::memory_order_release);
          continue;
        }
        if (corrupt) walk_corrupt.store(1, std::memory_order_relaxed);
        b->blk_owner[(size_t)i] = (uint16_t)tid;
        b->blk_first[(size_t)i] = (uint32_t)first;
        b->blk_count[(size_t)i] = (uint32_t)(mine.size() - first);
        chain[(size_t)i + 1].store(o, std::memory_order_release);
        // Off the hand-over path, while the block is still in this core's cache: the headers
        // that lie entirely in what is inflated so far (everything below u_end) are validated
        // and classified here, and a primary record is compared with the primary record before
        // it when that one is in this block too.  What cannot be decided yet (a header or QNAME
        // that continues in the next block, the first primary record of a block) is marked and
        // left to the passes after the barrier.
        try {
          std::vector<uint16_t>& fl = b->tl_flag[(size_t)tid];
          std::vector<uint32_t>& ls = b->tl_lseq[(size_t)tid];
          std::vector<uint8_t>& cl = b->tl_cls[(size_t)tid];
          const size_t cnt = mine.size() - first;
          fl.resize(first + cnt);
          ls.resize(first + cnt);
          cl.resize(first + cnt);
          const uint8_t* prev_name = nullptr;   // QNAME of the previous primary record of this block, if readable
          size_t prev_len = 0;
          bool prev_known = false;              // a primary record came before in this block
          for (size_t j = first; j < first + cnt; ++j) {
            const int64_t ro = mine[j];
            const int64_t next_ro = j + 1 < first + cnt ? mine[j + 1] : o;
            uint8_t cls = C_PENDING;
            uint16_t flag = 0;
            uint32_t lseq = 0;
            if (hot_classify && ro + 36 <= u_end) {
              const uint8_t* r = nbase + ro + 4;
              cls = classify_header(r, (uint64_t)(next_ro - ro) - 4, &flag, &lseq);
              if (cls & C_PRIMARY) {
                const size_t ql = r[8] ? (size_t)r[8] - 1 : 0;
                const bool name_here = ro + 36 + (int64_t)r[8] <= u_end;
                if (name_here && prev_known && prev_name) {
                  if (prev_len == ql && memcmp(prev_name, r + 32, ql) == 0) cls |= C_SAME;
                } else {
                  cls |= C_PENDING_SAME;
                }
                prev_known = true;
                prev_name = name_here ? r + 32 : nullptr;
                prev_len = ql;
              } else if (cls & C_BAD) {
                prev_known = false;   // (the passes behind stop at a corrupt record anyway)
                prev_name = nullptr;
              }
            } else {
              // unknown kind: it may be a primary record, so the next one cannot rely on prev_*
              prev_known = false;
              prev_name = nullptr;
            }
            fl[j] = flag;
            ls[j] = lseq;
            cl[j] = cls;
          }
        } catch (...) {
          bad.fetch_or(2, std::memory_order_relaxed);
        }
      }
#pragma omp barrier
      // ---- both selections are known: lay out the outputs (one thread), then fill them ----
#pragma omp single
      {
        const double t0 = omp_get_wtime();
        b->tm.wait += t0 - t_round;
        try {
          if (bad.load(std::memory_order_relaxed) & 2) throw std::bad_alloc();   // (the selection is incomplete)
          if (walk_corrupt.load(std::memory_order_relaxed)) werr = "corrupt BAM record (block_size out of range)";
          if (do_produce && !bad) {
            const int64_t ex = chain[(size_t)n_blk].load(std::memory_order_acquire);
            size_t n = 0;
            for (long i = 0; i < n_blk; ++i) {
              const uint32_t c = b->blk_count[(size_t)i];
              b->blk_count[(size_t)i] = (uint32_t)n;   // from here on: index of the block's first record
              n += c;
            }
            // what the last walker could not reach: records that lie entirely in the head (a file
            // whose records all followed the header in its first blocks), or behind empty blocks
            b->fin_off.clear();
            int64_t fo = ex;
            if (werr.empty())
              while (fo + 4 <= ntotal) {
                const int32_t bs = rd_i32(nbase + fo);
                if (bs < 32 || bs > (1 << 28)) {
                  werr = "corrupt BAM record (block_size out of range)";
                  break;
                }
                if (fo + 4 + (int64_t)bs > ntotal) break;
                b->fin_off.push_back(fo);
                fo += 4 + (int64_t)bs;
              }
            const size_t n_blocks_rec = n;
            n += b->fin_off.size();
            next.n_rec = n;
            next.sel = 0;
            next.exit_off = fo;
            next.walk_err = werr;
            next.w_off.resize(n + 1);
            next.w_off[n] = fo;
            if (!b->fin_off.empty())
              memcpy(next.w_off.data() + n_blocks_rec, b->fin_off.data(), b->fin_off.size() * sizeof(int64_t));
            blocks_rec = n_blocks_rec;
            next.w_flag.resize(n);
            next.w_lseq.resize(n);
            next.w_cls.resize(n);
            next.w_prevp.resize(n);
          }
          round_kept = s_idx.size();
          if (round_kept) {
            const size_t n1 = n_kept + round_kept;
            grow(im->read_starts, n1);
            grow(im->read_lens, n1);
            grow(im->rec_index, n1);
            grow(im->rec_uoff, n1);
            grow(im->fasta_keep, n1);
            if (want_meta) {
              grow(im->ref_id, n1);
              grow(im->pos, n1);
              grow(im->next_ref_id, n1);
              grow(im->next_pos, n1);
              grow(im->flag, n1);
              grow(im->mapq, n1);
              grow(im->qname_off, n1 + 1);
              grow(im->cigar_off, n1 + 1);
              grow(im->sa_off, n1 + 1);
              if (want_meta >= 2) grow(im->qual_off, n1 + 1);
              if (want_meta >= 3) grow(im->raw_off, n1 + 1);
              if (s_sa.size() < round_kept) s_sa.resize(round_kept);
            }
            w_first = words_zeroed;
            w_last = (n_bases + 31) / 32;
            if (w_last < w_first) w_last = w_first;
            grow(im->codes, (size_t)w_last + 1);
            grow(im->valid, (size_t)w_last + 1);
            words_zeroed = w_last;
          }
          part_sum.assign((size_t)team * 5 + 5, 0);
        } catch (...) {
          bad.fetch_or(2, std::memory_order_relaxed);
          round_kept = 0;
        }
        b->tm.select += omp_get_wtime() - t0;
      }
      // (implicit barrier)
      const double t_fill = omp_get_wtime();
      // gather the record offsets of the produced chunk, block by block
      if (do_produce && !bad) {
#pragma omp for schedule(static) nowait
        for (long i = 0; i < n_blk; ++i) {
          const size_t at = b->blk_count[(size_t)i];
          const size_t end = (i + 1 < n_blk) ? b->blk_count[(size_t)i + 1] : blocks_rec;
          const size_t ow = b->blk_owner[(size_t)i], f0 = b->blk_first[(size_t)i];
          if (end > at) {
            memcpy(next.w_off.data() + at, tl_off[ow].data() + f0, (end - at) * sizeof(int64_t));
            memcpy(next.w_flag.data() + at, b->tl_flag[ow].data() + f0, (end - at) * sizeof(uint16_t));
            memcpy(next.w_lseq.data() + at, b->tl_lseq[ow].data() + f0, (end - at) * sizeof(uint32_t));
            memcpy(next.w_cls.data() + at, b->tl_cls[ow].data() + f0, (end - at) * sizeof(uint8_t));
          }
        }
#pragma omp single nowait
        {
          for (size_t i = blocks_rec; i < next.n_rec; ++i) next.w_cls[i] = C_PENDING;   // (the finishing walk's)
        }
      }
      // initialise the stream words this round reaches
      if (round_kept) {
        uint64_t* cw = im->codes.data();
        uint32_t* vw = im->valid.data();
#pragma omp for schedule(static) nowait
        for (long w = (long)w_first; w < (long)w_last; ++w) {
          cw[w] = 0;
          vw[w] = 0;
        }
      }
#pragma omp barrier
      // classify what the walkers had to leave (a header that crossed a block boundary)
      if (do_produce && !bad) {
        const size_t n = next.n_rec;
#pragma omp for schedule(static) nowait
        for (long i = 0; i < (long)n; ++i) {
          if (!(next.w_cls[(size_t)i] & C_PENDING)) continue;
          const uint8_t* r = nbase + next.w_off[(size_t)i] + 4;
          const uint64_t bs = (uint64_t)(next.w_off[(size_t)i + 1] - next.w_off[(size_t)i]) - 4;
          uint16_t flag;
          uint32_t lseq;
          uint8_t cls = classify_header(r, bs, &flag, &lseq);
          if (cls & C_PRIMARY) cls |= C_PENDING_SAME;
          next.w_flag[(size_t)i] = flag;
          next.w_lseq[(size_t)i] = lseq;
          next.w_cls[(size_t)i] = cls;
        }
      }
      // pack the selected records of cur: bases, fixed fields, sizes of the variable ones
      if (round_kept) {
        const uint8_t* cbase = cur.data + cur.own;
        uint64_t sums[5] = {0, 0, 0, 0, 0};
#pragma omp for schedule(static) nowait
        for (long j = 0; j < (long)round_kept; ++j) {
          const size_t i = s_idx[(size_t)j];
          const size_t g = n_kept + (size_t)j;
          const uint8_t* r = cbase + cur.w_off[i] + 4;
          if ((size_t)j + 4 < round_kept) __builtin_prefetch(cbase + cur.w_off[s_idx[(size_t)j + 4]] + 4);
          const uint64_t bs = (uint64_t)(cur.w_off[i + 1] - cur.w_off[i]) - 4;
          const uint8_t l_name = r[8];
          const uint16_t n_cig = rd_u16(r + 12);
          const uint32_t l_seq = cur.w_lseq[i];
          const uint8_t* nib = r + 32 + l_name + 4 * (size_t)n_cig;
          pack_record(nib, l_seq, s_start[(size_t)j], im->codes.data(), im->valid.data(), (size_t)(r + bs - nib));
          im->read_starts[g] = s_start[(size_t)j];
          im->read_lens[g] = l_seq;
          im->rec_index[g] = cur.rec_base + i;
          im->rec_uoff[g] = cur.ustart + (uint64_t)cur.w_off[i];
          im->fasta_keep[g] = s_fk[(size_t)j];
          if (want_meta) {
            im->ref_id[g] = rd_i32(r);
            im->pos[g] = rd_i32(r + 4);
            im->mapq[g] = r[9];
            im->flag[g] = rd_u16(r + 14);
            im->next_ref_id[g] = rd_i32(r + 20);
            im->next_pos[g] = rd_i32(r + 24);
            const uint64_t ql = l_name ? (uint64_t)l_name - 1 : 0;
            size_t sl = 0;
            s_sa[(size_t)j] = find_sa_tag(nib + (l_seq + 1) / 2 + l_seq, r + bs, &sl);
            // (sizes for now; turned into offsets below)
            im->qname_off[g + 1] = ql;
            im->cigar_off[g + 1] = n_cig;
            im->sa_off[g + 1] = sl;
            sums[0] += ql, sums[1] += n_cig, sums[2] += sl;
            if (want_meta >= 2) im->qual_off[g + 1] = l_seq, sums[3] += l_seq;
            if (want_meta >= 3) im->raw_off[g + 1] = bs, sums[4] += bs;
          }
        }
        for (int q = 0; q < 5; ++q) part_sum[(size_t)tid * 5 + q] = sums[q];
      }
#pragma omp barrier
#pragma omp single
      {
        try {
          if (do_produce && !bad) {   // link every primary record to the one before it
            uint32_t prev = NO_PREV;
            std::string& lq = next.last_qname;
            lq = next.prev_qname;
            const size_t n = next.n_rec;
            for (size_t i = 0; i < n; ++i) {
              if (next.w_cls[i] & C_BAD) break;
              if (next.w_cls[i] & C_PRIMARY) {
                next.w_prevp[i] = prev;
                prev = (uint32_t)i;
              }
            }
            if (prev != NO_PREV) {
              const uint8_t* r = nbase + next.w_off[prev] + 4;
              lq.assign((const char*)r + 32, r[8] ? (size_t)r[8] - 1 : 0);
            }
          }
          if (round_kept && want_meta) {   // exclusive prefix of the per-thread sums; blob sizes
            uint64_t run[5] = {im->qname_off[n_kept], im->cigar_off[n_kept], im->sa_off[n_kept],
                               want_meta >= 2 ? im->qual_off[n_kept] : 0, want_meta >= 3 ? im->raw_off[n_kept] : 0};
            for (int t = 0; t < team; ++t)
              for (int q = 0; q < 5; ++q) {
                const uint64_t v = part_sum[(size_t)t * 5 + q];
                part_sum[(size_t)t * 5 + q] = run[q];
                run[q] += v;
              }
            grow(im->qname_blob, (size_t)run[0]);
            grow(im->cigar_blob, (size_t)run[1]);
            grow(im->sa_blob, (size_t)run[2]);
            if (want_meta >= 2) grow(im->qual_blob, (size_t)run[3]);
            if (want_meta >= 3) grow(im->raw_blob, (size_t)run[4]);
          }
        } catch (...) {
          bad.fetch_or(2, std::memory_order_relaxed);
        }
      }
      // (implicit barrier)
      if (do_produce && !bad) {   // "same QNAME as the primary record before it"
        const size_t n = next.n_rec;
#pragma omp for schedule(static) nowait
        for (long i = 0; i < (long)n; ++i) {
          if (!(next.w_cls[(size_t)i] & C_PENDING_SAME)) continue;
          next.w_cls[(size_t)i] &= (uint8_t)~C_PENDING_SAME;
          const uint8_t* r = nbase + next.w_off[(size_t)i] + 4;
          const char* qn = (const char*)r + 32;
          const size_t ql = r[8] ? (size_t)r[8] - 1 : 0;
          bool same;
          const uint32_t pv = next.w_prevp[(size_t)i];
          if (pv == NO_PREV) {   // the run may continue from the chunk before
            same = next.prev_qname.size() == ql && memcmp(next.prev_qname.data(), qn, ql) == 0;
          } else {
            const uint8_t* pr = nbase + next.w_off[pv] + 4;
            const size_t pl = pr[8] ? (size_t)pr[8] - 1 : 0;
            same = pl == ql && memcmp(pr + 32, qn, ql) == 0;
          }
          if (same) next.w_cls[(size_t)i] |= C_SAME;
        }
      }
      if (round_kept && want_meta && !(bad.load(std::memory_order_relaxed) & 2)) {   // sizes -> offsets, and the variable-length pieces to their places
        const uint8_t* cbase = cur.data + cur.own;
        uint64_t run[5];
        for (int q = 0; q < 5; ++q) run[q] = part_sum[(size_t)tid * 5 + q];
        // (the same static schedule as the sizing loop: thread `tid` sees the same records)
#pragma omp for schedule(static) nowait
        for (long j = 0; j < (long)round_kept; ++j) {
          const size_t i = s_idx[(size_t)j];
          const size_t g = n_kept + (size_t)j;
          const uint8_t* r = cbase + cur.w_off[i] + 4;
          const uint8_t l_name = r[8];
          const uint16_t n_cig = rd_u16(r + 12);
          const uint32_t l_seq = cur.w_lseq[i];
          const uint64_t ql = im->qname_off[g + 1], nc = im->cigar_off[g + 1], sl = im->sa_off[g + 1];
          if (ql) memcpy(&im->qname_blob[run[0]], r + 32, ql);
          const uint8_t* cg = r + 32 + l_name;
          if (nc) memcpy(&im->cigar_blob[run[1]], cg, 4 * (size_t)nc);
          if (sl) memcpy(&im->sa_blob[run[2]], s_sa[(size_t)j], sl);
          run[0] += ql, run[1] += nc, run[2] += sl;
          im->qname_off[g + 1] = run[0];
          im->cigar_off[g + 1] = run[1];
          im->sa_off[g + 1] = run[2];
          if (want_meta >= 2) {
            if (l_seq) memcpy(&im->qual_blob[run[3]], cg + 4 * (size_t)n_cig + (l_seq + 1) / 2, l_seq);
            run[3] += l_seq;
            im->qual_off[g + 1] = run[3];
          }
          if (want_meta >= 3) {
            const uint64_t bs = im->raw_off[g + 1];
            if (bs) memcpy(&im->raw_blob[run[4]], r, bs);
            run[4] += bs;
            im->raw_off[g + 1] = run[4];
          }
        }
      }
#pragma omp single nowait
      { b->tm.fill += omp_get_wtime() - t_fill; }
    }
    b->tm.rounds += omp_get_wtime() - t_round;
    b->tm.n_rounds++;
    n_kept += round_kept;
    if (bad.load() & 2) {
      if (do_produce) recycle_chunk(b, next);
      return bail("out of memory while decoding the BAM");
    }
    if (do_produce) {
      if (bad) {
        recycle_chunk(b, next);
        return bail("BGZF inflate failed (corrupt block or CRC mismatch)");
      }
      next.valid = true;
      b->record_index = next.rec_base + next.n_rec;
      b->ahead.valid = false;
    }
    if (!read_ok) return bail(rerr);
    if (do_read && b->ahead2.valid) {   // ahead is free by now (see do_read)
      std::swap(b->ahead, b->ahead2);
      b->ahead2.valid = false;
    }
    if (!perr.empty()) return bail(perr);
    if (hit_limit) {
      b->range_done = true;
      break;
    }
  }
  // the QNAME of the last primary record consumed is only needed when the pipeline is reset
  // (kdf_bam_seek clears it); chunks carry it between themselves
  // (an incomplete record at the end of the file keeps at_eof off: the next call reports it)
  b->eof = b->range_done ||
           (b->file_eof && !b->ahead.valid && !next.valid && b->carry.empty() &&
            (!cur.valid || (cur.sel == cur.n_rec && cur.exit_off == (int64_t)cur.total && cur.walk_err.empty())));
  const size_t n = n_kept;
  const uint64_t n_words = (n_bases + 31) / 32;
  im->n_bases = n_bases;
  double t_st = omp_get_wtime();
  auto lap = [&](double& acc) {
    const double t = omp_get_wtime();
    acc += t - t_st;
    t_st = t;
  };
  if (im->codes.size() < (n_words ? n_words : 1)) {   // (an empty batch)
    im->codes.assign(1, 0);
    im->valid.assign(1, 0);
  }
  im->codes.resize(n_words ? n_words : 1);
  im->valid.resize(n_words ? n_words : 1);
  im->read_starts.resize(n);
  im->read_lens.resize(n);
  im->rec_index.resize(n);
  im->rec_uoff.resize(n);
  im->fasta_keep.resize(n);
  lap(b->tm.index);
  if (n_bases <= 0xffffffffull) {   // sparse form of the validity bitmap (kdf_valid_from_invalid)
    // by word ranges: count, prefix, fill
    const int thr_inv = team_size((long)(n_words >> 12), 0);   // (a small batch is not worth a team)
    const int parts = thr_inv > 1 ? thr_inv * 4 : 1;
    const uint64_t wpp = (n_words + parts - 1) / parts;
    std::vector<uint64_t> cnt(parts + 1, 0);
    auto range = [&](int t, uint64_t* w0, uint64_t* w1) {
      *w0 = (uint64_t)t * wpp < n_words ? (uint64_t)t * wpp : n_words;
      *w1 = (uint64_t)(t + 1) * wpp < n_words ? (uint64_t)(t + 1) * wpp : n_words;
    };
    auto scan = [&](int t, uint32_t* o) -> uint64_t {
      uint64_t w0, w1, m = 0;
      range(t, &w0, &w1);
      const uint32_t* v = im->valid.data();
      for (uint64_t w = w0; w < w1; ++w) {
        uint32_t inv = ~v[w];
        if (w == n_words - 1 && (n_bases & 31)) inv &= ~0u << (32 - (n_bases & 31));
        while (inv) {
          int bit = __builtin_clz(inv);
          if (o) o[m] = (uint32_t)(w * 32 + (uint64_t)bit);
          ++m;
          inv &= ~(0x80000000u >> bit);
        }
      }
      return m;
    };
#pragma omp parallel for schedule(static) num_threads(thr_inv)
    for (int t = 0; t < parts; ++t) cnt[t + 1] = scan(t, nullptr);
    for (int t = 0; t < parts; ++t) cnt[t + 1] += cnt[t];
    grow(im->invalid, cnt[parts] ? cnt[parts] : 1);
#pragma omp parallel for schedule(static) num_threads(thr_inv)
    for (int t = 0; t < parts; ++t) scan(t, im->invalid.data() + cnt[t]);
    im->invalid.resize(cnt[parts]);
  }
  lap(b->tm.invalid);
  b->tm.n_batches++;
  b->tm.n_reads += n;
  memset(out, 0, sizeof(*out));
  out->impl = im;
  out->n_reads = n;
  out->n_bases = n_bases;
  out->codes = im->codes.data();
  out->valid = im->valid.data();
  out->read_starts = im->read_starts.data();
  out->read_lens = im->read_lens.data();
  out->rec_index = im->rec_index.data();
  out->rec_uoff = im->rec_uoff.data();
  out->fasta_keep = im->fasta_keep.data();
  if (want_meta) {
    out->ref_id = im->ref_id.data();
    out->pos = im->pos.data();
    out->next_ref_id = im->next_ref_id.data();
    out->next_pos = im->next_pos.data();
    out->flag = im->flag.data();
    out->mapq = im->mapq.data();
    out->qname_off = im->qname_off.data();
    out->qname_blob = im->qname_blob.data();
    out->cigar_off = im->cigar_off.data();
    out->cigar_blob = im->cigar_blob.data();
    out->sa_off = im->sa_off.data();
    out->sa_blob = im->sa_blob.data();
    if (want_meta >= 2) {
      out->qual_off = im->qual_off.data();
      out->qual_blob = im->qual_blob.data();
    }
    if (want_meta >= 3) {
      out->raw_off = im->raw_off.data();
      out->raw_blob = im->raw_blob.data();
    }
  }
  out->at_eof = b->eof ? 1 : 0;
  if (n_bases <= 0xffffffffull) {
    out->invalid_pos = im->invalid.data();
    out->n_invalid = im->invalid.size();
    out->has_invalid = 1;
  }
  return KDF_OK;
}

// ---- ranges ------------------------------------------------------------------
// Reposition the reader at a BGZF virtual offset (block file offset << 16 | offset inside
// the block) that is the start of a record — an entry of the .bai linear index, or
// kdf_bam_batch.rec_voff-style arithmetic of the caller — and forget the read-ahead.
// The QNAME-run state of the FASTA stream starts afresh.
int kdf_bam_seek(kdf_bam* h, uint64_t voffset) {
  Bam* b = reinterpret_cast<Bam*>(h);
  if (!b) {
    g_host_err = "kdf_bam_seek: NULL reader";
    return KDF_ERR_ARG;
  }
  const uint64_t coff = voffset >> 16;
  if (fseeko(b->fh, (off_t)coff, SEEK_SET) != 0) {
    g_host_err = "kdf_bam_seek: cannot seek";
    return KDF_ERR_ARG;
  }
  recycle_chunk(b, b->cur);
  recycle_chunk(b, b->next);
  b->range_done = false;
  b->ahead.valid = b->ahead2.valid = false;
  b->pending.clear();
  b->carry.clear();
  b->file_eof = false;
  b->eof = false;
  b->c_total = coff;
  b->skip_bytes = (size_t)(voffset & 0xffff);
  b->seen_parts = 0;
  b->u_limit = ~0ull;
  b->u_begin = 0;
  b->begin_coff = ~0ull;
  if (b->end_coff == coff) b->u_limit = b->u_total + b->end_in;   // (empty range)
  return KDF_OK;
}

// Bytes of records the decode pipeline reads ahead per chunk (default 64 MB: right for a
// sequential pass; a region fetch wants a few hundred KB).
int kdf_bam_set_chunk_bytes(kdf_bam* h, uint64_t n) {
  Bam* b = reinterpret_cast<Bam*>(h);
  if (!b || n < 4096) {
    g_host_err = "kdf_bam_set_chunk_bytes: bad argument";
    return KDF_ERR_ARG;
  }
  b->chunk_bytes = n;
  return KDF_OK;
}

// Records that start before this virtual offset are parsed (they update the QNAME-run
// state of the FASTA stream) but not delivered: a rank seeks a little before its range
// and sets its true start here, so that a run of same-QNAME records that straddles the
// boundary is collapsed exactly as in a sequential read.  Call right after kdf_bam_seek.
int kdf_bam_set_begin(kdf_bam* h, uint64_t voffset) {
  Bam* b = reinterpret_cast<Bam*>(h);
  if (!b) {
    g_host_err = "kdf_bam_set_begin: NULL reader";
    return KDF_ERR_ARG;
  }
  b->begin_coff = voffset >> 16;
  b->begin_in = (uint32_t)(voffset & 0xffff);
  b->u_begin = ~0ull;     // until the block is reached, nothing is delivered
  if (b->begin_coff == b->c_total) b->u_begin = b->u_total + b->begin_in;
  return KDF_OK;
}

// Deliver no record that starts at or after this virtual offset (the start of the next
// rank's range); ~0 removes the limit.  Call before the decode reaches that block.
int kdf_bam_set_end(kdf_bam* h, uint64_t voffset) {
  Bam* b = reinterpret_cast<Bam*>(h);
  if (!b) {
    g_host_err = "kdf_bam_set_end: NULL reader";
    return KDF_ERR_ARG;
  }
  if (voffset == ~0ull) {
    b->end_coff = ~0ull;
    b->end_in = 0;
    b->u_limit = ~0ull;
    return KDF_OK;
  }
  b->end_coff = voffset >> 16;
  b->end_in = (uint32_t)(voffset & 0xffff);
  b->u_limit = ~0ull;
  for (const auto& e : b->blk_index)      // the block may have been walked already
    if (e.second == b->end_coff) b->u_limit = e.first + b->end_in;
  return KDF_OK;
}

// ---- random access to records seen by the sequential decode ----------------
// kdf_bam_batch.rec_uoff names a record by its offset in the file's uncompressed
// stream; the reader remembers where every BGZF block it has walked starts (in both
// coordinates), so a record can be fetched again by inflating just its block(s).
int kdf_bam_fetch_records(kdf_bam* h, const uint64_t* uoffs, uint64_t n, uint8_t* out, uint64_t out_cap,
                          uint64_t* out_off, uint64_t* needed) {
  Bam* b = reinterpret_cast<Bam*>(h);
  if (!b || (n && !uoffs) || !out_off || !needed) {
    g_host_err = "kdf_bam_fetch_records: NULL argument";
    return KDF_ERR_ARG;
  }
  try {
    FILE* fh = fopen(b->path.c_str(), "rb");
    if (!fh) {
      g_host_err = "kdf_bam_fetch_records: cannot reopen " + b->path;
      return KDF_ERR_ARG;
    }
    struct Closer {
      FILE* f;
      ~Closer() { fclose(f); }
    } closer{fh};
    const auto& idx = b->blk_index;
    std::vector<uint8_t> comp(70000), cache;   // cache: inflated bytes from block `cache_first` on
    size_t cache_first = (size_t)-1, cache_next = 0;
    auto load_block = [&](size_t bi, std::vector<uint8_t>& dst) -> bool {   // append block bi to dst
      if (bi >= idx.size()) return false;
      if (fseeko(fh, (off_t)idx[bi].second, SEEK_SET) != 0) return false;
      if (fread(comp.data(), 1, 18, fh) != 18) return false;
      if (comp[0] != 31 || comp[1] != 139) return false;
      uint16_t xlen = (uint16_t)(comp[10] | (comp[11] << 8));
      if (xlen != 6) {   // the general case: re-read the whole extra field
        if (comp.size() < (size_t)12 + xlen + 8) comp.resize((size_t)12 + xlen + 70000);
        if (xlen > 6 && fread(comp.data() + 18, 1, (size_t)xlen - 6, fh) != (size_t)xlen - 6) return false;
      }
      int bsize = -1;
      for (uint32_t p = 0; p + 4 <= xlen;) {
        uint16_t slen = (uint16_t)(comp[12 + p + 2] | (comp[12 + p + 3] << 8));
        if (comp[12 + p] == 'B' && comp[12 + p + 1] == 'C' && slen == 2 && p + 6 <= xlen)
          bsize = (comp[12 + p + 4] | (comp[12 + p + 5] << 8)) + 1;
        p += 4 + slen;
      }
      if (bsize < 12 + (int)xlen + 8) return false;
      size_t have = (size_t)12 + xlen;
      if (comp.size() < (size_t)bsize) comp.resize((size_t)bsize);
      if (fread(comp.data() + have, 1, (size_t)bsize - have, fh) != (size_t)bsize - have) return false;
      const uint8_t* tl = comp.data() + bsize - 4;
      uint32_t isize = tl[0] | (tl[1] << 8) | (tl[2] << 16) | ((uint32_t)tl[3] << 24);
      if (isize > 65536) return false;
      size_t at = dst.size();
      dst.resize(at + isize);
      return isize == 0 || inflate_block(comp.data(), (uint32_t)bsize, dst.data() + at, isize, b->verify_crc);
    };
    uint64_t total = 0;
    out_off[0] = 0;
    for (uint64_t i = 0; i < n; ++i) {
      const uint64_t u = uoffs[i];
      // block holding offset u: the last one that starts at or before it
      size_t lo = 0, hi = idx.size();
      while (lo < hi) {
        size_t mid = (lo + hi) / 2;
        if (idx[mid].first <= u) lo = mid + 1; else hi = mid;
      }
      if (lo == 0) {
        g_host_err = "kdf_bam_fetch_records: offset precedes the blocks decoded so far";
        return KDF_ERR_ARG;
      }
      const size_t bi = lo - 1;
      if (cache_first != bi) {   // (records fetched in ascending order mostly share blocks)
        cache.clear();
        cache_first = bi;
        cache_next = bi;
      }
      const size_t in_off = (size_t)(u - idx[bi].first);
      auto ensure = [&](size_t need_bytes) -> bool {
        while (cache.size() < need_bytes) {
          if (!load_block(cache_next, cache)) return false;
          ++cache_next;
        }
        return true;
      };
      if (!ensure(in_off + 4)) {
        g_host_err = "kdf_bam_fetch_records: cannot read the record's block";
        return KDF_ERR_ARG;
      }
      int32_t bs = rd_i32(cache.data() + in_off);
      if (bs < 32 || bs > (1 << 28) || !ensure(in_off + 4 + (size_t)bs)) {
        g_host_err = "kdf_bam_fetch_records: no BAM record at this offset";
        return KDF_ERR_ARG;
      }
      if (out && total + (uint64_t)bs <= out_cap) memcpy(out + total, cache.data() + in_off + 4, (size_t)bs);
      total += (uint64_t)bs;
      out_off[i + 1] = total;
    }
    *needed = total;
    return KDF_OK;
  } catch (const std::exception& e) {
    g_host_err = std::string("kdf_bam_fetch_records: ") + e.what();
    return KDF_ERR_ARG;
  }
}

// ---- BGZF writer -------------------------------------------------------------
// `data` as a BGZF file (blocks of <= 0xff00 input bytes deflated by all threads,
// written in order, then the EOF marker): the container of the BAM and bgzip-VCF
// outputs.  block_coff (may be NULL) receives the file offset of every block, so that
// the caller can turn uncompressed offsets into BAI / TBI virtual offsets.
// ---- reference FASTA -> packed stream ------------------------------------------------------
// The reference genome as one stream (sequences separated by an invalid base), packed by all
// threads: header lines are located, the sequence bytes are cut into blocks, every block counts
// its bases (pass 1) and, once the prefix sums give its place in the stream, packs them (pass 2).
namespace {
struct FastaLayout {
  std::vector<uint64_t> seq_beg, seq_end;   // byte range of each record's sequence lines
  struct Block {
    uint64_t beg, end;   // bytes
    uint32_t seq;
    uint64_t bases;      // non-whitespace bytes in it
    uint64_t pos;        // stream position of its first base
  };
  std::vector<Block> blocks;
  std::vector<uint64_t> seq_len, seq_start;
  uint64_t total = 0;
};
inline bool fasta_skip(uint8_t c) { return c == '\n' || c == '\r' || c == ' ' || c == '\t' || c == '\v' || c == '\f'; }

void fasta_layout(const uint8_t* text, uint64_t n, int threads, FastaLayout& L) {
  // header lines: '>' at the start of a line
  const uint64_t SCAN = 4u << 20;
  const long n_scan = (long)((n + SCAN - 1) / SCAN);
  std::vector<std::vector<uint64_t>> found((size_t)(n_scan > 0 ? n_scan : 1));
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
  for (long b = 0; b < n_scan; ++b) {
    const uint64_t lo = (uint64_t)b * SCAN, hi = lo + SCAN < n ? lo + SCAN : n;
    const uint8_t* p = text + lo;
    while (p < text + hi) {
      const uint8_t* q = (const uint8_t*)memchr(p, '>', (size_t)(text + hi - p));
      if (!q) break;
      if (q == text || q[-1] == '\n') found[(size_t)b].push_back((uint64_t)(q - text));
      p = q + 1;
    }
  }
  std::vector<uint64_t> hdr;
  for (auto& v : found) hdr.insert(hdr.end(), v.begin(), v.end());
  const size_t ns = hdr.size();
  L.seq_beg.resize(ns);
  L.seq_end.resize(ns);
  for (size_t j = 0; j < ns; ++j) {
    const uint8_t* nl = (const uint8_t*)memchr(text + hdr[j], '\n', (size_t)(n - hdr[j]));
    L.seq_beg[j] = nl ? (uint64_t)(nl - text) + 1 : n;
    L.seq_end[j] = j + 1 < ns ? hdr[j + 1] : n;
  }
  const uint64_t BLK = 1u << 20;
  L.blocks.clear();
  for (size_t j = 0; j < ns; ++j)
    for (uint64_t o = L.seq_beg[j]; o < L.seq_end[j]; o += BLK)
      L.blocks.push_back({o, o + BLK < L.seq_end[j] ? o + BLK : L.seq_end[j], (uint32_t)j, 0, 0});
  const long nb = (long)L.blocks.size();
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads)
  for (long i = 0; i < nb; ++i) {
    FastaLayout::Block& k = L.blocks[(size_t)i];
    uint64_t c = 0;
    for (uint64_t o = k.beg; o < k.end; ++o) c += fasta_skip(text[o]) ? 0 : 1;
    k.bases = c;
  }
  L.seq_len.assign(ns, 0);
  L.seq_start.assign(ns, 0);
  for (auto& k : L.blocks) L.seq_len[k.seq] += k.bases;
  uint64_t p = 0;
  for (size_t j = 0; j < ns; ++j) {
    L.seq_start[j] = p;
    p += L.seq_len[j] + 1;
  }
  L.total = ns ? p - 1 : 0;
  uint32_t cur = ~0u;
  uint64_t at = 0;
  for (auto& k : L.blocks) {
    if (k.seq != cur) cur = k.seq, at = L.seq_start[cur];
    k.pos = at;
    at += k.bases;
  }
}
}  // namespace

int kdf_fasta_layout(const uint8_t* text, uint64_t n, int n_threads, uint64_t* n_seqs, uint64_t* n_bases) {
  if ((!text && n) || !n_seqs || !n_bases) {
    g_host_err = "kdf_fasta_layout: NULL argument";
    return KDF_ERR_ARG;
  }
  try {
    FastaLayout L;
    fasta_layout(text, n, n_threads > 0 ? n_threads : 1, L);
    *n_seqs = L.seq_len.size();
    *n_bases = L.total;
    return KDF_OK;
  } catch (const std::exception& e) {
    g_host_err = std::string("kdf_fasta_layout: ") + e.what();
    return KDF_ERR_ARG;
  }
}

int kdf_fasta_pack(const uint8_t* text, uint64_t n, int n_threads, uint64_t* codes, uint32_t* valid,
                   uint64_t* seq_starts, uint64_t* seq_lens) {
  if ((!text && n) || !codes || !valid || !seq_starts || !seq_lens) {
    g_host_err = "kdf_fasta_pack: NULL argument";
    return KDF_ERR_ARG;
  }
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
  try {
    const int threads = n_threads > 0 ? n_threads : 1;
    FastaLayout L;
    fasta_layout(text, n, threads, L);
    for (size_t j = 0; j < L.seq_len.size(); ++j) seq_starts[j] = L.seq_start[j], seq_lens[j] = L.seq_len[j];
    const long n_words = (long)((L.total + 31) / 32);
#pragma omp parallel for schedule(static) num_threads(threads)
    for (long w = 0; w < (n_words ? n_words : 1); ++w) codes[w] = 0, valid[w] = 0;
    const long nb = (long)L.blocks.size();
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads)
    for (long i = 0; i < nb; ++i) {
      const FastaLayout::Block& k = L.blocks[(size_t)i];
      if (!k.bases) continue;
      uint64_t p = k.pos;
      const uint64_t first_w = p >> 5, last_w = (p + k.bases - 1) >> 5;
      uint64_t cw = 0;
      uint32_t vw = 0;
      for (uint64_t o = k.beg; o < k.end; ++o) {
        const uint8_t ch = text[o];
        if (fasta_skip(ch)) continue;
        const uint8_t c = lut.v[ch];
        if (c < 4) {
          cw |= (uint64_t)c << (62 - 2 * (p & 31));
          vw |= 1u << (31 - (p & 31));
        }
        if ((p & 31) == 31) {   // the word is complete
          const uint64_t w = p >> 5;
          or_word64(codes + w, cw, w == first_w || w == last_w);
          or_word32(valid + w, vw, w == first_w || w == last_w);
          cw = 0, vw = 0;
        }
        ++p;
      }
      if (p & 31) {   // the last, partial word (shared with what follows)
        or_word64(codes + (p >> 5), cw, true);
        or_word32(valid + (p >> 5), vw, true);
      }
    }
    return KDF_OK;
  } catch (const std::exception& e) {
    g_host_err = std::string("kdf_fasta_pack: ") + e.what();
    return KDF_ERR_ARG;
  }
}

int kdf_bgzf_inflate_block(const uint8_t* src, uint32_t csize, uint8_t* dst, uint32_t usize, int verify_crc,
                           int impl) {
  if (!src || (!dst && usize) || usize > 65536) {
    g_host_err = "kdf_bgzf_inflate_block: bad argument";
    return KDF_ERR_ARG;
  }
  uint8_t dummy = 0;
  if (!inflate_block(src, csize, dst ? dst : &dummy, usize, verify_crc != 0, impl ? 1 : 0)) {
    g_host_err = "BGZF inflate failed (corrupt block or CRC mismatch)";
    return KDF_ERR_ARG;
  }
  return KDF_OK;
}

uint32_t kdf_crc32(const uint8_t* data, uint64_t n) { return kdf::crc32_of(data, (size_t)n); }

int kdf_bgzf_write(const char* path, const uint8_t* data, uint64_t n, int level, int n_threads,
                   uint64_t* block_coff, uint64_t block_cap, uint64_t* n_blocks) {
  if (!path || (n && !data)) {
    g_host_err = "kdf_bgzf_write: NULL argument";
    return KDF_ERR_ARG;
  }
  try {
    const uint64_t BLK = 0xff00;
    const uint64_t nb = (n + BLK - 1) / BLK;
    if (n_blocks) *n_blocks = nb;
    FILE* fh = fopen(path, "wb");
    if (!fh) {
      g_host_err = std::string("cannot create ") + path;
      return KDF_ERR_ARG;
    }
    if (level < 0 || level > 9) level = 6;
    if (n_threads < 1) n_threads = 1;
    const uint64_t GROUP = 1024;   // blocks compressed per round (64 MB of input)
    std::vector<std::vector<uint8_t>> outb(GROUP);
    uint64_t coff = 0;
    int bad = 0;
    for (uint64_t g0 = 0; g0 < nb && !bad; g0 += GROUP) {
      const long gn = (long)(nb - g0 < GROUP ? nb - g0 : GROUP);
#pragma omp parallel for schedule(dynamic, 4) num_threads(n_threads) reduction(| : bad)
      for (long j = 0; j < gn; ++j) {
        const uint64_t bi = g0 + (uint64_t)j;
        const uint8_t* src = data + bi * BLK;
        const uint32_t len = (uint32_t)(n - bi * BLK < BLK ? n - bi * BLK : BLK);
        std::vector<uint8_t>& ob = outb[(size_t)j];
        ob.resize(18 + compressBound(len) + 8);
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) {
          bad |= 1;
          continue;
        }
        zs.next_in = const_cast<uint8_t*>(src);
        zs.avail_in = len;
        zs.next_out = ob.data() + 18;
        zs.avail_out = (uInt)(ob.size() - 18 - 8);
        int rc = deflate(&zs, Z_FINISH);
        deflateEnd(&zs);
        if (rc != Z_STREAM_END || 18 + zs.total_out + 8 > 65536) {
          bad |= 1;   // (cannot happen for <= 0xff00 input bytes)
          continue;
        }
        const uint32_t clen = (uint32_t)zs.total_out;
        static const uint8_t HDR[16] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 'B', 'C', 2, 0};
        memcpy(ob.data(), HDR, 16);
        const uint16_t bsz = (uint16_t)(18 + clen + 8 - 1);
        ob[16] = (uint8_t)(bsz & 255);
        ob[17] = (uint8_t)(bsz >> 8);
        const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), src, len);
        uint8_t* tl = ob.data() + 18 + clen;
        for (int q = 0; q < 4; ++q) tl[q] = (uint8_t)(crc >> (8 * q));
        for (int q = 0; q < 4; ++q) tl[4 + q] = (uint8_t)(len >> (8 * q));
        ob.resize(18 + clen + 8);
      }
      for (long j = 0; j < gn && !bad; ++j) {
        if (block_coff && g0 + (uint64_t)j < block_cap) block_coff[g0 + (uint64_t)j] = coff;
        if (fwrite(outb[(size_t)j].data(), 1, outb[(size_t)j].size(), fh) != outb[(size_t)j].size()) bad |= 2;
        coff += outb[(size_t)j].size();
      }
    }
    static const uint8_t EOF_BLOCK[28] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67,
                                          2,  0,   27, 0, 3, 0, 0, 0, 0, 0, 0,   0, 0, 0};
    if (!bad && fwrite(EOF_BLOCK, 1, 28, fh) != 28) bad |= 2;
    if (block_coff && nb < block_cap) block_coff[nb] = coff;   // offset of the EOF block
    if (fclose(fh) != 0) bad |= 2;
    if (bad) {
      g_host_err = (bad & 2) ? std::string("write failed: ") + path : "deflate failed";
      return KDF_ERR_ARG;
    }
    return KDF_OK;
  } catch (const std::exception& e) {
    g_host_err = std::string("kdf_bgzf_write: ") + e.what();
    return KDF_ERR_ARG;
  }
}

uint64_t kdf_invalid_positions(const uint32_t* valid, uint64_t n_bases, uint32_t* out, uint64_t cap) {
  if (!valid || n_bases > 0xffffffffull) return ~0ull;
  return invalid_positions(valid, n_bases, out, cap);
}

void kdf_bam_batch_free(kdf_bam_batch* batch) {
  if (!batch || !batch->impl) return;
  impl_put(reinterpret_cast<kdf_bam_batch_impl*>(batch->impl));
  memset(batch, 0, sizeof(*batch));
}

}  // extern "C"
