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
#include <zlib.h>

#include <memory>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include "../../include/kdf.h"

namespace {

thread_local std::string g_host_err;

// std::vector whose resize() leaves new elements uninitialised (the packer zero-fills
// codes / valid itself, in parallel)
template <class T> struct NoInitAlloc : std::allocator<T> {
  template <class U> struct rebind { using other = NoInitAlloc<U>; };
  template <class U> void construct(U* p) noexcept { ::new ((void*)p) U; }
  template <class U, class... A> void construct(U* p, A&&... a) { ::new ((void*)p) U(std::forward<A>(a)...); }
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

// a record taken into the batch being built
struct Kept {
  const uint8_t* rec;  // the record body (inside one of the chunk buffers)
  uint64_t start;      // stream position of its first base
  uint64_t uoff;       // offset of the record (its block_size word) in the uncompressed stream
  uint32_t l_seq;
  uint8_t fasta_keep;  // KDF_BAM_FASTA would keep this record
};

// an inflated chunk: logical bytes data[begin .. size), `begin` leaves headroom for the
// unparsed tail of the chunk before it
struct Chunk {
  uint8_t* data = nullptr;
  size_t begin = 0, size = 0, cap = 0;
  size_t own = 0;        // where the chunk's own first byte sits (bytes before it: the previous chunk's tail)
  uint64_t ustart = 0;   // uncompressed-stream offset of data[own]
  bool valid = false;
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
  uint64_t u_limit = ~0ull;      // deliver no record at or past this uncompressed offset (kdf_bam_set_end)
  uint64_t u_begin = 0;          // records before this offset only update the QNAME-run state (kdf_bam_set_begin)
  uint64_t begin_coff = ~0ull;
  uint32_t begin_in = 0;
  size_t skip_bytes = 0;         // kdf_bam_seek: bytes of the first block that precede the target record
  uint64_t end_coff = ~0ull;     // kdf_bam_set_end: virtual offset (block, offset in block) to stop at
  uint32_t end_in = 0;
  int last_set_part = -1;        // read-part bit the last parsed record set in seen_parts (-1: none)
  std::string path;
  // collapse state of the FASTA stream (persists across batches)
  std::string cur_qname;
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
  std::vector<Kept> kept;   // records of the batch being built
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
    static const bool thp = getenv("KDF_BAM_THP") != nullptr;
    if (p && thp) madvise(p, *cap, MADV_HUGEPAGE);
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
    if (cur.valid) free(cur.data);
    if (next.valid) free(next.data);
    for (auto& e : pool) free(e.first);
  }
};

bool inflate_block(const uint8_t* src, uint32_t csize, uint8_t* dst, uint32_t usize, bool verify_crc = true) {
  // BGZF: 18-byte header (with BC subfield), deflate payload, crc32 + isize
  if (csize < 26) return false;
  uint16_t xlen = (uint16_t)(src[10] | (src[11] << 8));
  uint32_t hdr = 12 + xlen;
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
  if (verify_crc) {
    const uint8_t* tl = src + csize - 8;
    uint32_t want = tl[0] | (tl[1] << 8) | (tl[2] << 16) | ((uint32_t)tl[3] << 24);
    if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), dst, usize) != want) return false;
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
  const size_t SLAB = 8u << 20;
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
  c.begin = gap;
  c.own = gap;
  c.ustart = cb.ustart;
  c.size = gap + cb.total_u;
  c.cap = cap;
  c.valid = false;   // until inflated
  return true;
}

// Stage 2 (one block; any thread)
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
  out.insert(out.end(), c.data + c.begin, c.data + c.size);
  b->put_buf(c.data, c.cap);
  return true;
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

// pack l_seq bases (BAM nibbles) at stream position `start`: 32 bases at a time into an
// aligned word (eight table look-ups of four bases each), shifted into place; the first
// and the last stream word of a read are shared with its neighbours (atomic OR)
void pack_record(const uint8_t* nib, uint32_t l_seq, uint64_t start, uint64_t* codes,
                 uint32_t* valid) {
  if (!l_seq) return;
  const uint64_t first_w = start >> 5, last_w = (start + l_seq - 1) >> 5;
  const unsigned s = (unsigned)(start & 31);
  uint64_t carry_c = 0;
  uint32_t carry_v = 0;
  uint64_t w = first_w;
  for (uint32_t done = 0; done < l_seq; done += 32, ++w) {
    const uint32_t nb = l_seq - done < 32 ? l_seq - done : 32;
    const uint8_t* src = nib + (done >> 1);
    uint64_t lw = 0;
    uint32_t lv = 0;
    const unsigned groups = (nb + 3) >> 2;
    for (unsigned g = 0; g < groups; ++g) {
      uint16_t two;
      memcpy(&two, src + 2 * g, 2);   // past an odd end this reads into the qualities: masked below
      const uint16_t e = NIB4.v[two];
      lw |= (uint64_t)(e & 255) << (56 - 8 * g);
      lv |= (uint32_t)(e >> 8) << (28 - 4 * g);
    }
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
  std::vector<uint64_t, NoInitAlloc<uint64_t>> codes;
  std::vector<uint32_t, NoInitAlloc<uint32_t>> valid;
  std::vector<uint32_t> invalid;   // positions of the invalid bases (sparse form of `valid`)
  std::vector<uint64_t> read_starts;
  std::vector<uint32_t> read_lens;
  std::vector<uint64_t> rec_index;
  std::vector<uint64_t> rec_uoff;     // offset of every kept record in the uncompressed stream
  std::vector<uint8_t> fasta_keep;    // the record is part of the KDF_BAM_FASTA stream
  std::vector<int32_t> ref_id, pos, next_ref_id, next_pos;
  std::vector<uint16_t> flag;
  std::vector<uint8_t> mapq;
  std::vector<uint64_t> qname_off, cigar_off, sa_off;  // n+1 each
  std::vector<char> qname_blob, sa_blob;
  std::vector<uint32_t> cigar_blob;
  std::vector<uint64_t> raw_off;    // n+1 (want_meta >= 3): the BAM records themselves
  std::vector<uint8_t> raw_blob;    //   (bytes after block_size), for BAM output
  std::vector<uint64_t> qual_off;   // n+1 (want_meta >= 2)
  std::vector<uint8_t> qual_blob;   // Phred base qualities, l_seq bytes per read
  uint64_t n_bases = 0;
};

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
    std::unique_ptr<kdf_bam_batch_impl> im(new kdf_bam_batch_impl);
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

static int next_batch_impl(Bam* b, int mode, uint64_t max_bases, int want_meta, kdf_bam_batch* out,
                           kdf_bam_batch_impl* im) {
  // (the vector lives in the reader: its pages stay mapped from batch to batch, a fresh
  // one cost a page fault per 128 records in the middle of the serial walk)
  std::vector<Kept>& kept = b->kept;
  kept.clear();
  uint64_t n_bases = 0;
  // The decode is a three-stage pipeline over chunks of ~64 MB of records: while one
  // thread walks the records of chunk C (a serial chain: every header is a cache miss),
  // another reads the compressed bytes of chunk C+2 and the rest inflate chunk C+1.
  // Chunks whose records were kept stay where they are until the batch is packed; a
  // record that straddles two chunks is completed by copying the unparsed tail of the
  // old chunk into the headroom in front of the new one.  What is not consumed (the rest
  // of the current chunk, the look-ahead) stays in the reader for the next batch.
  const uint64_t CHUNK_BYTES = b->chunk_bytes;
  const size_t GAP = b->gap;
  std::vector<Chunk> retired;
  auto release_retired = [&]() {
    for (auto& c : retired) b->put_buf(c.data, c.cap);
    retired.clear();
  };
  auto bail = [&](const std::string& msg) {
    g_host_err = msg;
    release_retired();
    return KDF_ERR_ARG;
  };
  if (!b->cur.valid) {   // first call: the bytes that followed the header
    size_t cap = 0;
    uint8_t* p = b->get_buf(b->carry.size() + 1, &cap);
    if (!p) return bail("out of memory");
    if (!b->carry.empty()) memcpy(p, b->carry.data(), b->carry.size());
    b->cur.data = p;
    b->cur.begin = 0;
    b->cur.own = 0;
    b->cur.ustart = b->first_rec_uoff;
    b->cur.size = b->carry.size();
    b->cur.cap = cap;
    b->cur.valid = true;
    b->carry.clear();
  }
  im->rec_index.reserve(kept.capacity());
  const uint8_t* buf = b->cur.data;
  size_t buf_size = b->cur.size;
  size_t off = b->cur.begin;
  uint64_t cur_ustart = b->cur.ustart;
  size_t cur_own = b->cur.own;
  bool done = false, hit_limit = false;
  std::string perr, rerr;
  // stage 3: parse all complete records currently in the chunk
  auto parse = [&]() {
    while (true) {
      if (buf_size - off < 4) break;
      if (cur_ustart + (uint64_t)off - (uint64_t)cur_own >= b->u_limit) {   // end of this reader's range
        hit_limit = true;
        break;
      }
      int32_t bs = rd_i32(buf + off);
      if (bs < 32 || bs > (1 << 28)) {
        perr = "corrupt BAM record (block_size out of range)";
        return;
      }
      if (buf_size - off - 4 < (size_t)bs) break;
      const uint8_t* r = buf + off + 4;
      // the walk is a chain of one cache miss per record header: pull the headers of the
      // records a few hundred bytes ahead (the chunk buffer has slack past its end)
      __builtin_prefetch(r + bs + 1024);
      __builtin_prefetch(r + bs + 1088);
      uint16_t flag = rd_u16(r + 14);
      uint8_t l_name = r[8];
      uint16_t n_cig_v = rd_u16(r + 12);
      int32_t l_seq_s = rd_i32(r + 16);
      // the variable-length fields must lie inside the record: a corrupt l_seq / n_cigar /
      // l_read_name would otherwise send the packer (and the metadata pass) out of bounds
      if (l_seq_s < 0 || 32ull + l_name + 4ull * n_cig_v + ((uint64_t)l_seq_s + 1) / 2 + (uint64_t)l_seq_s >
                             (uint64_t)bs) {
        perr = "corrupt BAM record (field sizes exceed the record)";
        return;
      }
      uint32_t l_seq = (uint32_t)l_seq_s;
      // membership of the `samtools fasta -F 0xD00` stream, tracked in every mode (the
      // discovery pipeline decodes the child ONCE in scan mode and masks the counting
      // stream with this flag)
      bool fasta_keep = false;
      b->last_set_part = -1;
      if (!(flag & 0xD00)) {
        const char* qn = (const char*)r + 32;
        size_t ql = l_name ? (size_t)l_name - 1 : 0;
        if (b->cur_qname.size() != ql || memcmp(b->cur_qname.data(), qn, ql) != 0) {
          b->cur_qname.assign(qn, ql);
          b->seen_parts = 0;
        }
        bool r1 = flag & 0x40, r2 = flag & 0x80;
        unsigned part = (r1 && !r2) ? 1u : ((r2 && !r1) ? 2u : 0u);
        if (!(b->seen_parts & (1u << part))) {
          b->seen_parts |= 1u << part;
          b->last_set_part = (int)part;
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
      if (cur_ustart + (uint64_t)off - (uint64_t)cur_own < b->u_begin) keep = false;
      if (keep) {
        if (max_bases && !kept.empty() && n_bases + l_seq + 1 > max_bases) {
          done = true;
          break;  // leave this record for the next batch
        }
        uint64_t start = kept.empty() ? 0 : n_bases + 1;
        const uint64_t uoff = cur_ustart + (uint64_t)off - (uint64_t)cur_own;   // off may lie in the tail before `own`
        kept.push_back({buf + off + 4, start, uoff, l_seq, (uint8_t)(fasta_keep ? 1 : 0)});
        n_bases = start + l_seq;
        im->rec_index.push_back(b->record_index);
      }
      b->record_index++;
      off += 4 + (size_t)bs;
    }
  };
  while (true) {
    // the compressed blocks of the next chunk must be at hand (read here only when the
    // pipeline is cold: otherwise stage 1 of the previous round fetched them)
    if (!b->next.valid && !b->ahead.valid && !b->file_eof) {
      if (!read_comp(b, CHUNK_BYTES, b->ahead, rerr)) return bail(rerr);
    }
    const bool do_inflate = !b->next.valid && b->ahead.valid;
    Chunk nx;
    if (do_inflate && !alloc_chunk(b, b->ahead, GAP, nx)) return bail("out of memory");
    const bool do_read = !b->file_eof && (do_inflate || !b->ahead.valid);
    b->ahead2.valid = false;
    bool read_ok = true;
    int bad = 0;
    const long n_blk = do_inflate ? (long)b->ahead.blocks.size() : 0;
#pragma omp parallel num_threads(b->threads) reduction(| : bad)
    {
#pragma omp single nowait
      { parse(); }
#pragma omp single nowait
      {
        if (do_read) read_ok = read_comp(b, CHUNK_BYTES, b->ahead2, rerr);
      }
#pragma omp for schedule(dynamic, 4) nowait
      for (long i = 0; i < n_blk; ++i)
        if (!inflate_one(b->ahead, (size_t)i, nx, b->verify_crc)) bad |= 1;
    }
    if (do_inflate) {
      if (bad) {
        b->put_buf(nx.data, nx.cap);
        return bail("BGZF inflate failed (corrupt block or CRC mismatch)");
      }
      nx.valid = true;
      b->next = nx;
      b->ahead.valid = false;
    }
    if (!read_ok) return bail(rerr);
    if (!perr.empty()) return bail(perr);
    if (do_read && b->ahead2.valid) {   // ahead is free by now (see do_read)
      std::swap(b->ahead, b->ahead2);
      b->ahead2.valid = false;
    }
    if (done) break;   // batch full: the rest of this chunk and the look-ahead wait in the reader
    if (hit_limit) break;
    // this chunk is exhausted but for an incomplete record at its end
    const size_t tail = buf_size - off;
    if (!b->next.valid) {
      if (b->file_eof && !b->ahead.valid) {
        if (tail) return bail("truncated BAM file (incomplete record at the end)");
        break;
      }
      continue;   // the look-ahead is still to be inflated: go round again
    }
    Chunk& nxt = b->next;
    if (tail <= nxt.begin) {
      if (tail) memcpy(nxt.data + nxt.begin - tail, buf + off, tail);
      nxt.begin -= tail;
    } else {   // a record longer than the headroom: rebuild the next chunk behind the tail
      size_t cap = 0, n_next = nxt.size - nxt.begin;
      uint8_t* p = b->get_buf(tail + n_next + 1, &cap);
      if (!p) return bail("out of memory");
      memcpy(p, buf + off, tail);
      memcpy(p + tail, nxt.data + nxt.begin, n_next);
      b->put_buf(nxt.data, nxt.cap);
      nxt.own = tail + (nxt.own - nxt.begin);
      nxt.data = p;
      nxt.begin = 0;
      nxt.size = tail + n_next;
      nxt.cap = cap;
    }
    retired.push_back(b->cur);   // kept records point into it: released after packing
    b->cur = nxt;
    b->next = Chunk();
    buf = b->cur.data;
    buf_size = b->cur.size;
    off = b->cur.begin;
    cur_ustart = b->cur.ustart;
    cur_own = b->cur.own;
    if (b->skip_bytes) {   // first chunk after kdf_bam_seek: the target record starts inside its first block
      if (buf_size - off < b->skip_bytes) return bail("kdf_bam_seek: offset beyond the end of its block");
      off += b->skip_bytes;
      b->skip_bytes = 0;
    }
  }
  b->cur.begin = off;
  b->eof = (b->file_eof && !b->ahead.valid && !b->next.valid) || hit_limit;
  if (hit_limit) b->cur.begin = b->cur.size;   // nothing more to deliver from this range
  // A batch limit postponed a record whose parse had already updated the collapse state:
  // if it set a read-part bit, clear it, so that the record is kept when the next batch
  // parses it again (its QNAME is then still the current one).
  if (done && b->last_set_part >= 0) {
    b->seen_parts &= ~(1u << b->last_set_part);
    b->last_set_part = -1;
  }
  size_t n = kept.size();
  uint64_t n_words = (n_bases + 31) / 32;
  im->n_bases = n_bases;
  // zero-filled by all threads (also the first touch of these pages)
  im->codes.resize(n_words ? n_words : 1);
  im->valid.resize(n_words ? n_words : 1);
  {
    uint64_t* cw = im->codes.data();
    uint32_t* vw = im->valid.data();
    const long nw = (long)(n_words ? n_words : 1);
#pragma omp parallel for schedule(static) num_threads(b->threads)
    for (long w = 0; w < nw; ++w) {
      cw[w] = 0;
      vw[w] = 0;
    }
  }
  im->read_starts.resize(n);
  im->read_lens.resize(n);
#pragma omp parallel for schedule(static) num_threads(b->threads)
  for (long i = 0; i < (long)n; ++i) {
    const uint8_t* r = kept[i].rec;
    uint8_t l_name = r[8];
    uint16_t n_cig = rd_u16(r + 12);
    const uint8_t* nib = r + 32 + l_name + 4 * (size_t)n_cig;
    pack_record(nib, kept[i].l_seq, kept[i].start, im->codes.data(), im->valid.data());
    im->read_starts[i] = kept[i].start;
    im->read_lens[i] = kept[i].l_seq;
  }
  im->rec_uoff.resize(n);
  im->fasta_keep.resize(n);
  for (size_t i = 0; i < n; ++i) {
    im->rec_uoff[i] = kept[i].uoff;
    im->fasta_keep[i] = kept[i].fasta_keep;
  }
  if (n_bases <= 0xffffffffull) {   // sparse form of the validity bitmap (kdf_valid_from_invalid)
    // by word ranges: count, prefix, fill
    const int parts = b->threads > 1 ? b->threads * 4 : 1;
    const uint64_t wpp = ((n_words + parts - 1) / parts + 0) | 0;
    std::vector<uint64_t> cnt(parts + 1, 0);
    auto range = [&](int t, uint64_t* w0, uint64_t* w1) {
      *w0 = (uint64_t)t * wpp < n_words ? (uint64_t)t * wpp : n_words;
      *w1 = (uint64_t)(t + 1) * wpp < n_words ? (uint64_t)(t + 1) * wpp : n_words;
    };
    auto scan = [&](int t, uint32_t* out) -> uint64_t {
      uint64_t w0, w1, m = 0;
      range(t, &w0, &w1);
      const uint32_t* v = im->valid.data();
      for (uint64_t w = w0; w < w1; ++w) {
        uint32_t inv = ~v[w];
        if (w == n_words - 1 && (n_bases & 31)) inv &= ~0u << (32 - (n_bases & 31));
        while (inv) {
          int bit = __builtin_clz(inv);
          if (out) out[m] = (uint32_t)(w * 32 + (uint64_t)bit);
          ++m;
          inv &= ~(0x80000000u >> bit);
        }
      }
      return m;
    };
#pragma omp parallel for schedule(static) num_threads(b->threads)
    for (int t = 0; t < parts; ++t) cnt[t + 1] = scan(t, nullptr);
    for (int t = 0; t < parts; ++t) cnt[t + 1] += cnt[t];
    im->invalid.resize(cnt[parts] ? cnt[parts] : 1);
#pragma omp parallel for schedule(static) num_threads(b->threads)
    for (int t = 0; t < parts; ++t) scan(t, im->invalid.data() + cnt[t]);
    im->invalid.resize(cnt[parts]);
  }
  if (want_meta) {
    im->ref_id.resize(n);
    im->pos.resize(n);
    im->next_ref_id.resize(n);
    im->next_pos.resize(n);
    im->flag.resize(n);
    im->mapq.resize(n);
    im->qname_off.assign(n + 1, 0);
    im->cigar_off.assign(n + 1, 0);
    im->sa_off.assign(n + 1, 0);
    if (want_meta >= 2) im->qual_off.assign(n + 1, 0);
    if (want_meta >= 3) im->raw_off.assign(n + 1, 0);
    // pass 1 (parallel): fixed fields, and the size of every variable-length piece
    std::vector<const uint8_t*> sa_ptr(n, nullptr);
    auto find_sa = [](const uint8_t* t, const uint8_t* end, size_t* len) -> const uint8_t* {
      const uint8_t* found = nullptr;
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
            if (!z) return found;
            if (t0 == 'S' && t1 == 'A' && ty == 'Z') {
              *len = (size_t)(z - t);
              return t;
            }
            adv = (size_t)(z - t) + 1;
            break;
          }
          case 'B': {
            if (t + 5 > end) return found;
            char sub = (char)t[0];
            uint32_t cnt;
            memcpy(&cnt, t + 1, 4);
            size_t sz = (sub == 'c' || sub == 'C') ? 1 : ((sub == 's' || sub == 'S') ? 2 : 4);
            adv = 5 + sz * cnt;
            break;
          }
          default: return found;
        }
        if (t >= end) break;
        t += adv;
      }
      return found;
    };
#pragma omp parallel for schedule(static) num_threads(b->threads)
    for (long i = 0; i < (long)n; ++i) {
      const uint8_t* r = kept[i].rec;
      int32_t bs = rd_i32(r - 4);
      uint8_t l_name = r[8];
      uint16_t n_cig = rd_u16(r + 12);
      uint32_t l_seq = kept[i].l_seq;
      im->ref_id[i] = rd_i32(r);
      im->pos[i] = rd_i32(r + 4);
      im->mapq[i] = r[9];
      im->flag[i] = rd_u16(r + 14);
      im->next_ref_id[i] = rd_i32(r + 20);
      im->next_pos[i] = rd_i32(r + 24);
      im->qname_off[i + 1] = l_name ? (size_t)l_name - 1 : 0;
      im->cigar_off[i + 1] = n_cig;
      if (want_meta >= 3) im->raw_off[i + 1] = (uint64_t)bs;
      if (want_meta >= 2) im->qual_off[i + 1] = l_seq;
      const uint8_t* t = r + 32 + l_name + 4 * (size_t)n_cig + (l_seq + 1) / 2 + l_seq;
      size_t sl = 0;
      sa_ptr[i] = find_sa(t, r + bs, &sl);
      im->sa_off[i + 1] = sl;
    }
    // offsets
    for (size_t i = 0; i < n; ++i) {
      im->qname_off[i + 1] += im->qname_off[i];
      im->cigar_off[i + 1] += im->cigar_off[i];
      im->sa_off[i + 1] += im->sa_off[i];
      if (want_meta >= 2) im->qual_off[i + 1] += im->qual_off[i];
      if (want_meta >= 3) im->raw_off[i + 1] += im->raw_off[i];
    }
    im->qname_blob.resize(im->qname_off[n]);
    im->cigar_blob.resize(im->cigar_off[n]);
    im->sa_blob.resize(im->sa_off[n]);
    if (want_meta >= 2) im->qual_blob.resize(im->qual_off[n]);
    if (want_meta >= 3) im->raw_blob.resize(im->raw_off[n]);
    // pass 2 (parallel): copy the pieces to their places
#pragma omp parallel for schedule(static) num_threads(b->threads)
    for (long i = 0; i < (long)n; ++i) {
      const uint8_t* r = kept[i].rec;
      uint8_t l_name = r[8];
      uint16_t n_cig = rd_u16(r + 12);
      uint32_t l_seq = kept[i].l_seq;
      size_t ql = im->qname_off[i + 1] - im->qname_off[i];
      if (ql) memcpy(&im->qname_blob[im->qname_off[i]], r + 32, ql);
      const uint8_t* cg = r + 32 + l_name;
      if (n_cig) memcpy(&im->cigar_blob[im->cigar_off[i]], cg, 4 * (size_t)n_cig);
      if (want_meta >= 3) {
        size_t bs = im->raw_off[i + 1] - im->raw_off[i];
        if (bs) memcpy(&im->raw_blob[im->raw_off[i]], r, bs);
      }
      if (want_meta >= 2 && l_seq) {
        const uint8_t* q = cg + 4 * (size_t)n_cig + (l_seq + 1) / 2;
        memcpy(&im->qual_blob[im->qual_off[i]], q, l_seq);
      }
      size_t sl = im->sa_off[i + 1] - im->sa_off[i];
      if (sl) memcpy(&im->sa_blob[im->sa_off[i]], sa_ptr[i], sl);
    }
  }
  release_retired();   // packed: nothing points into the retired chunks any more
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
  out->at_eof = (b->eof && b->cur.begin == b->cur.size) ? 1 : 0;
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
  if (b->cur.valid) b->put_buf(b->cur.data, b->cur.cap);
  if (b->next.valid) b->put_buf(b->next.data, b->next.cap);
  b->cur = Chunk();
  b->next = Chunk();
  b->ahead.valid = b->ahead2.valid = false;
  b->pending.clear();
  b->carry.clear();
  b->file_eof = false;
  b->eof = false;
  b->c_total = coff;
  b->skip_bytes = (size_t)(voffset & 0xffff);
  b->cur_qname.clear();
  b->seen_parts = 0;
  b->last_set_part = -1;
  b->u_limit = ~0ull;
  b->u_begin = 0;
  b->begin_coff = ~0ull;
  if (b->end_coff == coff) b->u_limit = b->u_total + b->end_in;   // (empty range)
  // an empty current chunk whose "own" data starts where the stream continues
  size_t cap = 0;
  uint8_t* p = b->get_buf(1, &cap);
  if (!p) {
    g_host_err = "out of memory";
    return KDF_ERR_ARG;
  }
  b->cur.data = p;
  b->cur.begin = b->cur.size = b->cur.own = 0;
  b->cur.ustart = b->u_total;
  b->cur.cap = cap;
  b->cur.valid = true;
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
  delete reinterpret_cast<kdf_bam_batch_impl*>(batch->impl);
  memset(batch, 0, sizeof(*batch));
}

}  // extern "C"
