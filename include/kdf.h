/*
 * kdf.h — C ABI of libkdf_sm100.so, the B200 (sm_100a) k-mer engine that
 * replaces the Jellyfish / samtools subprocesses of jlanej/kmer_denovo_filter.
 *
 * The reference has no FFI: its seam is a set of Python functions that wrap
 * subprocesses.  Each entry point below names the reference call site it
 * replaces (paths relative to the reference's src/kmer_denovo_filter/).
 *
 * Conventions
 *   - plain C, no torch / C++ types in any signature;
 *   - every pointer marked DEV is device memory owned (allocated and freed) by
 *     the caller; the library never allocates or frees device memory;
 *   - every launch is ordered on the `stream` argument (a cudaStream_t passed
 *     as void*; NULL = legacy default stream) and returns without synchronising;
 *   - return value: 0 = KDF_OK, negative = error; kdf_last_error() gives the
 *     text of the last error on the calling thread;
 *   - no global mutable state besides that thread-local error string.
 *
 * Data layouts
 *   Stream ("packed read batch"): all sequences of a batch concatenated with
 *   exactly one invalid separator base between consecutive sequences.
 *     codes : uint64 words, 32 bases per word, base i of the stream in bits
 *             [62 - 2*(i%32), +2) of word i/32  (A=0 C=1 G=2 T=3; first base
 *             most significant, so a k-mer read out of the stream is already
 *             the Jellyfish/lexicographic integer key);
 *     valid : uint32 words, 32 bases per word, base i at bit 31 - (i%32);
 *             0 for N / IUPAC / separator / padding.
 *   Both arrays hold n_words = ceil(n_bases/32) words; bits past n_bases must
 *   be 0.  A window (k-mer) starting at p is counted iff valid[p..p+k) are all 1.
 *
 *   Table: open addressing over 4-slot buckets, linear probing by bucket; the
 *   caller provides one 32-byte-aligned device buffer of kdf_table_bytes():
 *     keys  : capacity slots of key_words u64 each (k <= 32: a bucket is 32
 *             bytes = one sector; k <= 64: 4 {lo, hi} keys = 64 bytes = two
 *             sectors of one line); empty = all-ones words (never a canonical
 *             k-mer); a probe reads one whole bucket with 256-bit loads;
 *     plane0: u32[capacity], plane1: u32[capacity] — two independent value
 *             planes (child count, parent count, reference flag ...), touched
 *             only when a key is found.
 *   capacity must be a multiple of 4.  Read-only tables whose keys fit in
 *   160 KB are copied to shared memory by the stream kernels.
 */
#ifndef KDF_H_
#define KDF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KDF_VERSION 2

/* status codes */
#define KDF_OK 0
#define KDF_ERR_ARG (-1)      /* invalid argument                               */
#define KDF_ERR_CUDA (-2)     /* CUDA runtime error (text in kdf_last_error)    */
#define KDF_ERR_NO_DEVICE (-3)
#define KDF_ERR_FULL (-4)     /* table full: Jellyfish would spill to .jf_N files
                                 (core/jellyfish_wrappers.py:59-70); we refuse  */
#define KDF_ERR_CAPACITY (-5) /* output buffer too small; *n_out has the need   */

/* update modes of kdf_count_stream / kdf_update_keys */
#define KDF_MODE_INSERT_COUNT 0     /* jellyfish count -C             : insert if absent, plane += arg */
#define KDF_MODE_INSERT_ONLY 1      /* jellyfish count --if priming   : insert if absent, planes untouched */
#define KDF_MODE_COUNT_IF_PRESENT 2 /* jellyfish count -C --if f.fa   : plane += arg only for existing keys */
#define KDF_MODE_MARK_IF_PRESENT 3  /* jellyfish query ref.jf (== 0?) : plane |= arg only for existing keys */

/* stats block written by the kernels (DEV, 4 x u64, caller zeroes it) */
#define KDF_STAT_WINDOWS 0 /* valid k-mer instances processed            */
#define KDF_STAT_FULL 1    /* != 0 : an insert found no free slot         */
#define KDF_STAT_HITS 2    /* instances that matched an existing key      */
#define KDF_STAT_NEW 3     /* keys newly inserted                         */
#define KDF_N_STATS 4

typedef struct kdf_table kdf_table; /* opaque host-side descriptor */

typedef struct kdf_stream {
  const uint64_t* codes; /* DEV */
  const uint32_t* valid; /* DEV */
  uint64_t n_bases;
} kdf_stream;

typedef struct kdf_device_props {
  int sm_count;
  int cc_major, cc_minor;
  int l2_bytes;
  uint64_t hbm_bytes;
  char name[128];
} kdf_device_props;

/* ---- housekeeping ------------------------------------------------------ */
int kdf_version(void);
const char* kdf_last_error(void);
int kdf_device_info(int device, kdf_device_props* out);

/* ---- table ------------------------------------------------------------- */
/* key_words for a k-mer size (1 for k<=32, 2 for k<=64, 0 = unsupported).   */
int kdf_key_words(int k);
/* bytes of device memory the caller must provide for `capacity` slots.      */
size_t kdf_table_bytes(uint64_t capacity, int key_words);
/* Capacity planning (replaces -s sizing: core/jellyfish_wrappers.py:73-107,
 * :155 `max(2n, 10M)`, :400 `max(2n, 1M)`): slots for n keys at load 0.5.    */
uint64_t kdf_table_capacity_for(uint64_t n_keys);

/* Wrap caller memory as a table for k-mer size k and clear it (async).
 * Replaces the hash that `jellyfish count -s` allocates
 * (discovery/pipeline.py:114-122; core/jellyfish_wrappers.py:167-176).      */
int kdf_table_create(kdf_table** out, int k, uint64_t capacity,
                     void* slots /*DEV*/, void* stream);
int kdf_table_destroy(kdf_table* t);
/* Attach a two-bit membership filter to a table whose keys are final (the primed
 * filter set of `count --if`, the proband-unique set): `words` is DEV memory of
 * n_words (a power of two) u32 owned by the caller and must outlive the table's use.
 * The probing stream ops (COUNT_IF_PRESENT, MARK_IF_PRESENT, the hit scan) then read
 * ONE 32-bit word per window — a filter of 4 bytes per key stays L2-resident where the
 * table does not — and only the windows it cannot rule out (the hits plus < 1 % false
 * positives) probe the table, in the queue drain with every lane busy.  Exact: a filter
 * has no false negatives.  Any inserting call or kdf_table_clear detaches it;
 * words == NULL detaches it explicitly.                                        */
int kdf_table_build_filter(kdf_table* t, uint32_t* words /*DEV*/, uint64_t n_words, void* stream);
int kdf_table_clear(kdf_table* t, void* stream);                 /* all slots -> empty */
int kdf_table_clear_plane(kdf_table* t, int plane, void* stream); /* one value plane -> 0 */
int kdf_table_info(const kdf_table* t, int* k, int* key_words, uint64_t* capacity);

/* ---- K1: rolling extract + canonicalise --------------------------------
 * One output per window start p in [0, n_bases): canonical key (lo[, hi]) and
 * a validity bit (same bit layout as `valid`).  Semantics of
 * kmer_utils.py:30-38 (canonicalize) and :91-121 (_extract_read_kmers), with
 * the Jellyfish rule that any non-ACGT base breaks the window.
 * out_hi may be NULL when k <= 32.                                          */
int kdf_extract_canonical(const kdf_stream* s, int k,
                          uint64_t* out_lo /*DEV n_bases*/,
                          uint64_t* out_hi /*DEV n_bases or NULL*/,
                          uint32_t* out_ok /*DEV n_words*/, void* stream);

/* ---- K1+K2 fused: extract, canonicalise and update the table -----------
 * mode INSERT_COUNT       : `jellyfish count -m k -C`
 *                           (discovery/pipeline.py:114-122,
 *                            core/jellyfish_wrappers.py:313-321, :411-419)
 * mode COUNT_IF_PRESENT   : `jellyfish count -C --if`
 *                           (core/jellyfish_wrappers.py:167-176,
 *                            discovery/pipeline.py:377-386)
 * mode MARK_IF_PRESENT    : reference subtraction, the dual of
 *                           `jellyfish query ref.jf -s cand.fa`
 *                           (discovery/pipeline.py:286-304): the reference is
 *                           streamed against the child table instead.
 * stats: DEV u64[KDF_N_STATS] accumulated (not cleared) by the kernel; may be NULL. */
int kdf_count_stream(kdf_table* t, const kdf_stream* s, int mode, int plane,
                     uint32_t arg, uint64_t* stats /*DEV*/, void* stream);

/* ---- K2 on explicit keys (filter priming, post all-to-all insert) ------ */
int kdf_update_keys(kdf_table* t, const uint64_t* lo /*DEV*/,
                    const uint64_t* hi /*DEV or NULL*/, uint64_t n, int mode,
                    int plane, uint32_t arg, uint64_t* stats /*DEV*/, void* stream);

/* Per-key accumulation: plane0[key_i] += add0[i], plane1[key_i] += add1[i] for
 * keys already in the table (add0/add1 may be NULL); keys not found are counted
 * in *n_missing (DEV u64, may be NULL).  Used to grow a table (re-hash) and to
 * merge per-rank counts; Jellyfish's analogue is `jellyfish merge`
 * (core/jellyfish_wrappers.py:335-366).                                      */
int kdf_add_planes(kdf_table* t, const uint64_t* lo /*DEV*/,
                   const uint64_t* hi /*DEV or NULL*/, uint64_t n,
                   const uint32_t* add0 /*DEV*/, const uint32_t* add1 /*DEV*/,
                   uint64_t* n_missing /*DEV*/, void* stream);

/* ---- K3: threshold + stream compaction ---------------------------------
 * Emits every occupied slot whose planes satisfy
 *   min0 <= plane0 <= max0  and  min1 <= plane1 <= max1.
 * Replaces `jellyfish dump -c -L n` (discovery/pipeline.py:207-211,
 * core/jellyfish_wrappers.py:262) and the `count == 0` / `<= parent_max_count`
 * line filters (discovery/pipeline.py:302, :528, :578).
 * Output order is unspecified (as is Jellyfish's dump order).  Any of the
 * output arrays may be NULL (count only).  n_out: DEV u64 counter, caller
 * zeroes it; on return-time it holds the number of matches even when
 * cap is too small (entries beyond cap are dropped).                        */
int kdf_threshold_compact(const kdf_table* t, uint32_t min0, uint32_t max0,
                          uint32_t min1, uint32_t max1,
                          uint64_t* out_lo /*DEV*/, uint64_t* out_hi /*DEV*/,
                          uint32_t* out_p0 /*DEV*/, uint32_t* out_p1 /*DEV*/,
                          uint64_t cap, uint64_t* n_out /*DEV*/, void* stream);

/* ---- K4: batched lookup -------------------------------------------------
 * `jellyfish query <jf> -s keys.fa` (kmer_utils.py:152-183,
 * discovery/pipeline.py:515-517, :565-567).  out_found[i] = 1/0; out_p0/out_p1
 * receive the planes (0 when absent); any may be NULL.                      */
int kdf_lookup_keys(const kdf_table* t, const uint64_t* lo /*DEV*/,
                    const uint64_t* hi /*DEV or NULL*/, uint64_t n,
                    uint8_t* out_found /*DEV*/, uint32_t* out_p0 /*DEV*/,
                    uint32_t* out_p1 /*DEV*/, void* stream);

/* ---- K4+K5: per-read membership scan + distinct reduction ---------------
 * For read r occupying stream bases [read_starts[r], read_starts[r]+read_lens[r]):
 *   out_ndistinct[r] = |{canonical k-mers of r present in the table}|
 *   out_nhits[r]     = number of window starts whose canonical k-mer is present
 * (core/bam_scanner.py:434-443; kmer_utils.py:209-238 scan_read).
 * Reads with out_ndistinct >= min_distinct and >= 1 additionally append every
 * hit as (stream position, slot index) to hit_pos/hit_slot (unordered);
 * *n_hits (DEV u64, caller zeroes) always counts them, entries beyond
 * hit_cap are dropped.  hit_pos/hit_slot may be NULL.  A read with more than
 * 1024 hit windows reports out_ndistinct = 0xFFFFFFFF and always emits its
 * hits; the caller finishes the distinct count from the emitted slot indices. */
#define KDF_NDISTINCT_OVERFLOW 0xFFFFFFFFu
int kdf_scan_reads(const kdf_table* t, const kdf_stream* s,
                   const uint64_t* read_starts /*DEV n_reads*/,
                   const uint32_t* read_lens /*DEV n_reads*/,
                   uint64_t n_reads, uint32_t min_distinct,
                   uint32_t* out_ndistinct /*DEV*/, uint32_t* out_nhits /*DEV*/,
                   uint64_t* hit_pos /*DEV*/, uint32_t* hit_slot /*DEV*/,
                   uint64_t hit_cap, uint64_t* n_hits /*DEV*/,
                   uint64_t* stats /*DEV*/, void* stream);

/* ---- K4+K5, sparse form: emit hits, then reduce per read -----------------
 * kdf_scan_stream_hits appends (stream position, slot) of every window whose
 * canonical k-mer is in the table (unordered; *n_hits counts all, entries
 * beyond hit_cap are dropped).  kdf_reduce_hits sorts the hits by position and
 * writes one record per read that has hits: read index, number of distinct
 * k-mers hit, number of hit windows, index of its first hit in the sorted
 * arrays (core/bam_scanner.py:434-443).  Records are unordered.  Scratch:
 * kdf_reduce_hits_scratch_bytes(n_hits) device bytes.  sorted_pos_out /
 * sorted_slot_out (n_hits entries) may be NULL.                              */
int kdf_scan_stream_hits(const kdf_table* t, const kdf_stream* s,
                         uint64_t* hit_pos /*DEV*/, uint32_t* hit_slot /*DEV*/,
                         uint64_t hit_cap, uint64_t* n_hits /*DEV*/,
                         uint64_t* stats /*DEV*/, void* stream);
size_t kdf_reduce_hits_scratch_bytes(uint64_t n_hits);
int kdf_reduce_hits(const uint64_t* hit_pos /*DEV*/, const uint32_t* hit_slot /*DEV*/,
                    uint64_t n_hits, const uint64_t* read_starts /*DEV, ascending*/,
                    uint64_t n_reads, void* scratch /*DEV*/, size_t scratch_bytes,
                    uint64_t* sorted_pos_out /*DEV*/, uint32_t* sorted_slot_out /*DEV*/,
                    uint64_t* rec_read /*DEV*/, uint32_t* rec_ndistinct /*DEV*/,
                    uint32_t* rec_nhits /*DEV*/, uint64_t* rec_first /*DEV*/,
                    uint64_t* n_recs /*DEV, caller zeroes*/, void* stream);

/* ---- K2p / K6: binning of canonical k-mers ---------------------------------
 * Bins are fixed-capacity regions of `bins`: bin b holds keys
 * [b*bin_cap, b*bin_cap + min(cursors[b], bin_cap)) (u64 keys for k <= 32,
 * {lo, hi} pairs for k <= 64).  cursors (DEV u64[n_parts]) and *overflow (DEV
 * u64) are zeroed by the caller once and accumulate over calls, so several
 * streams can be appended.  A key that does not fit sets *overflow != 0 and is
 * dropped: the caller must check it and retry with larger bins.
 *   by_owner == 0 : bin = hash range (top log2(n_parts) bits of the bucket
 *                   hash; n_parts a power of two <= 512) — input of
 *                   kdf_count_bins; replaces nothing in the reference, it is
 *                   how `jellyfish count` (discovery/pipeline.py:114-122) is
 *                   kept out of DRAM-random-access territory on the GPU;
 *   by_owner == 1 : bin = owner rank of a multi-GPU run (any n_parts <= 512),
 *                   in front of the all-to-all;
 *   by_owner == R >= 2 : composite, bin = owner * (n_parts / R) + hash range, for R
 *                   owner ranks (n_parts / R a power of two): what a rank receives is
 *                   already binned for kdf_count_bins_multi.                  */
int kdf_bin_stream(const kdf_stream* s, int k, int by_owner, int n_parts,
                   uint64_t* bins /*DEV*/, uint64_t bin_cap, uint64_t* cursors /*DEV*/,
                   uint64_t* overflow /*DEV*/, uint64_t* stats /*DEV or NULL*/, void* stream);
/* The same for the window starts of words [first_word, first_word + n_words) only
 * (bases of later words are still read for the windows that span them), so that a
 * stream can be binned chunk by chunk while its tail is still being copied in.  */
int kdf_bin_stream_range(const kdf_stream* s, uint64_t first_word, uint64_t n_words, int k,
                         int by_owner, int n_parts, uint64_t* bins /*DEV*/, uint64_t bin_cap,
                         uint64_t* cursors /*DEV*/, uint64_t* overflow /*DEV*/,
                         uint64_t* stats /*DEV or NULL*/, void* stream);
/* The same kernel with one base pointer per bin (bin_ptrs: DEV array of n_parts
 * device pointers, each to bin_cap keys).  The pointers may address PEER memory
 * mapped over NVLink / NVSwitch: with by_owner == 1 and bin r placed in rank r's
 * receive buffer, the shared-memory-staged flush of the binning kernel is the
 * all-to-all itself — compute and transfer in one kernel, no send buffer, no
 * NCCL call.  cursors / overflow stay local to the sender (each (sender, owner)
 * pair has its own region), the sender publishes min(cursors, bin_cap) afterwards. */
int kdf_bin_stream_to(const kdf_stream* s, int k, int by_owner, int n_parts,
                      uint64_t* const* bin_ptrs /*DEV*/, uint64_t bin_cap,
                      uint64_t* cursors /*DEV*/, uint64_t* overflow /*DEV*/,
                      uint64_t* stats /*DEV or NULL*/, void* stream);
int kdf_bin_stream_to_range(const kdf_stream* s, uint64_t first_word, uint64_t n_words, int k,
                            int by_owner, int n_parts, uint64_t* const* bin_ptrs /*DEV*/,
                            uint64_t bin_cap, uint64_t* cursors /*DEV*/,
                            uint64_t* overflow /*DEV*/, uint64_t* stats /*DEV or NULL*/,
                            void* stream);
int kdf_bin_keys(const uint64_t* lo /*DEV*/, const uint64_t* hi /*DEV or NULL*/, uint64_t n,
                 int k, int by_owner, int n_parts, uint64_t* bins /*DEV*/, uint64_t bin_cap,
                 uint64_t* cursors /*DEV*/, uint64_t* overflow /*DEV*/, void* stream);

/* Multi-pass binning — what the reference gets from Jellyfish's sized hash +
 * spill-and-merge (core/jellyfish_wrappers.py:73-107, 335-366;
 * discovery/pipeline.py:114-122, 186-189): the bins of a whole-genome sample
 * (8 bytes x every k-mer instance) do not fit HBM, so the hash ranges are split into
 * 2^pass_log2 groups by their TOP bits and the stream is re-extracted once per
 * group: this call bins only the keys of group pass_val (0 <= pass_val <
 * 2^pass_log2), into n_parts bins that are the group's own sub-ranges (by_owner 0 or
 * R >= 2; with by_owner == 1 the pass only filters).  kdf_count_bins_pass with the
 * same (pass_log2, pass_val) counts them: over all passes every key is counted
 * exactly once, in slices that cover 1 / (2^pass_log2 * n_parts) of the hash space.
 * Exactly one of bins / bin_ptrs is non-NULL (local regions, or one — possibly
 * peer-mapped — pointer per bin as in kdf_bin_stream_to).  pass_log2 == 0 is
 * kdf_bin_stream_range / kdf_bin_stream_to_range.                              */
int kdf_bin_stream_pass(const kdf_stream* s, uint64_t first_word, uint64_t n_words, int k,
                        int by_owner, int n_parts, uint64_t* bins /*DEV or NULL*/,
                        uint64_t* const* bin_ptrs /*DEV or NULL*/, uint64_t bin_cap,
                        uint64_t* cursors /*DEV*/, uint64_t* overflow /*DEV*/,
                        uint64_t* stats /*DEV or NULL*/, int pass_log2, uint32_t pass_val,
                        void* stream);
int kdf_bin_keys_pass(const uint64_t* lo /*DEV*/, const uint64_t* hi /*DEV or NULL*/, uint64_t n,
                      int k, int by_owner, int n_parts, int pass_log2, uint32_t pass_val,
                      uint64_t* bins /*DEV*/, uint64_t bin_cap, uint64_t* cursors /*DEV*/,
                      uint64_t* overflow /*DEV*/, void* stream);

/* K2 on hash-range bins: apply `mode` (kdf_update_keys semantics) to the keys of
 * every bin, bin after bin.  bins / cursors as written by kdf_bin_stream with
 * by_owner == 0 (for 128-bit keys the {lo, hi} pairs of a bin are interleaved).
 * The table's bucket index grows with the hash, so bin p only touches the p-th
 * n_parts-th of the table: a read-only table far larger than L2 (the filter set of
 * `jellyfish count --if`, discovery/pipeline.py:377-386, at whole-genome scale) is
 * probed with L2-resident traffic after one streaming binning pass, instead of one
 * random DRAM sector per k-mer.                                               */
int kdf_update_bins(kdf_table* t, int n_parts, const uint64_t* bins /*DEV*/, uint64_t bin_cap,
                    const uint64_t* cursors /*DEV*/, int mode, int plane, uint32_t arg,
                    uint64_t* stats /*DEV or NULL*/, void* stream);

/* Count every hash-range bin in an L2-resident table slice and emit.
 * For each bin p: clear `slice` (a table of slice_capacity slots in caller
 * memory of kdf_table_bytes(slice_capacity, key_words)); insert+count the child
 * keys of bin p into plane 0; OR 1 into plane 1 for every reference key of bin
 * p that is present (ref_bins may be NULL); then emit, exactly as
 * kdf_threshold_compact does, the slots with min0 <= plane0 <= max0 and
 * min1 <= plane1 <= max1.  This is `jellyfish count -C` + `dump -c -L` +
 * `query ref.jf` (discovery/pipeline.py:114-122, :207-211, :286-304) in one
 * call whose table never leaves L2.
 * counters: DEV u64[6], caller zeroes: [0] keys applied, [1] != 0 slice full
 * (results invalid: retry with a larger slice), [2] instances that found their
 * key, [3] distinct keys, [4] keys with plane0 >= count_min0, [5] occupied
 * slots seen by the emit pass.  n_out as in kdf_threshold_compact.
 *
 * Packed form.  The discovery chain discards the counts after thresholding
 * (discovery/pipeline.py:207-226 keeps only the k-mer column of `dump -c -L`),
 * and k is odd (utils.py:299-311), so a key leaves >= 2 spare bits in its most
 * significant word.  When the call asks only for "plane0 >= min0" (max0 ==
 * UINT32_MAX, out_p0 == out_p1 == NULL, min0 >= 1 fits the spare bits, min1 == 0,
 * max1 in {0, UINT32_MAX}, count_min0 <= 1 or == min0) the slice holds keys only:
 * a counter saturating at min0 lives in the spare bits, so every copy of a k-mer
 * after the min0-th is a plain bucket read (no atomic, no plane traffic), and
 * "in the reference" is one more state of that field.  Same outputs and counters;
 * keys must be canonical (the all-T key doubles as the empty marker).
 * kdf_count_bins_packed() says whether a call takes this form.               */
int kdf_count_bins_packed(int k, uint32_t min0, uint32_t max0, uint32_t min1, uint32_t max1,
                          uint32_t count_min0, int want_planes);
int kdf_count_bins(int k, int n_parts, const uint64_t* child_bins /*DEV*/,
                   uint64_t child_bin_cap, const uint64_t* child_cursors /*DEV*/,
                   const uint64_t* ref_bins /*DEV or NULL*/, uint64_t ref_bin_cap,
                   const uint64_t* ref_cursors /*DEV or NULL*/, void* slice /*DEV*/,
                   uint64_t slice_capacity, uint32_t min0, uint32_t max0, uint32_t min1,
                   uint32_t max1, uint64_t* out_lo /*DEV*/, uint64_t* out_hi /*DEV*/,
                   uint32_t* out_p0 /*DEV*/, uint32_t* out_p1 /*DEV*/, uint64_t out_cap,
                   uint64_t* n_out /*DEV*/, uint32_t count_min0, uint64_t* counters /*DEV*/,
                   void* stream);

/* The same with the bins of n_src sources (multi-GPU: one region per sending
 * rank, filled by kdf_bin_stream_to): bins are laid out [source][hash range]
 * [bin_cap] and cursors [source][hash range].  sub_split (a power of two) counts
 * every bin in that many passes, each over one sub-range of its hashes, so that the
 * table slice can stay L2-sized when there are few, large bins.                */
int kdf_count_bins_multi(int k, int n_parts, int n_src, int sub_split,
                         const uint64_t* child_bins /*DEV*/,
                         uint64_t child_bin_cap, const uint64_t* child_cursors /*DEV*/,
                         const uint64_t* ref_bins /*DEV or NULL*/, uint64_t ref_bin_cap,
                         const uint64_t* ref_cursors /*DEV or NULL*/, void* slice /*DEV*/,
                         uint64_t slice_capacity, uint32_t min0, uint32_t max0, uint32_t min1,
                         uint32_t max1, uint64_t* out_lo /*DEV*/, uint64_t* out_hi /*DEV*/,
                         uint32_t* out_p0 /*DEV*/, uint32_t* out_p1 /*DEV*/, uint64_t out_cap,
                         uint64_t* n_out /*DEV*/, uint32_t count_min0, uint64_t* counters /*DEV*/,
                         void* stream);
/* ... and for the bins of one pass of a multi-pass count (kdf_bin_stream_pass):
 * slice p of this call covers hash range (pass_val * n_parts + p) * sub_split .. of
 * 2^pass_log2 * n_parts * sub_split.  Outputs (n_out, counters) accumulate over the
 * passes when the caller does not zero them in between.                        */
int kdf_count_bins_pass(int k, int n_parts, int n_src, int sub_split, int pass_log2,
                        uint32_t pass_val, const uint64_t* child_bins /*DEV*/,
                        uint64_t child_bin_cap, const uint64_t* child_cursors /*DEV*/,
                        const uint64_t* ref_bins /*DEV or NULL*/, uint64_t ref_bin_cap,
                        const uint64_t* ref_cursors /*DEV or NULL*/, void* slice /*DEV*/,
                        uint64_t slice_capacity, uint32_t min0, uint32_t max0, uint32_t min1,
                        uint32_t max1, uint64_t* out_lo /*DEV*/, uint64_t* out_hi /*DEV*/,
                        uint32_t* out_p0 /*DEV*/, uint32_t* out_p1 /*DEV*/, uint64_t out_cap,
                        uint64_t* n_out /*DEV*/, uint32_t count_min0, uint64_t* counters /*DEV*/,
                        void* stream);

/* ---- K7: coverage of the reference by hit k-mers -------------------------
 * Replaces _collect_kmer_ref_positions (core/bam_scanner.py:97-117) and the
 * per-contig Counter merges of _anchor_and_cluster (discovery/pipeline.py:851-855)
 * for the informative reads of one batch.  Hits are (read index into the arrays
 * below, window start inside the read), sorted by (read, start).  Per read:
 * contig id, 0-based reference start, and BAM CIGAR words (len << 4 | op) at
 * cigar[read_cig_off[r] .. read_cig_off[r+1]).  Output: n_out runs of
 *   key   = contig << 40 | ref_pos << 1 | first      (sorted ascending)
 *   count = hit k-mers covering ref_pos with that `first` flag
 * `first` = 1 for the keys of a (read, position) seen for the first time in that read,
 * so sum(count) over both flags is the k-mer coverage and sum over first == 1 the number
 * of reads (the reference's read_coverage).  The last run may be the all-ones padding
 * key: ignore it.  out_keys / out_counts need n_hits * k entries.             */
size_t kdf_hit_coverage_scratch_bytes(uint64_t n_hits, int k);
int kdf_hit_coverage(const uint32_t* hit_read /*DEV*/, const uint32_t* hit_off /*DEV*/,
                     uint64_t n_hits, int k, const int32_t* read_contig /*DEV*/,
                     const int64_t* read_ref_start /*DEV*/, const uint64_t* read_cig_off /*DEV*/,
                     const uint32_t* cigar /*DEV*/, void* scratch /*DEV*/, size_t scratch_bytes,
                     uint64_t* out_keys /*DEV*/, uint32_t* out_counts /*DEV*/,
                     uint64_t* n_out /*DEV*/, void* stream);
/* the same expansion on the host (test hook, host pointers): keys[n_hits * k],
 * unused entries all ones, unsorted                                          */
int kdf_debug_hit_coverage_host(const uint32_t* hit_read, const uint32_t* hit_off, uint64_t n_hits,
                                int k, const int32_t* read_contig, const int64_t* read_ref_start,
                                const uint64_t* read_cig_off, const uint32_t* cigar,
                                uint64_t* keys);

/* ---- sparse validity ------------------------------------------------------
 * The validity bitmap is all ones except one separator per read and the rare N /
 * IUPAC base, so a stream can cross PCIe as codes + the ascending list of its invalid
 * positions (kdf_bam_batch.invalid_pos, or kdf_invalid_positions on the host: returns
 * the count, fills out[0 .. min(count, cap)); ~0 if n_bases > 2^32) and the bitmap be
 * rebuilt on the device: valid[w] = all ones inside n_bases, then the listed bits
 * cleared.  Same bitmap, a third less H2D traffic for the whole stream.          */
uint64_t kdf_invalid_positions(const uint32_t* valid /*HOST*/, uint64_t n_bases, uint32_t* out /*HOST or NULL*/,
                               uint64_t cap);
int kdf_valid_from_invalid(uint32_t* valid /*DEV n_words*/, uint64_t n_bases,
                           const uint32_t* invalid_pos /*DEV*/, uint64_t n_invalid, void* stream);

/* ---- host helpers (CPU, no device) --------------------------------------
 * Pack ASCII sequences into the stream layout.  seqs: concatenated bytes,
 * offsets[n_seqs+1].  Returns the stream length in bases (sum of lengths +
 * n_seqs-1 separators) and fills codes/valid (n_words each, caller-allocated,
 * pre-zeroed not required) and read_offsets[n_seqs+1] (may be NULL).
 * Upper/lower-case ACGT are valid; everything else is invalid.
 * With codes == NULL only the length is computed.                           */
uint64_t kdf_pack_sequences(const char* seqs, const uint64_t* offsets,
                            uint64_t n_seqs, uint64_t* codes, uint32_t* valid,
                            uint64_t* read_offsets);

/* A whole FASTA text (HOST memory, e.g. the reference genome) into the stream layout,
 * multi-threaded: header lines (">...") are dropped, line ends and blanks inside the
 * sequence lines skipped, records separated by one invalid base — what
 * `jellyfish count -C ref.fa` sees of the file (reference: kmer_utils.py / discovery
 * pipeline.py:286-332 build ref.k31.jf from it).  kdf_fasta_layout gives the number of
 * records and the stream length so that the caller can allocate codes / valid
 * ((n_bases+31)/32 words each, at least one) and seq_starts / seq_lens (n_seqs).   */
int kdf_fasta_layout(const uint8_t* text /*HOST*/, uint64_t n, int n_threads, uint64_t* n_seqs,
                     uint64_t* n_bases);
int kdf_fasta_pack(const uint8_t* text /*HOST*/, uint64_t n, int n_threads, uint64_t* codes /*HOST*/,
                   uint32_t* valid /*HOST*/, uint64_t* seq_starts /*HOST*/, uint64_t* seq_lens /*HOST*/);

/* ---- host BGZF/BAM decode -> packed batches (CPU, multi-threaded) --------
 * Replaces, for BAM input, `samtools fasta -F 0xD00` in front of Jellyfish
 * (core/jellyfish_wrappers.py:159-165; discovery/pipeline.py:106-112, 369-375)
 * and the pysam iteration of the anchoring scan (core/bam_scanner.py:405-414).
 *   KDF_BAM_FASTA : drop flag & 0xD00; inside a run of consecutive records with
 *                   the same QNAME keep the first record of each read-part
 *                   (READ1 / READ2 / other) — what samtools fasta emits.
 *   KDF_BAM_SCAN  : drop secondary (0x100) and duplicate (0x400) records only.
 *   KDF_BAM_ALL   : every record.
 * A batch owns its host memory until kdf_bam_batch_free.  Metadata arrays are
 * filled only when want_meta != 0, base qualities only when want_meta >= 2
 * (VCF-mode child reads, vcf/pipeline.py:671-690), the raw records only when
 * want_meta >= 3 (informative-reads BAM, discovery/pipeline.py:1979-2079).
 * max_bases == 0 reads to end of file.                                       */
#define KDF_BAM_FASTA 0
#define KDF_BAM_SCAN 1
#define KDF_BAM_ALL 2

typedef struct kdf_bam kdf_bam;

typedef struct kdf_bam_batch {
  void* impl;
  uint64_t n_reads;
  uint64_t n_bases;              /* stream length incl. separators            */
  const uint64_t* codes;         /* HOST, stream layout                       */
  const uint32_t* valid;         /* HOST                                      */
  const uint64_t* read_starts;   /* HOST n_reads                              */
  const uint32_t* read_lens;     /* HOST n_reads                              */
  const uint64_t* rec_index;     /* HOST n_reads: record number in the file   */
  const int32_t* ref_id;         /* --- metadata (want_meta) ---              */
  const int32_t* pos;
  const int32_t* next_ref_id;
  const int32_t* next_pos;
  const uint16_t* flag;
  const uint8_t* mapq;
  const uint64_t* qname_off;     /* n_reads+1 offsets into qname_blob         */
  const char* qname_blob;
  const uint64_t* cigar_off;     /* n_reads+1 offsets into cigar_blob         */
  const uint32_t* cigar_blob;    /* BAM encoding: len<<4 | op                 */
  const uint64_t* sa_off;        /* n_reads+1 offsets into sa_blob (SA:Z)     */
  const char* sa_blob;
  const uint64_t* qual_off;      /* n_reads+1 offsets into qual_blob (want_meta >= 2) */
  const uint8_t* qual_blob;      /* Phred base qualities (0xFF = absent)     */
  const uint64_t* raw_off;       /* n_reads+1 offsets into raw_blob (want_meta >= 3) */
  const uint8_t* raw_blob;       /* the BAM records (bytes after block_size) */
  int at_eof;
  int has_invalid;               /* invalid_pos / n_invalid are set (n_bases <= 2^32) */
  const uint32_t* invalid_pos;   /* HOST, ascending: positions of the invalid bases (separators,
                                    N / IUPAC) — the sparse form of `valid`, a fifth of its size at
                                    150-base reads; kdf_valid_from_invalid rebuilds the bitmap on
                                    the device so that only codes + this list cross PCIe          */
  uint64_t n_invalid;
  const uint64_t* rec_uoff;      /* HOST n_reads: offset of the record in the file's uncompressed
                                    stream (argument of kdf_bam_fetch_records)                   */
  const uint8_t* fasta_keep;     /* HOST n_reads: 1 = the record is part of the KDF_BAM_FASTA stream
                                    (set in every mode: a scan-mode decode + this mask gives the
                                    counting stream of the same file without a second pass)      */
} kdf_bam_batch;

int kdf_bam_open(const char* path, int n_threads, kdf_bam** out);
void kdf_bam_close(kdf_bam* b);
const char* kdf_bam_header_text(const kdf_bam* b, uint64_t* len); /* SAM header text */
int kdf_bam_n_refs(const kdf_bam* b);
const char* kdf_bam_ref_name(const kdf_bam* b, int i);
int64_t kdf_bam_ref_len(const kdf_bam* b, int i);
int kdf_bam_next_batch(kdf_bam* b, int mode, uint64_t max_bases, int want_meta,
                       kdf_bam_batch* out);
void kdf_bam_batch_free(kdf_bam_batch* batch);
const char* kdf_host_last_error(void);
/* Ranges of a BAM, for region fetches through the .bai (the reference's
 * bam.fetch(chrom, pos, pos + 1), vcf/pipeline.py:619-726) and for sharding a file over
 * the ranks of a multi-GPU run (rank r reads a contiguous BGZF range of
 * every BAM).  kdf_bam_seek repositions the reader at a BGZF virtual offset (block file
 * offset << 16 | offset in the block) that starts a record — an entry of the .bai linear
 * index — and drops its read-ahead; the FASTA stream's QNAME-run state starts afresh.
 * kdf_bam_set_end makes the reader report end of file at the first record that starts
 * at or after the given virtual offset (~0: no limit).
 * kdf_bam_set_begin (right after a seek to an EARLIER record): records that start before
 * this virtual offset only update the QNAME-run state and are not delivered, so a run
 * of same-QNAME records that straddles a rank boundary collapses as in a sequential read. */
int kdf_bam_seek(kdf_bam* b, uint64_t voffset);
int kdf_bam_set_chunk_bytes(kdf_bam* b, uint64_t n);   /* read-ahead per pipeline chunk (default 64 MB) */
int kdf_bam_set_begin(kdf_bam* b, uint64_t voffset);
int kdf_bam_set_end(kdf_bam* b, uint64_t voffset);
/* The raw BAM records (bytes after block_size) at the given uncompressed offsets
 * (kdf_bam_batch.rec_uoff of records this reader has already decoded), by inflating
 * only the blocks that hold them: what the informative-reads BAM writer needs
 * (discovery/pipeline.py:1979-2079) without decoding the file once more.
 * out_off[n+1] receives the record boundaries, *needed the total size; records are
 * copied only while they fit out_cap (call again with a larger buffer).          */
int kdf_bam_fetch_records(kdf_bam* b, const uint64_t* uoffs /*HOST*/, uint64_t n,
                          uint8_t* out /*HOST or NULL*/, uint64_t out_cap,
                          uint64_t* out_off /*HOST n+1*/, uint64_t* needed);
/* `data` as a BGZF file (multi-threaded deflate + EOF marker): the container of the
 * informative-reads BAM and of the bgzip'd annotated VCF (vcf/pipeline.py:1307-1357,
 * :1640-1700; the reference writes them through pysam / bgzip).  block_coff (HOST,
 * block_cap entries, may be NULL) receives the file offset of every block (and of the
 * EOF block) for BAI / TBI virtual offsets; blocks hold 0xff00 input bytes each.    */
int kdf_bgzf_write(const char* path, const uint8_t* data /*HOST*/, uint64_t n, int level,
                   int n_threads, uint64_t* block_coff /*HOST or NULL*/, uint64_t block_cap,
                   uint64_t* n_blocks);

/* One whole BGZF block (header .. CRC32/ISIZE trailer, `csize` bytes) -> its `usize`
 * inflated bytes, CRC-checked unless verify_crc == 0: the per-block step of every BAM
 * read above, exposed so it can be checked against zlib block by block
 * (tests/test_host_inflate.py).  impl 0: the library's own DEFLATE decoder
 * (csrc/kdf_inflate.cpp), 1: zlib's inflate().  htslib does this in bgzf.c
 * (bgzf_uncompress), which the reference reaches through `samtools fasta`
 * (kmer_utils.py:310-340) and pysam.  KDF_OK, or KDF_ERR_ARG for a corrupt block.     */
int kdf_bgzf_inflate_block(const uint8_t* src /*HOST*/, uint32_t csize, uint8_t* dst /*HOST*/,
                           uint32_t usize, int verify_crc, int impl);
/* CRC-32 (gzip polynomial) of a host buffer with the library's routine.              */
uint32_t kdf_crc32(const uint8_t* data /*HOST*/, uint64_t n);

/* Test hook: runs the device window-iterator templates on the CPU (host
 * instantiation of the same code) so the bit manipulation can be verified
 * without a GPU.  Not used by any product path.                             */
int kdf_debug_extract_host(const uint64_t* codes, const uint32_t* valid,
                           uint64_t n_bases, int k, int use_random_access,
                           uint64_t* out_lo, uint64_t* out_hi, uint8_t* out_ok);

/* Test hook: the device hash functions on the host (partition, bucket and
 * owner of each key); any output may be NULL.                               */
int kdf_debug_hash_host(const uint64_t* lo, const uint64_t* hi, uint64_t n, int key_words,
                        int log2_parts, uint32_t n_buckets, uint32_t n_ranks,
                        uint32_t* out_part, uint32_t* out_bucket, uint32_t* out_owner);

/* Random-access microbenchmark used for the "HBM random-access roofline"
 * (SURVEY §8d): n_ops uniformly random 32-byte sector reads (atomic == 0) or
 * sector read + 4-byte atomic add (atomic == 1) over buf_bytes.            */
int kdf_bench_random_access(void* buf /*DEV*/, uint64_t buf_bytes,
                            uint64_t n_ops, int atomic, uint64_t* sink /*DEV*/,
                            void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KDF_H_ */
