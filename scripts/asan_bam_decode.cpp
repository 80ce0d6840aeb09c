// Decode a BAM through the C ABI of the host decoder, touching every byte it returns: built with
// -fsanitize=address,undefined this is the memory-safety check of the decode pipeline (the GPU
// pool has no compute-sanitizer, and Python cannot load an ASan build of the whole library).
//   g++ -std=c++17 -O1 -g -fopenmp -fsanitize=address,undefined -o /tmp/asan_bam_decode \
//       scripts/asan_bam_decode.cpp kmer_denovo_filter_b200/csrc/kdf_host.cpp \
//       kmer_denovo_filter_b200/csrc/kdf_inflate.cpp -lz
//   /tmp/asan_bam_decode file.bam mode want_meta [threads] [max_bases]     -> "ok <reads> <digest>" | "kdferror <text>"
// scripts/fuzz_bam_corrupt.py runs it on every mutated file when KDF_FUZZ_CMD names the binary.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/kdf.h"

static uint64_t mix(uint64_t h, const void* p, size_t n) {
  const uint8_t* b = (const uint8_t*)p;
  for (size_t i = 0; i < n; ++i) h = (h ^ b[i]) * 0x100000001b3ull;
  return h;
}

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  const int mode = atoi(argv[2]), meta = atoi(argv[3]);
  const int threads = argc > 4 ? atoi(argv[4]) : 3;
  const uint64_t max_bases = argc > 5 ? strtoull(argv[5], nullptr, 10) : 0;
  kdf_bam* h = nullptr;
  if (kdf_bam_open(argv[1], threads, &h) != KDF_OK) {
    printf("kdferror %s\n", kdf_host_last_error());
    return 0;
  }
  uint64_t reads = 0, dig = 1469598103934665603ull;
  for (;;) {
    kdf_bam_batch b;
    if (kdf_bam_next_batch(h, mode, max_bases, meta, &b) != KDF_OK) {
      printf("kdferror %s\n", kdf_host_last_error());
      kdf_bam_close(h);
      return 0;
    }
    const uint64_t n = b.n_reads, nw = (b.n_bases + 31) / 32;
    dig = mix(dig, b.codes, nw * 8);
    dig = mix(dig, b.valid, nw * 4);
    dig = mix(dig, b.read_starts, n * 8);
    dig = mix(dig, b.read_lens, n * 4);
    dig = mix(dig, b.rec_index, n * 8);
    dig = mix(dig, b.rec_uoff, n * 8);
    dig = mix(dig, b.fasta_keep, n);
    if (b.has_invalid) dig = mix(dig, b.invalid_pos, b.n_invalid * 4);
    if (meta) {
      dig = mix(dig, b.ref_id, n * 4);
      dig = mix(dig, b.pos, n * 4);
      dig = mix(dig, b.next_ref_id, n * 4);
      dig = mix(dig, b.next_pos, n * 4);
      dig = mix(dig, b.flag, n * 2);
      dig = mix(dig, b.mapq, n);
      dig = mix(dig, b.qname_off, (n + 1) * 8);
      dig = mix(dig, b.cigar_off, (n + 1) * 8);
      dig = mix(dig, b.sa_off, (n + 1) * 8);
      dig = mix(dig, b.qname_blob, b.qname_off[n]);
      dig = mix(dig, b.cigar_blob, b.cigar_off[n] * 4);
      dig = mix(dig, b.sa_blob, b.sa_off[n]);
      if (meta >= 2) {
        dig = mix(dig, b.qual_off, (n + 1) * 8);
        dig = mix(dig, b.qual_blob, b.qual_off[n]);
      }
      if (meta >= 3) {
        dig = mix(dig, b.raw_off, (n + 1) * 8);
        dig = mix(dig, b.raw_blob, b.raw_off[n]);
      }
    }
    reads += n;
    const int last = b.at_eof;
    kdf_bam_batch_free(&b);
    if (last || n == 0) break;
  }
  kdf_bam_close(h);
  printf("ok %llu %016llx\n", (unsigned long long)reads, (unsigned long long)dig);
  return 0;
}
