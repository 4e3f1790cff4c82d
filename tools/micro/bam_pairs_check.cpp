// tools/micro/bam_pairs_check.cpp -> tools/bin/rg_pairs_check — test and timing driver for the input side of
// RUFUS.Filter (no GPU): reads the pairs either from two FASTQ files (tools/rg_fastq_batch.hpp:read_pair_batch, what
// tools/rufus_filter.cpp does) or from a BAM (tools/rg_bam_pairs.hpp:bam_pairs), writes the four lines of every pair
// to STUB.mate1.fastq / STUB.mate2.fastq and prints a digest of the arrays a batch hands to rg_filter_pairs
// (sequences, qualities, offsets, name lengths), batch by batch.
//   rg_pairs_check [-t THREADS] [-b BATCH_PAIRS] [-n] MATE1.fastq MATE2.fastq STUB
//   rg_pairs_check [-t THREADS] [-b BATCH_PAIRS] [-n] [-q] X.bam STUB     (-n: no files, digest only; -q: no digest either)
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <string>

#include "../rg_bam_pairs.hpp"

[[noreturn]] void rgbam::fail(const std::string& m) {
  fprintf(stderr, "rg_pairs_check: %s\n", m.c_str());
  fflush(stderr);
  _exit(1);
}

static uint64_t fnv(uint64_t h, const void* p, size_t n) {   // (zlib's CRC-32 of the block, folded into a running FNV hash: fast enough not to dominate the timing)
  const uint64_t c = crc32(crc32(0L, Z_NULL, 0), (const unsigned char*)p, (uInt)n);
  return (h ^ c ^ ((uint64_t)n << 32)) * 1099511628211ull;
}

int main(int argc, char** argv) {
  unsigned threads = 1;
  uint64_t batch = 1u << 16;
  bool write_files = true, digest = true;
  std::vector<const char*> pos;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "-t") && i + 1 < argc) threads = (unsigned)atoi(argv[++i]);
    else if (!strcmp(argv[i], "-b") && i + 1 < argc) batch = strtoull(argv[++i], nullptr, 10);
    else if (!strcmp(argv[i], "-n")) write_files = false;   // timing: digest only
    else if (!strcmp(argv[i], "-q")) digest = false;        // timing: not even the digest (what is left is the reader)
    else pos.push_back(argv[i]);
  }
  if (pos.size() != 2 && pos.size() != 3) return 2;
  const std::string stub = pos.back();
  FILE* o1 = fopen((stub + ".mate1.fastq").c_str(), "wb");
  FILE* o2 = fopen((stub + ".mate2.fastq").c_str(), "wb");
  FILE* f1 = fopen(pos[0], "rb");
  if (!o1 || !o2 || !f1) die("cannot open the files");
  uint64_t h = 1469598103934665603ull, pairs = 0, batches = 0;
  Batch b1, b2;
  b1.clear();
  b2.clear();
  auto take = [&](Batch& x, Batch& y) {
    for (Batch* b : {&x, &y}) {
      if (!digest) break;
      h = fnv(h, b->seq.data(), b->seq.size());
      h = fnv(h, b->qual.data(), b->qual.size());
      h = fnv(h, b->offs.data(), b->offs.size() * 8);
      h = fnv(h, b->text_offs.data(), b->text_offs.size() * 8);
      h = fnv(h, b->name_len.data(), b->name_len.size() * 4);
    }
    if (write_files) {
      fwrite(x.text.data(), 1, x.text.size(), o1);
      fwrite(y.text.data(), 1, y.text.size(), o2);
    }
    pairs += x.n();
    ++batches;
    x.clear();
    y.clear();
  };
  LineReader r1(f1);
  if (pos.size() == 2) {
    if (r1.peek() != 0x1f) die("one input that is not a BAM file");
    rgbam::bam_pairs(f1, r1.buf.data() + r1.pos, r1.len - r1.pos, threads, b1, b2, batch, take);
    if (b1.n()) take(b1, b2);
  } else {
    FILE* f2 = fopen(pos[1], "rb");
    if (!f2) die("cannot open the files");
    LineReader r2(f2);
    for (bool more = true; more;) {
      more = read_pair_batch(r1, r2, b1, b2, batch);
      if (b1.n()) take(b1, b2);
    }
  }
  fclose(o1);
  fclose(o2);
  printf("%llu pairs in %llu batches, digest %016llx\n", (unsigned long long)pairs, (unsigned long long)batches, (unsigned long long)h);
  return 0;
}
