// rg_bam_input.hpp — `jellyfish count X.bam`: a BAM file as a sequence input, decoded in process. What the
// reference feeds its counter for a BAM sample is `samtools view -F 3328 X.bam | PassThroughSamCheck X.chr` through
// a FIFO (runRufus.sh:599, RunJellyForRUFUS.sh:28): SAM text, re-parsed into FASTQ, re-parsed by the counter's
// reader. Here the records go from the BGZF members straight into the pinned upload buffer — the bases exactly as
// that chain would spell them (4-bit codes -> `=ACMGRSVTWYHKDBN`, `*` for an empty SEQ; PassThroughSamCheck copies
// field 10 as it is) followed by the 'N' that separates reads (mer_overlap_sequence_parser.hpp:179-207) — on
// `-t` threads (tools/rg_bam_pipeline.hpp). FLAG filter: -F 3328 as in the generator line, RG_BAM_EXCLUDE_FLAGS /
// RG_BAM_REQUIRE_FLAGS to change it. SURVEY §8f rank 2 ("host ingest feeding pinned buffers").
#pragma once
#include "rg_bam_pipeline.hpp"
#include "rg_cli.hpp"
#include "rg_fastq_parallel.hpp"

namespace rgbam {

struct BasePairs {
  char t[256][2];
  BasePairs() {
    for (int b = 0; b < 256; ++b) {
      t[b][0] = "=ACMGRSVTWYHKDBN"[b >> 4];
      t[b][1] = "=ACMGRSVTWYHKDBN"[b & 15];
    }
  }
};

// the bases of one record + 'N'
static void sequence_record(const unsigned char* r, uint32_t bs, const Header&, Text& o) {
  static const BasePairs pairs;
  if (bs < 32) fail("corrupt BAM record");
  const uint32_t l_name = r[8], n_cigar = le16(r + 12), l_seq = le32(r + 16);
  const uint64_t fixed = 32ull + l_name + 4ull * n_cigar + ((uint64_t)l_seq + 1) / 2 + l_seq;
  if (fixed > bs || l_name == 0) fail("corrupt BAM record (lengths)");
  char* w = o.room((size_t)l_seq + 4);
  if (l_seq == 0) {
    *w++ = '*';
  } else {
    const unsigned char* p = r + 32 + l_name + 4ull * n_cigar;
    for (uint32_t i = 0; i < (l_seq + 1) / 2; ++i) memcpy(w + 2 * i, pairs.t[p[i]], 2);   // (an odd length writes one too many: overwritten)
    w += l_seq;
  }
  *w++ = 'N';
  o.done(w);
}

static inline uint32_t flags_from_env(const char* name, uint32_t fallback) {
  const char* e = getenv(name);
  return e && *e ? (uint32_t)strtoul(e, nullptr, 0) : fallback;
}

// the BAM counterpart of read_sequences (tools/rg_cli.hpp): `head` = the bytes of the file already read
template <typename Buf, typename Flush>
static inline void bam_sequences(FILE* f, const char* head, size_t head_len, Buf& out, const size_t& chunk, Flush&& flush, unsigned threads) {
  Pipeline pl;
  pl.f = f;
  pl.prefix = (const unsigned char*)head;
  pl.prefix_len = head_len;
  pl.threads = threads ? threads : 1;
  pl.excl = flags_from_env("RG_BAM_EXCLUDE_FLAGS", 3328);
  pl.req = flags_from_env("RG_BAM_REQUIRE_FLAGS", 0);
  pl.render = sequence_record;
  setvbuf(f, nullptr, _IONBF, 0);   // the pipeline's reader asks for a megabyte at a time
  pl.run([&](const char* p, size_t n) { rgfq::append_in_pieces(p, n, out, chunk, flush); });
}

}  // namespace rgbam

// FASTA / FASTQ / BAM by the first byte of the input (0x1f = a gzip member: BGZF); `threads` serves the BAM decode
// and the scan of a large regular FASTQ file
template <typename Buf, typename Flush>
static inline void read_input(FILE* f, const char* path, Buf& out, const size_t& chunk, Flush&& flush, unsigned threads) {
  LineReader lr(f);
  const int c0 = lr.peek();
  if (c0 == 0x1f) {
    rgbam::bam_sequences(f, lr.buf.data() + lr.pos, lr.len - lr.pos, out, chunk, flush, threads);
    return;
  }
  if (rgfq::read_fastq_parallel(f, path, c0, out, chunk, flush, threads)) return;   // a large regular FASTQ file, -t > 1
  read_sequences(lr, path, out, chunk, flush);
}
