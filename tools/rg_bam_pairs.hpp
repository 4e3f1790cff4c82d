// rg_bam_pairs.hpp — `RUFUS.Filter HashList X.bam X.bam …`: the subject's BAM as the paired input of the read pull,
// decoded and re-paired in process. What the reference feeds RUFUS.Filter for a BAM sample is
// `samtools view -F 3328 X.bam | PassThroughSamCheck.stranded chr STUB` writing STUB.mate1.fastq / STUB.mate2.fastq
// (two FIFOs, runRufus.sh:963-967). Here the workers of tools/rg_bam_pipeline.hpp render every kept record as the
// three strings that tool would write for it — name, SEQ and QUAL as SAM text spells them, and for FLAG 0x10 the
// reverse complement with every character outside ACGTN dropped and the quality reversed
// (PassThroughSamCheck.stranded.cpp:197-279, restated in tools/passthrough_sam.cpp) — and the calling thread pairs
// them by name in file order: the first-seen record of a name is parked; when its mate arrives the SECOND-seen
// record is mate 1 and the parked one mate 2; leftovers are dropped. The pairs come out in the order, and with the
// bytes, of the two FASTQ files of the chain (tests/test_input_stream.py compares them on the host).
#pragma once
#include <string_view>
#include <unordered_map>

#include "rg_bam_input.hpp"
#include "rg_fastq_batch.hpp"

namespace rgbam {

struct Complement {
  char t[256];
  Complement() {
    memset(t, 0, sizeof(t));
    t[(unsigned char)'A'] = 'T'; t[(unsigned char)'C'] = 'G'; t[(unsigned char)'G'] = 'C';
    t[(unsigned char)'T'] = 'A'; t[(unsigned char)'N'] = 'N';
  }
};

// [u32 name_len][u32 seq_len][u32 qual_len] name seq qual
static void stranded_record(const unsigned char* r, uint32_t bs, const Header&, Text& o) {
  static const BasePairs pairs;
  static const Complement comp;
  if (bs < 32) fail("corrupt BAM record");
  const uint32_t l_name = r[8], n_cigar = le16(r + 12), flag = le16(r + 14), l_seq = le32(r + 16);
  const uint64_t fixed = 32ull + l_name + 4ull * n_cigar + ((uint64_t)l_seq + 1) / 2 + l_seq;
  if (fixed > bs || l_name == 0) fail("corrupt BAM record (lengths)");
  const uint32_t nl = l_name - 1;
  char* const head = o.room(12 + (size_t)nl + 2 * (size_t)l_seq + 8);
  char* w = head + 12;
  memcpy(w, r + 32, nl);
  w += nl;
  const unsigned char* ps = r + 32 + l_name + 4ull * n_cigar;
  const unsigned char* pq = ps + (l_seq + 1) / 2;
  const bool reverse = (flag & 16) != 0;
  char* const seq0 = w;
  if (l_seq == 0) {
    if (!reverse) *w++ = '*';                      // ("*" reverse-complemented: dropped)
  } else if (!reverse) {
    for (uint32_t i = 0; i < (l_seq + 1) / 2; ++i) memcpy(w + 2 * i, pairs.t[ps[i]], 2);   // (an odd length writes one too many: overwritten)
    w += l_seq;
  } else {
    for (uint32_t j = l_seq; j-- > 0;) {
      const unsigned code = (j & 1) ? (ps[j >> 1] & 15) : (ps[j >> 1] >> 4);
      const char c = comp.t[(unsigned char)"=ACMGRSVTWYHKDBN"[code]];
      *w = c;
      w += c != 0;
    }
  }
  const uint32_t sl = (uint32_t)(w - seq0);
  char* const qual0 = w;
  if (l_seq == 0 || pq[0] == 0xFF) {
    *w++ = '*';
  } else if (!reverse) {
    for (uint32_t i = 0; i < l_seq; ++i) w[i] = (char)(pq[i] + 33);
    w += l_seq;
  } else {
    for (uint32_t i = 0; i < l_seq; ++i) w[i] = (char)(pq[l_seq - 1 - i] + 33);
    w += l_seq;
  }
  const uint32_t ql = (uint32_t)(w - qual0);
  memcpy(head, &nl, 4);
  memcpy(head + 4, &sl, 4);
  memcpy(head + 8, &ql, 4);
  o.done(w);
}

// what read_record (tools/rg_fastq_batch.hpp) makes of the FASTQ record "@name\nseq\n+\nqual\n"
static inline void append_fastq_record(::Batch& b, const char* name, size_t nl, const char* seq, size_t sl, const char* qual, size_t ql) {
  b.name_len.push_back((uint32_t)(nl + 1));
  const size_t t0 = b.text.size();
  b.text.resize(t0 + nl + sl + ql + 6);
  char* w = b.text.data() + t0;
  *w++ = '@';
  memcpy(w, name, nl);
  w += nl;
  *w++ = '\n';
  memcpy(w, seq, sl);
  w += sl;
  *w++ = '\n';
  *w++ = '+';
  *w++ = '\n';
  memcpy(w, qual, ql);
  w += ql;
  *w++ = '\n';
  const size_t s0 = b.seq.size();
  b.seq.resize(s0 + sl);
  memcpy(b.seq.data() + s0, seq, sl);
  const size_t qn = ql < sl ? ql : sl;
  b.qual.resize(s0 + sl);                   // (seq and qual always have the same size: missing quality is padded)
  memcpy(b.qual.data() + s0, qual, qn);
  memset(b.qual.data() + s0 + qn, 0, sl - qn);
  b.offs.push_back(b.seq.size());
  b.text_offs.push_back(b.text.size());
}

// Every re-paired pair of the BAM, in the order of the chain's two FASTQ files: appended to (b1, b2); full(b1, b2)
// is called whenever they hold `max_reads` pairs (it empties them). `head` = the bytes of the file already read.
template <typename Full>
static inline void bam_pairs(FILE* f, const char* head, size_t head_len, unsigned threads, ::Batch& b1, ::Batch& b2, uint64_t max_reads, Full&& full) {
  struct Parked { char* blob; uint32_t name_len, seq_len, qual_len; };
  std::unordered_map<std::string_view, Parked> parked;
  Pipeline pl;
  pl.f = f;
  pl.prefix = (const unsigned char*)head;
  pl.prefix_len = head_len;
  pl.threads = threads ? threads : 1;
  pl.excl = flags_from_env("RG_BAM_EXCLUDE_FLAGS", 3328);
  pl.req = flags_from_env("RG_BAM_REQUIRE_FLAGS", 0);
  pl.render = stranded_record;
  setvbuf(f, nullptr, _IONBF, 0);
  pl.run([&](const char* p, size_t n) {
    const char* const end = p + n;
    while (p < end) {
      uint32_t nl, sl, ql;
      memcpy(&nl, p, 4);
      memcpy(&sl, p + 4, 4);
      memcpy(&ql, p + 8, 4);
      const char* name = p + 12;
      const char* seq = name + nl;
      const char* qual = seq + sl;
      p = qual + ql;
      auto it = parked.find(std::string_view(name, nl));
      if (it == parked.end()) {
        Parked pk;
        pk.name_len = nl; pk.seq_len = sl; pk.qual_len = ql;
        pk.blob = (char*)malloc((size_t)nl + sl + ql + 1);
        if (!pk.blob) fail("out of memory");
        memcpy(pk.blob, name, (size_t)nl + sl + ql);
        parked.emplace(std::string_view(pk.blob, nl), pk);
        continue;
      }
      const Parked pk = it->second;
      append_fastq_record(b1, name, nl, seq, sl, qual, ql);
      append_fastq_record(b2, name, nl, pk.blob + pk.name_len, pk.seq_len, pk.blob + pk.name_len + pk.seq_len, pk.qual_len);
      parked.erase(it);
      free(pk.blob);
      if (b1.n() >= max_reads) full(b1, b2);
    }
  });
  for (auto& kv : parked) free(kv.second.blob);
}

}  // namespace rgbam
