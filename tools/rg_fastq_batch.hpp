// rg_fastq_batch.hpp — 4-line FASTQ records into batches, and the lock-step reading of two mate
// files (src/RUFUS.Filter.cpp:196-231 takes one record from each file in turn). Shared by
// tools/rufus_filter.cpp and the CPU-only FIFO test harness tools/lockstep_check.cpp.
#pragma once
#include "rg_cli.hpp"

struct Batch {
  std::vector<char> seq, qual, text;       // text: the four lines of every record, '\n'-terminated
  std::vector<uint64_t> offs, text_offs;   // per read
  std::vector<uint32_t> name_len;          // length of line 1 (single-end: where ":MH" goes)
  void clear() {
    seq.clear(); qual.clear(); text.clear();
    offs.assign(1, 0); text_offs.assign(1, 0); name_len.clear();
  }
  uint64_t n() const { return offs.size() - 1; }
};

// one 4-line record; false at end of input (a truncated record counts as the end, as getline would
// leave empty strings the reference then scores as an empty read)
static bool read_record(LineReader& lr, Batch& b) {
  const char* l;
  size_t n;
  if (!lr.next(&l, &n)) return false;
  b.name_len.push_back((uint32_t)n);
  b.text.insert(b.text.end(), l, l + n);
  b.text.push_back('\n');
  size_t seq_len = 0;
  if (lr.next(&l, &n)) {
    b.seq.insert(b.seq.end(), l, l + n);
    seq_len = n;
    b.text.insert(b.text.end(), l, l + n);
  }
  b.text.push_back('\n');
  if (lr.next(&l, &n)) b.text.insert(b.text.end(), l, l + n);
  b.text.push_back('\n');
  size_t qn = 0;
  if (lr.next(&l, &n)) {
    qn = n < seq_len ? n : seq_len;
    b.qual.insert(b.qual.end(), l, l + qn);
    b.text.insert(b.text.end(), l, l + n);
  }
  b.qual.insert(b.qual.end(), seq_len - qn, '\0');   // missing quality = below any MinQ
  b.text.push_back('\n');
  b.offs.push_back(b.seq.size());
  b.text_offs.push_back(b.text.size());
  return true;
}

// up to max_reads pairs, one record from each file in turn; false when an input has ended
static bool read_pair_batch(LineReader& r1, LineReader& r2, Batch& b1, Batch& b2, uint64_t max_reads) {
  while (b1.n() < max_reads) {
    if (!read_record(r1, b1)) return false;
    if (!read_record(r2, b2)) {
      // mate 2 ended first: drop the unpaired mate-1 record
      b1.offs.pop_back(); b1.text_offs.pop_back(); b1.name_len.pop_back();
      b1.seq.resize(b1.offs.back()); b1.qual.resize(b1.offs.back()); b1.text.resize(b1.text_offs.back());
      return false;
    }
  }
  return true;
}
