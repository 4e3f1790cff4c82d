// tools/passthrough_sam.cpp — drop-ins for bin/PassThroughSamCheck (src/PassThroughSamCheck.cpp:51-154)
// and, compiled with -DRG_STRANDED, bin/PassThroughSamCheck.stranded
// (src/PassThroughSamCheck.stranded.cpp:78-280): the SAM text a `.generator` script prints on stdin
// -> FASTQ. Pure host text processing (no GPU work on this step); what changes against the
// reference is the I/O: block-buffered reads and writes instead of a flush per line (the reference's
// `endl` after every line is what makes this stage the bottleneck of RunJellyForRUFUS.sh:28).
//
//   PassThroughSamCheck CHRLOG < sam > fastq
//        record = "@" + field 1, field 10, "+", field 11 (no reverse complement); CHRLOG gets the
//        reference name every time it changes (starting with the literal "notachr")
//   PassThroughSamCheck.stranded CHRLOG STUB < sam
//        pairs records by read name: the first-seen record is parked; when its mate arrives the
//        SECOND-seen record goes to STUB.mate1.fastq and the parked one to STUB.mate2.fastq. Flag 0x10:
//        sequence reverse-complemented (A/C/G/T/N only, other characters dropped) and quality
//        reversed. Unpaired leftovers are dropped.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <string_view>
#include <unordered_map>
#include <utility>
#include <vector>

#include "rg_cli.hpp"

struct Fields {
  const char* name; size_t name_len;
  const char* flag; size_t flag_len;
  const char* chr; size_t chr_len;
  const char* seq; size_t seq_len;
  const char* qual; size_t qual_len;
};

// fields 1, 2, 3, 10, 11 of a SAM line; a field ends at a tab (or at the end of the line)
static bool split_sam(const char* l, size_t n, Fields& f) {
  const char* p = l;
  const char* end = l + n;
  const char* start[11];
  size_t len[11];
  for (int i = 0; i < 11; ++i) {
    if (p > end) return false;
    const char* t = (const char*)memchr(p, '\t', (size_t)(end - p));
    start[i] = p;
    len[i] = t ? (size_t)(t - p) : (size_t)(end - p);
    p = t ? t + 1 : end + 1;
  }
  f.name = start[0]; f.name_len = len[0];
  f.flag = start[1]; f.flag_len = len[1];
  f.chr = start[2]; f.chr_len = len[2];
  f.seq = start[9]; f.seq_len = len[9];
  f.qual = start[10]; f.qual_len = len[10];
  return true;
}

// reverse complement of A/C/G/T/N; every other character is dropped, like the reference's switch
// without a default (PassThroughSamCheck.stranded.cpp:209-215)
struct CompTable {
  char t[256];
  CompTable() {
    memset(t, 0, sizeof(t));
    t[(unsigned char)'A'] = 'T'; t[(unsigned char)'C'] = 'G'; t[(unsigned char)'G'] = 'C';
    t[(unsigned char)'T'] = 'A'; t[(unsigned char)'N'] = 'N';
  }
};
static const CompTable kComp;
static void revcomp_into(std::string& out, const char* s, size_t n) {
  const size_t at = out.size();
  out.resize(at + n);
  char* w = &out[at];
  for (size_t j = n; j-- > 0;) {
    const char c = kComp.t[(unsigned char)s[j]];
    *w = c;
    w += c != 0;
  }
  out.resize((size_t)(w - out.data()));
}

struct Out {
  FILE* f;
  std::vector<char> buf;
  explicit Out(FILE* fp) : f(fp) { buf.reserve(1 << 22); }
  void put(const char* s, size_t n) {
    buf.insert(buf.end(), s, s + n);
    if (buf.size() >= (1u << 22)) flush();
  }
  void put(char c) { buf.push_back(c); }
  void flush() {
    if (!buf.empty()) fwrite(buf.data(), 1, buf.size(), f);
    fflush(f);
    buf.clear();
  }
  void record(const char* name, size_t nl, const char* seq, size_t sl, const char* qual, size_t ql) {
    put('@'); put(name, nl); put('\n');
    put(seq, sl); put("\n+\n", 3);
    put(qual, ql); put('\n');
  }
};

int main(int argc, char** argv) {
#ifdef RG_STRANDED
  if (argc < 3) { printf("usage: PassThroughSamCheck.stranded CHRLOG STUB < sam\n"); return 0; }
#else
  if (argc < 2) { printf("usage: PassThroughSamCheck CHRLOG < sam > fastq\n"); return 0; }
#endif
  FILE* chr = fopen(argv[1], "w");
  if (!chr) { printf("ERROR, Output file could not be opened -%s\n", argv[1]); return 0; }
#ifdef RG_STRANDED
  FILE* f1 = fopen((std::string(argv[2]) + ".mate1.fastq").c_str(), "w");
  FILE* f2 = fopen((std::string(argv[2]) + ".mate2.fastq").c_str(), "w");
  if (!f1 || !f2) { printf("ERROR, Output file could not be opened -%s\n", argv[1]); return 0; }
  Out m1(f1), m2(f2);
  // a parked record = one allocation holding name, sequence and quality (as they will be written);
  // the map key is a view of the name inside it
  struct Parked { char* blob; uint32_t name_len, seq_len, qual_len; };
  std::unordered_map<std::string_view, Parked> parked;
  std::string seq, qual;
#else
  Out out(stdout);
#endif
  LineReader lr(stdin);
  std::string current = "notachr";
  const char* l;
  size_t n;
  while (lr.next(&l, &n)) {
    Fields f;
    if (!split_sam(l, n, f)) continue;
    if (f.chr_len != current.size() || memcmp(f.chr, current.data(), f.chr_len) != 0) {
      fprintf(chr, "%s\n", current.c_str());
      current.assign(f.chr, f.chr_len);
    }
#ifdef RG_STRANDED
    char flag_text[16];
    int flag;
    if (f.flag_len < sizeof(flag_text)) {
      memcpy(flag_text, f.flag, f.flag_len);
      flag_text[f.flag_len] = 0;
      flag = atoi(flag_text);
    } else {
      flag = atoi(std::string(f.flag, f.flag_len).c_str());
    }
    const char* sq = f.seq;
    const char* ql = f.qual;
    size_t sq_len = f.seq_len, ql_len = f.qual_len;
    if (flag & 16) {
      seq.clear();
      revcomp_into(seq, f.seq, f.seq_len);
      qual.assign(f.qual, f.qual_len);
      for (size_t a = 0, b = qual.size(); a + 1 < b; ++a, --b) std::swap(qual[a], qual[b - 1]);
      sq = seq.data(); sq_len = seq.size();
      ql = qual.data(); ql_len = qual.size();
    }
    auto it = parked.find(std::string_view(f.name, f.name_len));
    if (it != parked.end()) {
      const Parked pk = it->second;
      m1.record(f.name, f.name_len, sq, sq_len, ql, ql_len);
      m2.record(f.name, f.name_len, pk.blob + pk.name_len, pk.seq_len, pk.blob + pk.name_len + pk.seq_len, pk.qual_len);
      parked.erase(it);
      free(pk.blob);
      // the two outputs are usually FIFOs read in lock-step by RUFUS.Filter (runRufus.sh:963-967):
      // both are flushed together and in pieces well below the 64 KiB of a pipe, so that the writer
      // can never sit on a full mate-1 pipe while the reader waits for mate 2
      if (m1.buf.size() >= (1u << 14) || m2.buf.size() >= (1u << 14)) {
        m1.flush();
        m2.flush();
      }
    } else {
      Parked pk;
      pk.name_len = (uint32_t)f.name_len; pk.seq_len = (uint32_t)sq_len; pk.qual_len = (uint32_t)ql_len;
      pk.blob = (char*)malloc(f.name_len + sq_len + ql_len + 1);
      if (!pk.blob) die("out of memory");
      memcpy(pk.blob, f.name, f.name_len);
      memcpy(pk.blob + f.name_len, sq, sq_len);
      memcpy(pk.blob + f.name_len + sq_len, ql, ql_len);
      parked.emplace(std::string_view(pk.blob, f.name_len), pk);
    }
#else
    out.record(f.name, f.name_len, f.seq, f.seq_len, f.qual, f.qual_len);
#endif
  }
  fprintf(chr, "%s\n", current.c_str());
  fclose(chr);
#ifdef RG_STRANDED
  m1.flush(); m2.flush();
  fclose(f1); fclose(f2);
#else
  out.flush();
#endif
  return 0;
}
