// tools/bamview.cpp — the BAM side of a `.generator` script (runRufus.sh:599: `samtools view -F 3328 X.bam`):
// BGZF/BAM -> SAM text records on stdout, for PassThroughSamCheck[.stranded] to read. Host ingest
// (SURVEY §8f rank 2): the reference runs one single-threaded samtools per sample. Here the file flows through the
// batch pipeline of tools/rg_bam_pipeline.hpp (reader thread -> T inflate/format workers -> in-order sequencer ->
// this thread writing the text), with the DEFLATE decoder of tools/rg_inflate.hpp. Measured in
// profiles/r2_host_ingest.txt.
//
//   rg_bamview [view] [-F EXCL] [-f REQ] [-@ THREADS] [-h] FILE.bam
//        -F EXCL   skip records with any of these FLAG bits (samtools view -F; RUFUS uses 3328 =
//                  secondary 256 + duplicate 1024 + supplementary 2048)
//        -f REQ    keep only records with all of these FLAG bits
//        -h        print the header text first (samtools view -h); default: records only
//   A region argument (needs the .bai index) is not supported: the exit code is 2.
//
// Formats: BGZF = concatenated gzip members with a "BC" extra subfield holding the member size
// (SAM spec §4.1); BAM record layout §4.2; text rules §1.4 (POS/PNEXT 1-based, `=` for a mate on the
// same reference, `*` for absent RNAME/CIGAR/SEQ/QUAL, integer tag types all printed as `i`).
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "rg_bam_pipeline.hpp"

[[noreturn]] static void die(const std::string& m, int code = 1) {
  fprintf(stderr, "rg_bamview: %s\n", m.c_str());
  fflush(stderr);
  _exit(code);  // (several threads are running: no destructors, no atexit)
}
[[noreturn]] void rgbam::fail(const std::string& m) { die(m); }

using rgbam::Header;
using rgbam::Text;
using rgbam::le16;
using rgbam::le32;
using rgbam::le32s;

// ------------------------------------------------------------------------------------------- SAM text
static inline char* put_int(char* w, int64_t v) {
  char tmp[24];
  int n = 0;
  uint64_t u = v < 0 ? (uint64_t)(-(v + 1)) + 1 : (uint64_t)v;
  do { tmp[n++] = (char)('0' + u % 10); u /= 10; } while (u);
  if (v < 0) *w++ = '-';
  while (n) *w++ = tmp[--n];
  return w;
}

// one integer/float element of type `t` at p (at most 15 characters); advances p
static inline char* put_scalar(char* w, char t, const unsigned char*& p, const unsigned char* end) {
  size_t sz;
  switch (t) {
    case 'c': case 'C': sz = 1; break;
    case 's': case 'S': sz = 2; break;
    case 'i': case 'I': case 'f': sz = 4; break;
    default: die(std::string("unknown tag value type '") + t + "'");
  }
  if ((size_t)(end - p) < sz) die("corrupt BAM record (tag overruns the record)");
  switch (t) {
    case 'c': w = put_int(w, (int8_t)p[0]); break;
    case 'C': w = put_int(w, p[0]); break;
    case 's': w = put_int(w, (int16_t)le16(p)); break;
    case 'S': w = put_int(w, le16(p)); break;
    case 'i': w = put_int(w, le32s(p)); break;
    case 'I': w = put_int(w, le32(p)); break;
    default: {
      float f;
      const uint32_t u = le32(p);
      memcpy(&f, &u, 4);
      w += snprintf(w, 16, "%g", f);
    }
  }
  p += sz;
  return w;
}

struct SeqPairs {
  char t[256][2];
  SeqPairs() {
    for (int b = 0; b < 256; ++b) {
      t[b][0] = "=ACMGRSVTWYHKDBN"[b >> 4];
      t[b][1] = "=ACMGRSVTWYHKDBN"[b & 15];
    }
  }
};
static const SeqPairs kSeqPairs;

// the SAM line of the record at r (after its block_size word), appended to o. Every byte of a record prints as at
// most 5 characters (a `B:c` array element: "-128,"), so 6 * bs + the two reference names + 128 bounds the line.
static void format_record(const unsigned char* r, uint32_t bs, const Header& hd, Text& o) {
  if (bs < 32) die("corrupt BAM record");
  const int32_t ref_id = le32s(r), pos = le32s(r + 4);
  const uint32_t l_name = r[8], mapq = r[9], n_cigar = le16(r + 12), flag = le16(r + 14);
  const uint32_t l_seq = le32(r + 16);
  const int32_t next_ref = le32s(r + 20), next_pos = le32s(r + 24), tlen = le32s(r + 28);
  const uint64_t fixed = 32ull + l_name + 4ull * n_cigar + ((uint64_t)l_seq + 1) / 2 + l_seq;
  if (fixed > bs || l_name == 0) die("corrupt BAM record (lengths)");
  char* w = o.room(6ull * bs + 2 * hd.max_ref + 128);
  const unsigned char* p = r + 32;
  memcpy(w, p, l_name - 1);
  w += l_name - 1;
  p += l_name;
  *w++ = '\t';
  w = put_int(w, flag);
  *w++ = '\t';
  auto ref_name = [&](int32_t id) {
    if (id < 0) { *w++ = '*'; return; }
    if ((size_t)id >= hd.refs.size()) die("reference id out of range");
    const std::string& s = hd.refs[(size_t)id];
    memcpy(w, s.data(), s.size());
    w += s.size();
  };
  ref_name(ref_id);
  *w++ = '\t';
  w = put_int(w, (int64_t)pos + 1);
  *w++ = '\t';
  w = put_int(w, mapq);
  *w++ = '\t';
  if (n_cigar == 0) {
    *w++ = '*';
  } else {
    for (uint32_t c = 0; c < n_cigar; ++c) {
      const uint32_t v = le32(p + 4 * c);
      w = put_int(w, v >> 4);
      *w++ = (v & 15) < 9 ? "MIDNSHP=X"[v & 15] : '?';
    }
  }
  p += 4ull * n_cigar;
  *w++ = '\t';
  if (next_ref < 0) *w++ = '*';
  else if (next_ref == ref_id) *w++ = '=';
  else ref_name(next_ref);
  *w++ = '\t';
  w = put_int(w, (int64_t)next_pos + 1);
  *w++ = '\t';
  w = put_int(w, tlen);
  *w++ = '\t';
  if (l_seq == 0) {
    *w++ = '*';
  } else {
    // two bases per packed byte through a 256 x 2 table (an odd length writes one character too many: overwritten)
    for (uint32_t i = 0; i < (l_seq + 1) / 2; ++i) memcpy(w + 2 * i, kSeqPairs.t[p[i]], 2);
    w += l_seq;
  }
  p += (l_seq + 1) / 2;
  *w++ = '\t';
  if (l_seq == 0 || p[0] == 0xFF) {
    *w++ = '*';
  } else {
    for (uint32_t i = 0; i < l_seq; ++i) w[i] = (char)(p[i] + 33);
    w += l_seq;
  }
  p += l_seq;
  const unsigned char* end = r + bs;
  while (p < end) {
    if (end - p < 3) die("corrupt BAM record (tag)");
    *w++ = '\t';
    *w++ = (char)p[0];
    *w++ = (char)p[1];
    *w++ = ':';
    const char t = (char)p[2];
    p += 3;
    switch (t) {
      case 'A':
        if (p >= end) die("corrupt BAM record (tag overruns the record)");
        *w++ = 'A';
        *w++ = ':';
        *w++ = (char)*p++;
        break;
      case 'Z':
      case 'H': {
        *w++ = t;
        *w++ = ':';
        const void* z = memchr(p, 0, (size_t)(end - p));
        if (!z) die("corrupt BAM record (unterminated string tag)");
        const size_t len = (size_t)((const unsigned char*)z - p);
        memcpy(w, p, len);
        w += len;
        p = (const unsigned char*)z + 1;
        break;
      }
      case 'B': {
        if (end - p < 5) die("corrupt BAM record (array tag)");
        const char sub = (char)p[0];
        const uint32_t cnt = le32(p + 1);
        p += 5;
        *w++ = 'B';
        *w++ = ':';
        *w++ = sub;
        for (uint32_t i = 0; i < cnt; ++i) {
          if (p >= end) die("corrupt BAM record (array tag length)");
          *w++ = ',';
          w = put_scalar(w, sub, p, end);
        }
        break;
      }
      case 'f':
        *w++ = 'f';
        *w++ = ':';
        w = put_scalar(w, t, p, end);
        break;
      default:
        *w++ = 'i';
        *w++ = ':';
        w = put_scalar(w, t, p, end);
    }
  }
  *w++ = '\n';
  o.done(w);
}

int main(int argc, char** argv) {
  uint32_t excl = 0, req = 0;
  unsigned threads = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  bool with_header = false;
  std::vector<const char*> pos_args;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto value = [&](const char* opt) -> const char* {
      if (a.size() > 2) return argv[i] + 2;
      if (i + 1 >= argc) die(std::string("option ") + opt + " needs a value");
      return argv[++i];
    };
    if (a == "view" && pos_args.empty()) continue;
    if (a.rfind("-F", 0) == 0) excl = (uint32_t)strtoul(value("-F"), nullptr, 0);
    else if (a.rfind("-f", 0) == 0) req = (uint32_t)strtoul(value("-f"), nullptr, 0);
    else if (a.rfind("-@", 0) == 0) threads = (unsigned)std::min(256L, std::max(1L, strtol(value("-@"), nullptr, 10)));
    else if (a == "-h") with_header = true;
    else if (a[0] == '-' && a.size() > 1) die("unsupported option " + a, 2);
    else pos_args.push_back(argv[i]);
  }
  if (pos_args.empty()) die("usage: rg_bamview [view] [-F INT] [-f INT] [-@ THREADS] [-h] FILE.bam", 2);
  if (pos_args.size() > 1) die("region queries need the .bai index and are not supported", 2);
  FILE* f = strcmp(pos_args[0], "-") == 0 ? stdin : fopen(pos_args[0], "rb");
  if (!f) die(std::string("cannot open ") + pos_args[0]);
  setvbuf(f, nullptr, _IONBF, 0);   // the reader asks for a megabyte at a time

  rgbam::Pipeline pl;
  pl.f = f;
  pl.threads = threads;
  pl.excl = excl;
  pl.req = req;
  pl.with_header = with_header;
  pl.render = format_record;
  pl.run([](const char* p, size_t left) {
    while (left) {
      const ssize_t w = write(STDOUT_FILENO, p, left);
      if (w < 0) {
        if (errno == EINTR) continue;
        die("write error (stdout)");
      }
      p += w;
      left -= (size_t)w;
    }
  });
  return 0;
}
