// rg_cli.hpp — what the drop-in executables under tools/ share: argument helpers, the FASTA/FASTQ
// sequence reader (the parsing rules of mer_overlap_sequence_parser.hpp:124-207), k-mer text.
// Host side only: every computation goes through librufus_gpu.so (include/rufus_gpu.h).
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <errno.h>
#include <unistd.h>

#include <string>
#include <utility>
#include <vector>

#include "rufus_gpu.h"

[[noreturn]] static inline void die(const std::string& msg) {
  fprintf(stderr, "%s\n", msg.c_str());
  exit(1);
}
static inline void check(int64_t rc, const char* what) {
  if (rc < 0) die(std::string(what) + ": " + rg_last_error());
}

// yaggo's unsigned conversion: decimal SI suffixes (k M G T P E), e.g. -s 8G = 8 000 000 000
static inline uint64_t parse_size(const char* s) {
  char* end = nullptr;
  const double v = strtod(s, &end);
  if (end == s || v < 0) die(std::string("Invalid size '") + s + "'");
  double mult = 1;
  switch (*end) {
    case 0: break;
    case 'k': mult = 1e3; break;
    case 'M': mult = 1e6; break;
    case 'G': mult = 1e9; break;
    case 'T': mult = 1e12; break;
    case 'P': mult = 1e15; break;
    case 'E': mult = 1e18; break;
    default: die(std::string("Invalid size suffix in '") + s + "'");
  }
  return (uint64_t)(v * mult);
}

static inline std::string kmer_text(uint64_t key, int k) {
  std::string s((size_t)k, 'A');
  for (int i = 0; i < k; ++i) s[i] = "ACGT"[(key >> (2 * (k - 1 - i))) & 3];
  return s;
}

// growable byte buffer in PINNED host memory (rg_host_alloc), so that rg_count_push's copies run
// asynchronously and overlap both the kernels and the parsing of the next chunk
struct ByteBuf {
  char* p = nullptr;
  size_t n = 0, cap = 0;
  ByteBuf() = default;
  ByteBuf(const ByteBuf&) = delete;
  ByteBuf& operator=(const ByteBuf&) = delete;
  ~ByteBuf() { rg_host_free(p); }
  void reserve(size_t want) {
    if (want <= cap) return;
    size_t c = cap ? cap : (1u << 20);
    while (c < want) c += c / 2;
    char* q = (char*)rg_host_alloc(c);
    if (!q) die(std::string("out of pinned host memory: ") + rg_last_error());
    if (n) memcpy(q, p, n);
    rg_host_free(p);
    p = q;
    cap = c;
  }
  char* data() { return p; }
  const char* data() const { return p; }
  size_t size() const { return n; }
  bool empty() const { return n == 0; }
  char back() const { return p[n - 1]; }
  void clear() { n = 0; }
  void push_back(char c) { reserve(n + 1); p[n++] = c; }
  void append(const char* s, size_t len) { reserve(n + len); memcpy(p + n, s, len); n += len; }
  void swap(ByteBuf& o) { std::swap(p, o.p); std::swap(n, o.n); std::swap(cap, o.cap); }
  const char* begin() const { return p; }
  const char* end() const { return p + n; }
};

// the same interface over ordinary heap memory, for bytes that never travel to the device as they are (the FASTA of
// `jellyfish query -s`: a few MB whose k-mers are uploaded as packed keys; pinning and unpinning a growing buffer five
// times costs more than reading it)
struct HeapBuf {
  std::string s;
  void reserve(size_t want) { s.reserve(want); }
  const char* data() const { return s.data(); }
  size_t size() const { return s.size(); }
  bool empty() const { return s.empty(); }
  char back() const { return s.back(); }
  void clear() { s.clear(); }
  void push_back(char c) { s.push_back(c); }
  void append(const char* p, size_t len) { s.append(p, len); }
  const char* begin() const { return s.data(); }
  const char* end() const { return s.data() + s.size(); }
};

// buffered line reader over a FILE* (regular file, FIFO or stdin)
struct LineReader {
  FILE* f = nullptr;
  std::vector<char> buf;
  size_t pos = 0, len = 0;
  bool eof = false;
  explicit LineReader(FILE* fp, size_t cap = 1 << 22) : f(fp), buf(cap) {}
  bool fill() {
    if (eof) return false;
    if (pos < len) memmove(buf.data(), buf.data() + pos, len - pos);
    len -= pos;
    pos = 0;
    if (len == buf.size()) buf.resize(buf.size() * 2);
    // read(2), not fread: on a FIFO it returns what is there instead of waiting for the whole block,
    // which is what lets two inputs written in lock-step by one producer (PassThroughSamCheck.stranded
    // -> RUFUS.Filter, runRufus.sh:966-967) be read in lock-step without a deadlock
    ssize_t got;
    do got = read(fileno(f), buf.data() + len, buf.size() - len); while (got < 0 && errno == EINTR);
    if (got <= 0) { eof = true; return false; }
    len += (size_t)got;
    return true;
  }
  // next line without its '\n'; false at end of input. The pointer is valid until the next call.
  bool next(const char** line, size_t* n) {
    for (;;) {
      const char* nl = (const char*)memchr(buf.data() + pos, '\n', len - pos);
      if (nl) {
        *line = buf.data() + pos;
        *n = (size_t)(nl - (buf.data() + pos));
        pos += *n + 1;
        return true;
      }
      if (!fill()) {
        if (pos < len) {               // last line without a newline
          *line = buf.data() + pos;
          *n = len - pos;
          pos = len;
          return true;
        }
        return false;
      }
    }
  }
  int peek() {
    if (pos >= len && !fill()) return EOF;
    return (unsigned char)buf[pos];
  }
};

// Appends the sequence of every record of a FASTA/FASTQ stream to `out`, one 'N' after each
// record (what mer_overlap_sequence_parser.hpp:179-207 puts between reads: k-mers never span two
// records, and the lines of a multi-line record are joined). Format sniffed from the first byte
// (:124-153). flush(out) is called whenever `out` holds at least `chunk` bytes (read again after every flush: the
// caller may grow it).
template <typename Buf, typename Flush>
static inline void read_sequences(LineReader& lr, const char* path, Buf& out, const size_t& chunk, Flush&& flush) {
  const int c0 = lr.peek();
  if (c0 == EOF) return;
  if (c0 != '>' && c0 != '@') die(std::string("Unsupported format in '") + path + "'");
  const char* line;
  size_t n;
  if (c0 == '>') {
    while (lr.next(&line, &n)) {
      if (n && line[0] == '>') {
        if (!out.empty() && out.back() != 'N') out.push_back('N');
        if (out.size() >= chunk) flush(out);
      } else {
        out.append(line, n);
      }
    }
    if (!out.empty() && out.back() != 'N') out.push_back('N');
  } else {
    while (lr.next(&line, &n)) {       // header line '@...'
      size_t seq_len = 0;
      while (lr.next(&line, &n)) {     // sequence lines up to the '+' line
        if (n && line[0] == '+') break;
        out.append(line, n);
        seq_len += n;
      }
      size_t q = 0;
      do {                             // quality lines: as many characters as the sequence had
        if (!lr.next(&line, &n)) break;
        q += n;
      } while (q < seq_len);
      out.push_back('N');
      if (out.size() >= chunk) flush(out);
    }
  }
}
template <typename Buf, typename Flush>
static inline void read_sequences(FILE* f, const char* path, Buf& out, const size_t& chunk, Flush&& flush) {
  LineReader lr(f);
  read_sequences(lr, path, out, chunk, flush);
}
