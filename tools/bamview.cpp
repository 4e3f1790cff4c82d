// tools/bamview.cpp — the BAM side of a `.generator` script (runRufus.sh:599: `samtools view -F 3328 X.bam`):
// BGZF/BAM -> SAM text records on stdout, for PassThroughSamCheck[.stranded] to read. Host ingest
// (SURVEY §8f rank 2): the reference runs one single-threaded samtools per sample. Here the file flows through a
// pipeline of BATCHES (~2 MB of BGZF members each) that never stops for a serial step:
//
//   reader thread      file -> raw BGZF members of batch i+k            (one big read per batch)
//   T worker threads   inflate the members of a batch (tools/rg_inflate.hpp, CRC-32 checked, zlib as the
//                      second opinion for a member the fast decoder rejects), or format the records of a
//                      batch into SAM text — whichever is ready, oldest first
//   sequencer (main)   in file order: stitch the record that straddles two batches, find the record
//                      boundaries (one hop per record), apply -F / -f
//   writer thread      in file order: write(2) the text of batch i-k
//
// A ring of 2T+4 batch slots bounds the memory and recycles every buffer (no page faults after warm-up).
// Measured in profiles/r2_host_ingest.txt.
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
#include <zlib.h>

#include <algorithm>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "rg_inflate.hpp"

[[noreturn]] static void die(const std::string& m, int code = 1) {
  fprintf(stderr, "rg_bamview: %s\n", m.c_str());
  fflush(stderr);
  _exit(code);  // (several threads are running: no destructors, no atexit)
}

static inline uint16_t le16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static inline uint32_t le32(const unsigned char* p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static inline int32_t le32s(const unsigned char* p) { return (int32_t)le32(p); }

// uninitialised growable bytes (a std::vector would zero-fill every resize)
struct Bytes {
  unsigned char* p = nullptr;
  size_t n = 0, cap = 0;
  Bytes() = default;
  Bytes(const Bytes&) = delete;
  Bytes& operator=(const Bytes&) = delete;
  ~Bytes() { free(p); }
  void reserve(size_t want) {
    if (want <= cap) return;
    size_t c = cap ? cap : 4096;
    while (c < want) c += c / 2;
    unsigned char* q = (unsigned char*)realloc(p, c);
    if (!q) die("out of memory");
    p = q;
    cap = c;
  }
  void resize(size_t want) { reserve(want); n = want; }
  void append(const void* s, size_t len) {
    if (!len) return;
    reserve(n + len);
    memcpy(p + n, s, len);
    n += len;
  }
};

// ------------------------------------------------------------------------------------------- SAM text
struct Text : Bytes {
  inline char* room(size_t extra) { reserve(n + extra); return (char*)p + n; }
  inline void done(char* w) { n = (size_t)(w - (char*)p); }
};

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

struct Header {
  std::string text;
  std::vector<std::string> refs;
  size_t max_ref = 1;
};

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

// ------------------------------------------------------------------------------------------- the pipeline
struct Member {
  size_t off;       // of the deflate payload in Batch::raw
  uint32_t clen, isize, crc;
  size_t out;       // of the inflated bytes in Batch::data (behind the headroom)
};

enum State { FREE, READ, INFLATING, INFLATED, SCANNED, FORMATTING, FORMATTED };

struct Batch {
  State state = FREE;
  Bytes raw;                      // whole BGZF members as read from the file
  std::vector<Member> members;
  Bytes data;                     // [HEADROOM][inflated members]; the sequencer may start its view inside the headroom
  Bytes joined;                   // carry + data when the carry does not fit the headroom
  const unsigned char* base = nullptr;
  std::vector<std::pair<size_t, uint32_t>> recs;   // (offset of the record body from `base`, block_size)
  bool header_first = false;      // -h: the header text goes out in front of this batch's records
  Text text;
};

static const size_t HEADROOM = 1 << 16;

struct Pipeline {
  FILE* f;
  unsigned threads;
  uint32_t excl, req;
  bool with_header;
  std::vector<Batch> ring;
  std::mutex mu;
  std::condition_variable cv;
  size_t n_batches = SIZE_MAX;    // known once the reader has seen the end of the file
  size_t next_inflate = 0, next_format = 0;
  Header hd;

  Batch& slot(size_t seq) { return ring[seq % ring.size()]; }

  // ---- reader: batches of whole BGZF members
  void reader() {
    Bytes stash;                  // the bytes of a member cut by the end of the previous read
    bool eof = false;
    for (size_t seq = 0;; ++seq) {
      Batch& b = slot(seq);
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return b.state == FREE; });
      }
      const size_t target = 1 << 20;
      b.raw.n = 0;
      b.raw.append(stash.p, stash.n);
      stash.n = 0;
      b.members.clear();
      size_t p = 0, out = 0;
      for (;;) {
        // parse the complete members in raw[p, n)
        while (b.raw.n - p >= 18) {
          const unsigned char* h = b.raw.p + p;
          if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) die("not a BGZF file (bad gzip member header)");
          const uint32_t xlen = le16(h + 10);
          if (b.raw.n - p < 12 + (size_t)xlen) break;
          int64_t bsize = -1;
          for (uint32_t q = 0; q + 4 <= xlen;) {
            const uint32_t slen = le16(h + 12 + q + 2);
            if (h[12 + q] == 'B' && h[12 + q + 1] == 'C' && slen == 2 && q + 6 <= xlen) bsize = le16(h + 12 + q + 4);
            q += 4 + slen;
          }
          if (bsize < 0) die("BGZF member without a BC subfield");
          const int64_t clen = bsize + 1 - 12 - (int64_t)xlen - 8;
          if (clen < 0) die("corrupt BGZF block size");
          if (b.raw.n - p < (size_t)bsize + 1) break;
          Member m;
          m.off = p + 12 + xlen;
          m.clen = (uint32_t)clen;
          m.crc = le32(h + bsize + 1 - 8);
          m.isize = le32(h + bsize + 1 - 4);
          m.out = out;
          out += m.isize;
          b.members.push_back(m);
          p += (size_t)bsize + 1;
        }
        if (eof || (p >= target && !b.members.empty()) || b.members.size() >= 4096) break;
        b.raw.reserve(b.raw.n + target);
        const size_t got = fread(b.raw.p + b.raw.n, 1, target, f);
        if (got == 0) eof = true;
        b.raw.n += got;
      }
      if (eof && p != b.raw.n) {
        const unsigned char* h = b.raw.p + p;
        if (b.raw.n - p >= 4 && (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4))) die("not a BGZF file (bad gzip member header)");
        die("truncated BGZF block");
      }
      stash.append(b.raw.p + p, b.raw.n - p);
      b.raw.n = p;
      std::lock_guard<std::mutex> lk(mu);
      if (b.members.empty()) {        // nothing left: seq batches in total
        n_batches = seq;
        cv.notify_all();
        return;
      }
      b.state = READ;
      cv.notify_all();
    }
  }

  // ---- workers
  static void inflate_member(rg::Inflater& inf, const Batch& b, const Member& m, unsigned char* out) {
    if (m.isize == 0) return;
    const unsigned char* comp = b.raw.p + m.off;
    static const bool zlib_only = getenv("RG_BAMVIEW_ZLIB") != nullptr;   // measurement knob: zlib's inflate in the same pipeline
    if (!zlib_only && inf.run(comp, m.clen, out, m.isize) && (uint32_t)crc32(crc32(0L, Z_NULL, 0), out, m.isize) == m.crc) return;
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, -15) != Z_OK) die("zlib init failed");
    zs.next_in = const_cast<unsigned char*>(comp);
    zs.avail_in = m.clen;
    zs.next_out = out;
    zs.avail_out = m.isize;
    const int rc = inflate(&zs, Z_FINISH);
    inflateEnd(&zs);
    if (rc != Z_STREAM_END || zs.total_out != m.isize) die("corrupt BGZF block (inflate)");
    if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), out, m.isize) != m.crc) die("corrupt BGZF block (crc)");
  }

  void worker() {
    std::unique_ptr<rg::Inflater> inf(new rg::Inflater);
    for (;;) {
      Batch* b = nullptr;
      bool format = false;
      {
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
          if (next_format < n_batches && slot(next_format).state == SCANNED) {
            b = &slot(next_format++);
            b->state = FORMATTING;
            format = true;
            break;
          }
          if (next_inflate < n_batches && slot(next_inflate).state == READ) {
            b = &slot(next_inflate++);
            b->state = INFLATING;
            break;
          }
          if (next_format >= n_batches) return;     // every batch has been handed to a formatter
          cv.wait(lk);
        }
      }
      if (format) {
        b->text.n = 0;
        if (b->header_first) b->text.append(hd.text.data(), hd.text.size());
        for (const auto& r : b->recs) format_record(b->base + r.first, r.second, hd, b->text);
      } else {
        size_t total = 0;
        for (const Member& m : b->members) total += m.isize;
        b->data.resize(HEADROOM + total);
        for (const Member& m : b->members) inflate_member(*inf, *b, m, b->data.p + HEADROOM + m.out);
      }
      std::lock_guard<std::mutex> lk(mu);
      b->state = format ? FORMATTED : INFLATED;
      cv.notify_all();
    }
  }

  // ---- writer
  void writer() {
    for (size_t seq = 0;; ++seq) {
      Batch& b = slot(seq);
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return seq >= n_batches || b.state == FORMATTED; });
        if (seq >= n_batches) return;
      }
      const char* p = (const char*)b.text.p;
      size_t left = b.text.n;
      while (left) {
        const ssize_t w = write(STDOUT_FILENO, p, left);
        if (w < 0) {
          if (errno == EINTR) continue;
          die("write error (stdout)");
        }
        p += w;
        left -= (size_t)w;
      }
      std::lock_guard<std::mutex> lk(mu);
      b.state = FREE;
      cv.notify_all();
    }
  }

  // the BAM header at v[0, n): magic, l_text, text, n_ref, then (l_name, name, l_ref) per reference.
  // Returns its size, or 0 while it is incomplete.
  size_t parse_header(const unsigned char* v, size_t n) {
    if (n < 12) return 0;
    if (memcmp(v, "BAM\1", 4) != 0) die("not a BAM file (bad magic)");
    const uint32_t l_text = le32(v + 4);
    const size_t q = 8 + (size_t)l_text;
    if (n < q + 4) return 0;
    const uint32_t n_ref = le32(v + q);
    size_t w = q + 4;
    std::vector<std::string> refs;
    for (uint32_t r = 0; r < n_ref; ++r) {
      if (n < w + 4) return 0;
      const uint32_t l = le32(v + w);
      if (n < w + 4 + (size_t)l + 4) return 0;
      refs.emplace_back((const char*)v + w + 4, l ? l - 1 : 0);
      w += 4 + (size_t)l + 4;
    }
    hd.text.assign((const char*)v + 8, l_text);
    while (!hd.text.empty() && hd.text.back() == 0) hd.text.pop_back();
    hd.refs.swap(refs);
    for (const std::string& s : hd.refs) hd.max_ref = std::max(hd.max_ref, s.size());
    return w;
  }

  // ---- sequencer (the calling thread): batches in file order
  void run() {
    std::thread rd(&Pipeline::reader, this), wr(&Pipeline::writer, this);
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < threads; ++t) pool.emplace_back(&Pipeline::worker, this);
    Bytes carry;                  // the inflated bytes behind the last complete record of the previous batch
    bool header_done = false;
    for (size_t seq = 0;; ++seq) {
      Batch& b = slot(seq);
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return seq >= n_batches || b.state == INFLATED; });
        if (seq >= n_batches) break;
      }
      const unsigned char* v = b.data.p + HEADROOM;
      size_t n = b.data.n - HEADROOM;
      if (carry.n) {
        if (carry.n <= HEADROOM) {
          memcpy(b.data.p + HEADROOM - carry.n, carry.p, carry.n);
          v -= carry.n;
        } else {                  // a header or a record larger than the headroom
          b.joined.n = 0;
          b.joined.append(carry.p, carry.n);
          b.joined.append(v, n);
          v = b.joined.p;
        }
        n += carry.n;
        carry.n = 0;
      }
      b.base = v;
      b.recs.clear();
      b.header_first = false;
      size_t p = 0;
      if (!header_done) {
        p = parse_header(v, n);
        header_done = p != 0;
        b.header_first = header_done && with_header;
      }
      if (header_done) {
        while (p + 4 <= n) {
          const uint32_t bs = le32(v + p);
          if (p + 4 + (size_t)bs > n) break;
          const uint32_t flag = bs >= 16 ? le16(v + p + 4 + 14) : 0;
          if (!(flag & excl) && (flag & req) == req) b.recs.emplace_back(p + 4, bs);
          p += 4 + (size_t)bs;
        }
      }
      carry.append(v + p, n - p);
      std::lock_guard<std::mutex> lk(mu);
      b.state = SCANNED;
      cv.notify_all();
    }
    rd.join();
    for (auto& th : pool) th.join();
    wr.join();
    if (!header_done) die(carry.n ? "truncated BAM header" : "empty input");
    if (carry.n) die("truncated BAM record at the end of the file");
  }
};

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

  Pipeline pl;
  pl.f = f;
  pl.threads = threads;
  pl.excl = excl;
  pl.req = req;
  pl.with_header = with_header;
  pl.ring = std::vector<Batch>(2 * (size_t)threads + 4);
  pl.run();
  return 0;
}
