// tools/bamview.cpp — the BAM side of a `.generator` script (runRufus.sh:599: `samtools view -F 3328 X.bam`):
// BGZF/BAM -> SAM text records on stdout, for PassThroughSamCheck[.stranded] to read. Host ingest
// (SURVEY §8f rank 2): the reference runs one single-threaded samtools per sample; here the BGZF
// blocks of a batch are inflated by T threads, the records of the batch are formatted by T threads,
// and the text is written in file order.
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
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <string>
#include <thread>
#include <vector>

[[noreturn]] static void die(const std::string& m, int code = 1) {
  fprintf(stderr, "rg_bamview: %s\n", m.c_str());
  exit(code);
}

static inline uint16_t le16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static inline uint32_t le32(const unsigned char* p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static inline int32_t le32s(const unsigned char* p) { return (int32_t)le32(p); }

struct Block {
  std::vector<unsigned char> comp;   // the deflate payload
  uint32_t isize = 0;                // inflated size (gzip trailer)
  uint32_t crc = 0;
};

// one BGZF member from the file; false at a clean end of file
static bool read_block(FILE* f, Block& b) {
  unsigned char h[12];
  const size_t got = fread(h, 1, 12, f);
  if (got == 0) return false;
  if (got != 12 || h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) die("not a BGZF file (bad gzip member header)");
  const uint32_t xlen = le16(h + 10);
  std::vector<unsigned char> extra(xlen);
  if (fread(extra.data(), 1, xlen, f) != xlen) die("truncated BGZF header");
  int64_t bsize = -1;
  for (uint32_t q = 0; q + 4 <= xlen;) {
    const uint32_t slen = le16(extra.data() + q + 2);
    if (extra[q] == 'B' && extra[q + 1] == 'C' && slen == 2 && q + 6 <= xlen) bsize = le16(extra.data() + q + 4);
    q += 4 + slen;
  }
  if (bsize < 0) die("BGZF member without a BC subfield");
  const int64_t clen = bsize + 1 - 12 - (int64_t)xlen - 8;
  if (clen < 0) die("corrupt BGZF block size");
  b.comp.resize((size_t)clen);
  unsigned char tail[8];
  if (fread(b.comp.data(), 1, (size_t)clen, f) != (size_t)clen || fread(tail, 1, 8, f) != 8) die("truncated BGZF block");
  b.crc = le32(tail);
  b.isize = le32(tail + 4);
  return true;
}

static void inflate_block(const Block& b, unsigned char* out) {
  if (b.isize == 0) return;
  z_stream zs;
  memset(&zs, 0, sizeof(zs));
  if (inflateInit2(&zs, -15) != Z_OK) die("zlib init failed");
  zs.next_in = const_cast<unsigned char*>(b.comp.data());
  zs.avail_in = (uInt)b.comp.size();
  zs.next_out = out;
  zs.avail_out = b.isize;
  const int rc = inflate(&zs, Z_FINISH);
  inflateEnd(&zs);
  if (rc != Z_STREAM_END || zs.total_out != b.isize) die("corrupt BGZF block (inflate)");
  if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), out, b.isize) != b.crc) die("corrupt BGZF block (crc)");
}

template <typename F>
static void parallel_for(size_t n, unsigned threads, F&& fn) {
  if (threads <= 1 || n <= 1) {
    for (size_t i = 0; i < n; ++i) fn(i);
    return;
  }
  std::vector<std::thread> pool;
  const unsigned t_n = (unsigned)std::min<size_t>(threads, n);
  for (unsigned t = 0; t < t_n; ++t)
    pool.emplace_back([&, t]() {
      for (size_t i = t; i < n; i += t_n) fn(i);
    });
  for (auto& th : pool) th.join();
}

static inline void put_int(std::string& o, int64_t v) {
  char tmp[24];
  int n = 0;
  uint64_t u = v < 0 ? (uint64_t)(-(v + 1)) + 1 : (uint64_t)v;
  do { tmp[n++] = (char)('0' + u % 10); u /= 10; } while (u);
  if (v < 0) o.push_back('-');
  while (n) o.push_back(tmp[--n]);
}

static void put_float(std::string& o, float f) {
  char tmp[40];
  snprintf(tmp, sizeof(tmp), "%g", f);
  o += tmp;
}

// one integer/float element of type `t` at p; returns its size
static size_t put_scalar(std::string& o, char t, const unsigned char* p) {
  switch (t) {
    case 'c': put_int(o, (int8_t)p[0]); return 1;
    case 'C': put_int(o, p[0]); return 1;
    case 's': put_int(o, (int16_t)le16(p)); return 2;
    case 'S': put_int(o, le16(p)); return 2;
    case 'i': put_int(o, le32s(p)); return 4;
    case 'I': put_int(o, le32(p)); return 4;
    case 'f': { float f; uint32_t u = le32(p); memcpy(&f, &u, 4); put_float(o, f); return 4; }
    default: die(std::string("unknown tag value type '") + t + "'");
  }
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
};

// the SAM line of the record at r (after its block_size word), appended to o
static void format_record(const unsigned char* r, uint32_t bs, const Header& hd, std::string& o) {
  if (bs < 32) die("corrupt BAM record");
  const int32_t ref_id = le32s(r), pos = le32s(r + 4);
  const uint32_t l_name = r[8], mapq = r[9], n_cigar = le16(r + 12), flag = le16(r + 14);
  const uint32_t l_seq = le32(r + 16);
  const int32_t next_ref = le32s(r + 20), next_pos = le32s(r + 24), tlen = le32s(r + 28);
  const uint64_t fixed = 32ull + l_name + 4ull * n_cigar + (l_seq + 1) / 2 + l_seq;
  if (fixed > bs || l_name == 0) die("corrupt BAM record (lengths)");
  const unsigned char* p = r + 32;
  o.append((const char*)p, l_name - 1);
  p += l_name;
  o.push_back('\t');
  put_int(o, flag);
  o.push_back('\t');
  auto ref_name = [&](int32_t id) -> const std::string& {
    static const std::string star = "*";
    if (id < 0) return star;
    if ((size_t)id >= hd.refs.size()) die("reference id out of range");
    return hd.refs[(size_t)id];
  };
  o += ref_name(ref_id);
  o.push_back('\t');
  put_int(o, (int64_t)pos + 1);
  o.push_back('\t');
  put_int(o, mapq);
  o.push_back('\t');
  if (n_cigar == 0) {
    o.push_back('*');
  } else {
    for (uint32_t c = 0; c < n_cigar; ++c) {
      const uint32_t v = le32(p + 4 * c);
      put_int(o, v >> 4);
      o.push_back((v & 15) < 9 ? "MIDNSHP=X"[v & 15] : '?');
    }
  }
  p += 4ull * n_cigar;
  o.push_back('\t');
  if (next_ref < 0) o.push_back('*');
  else if (next_ref == ref_id) o.push_back('=');
  else o += ref_name(next_ref);
  o.push_back('\t');
  put_int(o, (int64_t)next_pos + 1);
  o.push_back('\t');
  put_int(o, tlen);
  o.push_back('\t');
  if (l_seq == 0) {
    o.push_back('*');
  } else {
    // two bases per packed byte through a 256 x 2 table
    const size_t at = o.size();
    o.resize(at + l_seq + 1);
    char* w = &o[at];
    for (uint32_t i = 0; i < (l_seq + 1) / 2; ++i) memcpy(w + 2 * i, kSeqPairs.t[p[i]], 2);
    o.resize(at + l_seq);
  }
  p += (l_seq + 1) / 2;
  o.push_back('\t');
  if (l_seq == 0 || p[0] == 0xFF) {
    o.push_back('*');
  } else {
    const size_t at = o.size();
    o.resize(at + l_seq);
    char* w = &o[at];
    for (uint32_t i = 0; i < l_seq; ++i) w[i] = (char)(p[i] + 33);
  }
  p += l_seq;
  const unsigned char* end = r + bs;
  while (p < end) {
    if (end - p < 3) die("corrupt BAM record (tag)");
    o.push_back('\t');
    o.push_back((char)p[0]);
    o.push_back((char)p[1]);
    o.push_back(':');
    const char t = (char)p[2];
    p += 3;
    switch (t) {
      case 'A':
        o += "A:";
        o.push_back((char)*p++);
        break;
      case 'Z':
      case 'H': {
        o.push_back(t);
        o.push_back(':');
        const void* z = memchr(p, 0, (size_t)(end - p));
        if (!z) die("corrupt BAM record (unterminated string tag)");
        o.append((const char*)p, (size_t)((const unsigned char*)z - p));
        p = (const unsigned char*)z + 1;
        break;
      }
      case 'B': {
        if (end - p < 5) die("corrupt BAM record (array tag)");
        const char sub = (char)p[0];
        const uint32_t cnt = le32(p + 1);
        p += 5;
        o += "B:";
        o.push_back(sub);
        for (uint32_t i = 0; i < cnt; ++i) {
          if (p >= end) die("corrupt BAM record (array tag length)");
          o.push_back(',');
          p += put_scalar(o, sub, p);
        }
        break;
      }
      case 'f':
        o += "f:";
        p += put_scalar(o, t, p);
        break;
      default:
        o += "i:";
        p += put_scalar(o, t, p);
    }
    if (p > end) die("corrupt BAM record (tag overruns the record)");
  }
  o.push_back('\n');
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
    else if (a.rfind("-@", 0) == 0) threads = (unsigned)std::max(1L, strtol(value("-@"), nullptr, 10));
    else if (a == "-h") with_header = true;
    else if (a[0] == '-' && a.size() > 1) die("unsupported option " + a, 2);
    else pos_args.push_back(argv[i]);
  }
  if (pos_args.empty()) die("usage: rg_bamview [view] [-F INT] [-f INT] [-@ THREADS] [-h] FILE.bam", 2);
  if (pos_args.size() > 1) die("region queries need the .bai index and are not supported", 2);
  FILE* f = strcmp(pos_args[0], "-") == 0 ? stdin : fopen(pos_args[0], "rb");
  if (!f) die(std::string("cannot open ") + pos_args[0]);
  static char outbuf[1 << 20];
  setvbuf(stdout, outbuf, _IOFBF, sizeof(outbuf));

  // `pending` = inflated bytes not yet consumed (a record or the header may straddle batches)
  std::vector<unsigned char> pending;
  Header hd;
  bool header_done = false;
  const size_t batch_blocks = 64 * (size_t)threads;
  std::vector<Block> blocks(batch_blocks);                 // reused: the payload buffers keep their capacity
  std::vector<std::string> text(threads);
  std::vector<std::pair<size_t, uint32_t>> recs;
  bool eof = false;
  while (!eof) {
    size_t n_blocks = 0;
    while (n_blocks < batch_blocks) {
      if (!read_block(f, blocks[n_blocks])) { eof = true; break; }
      ++n_blocks;
    }
    // inflate the batch behind what is pending
    std::vector<size_t> at(n_blocks + 1, pending.size());
    for (size_t i = 0; i < n_blocks; ++i) at[i + 1] = at[i] + blocks[i].isize;
    pending.resize(at[n_blocks]);
    parallel_for(n_blocks, threads, [&](size_t i) { inflate_block(blocks[i], pending.data() + at[i]); });

    size_t p = 0;
    const size_t n = pending.size();
    if (!header_done) {
      // magic, l_text, text, n_ref, then (l_name, name, l_ref) per reference
      bool ok = n >= 12;
      size_t q = 0;
      if (ok) {
        if (memcmp(pending.data(), "BAM\1", 4) != 0) die("not a BAM file (bad magic)");
        const uint32_t l_text = le32(pending.data() + 4);
        q = 8 + (size_t)l_text;
        ok = n >= q + 4;
        if (ok) {
          const uint32_t n_ref = le32(pending.data() + q);
          size_t w = q + 4;
          std::vector<std::string> refs;
          for (uint32_t r = 0; ok && r < n_ref; ++r) {
            ok = n >= w + 4;
            if (!ok) break;
            const uint32_t l = le32(pending.data() + w);
            ok = n >= w + 4 + l + 4;
            if (!ok) break;
            refs.emplace_back((const char*)pending.data() + w + 4, l ? l - 1 : 0);
            w += 4 + (size_t)l + 4;
          }
          if (ok) {
            hd.text.assign((const char*)pending.data() + 8, l_text);
            while (!hd.text.empty() && hd.text.back() == 0) hd.text.pop_back();
            hd.refs.swap(refs);
            header_done = true;
            p = w;
            if (with_header) fwrite(hd.text.data(), 1, hd.text.size(), stdout);
          }
        }
      }
      if (!header_done) {
        if (eof) die("truncated BAM header");
        continue;
      }
    }
    // record boundaries of the complete records in `pending`
    recs.clear();
    while (p + 4 <= n) {
      const uint32_t bs = le32(pending.data() + p);
      if (p + 4 + (size_t)bs > n) break;
      const uint32_t flag = bs >= 16 ? le16(pending.data() + p + 4 + 14) : 0;
      if (!(flag & excl) && (flag & req) == req) recs.emplace_back(p + 4, bs);
      p += 4 + (size_t)bs;
    }
    if (eof && p != n) die("truncated BAM record at the end of the file");
    const size_t parts = std::min<size_t>(threads, (recs.size() + 255) / 256);
    // (the text buffers live across batches: fresh allocations would fault their pages in again
    // every batch, which serialises the threads in the kernel)
    parallel_for(parts, threads, [&](size_t t) {
      const size_t lo = recs.size() * t / parts, hi = recs.size() * (t + 1) / parts;
      std::string& o = text[t];
      o.clear();
      for (size_t i = lo; i < hi; ++i) format_record(pending.data() + recs[i].first, recs[i].second, hd, o);
    });
    for (size_t t = 0; t < parts; ++t)
      if (fwrite(text[t].data(), 1, text[t].size(), stdout) != text[t].size()) die("write error (stdout)");
    pending.erase(pending.begin(), pending.begin() + (ptrdiff_t)p);
  }
  if (!header_done) die("empty input");
  if (fflush(stdout) != 0) die("write error (stdout)");
  return 0;
}
