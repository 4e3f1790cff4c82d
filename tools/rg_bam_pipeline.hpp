// rg_bam_pipeline.hpp — BGZF/BAM decoding as a pipeline of BATCHES (~2 MB of BGZF members each) that never stops
// for a serial step; shared by tools/bamview.cpp (records -> SAM text -> stdout) and tools/jellyfish.cpp (`jellyfish
// count X.bam`: records -> sequences straight into the pinned upload buffers, no SAM text and no pipe in between):
//
//   reader thread      file -> raw BGZF members of batch i+k            (one big read per batch)
//   T worker threads   inflate the members of a batch (tools/rg_inflate.hpp, CRC-32 checked, zlib as the
//                      second opinion for a member the fast decoder rejects), or format the records of a
//                      batch with the caller's `Format` — whichever is ready, oldest first
//   sequencer thread   in file order: stitch the record that straddles two batches, find the record
//                      boundaries (one hop per record), apply the FLAG filters
//   the calling thread in file order: hands the formatted bytes of batch i-k to the caller's sink
//
// A ring of 2T+4 batch slots bounds the memory and recycles every buffer (no page faults after warm-up).
// Formats: BGZF = concatenated gzip members with a "BC" extra subfield holding the member size (SAM spec §4.1);
// BAM record layout §4.2. The includer defines rgbam::fail (message + exit; several threads are running).
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "rg_crc32.hpp"
#include "rg_inflate.hpp"

namespace rgbam {

[[noreturn]] void fail(const std::string& message);   // defined by the includer

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
    if (!q) fail("out of memory");
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

struct Text : Bytes {
  inline char* room(size_t extra) { reserve(n + extra); return (char*)p + n; }
  inline void done(char* w) { n = (size_t)(w - (char*)p); }
};


struct Header {
  std::string text;
  std::vector<std::string> refs;
  size_t max_ref = 1;
};


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

// what a worker does with one kept record: append its rendering to `out` (r = the record behind its block_size word)
typedef void (*Format)(const unsigned char* r, uint32_t bs, const Header& hd, Text& out);

struct Pipeline {
  FILE* f = nullptr;
  const unsigned char* prefix = nullptr;   // bytes of the file the caller has already read (format sniffing)
  size_t prefix_len = 0;
  unsigned threads = 1;
  uint32_t excl = 0, req = 0;              // skip records with any of `excl`, keep only records with all of `req`
  bool with_header = false;                // the header text goes out in front of the first records
  Format render = nullptr;
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
    stash.append(prefix, prefix_len);
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
          if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) fail("not a BGZF file (bad gzip member header)");
          const uint32_t xlen = le16(h + 10);
          if (b.raw.n - p < 12 + (size_t)xlen) break;
          int64_t bsize = -1;
          for (uint32_t q = 0; q + 4 <= xlen;) {
            const uint32_t slen = le16(h + 12 + q + 2);
            if (h[12 + q] == 'B' && h[12 + q + 1] == 'C' && slen == 2 && q + 6 <= xlen) bsize = le16(h + 12 + q + 4);
            q += 4 + slen;
          }
          if (bsize < 0) fail("BGZF member without a BC subfield");
          const int64_t clen = bsize + 1 - 12 - (int64_t)xlen - 8;
          if (clen < 0) fail("corrupt BGZF block size");
          if (b.raw.n - p < (size_t)bsize + 1) break;
          Member m;
          m.off = p + 12 + xlen;
          m.clen = (uint32_t)clen;
          m.crc = le32(h + bsize + 1 - 8);
          m.isize = le32(h + bsize + 1 - 4);
          if (m.isize > (1u << 16)) fail("corrupt BGZF block (inflated size above 64 KiB)");
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
        if (b.raw.n - p >= 4 && (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4))) fail("not a BGZF file (bad gzip member header)");
        fail("truncated BGZF block");
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
    if (!zlib_only && inf.run(comp, m.clen, out, m.isize) && rg::crc32_of(out, m.isize) == m.crc) return;
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, -15) != Z_OK) fail("zlib init failed");
    zs.next_in = const_cast<unsigned char*>(comp);
    zs.avail_in = m.clen;
    zs.next_out = out;
    zs.avail_out = m.isize;
    const int rc = inflate(&zs, Z_FINISH);
    inflateEnd(&zs);
    if (rc != Z_STREAM_END || zs.total_out != m.isize) fail("corrupt BGZF block (inflate)");
    if (rg::crc32_of(out, m.isize) != m.crc) fail("corrupt BGZF block (crc)");
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
        for (const auto& r : b->recs) render(b->base + r.first, r.second, hd, b->text);
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

  // ---- consumer (the calling thread): the formatted bytes of every batch, in file order
  template <typename Sink>
  void consume(Sink&& sink) {
    for (size_t seq = 0;; ++seq) {
      Batch& b = slot(seq);
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return seq >= n_batches || b.state == FORMATTED; });
        if (seq >= n_batches) return;
      }
      sink((const char*)b.text.p, b.text.n);
      std::lock_guard<std::mutex> lk(mu);
      b.state = FREE;
      cv.notify_all();
    }
  }

  // the BAM header at v[0, n): magic, l_text, text, n_ref, then (l_name, name, l_ref) per reference.
  // Returns its size, or 0 while it is incomplete.
  size_t parse_header(const unsigned char* v, size_t n) {
    if (n < 12) return 0;
    if (memcmp(v, "BAM\1", 4) != 0) fail("not a BAM file (bad magic)");
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

  // ---- sequencer: batches in file order
  void sequencer() {
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
    if (!header_done) fail(carry.n ? "truncated BAM header" : "empty input");
    if (carry.n) fail("truncated BAM record at the end of the file");
  }

  // Runs the whole file through the pipeline; sink(bytes, n) is called on THIS thread, in file order.
  template <typename Sink>
  void run(Sink&& sink) {
    ring = std::vector<Batch>(2 * (size_t)threads + 4);
    std::thread rd(&Pipeline::reader, this), sq(&Pipeline::sequencer, this);
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < threads; ++t) pool.emplace_back(&Pipeline::worker, this);
    consume(sink);
    rd.join();
    for (auto& th : pool) th.join();
    sq.join();
  }
};


}  // namespace rgbam
