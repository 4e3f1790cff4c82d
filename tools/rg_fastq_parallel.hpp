// rg_fastq_parallel.hpp — `jellyfish count -t T reads.fastq`: the FASTQ scan on T threads for a REGULAR file (the
// `-t` the reference hands to its own parser threads, mer_overlap_sequence_parser.hpp). One thread scans ~0.8 Gbases/s
// (tools/micro/parse_bench.cpp), a fraction of what the GPU counts; FIFOs stay on the serial reader (their producer is
// the limit), files are cut into segments that worker threads pread and parse on their own, and the calling thread
// appends the segments IN FILE ORDER to the pinned chunk and pushes it — the byte stream handed to the counter is the
// serial reader's (tests/test_input_stream.py), and every CUDA call stays on the calling thread.
//
// A segment starts at the first record start at or after its nominal offset. In a FASTQ file whose records are the
// usual four lines that is unambiguous: a line starting with '@' whose second-next line starts with '+' is a header
// (were it a quality line, its second-next line would be the next record's bases, which never start with '+').
// Wrapped (multi-line) records are legal FASTQ and the serial reader joins them; here every record is checked to be
// four lines with at least as many quality characters as bases, and anything else stops the run with a message that says
// to use -t 1 — never a silently different count.
#pragma once
#include <sys/stat.h>
#include <unistd.h>

#include <condition_variable>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "rg_cli.hpp"

namespace rgfq {

[[noreturn]] static inline void irregular(const char* path, uint64_t offset, const char* what) {
  fprintf(stderr, "%s: byte %llu: %s: not a plain four-line FASTQ file; run with -t 1 (the serial reader joins wrapped records)\n", path,
          (unsigned long long)offset, what);
  fflush(stderr);
  _exit(1);
}

static inline size_t env_size(const char* name, size_t fallback) {
  const char* e = getenv(name);
  return e && *e ? (size_t)strtoull(e, nullptr, 10) : fallback;
}

struct Segment {
  std::string seq;      // bases + 'N' per record
  bool done = false;
};

struct Scanner {
  int fd;
  const char* path;
  uint64_t size, seg_bytes, n_seg;

  bool read_at(uint64_t off, char* buf, size_t want, size_t& got) const {
    got = 0;
    while (got < want) {
      const ssize_t r = pread(fd, buf + got, want - got, (off_t)(off + got));
      if (r < 0) {
        if (errno == EINTR) continue;
        return false;
      }
      if (r == 0) break;
      got += (size_t)r;
    }
    return true;
  }

  // offset of the first record start at or after `o` (o > 0), or `size` when there is none
  uint64_t record_start(uint64_t o, std::vector<char>& buf) const {
    if (o >= size) return size;
    uint64_t base = o - 1;                    // (look at the byte before: o itself may be a line start)
    size_t want = 1 << 16;
    for (;;) {
      if (buf.size() < want) buf.resize(want);
      size_t got;
      if (!read_at(base, buf.data(), want, got)) die(std::string("read error on '") + path + "'");
      const bool at_eof = base + got >= size;
      // line starts inside the window: positions after a '\n'
      size_t p = 0;
      const char* nl = (const char*)memchr(buf.data(), '\n', got);
      if (!nl) {
        if (at_eof) return size;
        want *= 4;
        continue;
      }
      p = (size_t)(nl - buf.data()) + 1;
      bool need_more = false;
      while (p < got) {
        // lines L0, L1, L2 from p
        const char* e0 = (const char*)memchr(buf.data() + p, '\n', got - p);
        if (!e0) { need_more = !at_eof; break; }
        const size_t p1 = (size_t)(e0 - buf.data()) + 1;
        const char* e1 = p1 < got ? (const char*)memchr(buf.data() + p1, '\n', got - p1) : nullptr;
        if (!e1) { need_more = !at_eof; break; }
        const size_t p2 = (size_t)(e1 - buf.data()) + 1;
        if (p2 >= got) { need_more = !at_eof; break; }
        if (buf[p] == '@' && buf[p2] == '+') return base + p;
        p = p1;
      }
      if (!need_more) return size;            // the tail of the file holds no further record start
      want *= 4;
    }
  }

  // the records that START in [from, to): bases + 'N' each, checked to be plain four-line records
  void parse(uint64_t from, uint64_t to, std::vector<char>& buf, std::string& out) const {
    out.clear();
    if (from >= to) return;
    size_t want = (size_t)(to - from) + (1 << 16);
    size_t got = 0;
    bool at_eof = false;
    auto load = [&]() {
      if (buf.size() < want) buf.resize(want);
      if (!read_at(from, buf.data(), want, got)) die(std::string("read error on '") + path + "'");
      at_eof = from + got >= size;
    };
    load();
    out.reserve((size_t)(to - from) / 2 + 4096);
    size_t p = 0;
    while (from + p < to && p < got) {
      const char* l[4];
      size_t n[4];
      int have = 0;
      size_t q = p;
      while (have < 4) {
        const char* e = q < got ? (const char*)memchr(buf.data() + q, '\n', got - q) : nullptr;
        if (e) {
          l[have] = buf.data() + q;
          n[have] = (size_t)(e - (buf.data() + q));
          q += n[have] + 1;
          ++have;
          continue;
        }
        if (!at_eof) {                        // the record runs past what was read: read more and start it over
          want = want + (want >> 1) + (1 << 20);
          load();
          have = -1;
          break;
        }
        if (q < got) {                        // last line of the file without a newline
          l[have] = buf.data() + q;
          n[have] = got - q;
          q = got;
          ++have;
        }
        break;
      }
      if (have < 0) continue;
      if (have == 0) break;
      if (l[0][0] != '@' || n[0] == 0) irregular(path, from + p, "a record that does not start with '@'");
      if (have == 4) {
        if (n[2] == 0 || l[2][0] != '+') irregular(path, from + p, "no '+' line after one line of bases");
        // (more quality characters than bases: the serial reader takes the one line too; fewer: it would go on reading
        // quality lines, except at the very end of the file)
        if (n[3] < n[1] && !(at_eof && q >= got)) irregular(path, from + p, "fewer quality characters than bases");
        out.append(l[1], n[1]);
      } else {
        // the file ends inside this record: what the serial reader makes of it
        if (have >= 3 && (n[2] == 0 || l[2][0] != '+')) irregular(path, from + p, "no '+' line after one line of bases");
        if (have >= 2 && !(n[1] && l[1][0] == '+')) out.append(l[1], n[1]);
      }
      out.push_back('N');
      p = q;
    }
  }
};

// Appends a run of "bases + 'N'" records to the chunk in pieces of ~4 MB, flushing whenever the chunk is full, so that a
// chunk never grows far beyond `chunk` (the multi-GPU count stages a chunk in a device buffer of chunk + 16 MB). A piece
// ends behind an 'N': no k-mer spans one, so none is lost between two chunks.
template <typename Buf, typename Flush>
static inline void append_in_pieces(const char* p, size_t n, Buf& out, const size_t& chunk, Flush&& flush) {
  for (size_t at = 0; at < n;) {
    size_t take = std::min(n - at, (size_t)(4u << 20));
    if (at + take < n) {
      const char* cut = (const char*)memrchr(p + at, 'N', take);
      if (!cut) cut = (const char*)memchr(p + at + take, 'N', n - at - take);
      take = cut ? (size_t)(cut - (p + at)) + 1 : n - at;
    }
    out.append(p + at, take);
    at += take;
    if (out.size() >= chunk) flush(out);
  }
}

// true: the whole file has been read and flushed through (out, flush). false: not applicable (not a regular file,
// too small, one thread, not FASTQ) — nothing was consumed, use the serial reader.
template <typename Buf, typename Flush>
static inline bool read_fastq_parallel(FILE* f, const char* path, int first_byte, Buf& out, const size_t& chunk, Flush&& flush, unsigned threads) {
  if (threads < 2 || first_byte != '@') return false;
  struct stat st;
  if (fstat(fileno(f), &st) != 0 || !S_ISREG(st.st_mode)) return false;
  const uint64_t size = (uint64_t)st.st_size;
  if (size < env_size("RG_PARALLEL_FASTQ_MIN_BYTES", 64u << 20)) return false;
  Scanner sc;
  sc.fd = fileno(f);
  sc.path = path;
  sc.size = size;
  sc.seg_bytes = std::max<size_t>(1024, env_size("RG_PARALLEL_FASTQ_SEGMENT", 8u << 20));
  sc.n_seg = (size + sc.seg_bytes - 1) / sc.seg_bytes;
  const uint64_t window = (uint64_t)threads + 2;   // segments in flight
  std::vector<Segment> ring(window);
  std::mutex mu;
  std::condition_variable cv;
  uint64_t next = 0, consumed = 0;
  auto worker = [&]() {
    std::vector<char> buf, probe;
    std::string seq;
    for (;;) {
      uint64_t i;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return next >= sc.n_seg || next < consumed + window; });
        if (next >= sc.n_seg) return;
        i = next++;
      }
      const uint64_t from = i == 0 ? 0 : sc.record_start(i * sc.seg_bytes, probe);
      const uint64_t to = i + 1 >= sc.n_seg ? size : sc.record_start((i + 1) * sc.seg_bytes, probe);
      sc.parse(from, to, buf, seq);
      std::lock_guard<std::mutex> lk(mu);
      Segment& s = ring[i % window];
      s.seq.swap(seq);
      s.done = true;
      cv.notify_all();
    }
  };
  std::vector<std::thread> pool;
  const unsigned n_workers = std::min(threads, 6u);   // beyond ~4 the calling thread's copy into the chunk is the limit
  for (unsigned t = 0; t < n_workers; ++t) pool.emplace_back(worker);
  std::string seq;
  for (uint64_t i = 0; i < sc.n_seg; ++i) {
    Segment& s = ring[i % window];
    {
      std::unique_lock<std::mutex> lk(mu);
      cv.wait(lk, [&] { return s.done; });
      seq.swap(s.seq);
      s.done = false;
      ++consumed;
      cv.notify_all();
    }
    append_in_pieces(seq.data(), seq.size(), out, chunk, flush);
  }
  for (auto& th : pool) th.join();
  return true;
}

}  // namespace rgfq
