// tools/fastq_to_sam.cpp — drop-in for `perl scripts/FastqToSam.pl FILE`, the line runRufus.sh:159-161 writes into
// the `.generator` of a FASTQ input (`perl $RDIR/scripts/FastqToSam.pl <(cat X.fastq)`): FASTQ -> unaligned SAM
// text for PassThroughSamCheck[.stranded]. Host text processing, no GPU step; the Perl script manages ~215 MB/s of
// FASTQ (this: 1.5 GB/s) (profiles/r2_host_ingest.txt).
//
//   rg_fastq_to_sam FILE        (FILE may be a FIFO or /dev/fd/N)
//
// Per record of four lines (scripts/FastqToSam.pl:26-39), byte for byte what the script prints:
//   NAME <tab> 0 * 0 * * * 0 0 <tab> line 2 <tab> line 4 <tab> <newline>
// NAME = the first whitespace-separated word of line 1 (Perl's `split ' '`: leading whitespace skipped,
// separators are space, \t, \n, \r, \f, \v) without its first character; lines 2 and 4 lose only their '\n'
// (a '\r' stays, as with chomp); line 3 is ignored; a record cut short by the end of the file prints with the
// missing lines empty.
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <fcntl.h>

#include <vector>

static void write_all(const char* p, size_t n) {
  while (n) {
    const ssize_t w = write(STDOUT_FILENO, p, n);
    if (w < 0) {
      if (errno == EINTR) continue;
      _exit(1);
    }
    p += w;
    n -= (size_t)w;
  }
}

static inline bool perl_space(unsigned char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

int main(int argc, char** argv) {
  const int fd = argc > 1 ? open(argv[1], O_RDONLY) : -1;
  if (fd < 0) {
    fprintf(stderr, "ERROR could not open sam file at %s line 16.\n", argv[0]);
    return 2;
  }
  std::vector<char> in(8u << 20), out;
  out.reserve(10u << 20);
  size_t len = 0;          // bytes of `in` in use
  bool eof = false;
  // the four lines of the record being assembled: (offset, length without '\n') into `in`
  size_t pos = 0;
  for (;;) {
    // make sure `in[pos, len)` holds a whole record (four newlines) or everything up to the end of the file
    const char* l[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t n[4] = {0, 0, 0, 0};
    int have = 0;
    size_t p = pos;
    while (have < 4) {
      const char* nl = (const char*)memchr(in.data() + p, '\n', len - p);
      if (nl) {
        l[have] = in.data() + p;
        n[have] = (size_t)(nl - (in.data() + p));
        p += n[have] + 1;
        ++have;
        continue;
      }
      if (eof) {
        if (p < len) {         // last line without a newline
          l[have] = in.data() + p;
          n[have] = len - p;
          p = len;
          ++have;
        }
        break;
      }
      // refill: keep the unconsumed tail, read more, start the record over
      if (pos) {
        memmove(in.data(), in.data() + pos, len - pos);
        len -= pos;
        pos = 0;
      }
      if (len == in.size()) in.resize(in.size() * 2);
      ssize_t got;
      do got = read(fd, in.data() + len, in.size() - len); while (got < 0 && errno == EINTR);
      if (got <= 0) eof = true;
      else len += (size_t)got;
      have = 0;
      p = pos;
    }
    if (have == 0) break;
    // NAME
    const char* a = l[0];
    const char* e = l[0] + n[0];
    while (a < e && perl_space((unsigned char)*a)) ++a;
    const char* b = a;
    while (b < e && !perl_space((unsigned char)*b)) ++b;
    if (a < b) ++a;            // substr $t[0], 1
    const size_t need = (size_t)(b - a) + n[1] + n[3] + 32;
    if (out.size() + need > out.capacity()) {
      write_all(out.data(), out.size());
      out.clear();
      if (need > out.capacity()) out.reserve(need);
    }
    const size_t at = out.size();
    out.resize(at + need);
    char* w = out.data() + at;
    memcpy(w, a, (size_t)(b - a));
    w += b - a;
    static const char kMiddle[] = "\t0\t*\t0\t*\t*\t*\t0\t0\t";
    memcpy(w, kMiddle, sizeof(kMiddle) - 1);
    w += sizeof(kMiddle) - 1;
    if (n[1]) memcpy(w, l[1], n[1]);
    w += n[1];
    *w++ = '\t';
    if (n[3]) memcpy(w, l[3], n[3]);
    w += n[3];
    *w++ = '\t';
    *w++ = '\n';
    out.resize((size_t)(w - out.data()));
    pos = p;
  }
  write_all(out.data(), out.size());
  return 0;
}
