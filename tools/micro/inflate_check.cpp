// tools/micro/inflate_check.cpp — test and timing driver for tools/rg_inflate.hpp (not product code).
//   inflate_check CASES            CASES = repeated [u32 comp_len][u32 out_len][comp bytes]; writes, per case, one
//                                  status byte (1 = inflated, 0 = rejected) followed by the inflated bytes when 1
//   inflate_check --time FILE.bam  inflates every BGZF member of the file with rg::Inflater and with zlib
//   inflate_check --crc FILE       rg::crc32_of against zlib's crc32 on pseudo-random slices of the file
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <chrono>
#include <vector>

#include <algorithm>

#include "../rg_crc32.hpp"
#include "../rg_inflate.hpp"

static std::vector<unsigned char> slurp(const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) { perror(path); exit(1); }
  std::vector<unsigned char> d;
  unsigned char buf[1 << 16];
  size_t n;
  while ((n = fread(buf, 1, sizeof buf, f)) > 0) d.insert(d.end(), buf, buf + n);
  fclose(f);
  return d;
}

int main(int argc, char** argv) {
  static rg::Inflater inf;
  if (argc == 3 && !strcmp(argv[1], "--time")) {
    const std::vector<unsigned char> d = slurp(argv[2]);
    std::vector<unsigned char> out(1 << 16), ref(1 << 16);
    double t_mine = 0, t_zlib = 0;
    size_t total = 0, p = 0, bad = 0;
    while (p + 18 <= d.size()) {
      const unsigned xlen = d[p + 10] | (d[p + 11] << 8), bsize = (d[p + 16] | (d[p + 17] << 8)) + 1u;
      const unsigned char* c = &d[p + 12 + xlen];
      const unsigned clen = bsize - 12 - xlen - 8;
      unsigned isize;
      memcpy(&isize, &d[p + bsize - 4], 4);
      auto t0 = std::chrono::steady_clock::now();
      const bool ok = inf.run(c, clen, out.data(), isize);
      auto t1 = std::chrono::steady_clock::now();
      z_stream zs;
      memset(&zs, 0, sizeof zs);
      inflateInit2(&zs, -15);
      zs.next_in = const_cast<unsigned char*>(c);
      zs.avail_in = clen;
      zs.next_out = ref.data();
      zs.avail_out = isize;
      inflate(&zs, Z_FINISH);
      inflateEnd(&zs);
      auto t2 = std::chrono::steady_clock::now();
      t_mine += std::chrono::duration<double>(t1 - t0).count();
      t_zlib += std::chrono::duration<double>(t2 - t1).count();
      if (!ok || memcmp(out.data(), ref.data(), isize) != 0) ++bad;
      total += isize;
      p += bsize;
    }
    printf("%s: %.1f MB inflated, rg::Inflater %.3f s (%.0f MB/s), zlib %s %.3f s (%.0f MB/s), mismatching members: %zu\n", argv[2],
           total / 1e6, t_mine, total / 1e6 / t_mine, zlibVersion(), t_zlib, total / 1e6 / t_zlib, bad);
    return bad ? 1 : 0;
  }
  if (argc == 3 && !strcmp(argv[1], "--crc")) {
    // rg::crc32_of (tools/rg_crc32.hpp: carry-less-multiply folding where the CPU has it) against zlib's crc32 on
    // pseudo-random slices of the file: every length class (below one fold, not a multiple of 16, many folds), every
    // alignment
    const std::vector<unsigned char> d = slurp(argv[2]);
    unsigned s = 12345;
    size_t bad = 0, cases = 0;
    for (int it = 0; it < 60000 && d.size() > 1; ++it) {
      s = s * 1664525u + 1013904223u;
      const size_t off = (s >> 8) % d.size();
      s = s * 1664525u + 1013904223u;
      const size_t n = std::min<size_t>(d.size() - off, (s >> 8) % (it % 3 == 0 ? 300 : 70000));
      bad += rg::crc32_of(d.data() + off, n) != (uint32_t)crc32(crc32(0L, Z_NULL, 0), d.data() + off, (uInt)n);
      ++cases;
    }
    printf("%zu cases, %zu bad\n", cases, bad);
    return bad ? 1 : 0;
  }
  if (argc != 2) return 2;
  const std::vector<unsigned char> d = slurp(argv[1]);
  size_t p = 0;
  std::vector<unsigned char> out;
  while (p + 8 <= d.size()) {
    unsigned clen, olen;
    memcpy(&clen, &d[p], 4);
    memcpy(&olen, &d[p + 4], 4);
    p += 8;
    out.assign((size_t)olen + 1, 0xAB);
    const std::vector<unsigned char> comp(d.begin() + p, d.begin() + p + clen);   // own allocation: a sanitizer sees over-reads
    const bool ok = inf.run(comp.data(), clen, out.data(), olen);
    if (out[olen] != 0xAB) { fprintf(stderr, "wrote past the end of the output\n"); return 3; }
    putchar(ok ? 1 : 0);
    if (ok) fwrite(out.data(), 1, olen, stdout);
    p += clen;
  }
  return 0;
}
