// tools/micro/parse_bench.cpp — throughput of the FASTA/FASTQ scan of `jellyfish count` (tools/rg_cli.hpp:
// read_sequences) on one host thread, into an ordinary heap buffer instead of pinned memory (no GPU needed).
// Measurement driver, not product code.
//   g++ -O2 -std=c++17 -Itools -Iinclude tools/micro/parse_bench.cpp -o /tmp/parse_bench \
//       -Lrufus_b200 -lrufus_gpu -Wl,-rpath,$PWD/rufus_b200 -Wl,-rpath-link,/usr/local/cuda/lib64
//   /tmp/parse_bench reads.fastq
#include <chrono>
#include <string>

#include "rg_cli.hpp"

// (HeapBuf: tools/rg_cli.hpp)

int main(int argc, char** argv) {
  if (argc != 2) return 2;
  for (int rep = 0; rep < 3; rep++) {
    FILE* f = fopen(argv[1], "rb");
    if (!f) die(std::string("cannot open ") + argv[1]);
    HeapBuf b;
    b.s.reserve(80u << 20);
    const size_t chunk = 64u << 20;
    uint64_t total = 0;
    const auto t0 = std::chrono::steady_clock::now();
    read_sequences(f, argv[1], b, chunk, [&](HeapBuf& x) { total += x.size(); x.clear(); });
    total += b.size();
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    fseek(f, 0, SEEK_END);
    const double sz = (double)ftell(f);
    fclose(f);
    printf("sequence bytes %lu of a %.0f MB file in %.3f s: %.2f GB/s of file, %.0f Mbases/s\n", (unsigned long)total, sz / 1e6, dt, sz / 1e9 / dt, total / 1e6 / dt);
  }
  return 0;
}
