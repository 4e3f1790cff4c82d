// tools/micro/bam_seq_check.cpp -> tools/bin/rg_input_to_seq — test and timing driver for the input side of `jellyfish
// count` (tools/rg_bam_input.hpp:read_input): prints the byte stream the counter would be handed (every chunk, in
// order) for a FASTA / FASTQ / BAM input, using an ordinary heap buffer instead of pinned memory: no GPU needed.
//   rg_input_to_seq [-t THREADS] [-c CHUNK_BYTES] FILE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <string>

#include "../rg_bam_input.hpp"

[[noreturn]] void rgbam::fail(const std::string& m) {
  fprintf(stderr, "rg_input_to_seq: %s\n", m.c_str());
  fflush(stderr);
  _exit(1);
}

// (HeapBuf: tools/rg_cli.hpp)

int main(int argc, char** argv) {
  unsigned threads = 1;
  size_t chunk = 8u << 20;
  const char* path = nullptr;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "-t") && i + 1 < argc) threads = (unsigned)atoi(argv[++i]);
    else if (!strcmp(argv[i], "-c") && i + 1 < argc) chunk = (size_t)atoll(argv[++i]);
    else path = argv[i];
  }
  if (!path) return 2;
  FILE* f = fopen(path, "rb");
  if (!f) die(std::string("Failed to open input file '") + path + "'");
  HeapBuf b;
  auto flush = [&](HeapBuf& x) {
    if (fwrite(x.s.data(), 1, x.s.size(), stdout) != x.s.size()) _exit(1);
    x.clear();
  };
  read_input(f, path, b, chunk, flush, threads);
  if (!b.empty()) flush(b);
  fflush(stdout);
  return 0;
}
