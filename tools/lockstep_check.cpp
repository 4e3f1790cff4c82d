// tools/lockstep_check.cpp — TEST HARNESS (no GPU): reads two mate files exactly the way
// tools/rufus_filter.cpp does (tools/rg_fastq_batch.hpp, small batches) and prints every pair's two
// header lines. tests/test_tools_host.py uses it to prove that FIFO pipelines with a lock-step
// producer (the reference's or our PassThroughSamCheck.stranded) do not deadlock.
//   rg_lockstep_check MATE1 MATE2 [batch_reads]
#include "rg_fastq_batch.hpp"

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: rg_lockstep_check MATE1 MATE2 [batch_reads]\n"); return 2; }
  FILE* f1 = fopen(argv[1], "rb");
  FILE* f2 = fopen(argv[2], "rb");
  if (!f1 || !f2) { fprintf(stderr, "cannot open the inputs\n"); return 1; }
  const uint64_t batch = argc > 3 ? strtoull(argv[3], nullptr, 10) : 1000;
  LineReader r1(f1), r2(f2);
  Batch b1, b2;
  uint64_t n = 0;
  for (bool more = true; more;) {
    b1.clear();
    b2.clear();
    more = read_pair_batch(r1, r2, b1, b2, batch);
    for (uint64_t i = 0; i < b1.n(); ++i) {
      fwrite(b1.text.data() + b1.text_offs[i], 1, b1.name_len[i], stdout);
      fputc('\t', stdout);
      fwrite(b2.text.data() + b2.text_offs[i], 1, b2.name_len[i], stdout);
      fputc('\n', stdout);
    }
    n += b1.n();
  }
  fprintf(stderr, "%llu pairs\n", (unsigned long long)n);
  return 0;
}
