// tools/rufus_filter.cpp — drop-ins for bin/RUFUS.Filter (paired; src/RUFUS.Filter.cpp:26-301) and,
// compiled with -DRG_SINGLE, bin/RUFUS.Filter.single (src/RUFUS.Filter.ss.cpp:25-211), with the
// per-read k-mer lookups done on the GPU through librufus_gpu.so.
//
//   RUFUS.Filter        HashList Mate1.fq Mate2.fq OutStub K MinQ Threshold Threads
//        -> OutStub.Mutations.Mate1.fastq / OutStub.Mutations.Mate2.fastq (the 4 lines of both mates,
//           verbatim, of every pair where one mate has >= Threshold mutant k-mer windows)
//   RUFUS.Filter.single HashList Reads.fq|stdin OutStub K MinQ Threshold Threads
//        -> OutStub.Mutations.fastq (read name + ":MH<hits>")
// The two mate files may be FIFOs: records are read 4 lines at a time, alternating between the
// files, like the reference does (a lock-step reader never blocks the producer pair).
// An extension (paired tool): when Mate1.fq is a BAM file (and Mate2.fq names the same file or is "-"), the BAM is
// decoded and re-paired in process on `Threads` threads — the pairs PassThroughSamCheck.stranded would have written
// to the two files, in the same order (tools/rg_bam_pairs.hpp).
// Like the reference, problems opening files are reported on stdout and the exit code stays 0
// for the cases the reference returns 0 for (runRufus.sh only looks at the outputs).
#include "rg_cli.hpp"

#include "rg_fastq_batch.hpp"
#ifndef RG_SINGLE
#include "rg_bam_pairs.hpp"

[[noreturn]] void rgbam::fail(const std::string& m) {   // (raised on a pipeline thread: leave without unwinding the others)
  fprintf(stderr, "RUFUS.Filter: BAM input: %s\n", m.c_str());
  fflush(stderr);
  _exit(1);
}
#endif

static std::string slurp(const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) return std::string();
  std::string s;
  char buf[1 << 16];
  size_t got;
  while ((got = fread(buf, 1, sizeof(buf), f)) > 0) s.append(buf, got);
  fclose(f);
  return s;
}

int main(int argc, char** argv) {
#ifdef RG_SINGLE
  const int need = 8;
  printf("Call is PreBuiltMutHash Mutant.fq  firstpassfile hashsize MinQ HashCountThreshold threads \n");
#else
  const int need = 9;
  printf("Call is PreBuiltMutHash Mutant.fq1 Mutant.fq2 firstpassfile hashsize MinQ HashCountThreshold threads \n");
#endif
  if (argc < need) {
    printf("Error, not enough arguments\n");
    return 1;
  }
  int a = 1;
  const char* hashlist = argv[a++];
  const char* in1 = argv[a++];
#ifndef RG_SINGLE
  const char* in2 = argv[a++];
#endif
  const std::string stub = argv[a++];
  const int k = atoi(argv[a++]);
  const int minq = atoi(argv[a++]);
  const int thresh = atoi(argv[a++]);
  const int threads = atoi(argv[a++]);   // (BAM input: decode threads; the GPU needs none)
  (void)threads;
  if (k < 1 || k > 32) die("HashSize must be 1..32");

  FILE* hf = fopen(hashlist, "rb");
  if (!hf) {
    printf("Error, ParentHashFile could not be opened");
    return 0;
  }
  fclose(hf);
#ifdef RG_SINGLE
  FILE* f1 = fopen(!strcmp(in1, "stdin") ? "/dev/stdin" : in1, "rb");
#else
  FILE* f1 = fopen(in1, "rb");
  FILE* f2 = !strcmp(in2, "-") ? nullptr : fopen(in2, "rb");
  if (!f2 && strcmp(in2, "-") != 0) {
    printf("Error, MutFile2 could not be opened");
    return 0;
  }
#endif
  if (!f1) {
    printf("Error, MutFile could not be opened");
    return 0;
  }
#ifdef RG_SINGLE
  FILE* o1 = fopen((stub + ".Mutations.fastq").c_str(), "w");
#else
  FILE* o1 = fopen((stub + ".Mutations.Mate1.fastq").c_str(), "w");
  FILE* o2 = fopen((stub + ".Mutations.Mate2.fastq").c_str(), "w");
  if (!o2) {
    printf("ERROR, Output file could not be opened -%s\n", stub.c_str());
    return 0;
  }
#endif
  if (!o1) {
    printf("ERROR, Output file could not be opened -%s\n", stub.c_str());
    return 0;
  }

  rg_ctx* ctx = rg_init(0);
  if (!ctx) die(rg_last_error());
  const std::string text = slurp(hashlist);
#ifdef RG_SINGLE
  rg_set* set = rg_set_from_hashlist_text(ctx, text.data(), text.size(), k, 0);
#else
  rg_set* set = rg_set_from_hashlist_text(ctx, text.data(), text.size(), k, 1);
#endif
  if (!set) die(rg_last_error());
  printf("Mutant k-mers loaded: %llu\n", (unsigned long long)rg_set_size(set));

  // 64 Ki reads (pairs) per call: ~40 MB of host buffers that stay resident and are reused, instead of 1 Mi reads whose
  // growing vectors fault in 1.2 GB of fresh pages per batch (measured on the host side alone, 500 k pairs:
  // 1.19 s at 1 Mi, 0.46 s at 64 Ki; profiles/r2_host_ingest.txt); still ~20 Mbases per kernel launch
  const uint64_t batch_reads = 1u << 16;
  uint64_t n_in = 0, n_out = 0;
  LineReader r1(f1);
  Batch b1;
#ifndef RG_SINGLE
  LineReader r2(f2 ? f2 : f1);   // (f2 is absent only for a BAM input, where r2 is never read)
  Batch b2;
  const bool is_bam = r1.peek() == 0x1f;   // (the first read of mate 1, which the loop below would do next anyway)
  if (!f2 && !is_bam) {
    printf("Error, MutFile2 could not be opened");
    return 0;
  }
  bool first = true;
  auto filter_and_write = [&](Batch& x, Batch& y) {
    std::vector<uint8_t> keep(x.n());
    check(rg_filter_pairs(set, x.seq.data(), x.qual.data(), x.offs.data(), y.seq.data(), y.qual.data(), y.offs.data(), x.n(),
                          minq, thresh, keep.data()), "filter");
    for (uint64_t i = 0; i < x.n(); ++i)
      if (keep[i]) {
        fwrite(x.text.data() + x.text_offs[i], 1, x.text_offs[i + 1] - x.text_offs[i], o1);
        fwrite(y.text.data() + y.text_offs[i], 1, y.text_offs[i + 1] - y.text_offs[i], o2);
        ++n_out;
      }
  };
#endif
  for (bool more = true; more;) {
    b1.clear();
#ifdef RG_SINGLE
    while (b1.n() < batch_reads && (more = read_record(r1, b1))) {}
    if (b1.n() == 0) break;
    std::vector<int32_t> hits(b1.n());
    check(rg_filter_single(set, b1.seq.data(), b1.qual.data(), b1.offs.data(), b1.n(), minq, hits.data()), "filter");
    for (uint64_t i = 0; i < b1.n(); ++i)
      if (hits[i] >= thresh) {
        const char* rec = b1.text.data() + b1.text_offs[i];
        fwrite(rec, 1, b1.name_len[i], o1);
        fprintf(o1, ":MH%d", hits[i]);
        fwrite(rec + b1.name_len[i], 1, b1.text_offs[i + 1] - b1.text_offs[i] - b1.name_len[i], o1);
        ++n_out;
      }
#else
    b2.clear();
    if (first && is_bam) {                 // a BAM file: every pair comes from it, re-paired in process
      rgbam::bam_pairs(f1, r1.buf.data() + r1.pos, r1.len - r1.pos, (unsigned)(threads > 0 ? threads : 1), b1, b2, batch_reads,
                       [&](Batch& x, Batch& y) {
                         filter_and_write(x, y);
                         n_in += x.n();
                         x.clear();
                         y.clear();
                       });
      more = false;
    } else {
      more = read_pair_batch(r1, r2, b1, b2, batch_reads);
    }
    first = false;
    if (b1.n() == 0) break;
    filter_and_write(b1, b2);
#endif
    n_in += b1.n();
  }
  fclose(o1);
  fclose(f1);
#ifndef RG_SINGLE
  fclose(o2);
  if (f2) fclose(f2);
#endif
  printf("Read in %llu records: Found %llu\n\nDone running RUFUS.Filter (GPU)\n", (unsigned long long)n_in,
         (unsigned long long)n_out);
  rg_set_free(set);
  rg_shutdown(ctx);
  return 0;
}
