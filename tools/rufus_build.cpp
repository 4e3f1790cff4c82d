// tools/rufus_build.cpp — drop-in for bin/RUFUS.Build (src/RUFUS.Build.cpp:33-294), the legacy
// selection of subject-specific k-mers from sorted text dumps, with the table lookups on the GPU.
//
//   RUFUS.Build -c CONTROL.tab [-c ...] -s SUBJECT.tab -o OUT -hs K -mS minSubject -mC maxControlSum
//               [-max maxSubject] [-t threads] [-d controlDelimiter]
//   SUBJECT.tab : `KMER<TAB>COUNT` lines; CONTROL.tab : `KMER<d>COUNT` lines (d = tab unless -d)
//   OUT         : `HashToLong(KMER)<TAB>sum of control counts<TAB>subject count<TAB>KMER`, in subject
//                 order, for every subject k-mer with minSubject <= count <= maxSubject whose summed
//                 control count is <= maxControlSum (:224-262); OUT.Buld.Log is created empty (:186-196).
// The reference walks the files as a merge join and therefore needs them sorted; the result below is
// the same for sorted inputs without duplicate k-mers (what `jellyfish dump | sort` produces).
// Like the reference, problems are reported on stdout and the exit code stays 0.
#include <math.h>

#include <unordered_map>

#include "rg_cli.hpp"

struct Dump {
  std::vector<uint64_t> keys;
  std::vector<uint32_t> counts;
  std::vector<std::string> text;   // subject only: the k-mer column as written
};

static bool pack_kmer(const char* s, size_t n, int k, uint64_t* key) {
  if ((int)n != k) return false;
  uint64_t v = 0;
  for (size_t i = 0; i < n; ++i) {
    int c;
    switch (s[i]) {
      case 'A': c = 0; break;
      case 'C': c = 1; break;
      case 'G': c = 2; break;
      case 'T': c = 3; break;
      default: return false;
    }
    v = (v << 2) | (uint64_t)c;
  }
  *key = v;
  return true;
}

static bool read_dump(const char* path, char delim, int k, bool keep_text, Dump& d) {
  FILE* f = fopen(path, "rb");
  if (!f) return false;
  LineReader lr(f);
  const char* l;
  size_t n;
  while (lr.next(&l, &n)) {
    const char* sep = (const char*)memchr(l, delim, n);
    if (!sep) continue;
    uint64_t key;
    if (!pack_kmer(l, (size_t)(sep - l), k, &key)) continue;
    d.keys.push_back(key);
    d.counts.push_back((uint32_t)atoi(std::string(sep + 1, l + n).c_str()));
    if (keep_text) d.text.emplace_back(l, sep);
  }
  fclose(f);
  return true;
}

int main(int argc, char** argv) {
  std::vector<const char*> controls;
  const char* subject = nullptr;
  std::string out;
  int k = -1, threads = 1;
  double min_subject = -1.0, max_control = -1.0;
  long max_cov = 100000000;
  char delim = '\t';
  for (int i = 1; i < argc; ++i) {
    const std::string p = argv[i];
    auto need = [&](int n) {
      if (i + n >= argc) {
        printf("ERROR: missing value for %s\n", p.c_str());
        exit(0);
      }
    };
    if (p == "-h") { printf("RUFUS.Build (GPU): -c control.tab ... -s subject.tab -o out -hs K -mS min -mC maxControlSum [-max max] [-t T] [-d delim]\n"); return 0; }
    else if (p == "-c") { need(1); controls.push_back(argv[++i]); }
    else if (p == "-s") { need(1); subject = argv[++i]; }
    else if (p == "-o") { need(1); out = argv[++i]; }
    else if (p == "-hs") { need(1); k = atoi(argv[++i]); }
    else if (p == "-mS") { need(1); min_subject = atof(argv[++i]); }
    else if (p == "-mC") { need(1); max_control = atof(argv[++i]); }
    else if (p == "-max") { need(1); max_cov = atol(argv[++i]); }
    else if (p == "-t") { need(1); threads = atoi(argv[++i]); }
    else if (p == "-d") { need(1); delim = argv[++i][0]; }
    else { printf("ERROR: unkown command line paramater -%s-\n", argv[i]); return 0; }
  }
  (void)threads;
  if (controls.empty()) { printf("Error in control file inputs, either empty or number of file not the same for both\n"); return 0; }
  if (!subject) { printf("Error subject file required\n"); return 0; }
  if (out.empty()) out = subject;
  if (k == -1) { printf("Error Hash Size required\n"); return 0; }
  if (k < 1 || k > 32) { printf("Error Hash Size must be 1..32\n"); return 0; }
  if (min_subject < 0) { printf("Error Minimum coverage in the subject required and must be greather than 0\n"); return 0; }
  if (max_control < -1) { printf("Error Maximum Coverage in the control must be set and must be greaterthan or equal to 0\n"); return 0; }

  std::vector<Dump> ctl(controls.size());
  for (size_t i = 0; i < controls.size(); ++i)
    if (!read_dump(controls[i], delim, k, false, ctl[i])) {
      printf("Error, ParentHashFile could not be opened\n%s\n", controls[i]);
      return 0;
    }
  Dump sub;
  if (!read_dump(subject, '\t', k, true, sub)) { printf("Error, MutHashFile could not be opened"); return 0; }
  FILE* fo = fopen(out.c_str(), "w");
  if (!fo) { printf("Error MutHashTableFilter file couldnt be opened - %s\n", out.c_str()); return 0; }
  FILE* flog = fopen((out + ".Buld.Log").c_str(), "w");
  if (flog) fclose(flog);

  rg_ctx* ctx = rg_init(0);
  if (!ctx) die(rg_last_error());
  rg_table* st = rg_table_from_pairs(ctx, k, 1, sub.keys.data(), sub.counts.data(), sub.keys.size());
  if (!st) die(rg_last_error());
  std::vector<rg_table*> ct;
  for (Dump& d : ctl) {
    rg_table* t = rg_table_from_pairs(ctx, k, 1, d.keys.data(), d.counts.data(), d.keys.size());
    if (!t) die(rg_last_error());
    ct.push_back(t);
  }
  // `Acount >= MinCovSubject` / `TotalParentDepth <= MaxCovControl` compare ints with doubles (:226,257)
  const int64_t lo = (int64_t)ceil(min_subject), hi = max_cov, mc = (int64_t)floor(max_control);
  uint64_t cap = sub.keys.size() ? sub.keys.size() : 1, need = 0;
  std::vector<uint64_t> keys(cap);
  std::vector<uint32_t> counts(cap), sums(cap);
  const int64_t n = rg_build_select(st, lo, hi, ct.data(), (int)ct.size(), mc, keys.data(), counts.data(), sums.data(),
                                    cap, &need);
  check(n, "build");
  std::unordered_map<uint64_t, uint32_t> sum_of;
  sum_of.reserve((size_t)n * 2 + 1);
  for (int64_t i = 0; i < n; ++i) sum_of[keys[(size_t)i]] = sums[(size_t)i];
  uint64_t found = 0;
  for (size_t i = 0; i < sub.keys.size(); ++i) {      // subject order, like the reference's single pass
    auto it = sum_of.find(sub.keys[i]);
    if (it == sum_of.end()) continue;
    if ((int64_t)sub.counts[i] < lo || (int64_t)sub.counts[i] > hi) continue;
    fprintf(fo, "%llu\t%u\t%u\t%s\n", (unsigned long long)rg_rufus_pack(sub.text[i].c_str(), k), it->second,
            sub.counts[i], sub.text[i].c_str());
    ++found;
  }
  fclose(fo);
  printf("\nDone reading Parent File\nmuts = %llu\n\nreally done\n", (unsigned long long)found);
  rg_table_free(st);
  for (rg_table* t : ct) rg_table_free(t);
  rg_shutdown(ctx);
  return 0;
}
