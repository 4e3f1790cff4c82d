// tools/annotate_overlap.cpp — drop-in for bin/AnnotateOverlap (src/AnnotateOverlap.cpp:22-161), the tool
// that writes `<contigs>.hashcount.fastq` (Overlap.shorter.sh): every contig gets ":MH0" appended to
// its name and its quality line replaced by the per-base depth of mutant k-mer windows,
// char(33 + min(depth, 93)); the k-mer lookups run on the GPU (rg_annotate).
//
//   AnnotateOverlap HashList Contigs.fastq|stdin HashOut > annotated.fastq
//   HashOut: for every window i < length-K of every contig, the lexicographically smaller of the k-mer
//            and its reverse complement, followed by " 1" (:149-156)
#include "rg_cli.hpp"

static std::string revcomp(const std::string& s) {     // Util.cpp:187-210: other characters are dropped
  std::string r;
  r.reserve(s.size());
  for (size_t i = s.size(); i-- > 0;) {
    switch (s[i]) {
      case 'A': r.push_back('T'); break;
      case 'C': r.push_back('G'); break;
      case 'G': r.push_back('C'); break;
      case 'T': r.push_back('A'); break;
      case 'N': r.push_back('N'); break;
      default: break;
    }
  }
  return r;
}

int main(int argc, char** argv) {
  if (argc < 4) die("usage: AnnotateOverlap HashList FASTQ|stdin HashOut > annotated.fastq");
  std::string text;
  int k = 0;
  {
    FILE* f = fopen(argv[1], "rb");
    if (!f) die(std::string("cannot open ") + argv[1]);
    char buf[1 << 16];
    size_t got;
    while ((got = fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, got);
    fclose(f);
    // HashSize = length of the k-mer column of the LAST line (:60-76)
    size_t pos = 0;
    while (pos < text.size()) {
      size_t nl = text.find('\n', pos);
      if (nl == std::string::npos) nl = text.size();
      const std::string line = text.substr(pos, nl - pos);
      pos = nl + 1;
      if (line.empty()) continue;
      std::vector<std::string> sp;
      size_t a = 0;
      for (;;) {
        const size_t b = line.find(' ', a);
        if (b == std::string::npos) { if (a < line.size()) sp.push_back(line.substr(a)); break; }
        sp.push_back(line.substr(a, b - a));
        a = b + 1;
      }
      if (sp.size() == 2) k = (int)sp[0].size();
      else if (sp.size() == 4) k = (int)sp[3].size();
      else if (sp.size() == 1) k = (int)line.substr(0, line.find('\t')).size();
    }
  }
  if (k < 1 || k > 32) die("AnnotateOverlap: HashList without k-mers (or k > 32)");
  FILE* in = fopen(!strcmp(argv[2], "stdin") ? "/dev/stdin" : argv[2], "rb");
  if (!in) die(std::string("cannot open ") + argv[2]);
  FILE* hash_out = fopen(argv[3], "w");
  if (!hash_out) die(std::string("cannot open ") + argv[3]);

  rg_ctx* ctx = rg_init(0);
  if (!ctx) die(rg_last_error());
  rg_set* set = rg_set_from_hashlist_text(ctx, text.data(), text.size(), k, 1);
  if (!set) die(rg_last_error());

  LineReader lr(in);
  std::vector<std::string> l1, l2, l3;
  std::vector<char> seq, qual;
  std::vector<uint64_t> offs;
  std::vector<uint8_t> depth;
  const char* l;
  size_t n;
  for (bool more = true; more;) {
    l1.clear(); l2.clear(); l3.clear(); seq.clear(); qual.clear();
    offs.assign(1, 0);
    while (l1.size() < (1u << 18)) {
      if (!lr.next(&l, &n)) { more = false; break; }
      l1.emplace_back(l, n);
      std::string s, p, q;
      if (lr.next(&l, &n)) s.assign(l, n);
      if (lr.next(&l, &n)) p.assign(l, n);
      if (lr.next(&l, &n)) q.assign(l, n);
      seq.insert(seq.end(), s.begin(), s.end());
      q.resize(s.size(), '\0');              // a missing quality reads as '\0': below any threshold
      qual.insert(qual.end(), q.begin(), q.end());
      offs.push_back(seq.size());
      l2.push_back(std::move(s));
      l3.push_back(std::move(p));
    }
    if (l1.empty()) break;
    depth.assign(seq.size() + 1, 0);
    check(rg_annotate(set, seq.data(), qual.data(), offs.data(), l1.size(), 3, depth.data()), "annotate");
    std::string out;
    for (size_t r = 0; r < l1.size(); ++r) {
      out.clear();
      out += l1[r];
      out += ":MH0\n";
      out += l2[r];
      out += "\n";
      out += l3[r];
      out += "\n";
      for (uint64_t j = offs[r]; j < offs[r + 1]; ++j) out.push_back(depth[j] < 93 ? (char)(depth[j] + 33) : (char)126);
      out += "\n";
      fwrite(out.data(), 1, out.size(), stdout);
      const std::string& s = l2[r];
      for (size_t i = 0; i + (size_t)k < s.size(); ++i) {
        const std::string h = s.substr(i, (size_t)k), rc = revcomp(h);
        fprintf(hash_out, "%s 1\n", (h < rc ? h : rc).c_str());
      }
    }
  }
  fclose(hash_out);
  fclose(in);
  rg_set_free(set);
  rg_shutdown(ctx);
  return 0;
}
