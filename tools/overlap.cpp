// tools/overlap.cpp — drop-in for bin/Overlap (src/Overlap.cpp), the greedy overlap assembly of the
// pulled reads, for its production input: a `.fastqd` file (scripts/Overlap.shorter.sh:141,153,165).
//
//   Overlap IN.fastqd MinPercent MinOverlap MinCoverage ReportStub SearchHash ACT OutStub LCtrim Threads
//     -> OutStub.fastq, OutStub.fastqd (contigs of length >= 95 whose max depth >= MinCoverage, named
//        @NODE_<SearchHash>_<idx>_L<len>_D<maxdep>:<F>:<R>:), OutStub.fastqgood.fastq (empty),
//        OutStub.fastqbad.fastq (reads too short after trimming)
//
// The greedy loop is serial by construction (step i merges read i into its best partner, and later
// steps see the merged contig; :935-1126) and runs here; the seed index (O2), the seed lookups of
// the candidate search (O3) and the Align3 placement search (O4) run on the GPU (rg_overlap_*).
// Results equal the reference run with Threads = 1 (with more threads its tie-breaking races);
// Threads is accepted and ignored. Not supported: the raw-FASTQ input branch (:763-846), which the
// pipeline never uses and which traps in the reference (AdjustBases falls off its end).
#include <algorithm>
#include <map>
#include <unordered_map>

#include "rg_cli.hpp"

static std::string revcomp(const std::string& s) {        // Util.cpp:187-210
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
static std::string rev_drop_nul(const std::string& s) {   // Util::RevQual, Util.cpp:212-222
  std::string r;
  r.reserve(s.size());
  for (size_t i = s.size(); i-- > 0;)
    if (s[i] != '\0') r.push_back(s[i]);
  return r;
}
static std::string flip_strands(const std::string& s) {   // :581-593
  std::string r;
  for (char c : s) {
    if (c == '+') r.push_back('-');
    else if (c == '-') r.push_back('+');
    else if (c == '.') r.push_back('.');
  }
  return r;
}

struct Read {
  std::string seq, qual, depth, strand;                   // depth: one byte per base
};

// :510-552: drop the ends whose depth is <= cutoff (depth bytes compared as char)
static void trim_low_coverage_ends(Read& r, int cutoff) {
  size_t hi = r.seq.size();
  while (hi > 0 && !((int)(signed char)r.depth[hi - 1] > cutoff)) --hi;
  r.seq.resize(hi); r.qual.resize(std::min(r.qual.size(), hi)); r.depth.resize(hi);
  if (r.seq.size() <= 1) { r.seq.clear(); r.qual.clear(); r.depth.clear(); return; }
  size_t lo = 0;
  while (lo < r.seq.size() && !((int)(signed char)r.depth[lo] > cutoff)) ++lo;
  r.seq.erase(0, lo);
  r.qual.erase(0, std::min(lo, r.qual.size()));
  r.depth.erase(0, lo);
}

// ColapsContigs (:362-466): merge A into B at offset k
static void collapse(const Read& A, int k, Read& B) {
  const int aoff = k > 0 ? k : 0, boff = k > 0 ? 0 : -k;
  const int asz = (int)A.seq.size(), bsz = (int)B.seq.size();
  std::string seq, qual, dep;
  auto at = [](const std::string& s, int i) -> char { return i < (int)s.size() ? s[(size_t)i] : '\0'; };
  for (int i = 0; i < asz + bsz; ++i) {
    const int ia = i - aoff, ib = i - boff;
    const bool in_a = ia >= 0 && ia < asz, in_b = ib >= 0 && ib < bsz;
    const char ab = in_a ? A.seq[(size_t)ia] : 'Z', bb = in_b ? B.seq[(size_t)ib] : 'Z';
    const char aq = in_a ? at(A.qual, ia) : '!', bq = in_b ? at(B.qual, ib) : '!';
    const unsigned char ad = in_a ? (unsigned char)at(A.depth, ia) : 0, bd = in_b ? (unsigned char)at(B.depth, ib) : 0;
    if (ab == bb && ab != 'Z') {
      seq.push_back(ab);
      qual.push_back(aq >= bq ? aq : bq);
      dep.push_back((int)ad + (int)bd < 250 ? (char)(ad + bd) : (char)250);
    } else if (ab == 'Z' && bb != 'Z') {
      seq.push_back(bb); qual.push_back(bq); dep.push_back((char)bd);
    } else if (ab != 'Z' && bb == 'Z') {
      seq.push_back(ab); qual.push_back(aq); dep.push_back((char)ad);
    } else if (ab != 'Z' && bb != 'Z') {
      if (ad > bd) { seq.push_back(ab); qual.push_back(aq); dep.push_back((char)ad); }
      else if (ad < bd) { seq.push_back(bb); qual.push_back(bq); dep.push_back((char)bd); }
      else if (aq >= bq) { seq.push_back(ab); qual.push_back(aq); dep.push_back((char)ad); }
      else { seq.push_back(bb); qual.push_back(bq); dep.push_back((char)bd); }
    } else {
      break;
    }
  }
  B.seq.swap(seq);
  B.qual.swap(qual);
  B.depth.swap(dep);
  B.strand += A.strand;
}

int main(int argc, char** argv) {
  printf("you gave %d Arguments\n", argc);
  if (argc != 11) {
    printf("ERROR, wrong numbe of arguemnts\nCall is: FASTQ, MinPercent, MinOverlap, MinCoverage, ReportStub, "
           "SearchHashSize, ACT, OutFile LCendTrimEpth Threads\n");
    return 0;
  }
  const std::string in_path = argv[1];
  const float min_percent = (float)atof(argv[2]);
  const int min_overlap = atoi(argv[3]), min_coverage = atoi(argv[4]), search_hash = atoi(argv[6]), act = atoi(argv[7]);
  const int trim_cutoff = atoi(argv[9]);
  const std::string out_stub = argv[8];
  if (in_path.find(".fastqd") == std::string::npos) die("Overlap (GPU): only .fastqd input is supported");
  FILE* in = fopen(in_path.c_str(), "rb");
  if (!in) { printf("Error, ParentHashFile could not be opened"); return 0; }
  FILE* report = fopen((out_stub + ".fastq").c_str(), "w");
  FILE* dep_report = fopen((out_stub + ".fastqd").c_str(), "w");
  FILE* good = fopen((out_stub + ".fastqgood.fastq").c_str(), "w");
  FILE* bad = fopen((out_stub + ".fastqbad.fastq").c_str(), "w");
  if (!report || !dep_report || !good || !bad) { printf("ERROR, Mut-Output file could not be opened - %s\n", out_stub.c_str()); return 0; }

  // ---- O1: 6-line records (:717-762)
  std::vector<Read> reads;
  {
    LineReader lr(in);
    const char* l;
    size_t n;
    while (lr.next(&l, &n)) {
      std::string l1(l, n), l3, l6;
      Read r;
      if (lr.next(&l, &n)) r.seq.assign(l, n);
      if (lr.next(&l, &n)) l3.assign(l, n);
      if (lr.next(&l, &n)) r.qual.assign(l, n);
      if (lr.next(&l, &n)) r.strand.assign(l, n);
      if (lr.next(&l, &n)) l6.assign(l, n);
      bool multiple = false;
      for (size_t a = 0; a <= l6.size();) {               // Util::Split on ' ' + atoi per token
        size_t b = l6.find(' ', a);
        if (b == std::string::npos) b = l6.size();
        if (!(b == l6.size() && a == b)) {                // a trailing empty token is dropped
          const unsigned char c = (unsigned char)atoi(l6.substr(a, b - a).c_str());
          r.depth.push_back((char)c);
          if ((int)c > 1) multiple = true;
        }
        a = b + 1;
      }
      if (multiple) trim_low_coverage_ends(r, trim_cutoff);
      if ((int)r.seq.size() > search_hash + 1) {
        reads.push_back(std::move(r));
      } else {
        fprintf(bad, "%s\n%s\n%s\n%s\n", l1.c_str(), r.seq.c_str(), l3.c_str(), r.qual.c_str());
      }
    }
    fclose(in);
  }
  fclose(good);
  fclose(bad);
  const size_t n_reads = reads.size();

  rg_ctx* ctx = rg_init(0);
  if (!ctx) die(rg_last_error());
  rg_overlap* ov = rg_overlap_create(ctx, search_hash);
  if (!ov) die(rg_last_error());
  auto build_index = [&](size_t first) {                  // O2 (:34-76)
    std::string cat;
    std::vector<uint64_t> offs(n_reads + 1, 0);
    for (size_t i = 0; i < n_reads; ++i) {
      offs[i] = cat.size();
      cat += reads[i].seq;
    }
    offs[n_reads] = cat.size();
    check(rg_overlap_build(ov, cat.data(), offs.data(), n_reads, first), "overlap build");
  };
  build_index(0);

  // O3: candidates of a batch of query sequences (:78-167); the device returns the raw seed hits,
  // the counting and the ranking (count descending, ties by descending index, at most 1001) is done here
  auto candidates = [&](const std::vector<std::string>& q, const std::vector<uint64_t>& ai,
                        std::vector<std::vector<uint32_t>>& out) {
    std::string cat;
    std::vector<uint64_t> offs(q.size() + 1, 0), hoffs(q.size() + 1, 0);
    for (size_t i = 0; i < q.size(); ++i) {
      offs[i] = cat.size();
      cat += q[i];
    }
    offs[q.size()] = cat.size();
    std::vector<uint32_t> hits(1 << 16);
    for (;;) {
      const int64_t n = rg_overlap_seed_hits(ov, cat.data(), offs.data(), ai.data(), q.size(), hits.data(), hits.size(),
                                             hoffs.data());
      if (n == RG_ERR_SMALL) { hits.resize(hoffs[q.size()] + 1024); continue; }
      check(n, "overlap seed hits");
      break;
    }
    out.assign(q.size(), {});
    for (size_t i = 0; i < q.size(); ++i) {
      std::map<uint32_t, int> positions;
      for (uint64_t h = hoffs[i]; h < hoffs[i + 1]; ++h) positions[hits[h]]++;
      std::vector<std::pair<int, uint32_t>> ranked;
      for (auto& kv : positions)
        if (kv.second > act) ranked.emplace_back(kv.second, kv.first);
      std::sort(ranked.begin(), ranked.end(), [](const std::pair<int, uint32_t>& a, const std::pair<int, uint32_t>& b) {
        return a.first != b.first ? a.first > b.first : a.second > b.second;
      });
      for (auto& cr : ranked) {
        if (cr.first >= act) {
          out[i].push_back(cr.second);
          if (out[i].size() > 1000) break;
        }
      }
    }
  };

  // O4 results of one candidate list, with the PerfectMatch short-circuit in candidate order (:203-345):
  // once a candidate has a perfect "inside" placement, the suffix/prefix searches are skipped for it
  // and for every later candidate
  std::vector<rg_align_result> res_f, res_r;
  auto resolve = [&](const std::vector<rg_align_result>& res, const std::vector<uint32_t>& cand, bool& perfect,
                     int& best_index, int& offset) {
    int best = 0;
    for (size_t c = 0; c < cand.size(); ++c) {
      const rg_align_result& r = res[c];
      const bool only_inside = perfect || r.perfect1;
      const int score = only_inside ? r.score1 : r.score_all, off = only_inside ? r.overlap1 : r.overlap_all;
      if (r.perfect1) perfect = true;
      if (score > best) {
        best = score;
        best_index = (int)cand[c];
        offset = off;
      }
    }
    return best;
  };

  const size_t buffer = 100;                              // Buffer = 100 * Threads with Threads = 1 (:648)
  long since_build = 1;
  for (size_t b0 = 0; b0 < n_reads; b0 += buffer) {
    since_build += (long)buffer;
    if (since_build > 1000000) {
      build_index(b0);
      since_build = 0;
    }
    const size_t hi = std::min(n_reads, b0 + buffer);
    std::vector<std::string> q;
    std::vector<uint64_t> ai;
    for (size_t i = b0; i < hi; ++i) { q.push_back(reads[i].seq); ai.push_back(i); }
    for (size_t i = b0; i < hi; ++i) { q.push_back(revcomp(reads[i].seq)); ai.push_back(i); }
    std::vector<std::vector<uint32_t>> lists;
    candidates(q, ai, lists);
    for (size_t i = b0; i < hi; ++i) {
      Read A = reads[i];
      // both strands in one launch (the reference only searches the reverse strand when the forward
      // search found no perfect match, :973-1001; computing it anyway changes nothing: it is ignored then)
      Read R;
      R.seq = revcomp(A.seq);
      const std::vector<uint32_t>& cf = lists[i - b0];
      const std::vector<uint32_t>& cr = lists[(hi - b0) + (i - b0)];
      const bool same_len = R.seq.size() == A.seq.size();   // (RevComp drops characters other than ACGTN)
      res_f.resize(cf.size());
      res_r.resize(cr.size());
      if (same_len) {
        check(rg_overlap_align_both(ov, A.seq.data(), R.seq.data(), A.seq.size(), cf.data(), cf.size(), cr.data(), cr.size(),
                                    min_percent, min_overlap, res_f.data(), res_r.data()), "align");
      } else {
        check(rg_overlap_align(ov, A.seq.data(), A.seq.size(), cf.data(), cf.size(), min_percent, min_overlap, res_f.data()), "align");
        check(rg_overlap_align(ov, R.seq.data(), R.seq.size(), cr.data(), cr.size(), min_percent, min_overlap, res_r.data()), "align");
      }
      bool perfect = false;
      int best_index = -1, k = 0;
      int score = resolve(res_f, cf, perfect, best_index, k);
      if (!perfect) {
        R.qual = rev_drop_nul(A.qual);
        R.depth = rev_drop_nul(A.depth);
        R.strand = flip_strands(A.strand);
        int rbest = -1, rk = -1;
        const int rscore = resolve(res_r, cr, perfect, rbest, rk);
        if (rscore > score) {
          A = R;
          score = rscore;
          best_index = rbest;
          k = rk;
        }
      }
      if (score < min_overlap || best_index < 0) continue;
      collapse(A, k, reads[(size_t)best_index]);
      reads[i].seq = "moved";
      const std::string& merged = reads[(size_t)best_index].seq;
      check(rg_overlap_update(ov, (uint64_t)best_index, merged.data(), merged.size()), "overlap update");
      // (read i itself is never looked at again: candidates always have a larger index than the query)
    }
  }

  // ---- output (:1128-1174)
  printf("\nRESULTS\n");
  size_t count = 0;
  for (size_t i = 0; i < n_reads; ++i) {
    const Read& r = reads[i];
    if (r.seq == "moved" || r.seq.size() < 95) continue;
    int max_dep = -1;
    for (char c : r.depth) max_dep = std::max(max_dep, (int)(unsigned char)c);
    if (max_dep < min_coverage) continue;
    ++count;
    int f = 0, rv = 0;
    for (char c : r.strand) { if (c == '+') ++f; else if (c == '-') ++rv; }
    char name[256];
    snprintf(name, sizeof(name), "@NODE_%s_%zu_L%zu_D%d:%d:%d:", argv[6], i, r.seq.size(), max_dep, f, rv);
    fprintf(report, "%s\n%s\n+\n%s\n", name, r.seq.c_str(), r.qual.c_str());
    fprintf(dep_report, "%s\n%s\n+\n%s\n%s\n", name, r.seq.c_str(), r.qual.c_str(), r.strand.c_str());
    for (size_t w = 0; w < r.depth.size(); ++w) fprintf(dep_report, w ? " %d" : "%d", (int)(unsigned char)r.depth[w]);
    if (r.depth.empty()) fprintf(dep_report, "0");
    fprintf(dep_report, "\n");
  }
  printf("Wrote %zu sequences\n", count);
  fclose(report);
  fclose(dep_report);
  rg_overlap_free(ov);
  rg_shutdown(ctx);
  return 0;
}
