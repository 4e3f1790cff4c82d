// tools/model_dist.cpp — drop-in for bin/ModelDist (src/ModelDist.cpp:375-973), the copy-number model
// fit that runRufus.sh:849-867 runs on `<generator>.Jhash.histo` to obtain MutantMinCov (line 2 of the
// `.7.7.model` file) and MutantSC (line 4). Pure host FP64 work on a 10 k-line histogram: no GPU step.
//
//   ModelDist HISTO K READLEN THREADS      -> HISTO.7.7.model, HISTO.7.7.dist, HISTO.7.7.prob + a log on stdout
//
// Same arithmetic, same order of every floating-point sum as the reference, so the three files and the log
// come out byte-identical (tests/test_model_dist.py compares them with the reference binary built by
// oracle/Makefile); what changes is the amount of work:
//   * the reference rebuilds, for each of the ~760 candidate parameter sets of the coordinate search, three
//     dense (rows x copy-number) matrices of vector<vector<double>> — 3.5 M exp() calls and 10 k allocations per
//     candidate. Here a candidate costs one pass over the NON-ZERO band of each column: exp(-t) is exactly
//     +0 in IEEE double for t > 746, a +0 term leaves a sum unchanged, so rows outside |x - mu| <= sqrt(746 *
//     2 sigma^2) are skipped without changing a bit of any sum; the per-column constants are hoisted; only the
//     rows the residual needs (Inflection .. 5*SC and the SC*a anchors) are normalised.
//   * the skew search of the reference never runs (its loop condition `SKhigh < 1e-50` is false from the
//     start, ModelDist.cpp:636), so skew stays 0, pow(0, power) = 0 and the "power" search compares eleven
//     identical numbers: evaluated once.
//   * THREADS is honoured (the reference ignores it and always asks OpenMP for 11 threads).
// Reference behaviour kept on purpose: the histogram is indexed from the first non-zero line (with -L 2 the
// count of multiplicity 2 sits at index 1, ModelDist.cpp:432-449), `values[9]` is treated as the 11-entry
// array the loops use, and the two reads the reference makes one element past a vector (ErrorDist[size] in
// the last `.prob` line, prob[SC*a][a+1] for the last copy number) are taken as 0, which is what the
// reference binary produces on freshly allocated memory.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <thread>
#include <vector>

static double pi = 3.14159L;  // ModelDist.cpp:28 (long double literal rounded to double)

// ModelDist.cpp:31-37
static double norm_ref(double x, double mu, double sigma, double skew, double p) {
  if (x < mu) sigma = sigma + pow((mu - x) * skew, p);
  return (1 / (sqrt(2 * pi * pow(sigma, 2)))) * exp(-(pow(x - mu, 2) / (2 * pow(sigma, 2))));
}

struct Column {
  double mu, sigma, coef, den;
  long lo, hi;  // rows outside [lo, hi] are exactly 0 (skew == 0 only)
};

struct Shape {  // the columns of one candidate (SC, stdev, factor): index 1 = het peak, index j + 1 = j copies
  long N, J;    // rows, number of j with j < N / SC
  double skew, power;
  std::vector<Column> col;  // 1 .. J + 1
  bool banded;

  Shape(long n, double SC, double stdev, double factor, double skew_, double power_) : N(n), skew(skew_), power(power_) {
    J = 0;
    for (long j = 1; j < (double)(size_t)N / SC; j++) J++;
    col.resize(J + 2);
    banded = (skew == 0 && power > 0);
    set(1, SC / 2, stdev * (1 - ((1 - (stdev / (stdev + ((1) * factor)))) / 2)));
    for (long j = 1; j <= J; j++) set(j + 1, SC * j, stdev + ((j - 1) * factor));
  }
  void set(long c, double mu, double sigma) {
    Column& k = col[c];
    k.mu = mu;
    k.sigma = sigma;
    k.coef = (1 / (sqrt(2 * pi * pow(sigma, 2))));
    k.den = (2 * pow(sigma, 2));
    k.lo = 0;
    k.hi = N - 1;
    if (!(k.den > 0) || !(k.coef > 0) || !std::isfinite(k.coef) || !std::isfinite(k.den) || !std::isfinite(mu)) {
      banded = false;
      return;
    }
    double w = sqrt(746.0 * k.den) * (1 + 1e-9) + 2;  // |x - mu| > w  =>  (x - mu)^2 / den > 746
    if (mu - w > 0) k.lo = (long)(mu - w);
    if (mu + w < (double)(N - 1)) k.hi = (long)(mu + w) + 1;
    if (k.hi > N - 1) k.hi = N - 1;
  }
  inline double raw(long i, long c) const {
    const Column& k = col[c];
    double x = (double)i;
    if (!banded) return norm_ref(x, k.mu, k.sigma, skew, power);
    double t = pow(x - k.mu, 2) / k.den;
    if (t > 746.0) return 0.0;  // exp(-t) rounds to +0 and coef is finite
    return k.coef * exp(-t);
  }
  double column_sum(long c, long first_row) const {
    double sum = 0.0;
    long lo = banded ? (col[c].lo > first_row ? col[c].lo : first_row) : first_row;
    long hi = banded ? col[c].hi : N - 1;
    for (long i = lo; i <= hi; i++) sum += raw(i, c);
    return sum;
  }
};

// testModel / testModelLog (ModelDist.cpp:72-189 and 191-306): residual of one candidate against the
// error-corrected histogram over rows [Inflection, SC * MaxCopy).
static double test_model(bool log_residual, double SC, double stdev, double factor, double skew, double power,
                         const std::vector<long>& histo, int Inflection, int MaxCopy) {
  const long N = (long)histo.size();
  Shape s(N, SC, stdev, factor, skew, power);
  const long J = s.J;
  std::vector<double> colsum(J + 2, 1.0);
  for (long c = 1; c <= J; c++) colsum[c] = s.column_sum(c, 1);
  auto dist = [&](long i, long c) { return c <= J ? s.raw(i, c) / colsum[c] : s.raw(i, c); };  // last column: never normalised
  auto row_total = [&](long i) {
    double total = 0;
    for (long c = 1; c <= J; c++) total += dist(i, c);
    return total;
  };
  std::vector<double> RC(1, 0.0);
  {  // Diploid branch (ModelDist.cpp:127-140)
    size_t iSC = (size_t)SC, iH = (size_t)(SC / 2);
    double tSC = histo[iSC] / dist(iSC, 2);
    double hetKmers = (histo[iH] - (dist(iH, 2) * tSC)) / dist(iH, 1);
    RC.push_back(hetKmers > 0 ? hetKmers : 0);
    RC.push_back(histo[iSC] / dist(iSC, 2));
  }
  for (long a = 2; a < (double)(size_t)N / SC; a++) {
    size_t r = (size_t)(SC * a);
    double d = dist(r, a + 1);
    double prob = d / row_total(r);
    RC.push_back((double)histo[r] / d * prob);
  }
  double ss = 0;
  for (double i = Inflection; i < SC * MaxCopy; i++) {
    if (i < 0 || i >= (double)N) continue;  // the reference reads out of bounds here
    size_t r = (size_t)i;
    double sum = 0;
    for (long j = 1; j <= J; j++) sum += dist(r, j) * RC[j];
    if (log_residual)
      ss += (pow(log((double)histo[r]) - log((double)sum), 2));
    else
      ss += (pow((double)histo[r] - sum, 2));
  }
  return ss;
}

// FitErrorModel (ModelDist.cpp:333-361): error k-mers fall as 1/x^p; p by a float grid search on log residuals
static float fit_error_model(std::vector<double>& Error, double& Total, int max, std::string& log) {
  double LastSSQ = 0;
  float bestP = 0;
  std::vector<double> logE(max > 1 ? max : 1, 0.0);
  for (int i = 1; i < max; i++) logE[i] = ::log(Error[i]);
  for (int i = 1; i < max; i++) LastSSQ += pow(logE[i] - ::log((1 / (pow((double)i, (double)100))) * Error[1]), 2);
  for (float p = 7; p > .1; p += -.001) {
    double SSQ = 0;
    for (int i = 1; i < max; i++) SSQ += pow(logE[i] - ::log((1 / (pow((double)i, (double)p))) * Error[1]), 2);
    if (SSQ < LastSSQ) {
      LastSSQ = SSQ;
      bestP = p;
    }
  }
  Total = 0;
  for (size_t i = 1; i < Error.size(); i++) {
    Error[i] = (1 / (pow((double)i, (double)bestP))) * Error[1];
    Total += Error[i];
  }
  char b[64];
  snprintf(b, sizeof b, "best error is 1/x^%g\n", (double)bestP);
  log += b;
  return bestP;
}

struct Out {  // text exactly as an ostream with default flags prints it
  std::string s;
  void d(double v) {
    if (v == 0 && !signbit(v)) { s += '0'; return; }
    char b[40];
    s.append(b, snprintf(b, sizeof b, "%g", v));
  }
  void l(long v) {
    char b[24];
    s.append(b, snprintf(b, sizeof b, "%ld", v));
  }
  void t() { s += '\t'; }
  void nl() { s += '\n'; }
  void str(const char* p) { s += p; }
  bool save(const std::string& path) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    bool ok = fwrite(s.data(), 1, s.size(), f) == s.size();
    return fclose(f) == 0 && ok;
  }
};

static std::vector<std::string> split_tab(const std::string& line) {  // Util::Split(line, '\t')
  std::vector<std::string> t;
  size_t a = 0;
  if (line.empty()) return t;
  for (;;) {
    size_t b = line.find('\t', a);
    if (b == std::string::npos) {
      if (a < line.size()) t.push_back(line.substr(a));
      break;
    }
    t.push_back(line.substr(a, b - a));
    a = b + 1;
  }
  return t;
}

static bool next_line(FILE* f, std::string& line) {
  line.clear();
  int c;
  bool any = false;
  while ((c = getc(f)) != EOF) {
    any = true;
    if (c == '\n') return true;
    line += (char)c;
  }
  return any;
}

int main(int argc, char** argv) {
  Out log;
  auto flush_log = [&] { fwrite(log.s.data(), 1, log.s.size(), stdout); fflush(stdout); log.s.clear(); };
  log.str("Call is histoFile HS ReadLength Threads\n");
  if (argc < 5) {
    flush_log();
    fprintf(stderr, "usage: ModelDist histoFile HS ReadLength Threads\n");
    return 1;
  }
  const std::string base = argv[1];
  FILE* hf = fopen(argv[1], "rb");
  if (!hf) {
    log.str("Error, HistoFile could not be opened");
    flush_log();
    return 0;
  }
  log.str("Parent File open - "); log.str(argv[1]); log.nl();
  Out model_f, dist_f, prob_f;
  {  // the reference creates (truncates) .model and .dist before it parses anything
    FILE* a = fopen((base + ".7.7.model").c_str(), "wb");
    FILE* b = a ? fopen((base + ".7.7.dist").c_str(), "wb") : nullptr;
    if (a) fclose(a);
    if (b) fclose(b);
    if (!a || !b) {
      log.str("Error, Model file could not be opened");
      flush_log();
      return 0;
    }
  }
  const int HS = atoi(argv[2]);
  const int ReadLength = atoi(argv[3]);
  int Threads = atoi(argv[4]);
  if (Threads < 1) Threads = 1;
  if (Threads > 11) Threads = 11;

  // ---- histogram (ModelDist.cpp:412-479)
  double SC = 1, SCvalue = -1, stdvalue = -1, value = -1, HistoSum = 0;
  int Inflection = -1;
  std::vector<long> histo(1, 0);
  std::string line;
  next_line(hf, line);
  std::vector<std::string> temp = split_tab(line);
  auto field = [&](size_t k) { return k < temp.size() ? temp[k] : std::string(); };
  log.str("first line = "); log.str(field(0).c_str()); log.str(" - "); log.str(field(1).c_str()); log.nl();
  int tries = 0;
  while (atoi(field(1).c_str()) == 0 || atoi(field(0).c_str()) == 0) {
    log.str("getting another \n");
    next_line(hf, line);
    temp = split_tab(line);
    log.str("got "); log.str(field(0).c_str()); log.str(" - "); log.str(field(1).c_str()); log.nl();
    if (++tries > 10) {
      log.str("ERROR there are no kmers in this file\n");
      flush_log();
      return 1;
    }
  }
  log.str("going with "); log.str(field(0).c_str()); log.str(" - "); log.str(field(1).c_str()); log.nl();
  value = atol(field(1).c_str());
  histo.push_back((long)value);
  long last = (long)value;
  bool PastInflection = false;
  long i = 1, total = 0, TotalKmers = 0;
  while (next_line(hf, line)) {
    i++;
    temp = split_tab(line);
    value = atol(field(1).c_str());
    histo.push_back((long)value);
    total = (long)(total + value);
    TotalKmers = (long)(TotalKmers + value * atoi(field(0).c_str()));
    HistoSum += value;
    if (value - last > 0 && !PastInflection) {
      Inflection = (int)(i - 1);
      PastInflection = true;
    }
    if (PastInflection && SCvalue < value) {
      SCvalue = value;
      SC = i;
    }
    last = histo[i];
  }
  fclose(hf);
  const long N = (long)histo.size();
  int NumberOfReads = (int)(TotalKmers / (ReadLength - HS + 1));
  log.str("Number of reads = "); log.l(NumberOfReads); log.nl();
  for (i = 0; i < 10; i++) {
    log.str("I = "); log.l(i); log.str(" \t "); log.l(i < N ? histo[i] : 0); log.nl();
  }
  log.str("SC = "); log.d(SC); log.str(" vlaue = "); log.d(SCvalue); log.nl();
  const int rawSC = (int)SC;
  if (!((double)N / (SC * 1.1) > 2)) {   // the model needs at least the het, 1-copy and 2-copy columns (the reference reads out of bounds here)
    log.str("ERROR the histogram is too short to fit a model\n");
    flush_log();
    return 1;
  }
  stdvalue = SCvalue * exp(-.5);
  for (i = (long)SC; i < N; i++)
    if (histo[i] - stdvalue < 0) break;
  double stdev = i - SC;
  log.str("stdi = "); log.l(i); log.str(" stdev = "); log.d(stdev); log.nl();

  // ---- error model and the error-corrected histogram (ModelDist.cpp:481-499)
  std::vector<long> histo2(histo);
  std::vector<double> ErrorModel(histo.begin(), histo.end()), ErrorDist;
  double burner;
  fit_error_model(ErrorModel, burner, Inflection, log.s);
  for (long k = 0; k < N; k++) {
    ErrorDist.push_back(ErrorModel[k] / burner);
    histo2[k] = histo[k] - ErrorModel[k] > 0 ? (long)(histo[k] - ErrorModel[k]) : 0;
  }
  flush_log();

  // ---- coordinate search, three passes (ModelDist.cpp:501-717)
  double factor = 1, skew = 0, Power = 1;
  double bestS = stdev, bestF = factor, bestSC = SC, bestSK = skew, bestP = Power;
  enum Which { W_F, W_SC, W_ST, W_P };
  auto candidate = [&](Which w, double v) {
    switch (w) {
      case W_F: return test_model(true, bestSC, bestS, v, bestSK, bestP, histo2, Inflection, 5);
      case W_SC: return test_model(false, v, bestS, bestF, bestSK, bestP, histo2, Inflection, 5);
      case W_ST: return test_model(false, bestSC, v, bestF, bestSK, bestP, histo2, Inflection, 5);
      default: return test_model(true, bestSC, bestS, bestF, bestSK, v, histo2, Inflection, 5);
    }
  };
  // one refinement loop: 11 candidates on [lo, hi], keep the two intervals around the lowest
  auto refine = [&](Which w, double lo, double hi, double stop_ratio, double hi_floor, double lo_floor, double& best) {
    double count = 0;
    while (lo / hi < stop_ratio && hi > hi_floor) {
      count++;
      double values[11];
      const bool flat = (w == W_P && bestSK == 0 && lo > 0);  // pow(0, p) = 0 for every p > 0: eleven equal values
      if (flat) {
        double v = candidate(w, lo);
        for (int x = 0; x <= 10; x++) values[x] = v;
      } else {
        std::vector<std::thread> pool;
        auto work = [&](int t) {
          for (int x = t; x <= 10; x += Threads) values[x] = candidate(w, lo + (((hi - lo) / 10) * x));
        };
        for (int t = 1; t < Threads; t++) pool.emplace_back(work, t);
        work(0);
        for (auto& th : pool) th.join();
      }
      double lowest = values[0];
      int lowestX = 0;
      for (int x = 1; x <= 10; x++)
        if (values[x] < lowest) {
          lowestX = x;
          lowest = values[x];
        }
      if (lo + ((hi - lo) / 10) * (lowestX - 1) >= lo_floor)
        lo = lo + ((hi - lo) / 10) * (lowestX - 1);
      else
        lo = lo_floor;
      hi = lo + ((hi - lo) / 10) * (lowestX + 1);
      best = lo + ((hi - lo) / 10) * (lowestX);
    }
    return count;
  };
  for (int pass = 0; pass <= 2; pass++) {
    log.str("On "); log.l(pass + 1); log.str(" pass\n");
    double steps = refine(W_F, 1, 20, .999, 1e-10, 0, bestF);
    log.str("\t best Factor = "); log.d(bestF); log.str(" steps = "); log.d(steps); log.nl();
    steps = refine(W_SC, SC * .9, SC * 1.1, .999, 1e-50, 0, bestSC);
    log.str("\t\tbestSC = "); log.d(bestSC); log.str(" steps = "); log.d(steps); log.nl();
    steps = refine(W_ST, stdev * .9, stdev * 1.1, .99, 1e-50, 0, bestS);
    log.str("\t\tbest StdDev = "); log.d(bestS); log.str(" steps = "); log.d(steps); log.nl();
    // the skew loop of the reference (`SKlow / SKhigh < .999 and SKhigh < 1e-50` with SKhigh = 2) never runs
    log.str("\t\tbest skew factor = "); log.d(bestSK); log.str(" steps = "); log.d(0); log.nl();
    steps = refine(W_P, 1, 2, .999, 1e-50, 1, bestP);
    log.str("\t\tbest Power factor = "); log.d(bestP); log.str(" steps = "); log.d(steps); log.nl();
    stdev = bestS;
    factor = bestF;
    SC = bestSC;
    skew = bestSK;
    Power = bestP;
    flush_log();
  }
  log.str("Best Model is SC = "); log.d(SC); log.str(" StdDev = "); log.d(stdev); log.str(" F = "); log.d(factor);
  log.str(" skew = "); log.d(skew); log.str(" bestP = "); log.d(Power); log.nl();

  // ---- the final model on the RAW histogram, rows 0 .. N-1 (ModelDist.cpp:719-823)
  Shape s(N, SC, stdev, factor, skew, Power);
  const long J = s.J, W = J + 2;  // a dist row: [0] = 0, [1] het, [2 .. J + 1] copies
  std::vector<double> dist((size_t)N * W, 0.0);
  for (long c = 1; c <= J + 1; c++) {
    long lo = s.banded ? s.col[c].lo : 0, hi = s.banded ? s.col[c].hi : N - 1;
    for (long r = lo; r <= hi; r++) dist[(size_t)r * W + c] = s.raw(r, c);
  }
  for (long c = 1; c <= J; c++) {
    double sum = 0.0;
    for (long r = 0; r < N; r++) sum += dist[(size_t)r * W + c];
    for (long r = 0; r < N; r++) dist[(size_t)r * W + c] = dist[(size_t)r * W + c] / sum;
  }
  // prob row k of the reference = marginal probabilities of histogram row k - 1 (a placeholder sits at index 0),
  // J + 1 entries per row: [0] = 0, [1 .. J]
  const long PW = J + 1;
  std::vector<double> prob((size_t)(N + 1) * PW, 0.0);
  for (long r = 0; r < N; r++) {
    double tot = 0;
    for (long c = 1; c <= J; c++) tot += dist[(size_t)r * W + c];
    for (long c = 1; c <= J; c++) prob[(size_t)(r + 1) * PW + c] = dist[(size_t)r * W + c] / tot;
  }
  std::vector<double> RC(1, 0.0);
  {
    size_t iSC = (size_t)SC, iH = (size_t)(SC / 2);
    double tSC = histo[iSC] / dist[iSC * W + 2];
    double hetKmers = (histo[iH] - (dist[iH * W + 2] * tSC)) / dist[iH * W + 1];
    RC.push_back(hetKmers > 0 ? hetKmers : 0);
    RC.push_back(histo[iSC] / dist[iSC * W + 2]);
  }
  for (long a = 2; a < (double)(size_t)N / SC; a++) {
    size_t r = (size_t)(SC * a);
    double p = a + 1 < PW ? prob[r * PW + (a + 1)] : 0.0;  // one past the row for the last copy number: taken as 0
    RC.push_back((double)histo[r] / dist[r * W + (a + 1)] * p);
  }
  std::vector<double> model((size_t)N * PW, 0.0), ModelSums(N, 0.0);
  for (long r = 0; r < N; r++) {
    double sum = 0;
    for (long c = 1; c <= J; c++) {
      double m = dist[(size_t)r * W + c] * RC[c];
      model[(size_t)r * PW + c] = m;
      sum += m;
    }
    ModelSums[r] = sum;
  }
  double GenomeSize = 0;
  for (size_t k = 1; k < RC.size(); k++) GenomeSize += RC[k] * k;
  log.str("GenomeSize = "); log.d(GenomeSize); log.nl();

  int kcutoff = -1;  // first multiplicity at which a k-mer is more likely genomic than an error
  for (long k = 1; k < N; k++) {
    float num = 0;
    for (long c = 1; c < W; c++) num += dist[(size_t)k * W + c];
    double pne = num / (num + ErrorDist[k]);
    log.str("prob not error = "); log.d(pne); log.nl();
    if (pne > 0.5) {
      kcutoff = (int)k;
      log.str("this one\n");
      break;
    }
  }
  log.str("here1\n");
  for (Out* f : {&model_f, &dist_f, &prob_f}) {
    f->l(3); f->nl(); f->l(kcutoff); f->nl();
  }
  for (Out* f : {&model_f, &dist_f, &prob_f}) { f->d(HistoSum); f->nl(); }
  for (Out* f : {&model_f, &dist_f, &prob_f}) { f->l(rawSC); f->nl(); }
  log.str("here\n");
  flush_log();

  // ---- .model (ModelDist.cpp:880-921): share of each copy number, then the first 5*SC rows
  for (long c = 1; c < PW; c++) {
    long LocalSum = 0;
    for (long k = 1; k < N; k++) LocalSum = (long)(LocalSum + model[(size_t)k * PW + c]);
    model_f.d(((double)LocalSum) / ((double)HistoSum)); model_f.t();
  }
  model_f.nl();
  model_f.str("K\tRawCount\tErrorModel\tContSubtract\tModelSum\t1x\t2x\t3x\t4x\t5x\t6x\t7x\t8x\t9x\n");
  auto model_at = [&](long k, long c) { return c < PW ? model[(size_t)k * PW + c] : 0.0; };
  model_f.str("0\t0\t0\t0\t0");
  for (long c = 1; c < 10; c++) { model_f.t(); model_f.d(model_at(0, c)); }
  model_f.nl();
  for (long k = 1; k < SC * 5 && k < N; k++) {
    model_f.l(k); model_f.t(); model_f.l(histo[k]); model_f.t(); model_f.d(ErrorModel[k]); model_f.t(); model_f.l(0);
    model_f.t(); model_f.d(ModelSums[k]);
    for (long c = 1; c < 10; c++) { model_f.t(); model_f.d(model_at(k, c)); }
    model_f.nl();
  }
  // ---- .dist (ModelDist.cpp:893-947)
  dist_f.d(burner); dist_f.t(); dist_f.l(0); dist_f.t();
  for (size_t k = 1; k < RC.size(); k++) { dist_f.d(RC[k]); dist_f.t(); }
  dist_f.nl();
  dist_f.d(SC); dist_f.nl();
  dist_f.str("0\t0\t0");
  for (long c = 1; c < W; c++) { dist_f.t(); dist_f.d(dist[c]); }
  dist_f.nl();
  for (long k = 1; k < N; k++) {
    dist_f.l(k); dist_f.t(); dist_f.d(ErrorDist[k]); dist_f.t(); dist_f.l(0);
    for (long c = 1; c < W; c++) { dist_f.t(); dist_f.d(dist[(size_t)k * W + c]); }
    dist_f.nl();
  }
  for (long c = 1; c < W; c++) { dist_f.t(); dist_f.d(dist[c]); }
  // ---- .prob (ModelDist.cpp:949-960); its last row reads ErrorDist one past the end: 0
  prob_f.nl(); prob_f.nl();
  for (long k = 1; k < N + 1; k++) {
    prob_f.l(k); prob_f.t(); prob_f.d(k < N ? ErrorDist[k] : 0.0); prob_f.t(); prob_f.l(0);
    for (long c = 1; c < PW; c++) { prob_f.t(); prob_f.d(prob[(size_t)k * PW + c]); }
    prob_f.nl();
  }
  if (!model_f.save(base + ".7.7.model") || !dist_f.save(base + ".7.7.dist")) {
    log.str("Error, Model file could not be opened");
    flush_log();
    return 0;
  }
  if (!prob_f.save(base + ".7.7.prob")) {
    log.str("Error, Prob file could not be opened");
    flush_log();
    return 0;
  }
  log.str("GenomeSize = "); log.d(GenomeSize); log.nl();
  log.str("Inflection point = "); log.l(Inflection); log.nl();
  log.str("Recomended RUFUS cutoff = "); log.d(SC - (5 * stdev)); log.nl();
  log.str("-1std = "); log.d(SC - (1 * stdev)); log.str("\t-2std = "); log.d(SC - (2 * stdev));
  log.str("\t-3std = "); log.d(SC - (3 * stdev)); log.str("\t-4std = "); log.d(SC - (4 * stdev)); log.nl();
  flush_log();
  return 0;
}
