// tools/jellyfish.cpp — drop-in for the `jellyfish` binary RUFUS calls (runRufus.sh:748-754,
// RunJellyForRUFUS.sh:29-37, CheckJellyHashList.sh:12, Overlap.shorter.sh:245-255), with the work
// done on the GPU through librufus_gpu.so. Same argv, same files:
//
//   jellyfish count [--disk] -m K [-C] [-L min] [-U max] -s SIZE [-t T] [-o OUT] IN...
//        IN: FASTA/FASTQ files, FIFOs or /dev/stdin. OUT: a .Jhash byte-compatible with
//        Jellyfish 2.2.5 binary/sorted (count_main.cc:214-354). An extension: IN may also be a BAM file, which
//        is decoded in process on T threads and counted as `samtools view -F 3328 IN | PassThroughSamCheck`
//        would present it (tools/rg_bam_input.hpp)
//        --disk (hash_counter.hpp:182-202, count_main.cc:316-338): a table that would outgrow the device is written
//        out as the intermediate file OUT<N> and counting goes on with an empty table; the parts are merged (counts
//        summed, then -L/-U) and unlinked at the end (tools/rg_jf_parts.hpp). RG_COUNT_SPILL_DISTINCT=<n> sets the
//        number of distinct k-mers a table may hold (default: device memory / 128 bytes); --no-merge / --no-unlink
//        keep the parts, as they do in Jellyfish.
//   jellyfish histo [-f] [-l low] [-h high] [-o OUT] DB          (histo_main.cc:46-110)
//   jellyfish dump  [-c] [-t] [-L min] [-U max] [-o OUT] DB      (dump_main.cc)
//   jellyfish query [-s FASTA]... [-o OUT] DB [MER...]           (query_main.cc:44-129)
//   jellyfish merge DB DB...     the MODIFIED merge of src/modifiedJellyfish/jellyfish/merge_files.cc:86-154:
//        prints `KMER<TAB>COUNT` for the k-mers present in exactly one input with count >= 5
// Options RUFUS never passes are accepted where they are harmless and rejected otherwise.
#include <algorithm>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>

#include <sys/stat.h>

#include "rg_bam_input.hpp"
#include "rg_cli.hpp"
#include "rg_jf_parts.hpp"

[[noreturn]] void rgparts::fail(const std::string& m) { die("jellyfish count: " + m); }

[[noreturn]] void rgbam::fail(const std::string& m) {   // (raised on a pipeline thread: leave without unwinding the others)
  fprintf(stderr, "jellyfish count: BAM input: %s\n", m.c_str());
  fflush(stderr);
  _exit(1);
}

static const char* next_arg(int& i, int argc, char** argv, const char* opt) {
  if (i + 1 >= argc) die(std::string("Missing argument for ") + opt);
  return argv[++i];
}
// "-m25" / "-m 25" / "--mer-len=25" / "--mer-len 25"
static bool opt_value(const char* a, const char* shortf, const char* longf, int& i, int argc, char** argv,
                      const char** val) {
  const size_t ls = strlen(shortf), ll = strlen(longf);
  if (strncmp(a, longf, ll) == 0 && ll > 0) {
    if (a[ll] == '=') { *val = a + ll + 1; return true; }
    if (a[ll] == 0) { *val = next_arg(i, argc, argv, longf); return true; }
    return false;
  }
  if (ls > 0 && strncmp(a, shortf, ls) == 0 && a[1] != '-') {
    if (a[ls] == 0) { *val = next_arg(i, argc, argv, shortf); return true; }
    *val = a + ls;
    return true;
  }
  return false;
}

static FILE* open_out(const char* path) {
  if (!path) return stdout;
  FILE* f = fopen(path, "w");
  if (!f) die(std::string("Error opening output file '") + path + "'");
  return f;
}

// ---------------------------------------------------------------------------- histogram sidecar
// `jellyfish count` is always followed by `jellyfish histo -f` on the file it wrote (RunJellyForRUFUS.sh:29,37),
// a second process that would create a CUDA context and re-load the whole .Jhash for 10 002 numbers the counting
// process can produce from the table it already holds (histo_kernel). count therefore leaves `<out>.rg_histo`
// next to the .Jhash: magic, the .Jhash's size and mtime (a stale sidecar is ignored), then the rows of
// `histo -l 1 -h 10000`; histo uses it when it is asked for exactly that and the .Jhash is the one it describes.
static const uint64_t RG_HISTO_MAGIC = 0x3130747369486752ull;   // "RgHist01"
static bool file_id(const char* path, uint64_t id[2]) {
  struct stat st;
  if (stat(path, &st) != 0) return false;
  id[0] = (uint64_t)st.st_size;
  id[1] = (uint64_t)st.st_mtim.tv_sec * 1000000000ull + (uint64_t)st.st_mtim.tv_nsec;
  return true;
}
static void write_histo_sidecar(const char* jhash, const std::vector<uint64_t>& rows) {
  uint64_t id[2];
  if (!file_id(jhash, id)) return;
  const std::string p = std::string(jhash) + ".rg_histo";
  FILE* f = fopen(p.c_str(), "wb");
  if (!f) return;                                    // (a convenience: never an error)
  const uint64_t head[4] = {RG_HISTO_MAGIC, id[0], id[1], (uint64_t)rows.size()};
  const bool ok = fwrite(head, 8, 4, f) == 4 && fwrite(rows.data(), 8, rows.size(), f) == rows.size();
  fclose(f);
  if (!ok) remove(p.c_str());
}
static bool read_histo_sidecar(const char* jhash, std::vector<uint64_t>& rows) {
  uint64_t id[2], head[4];
  if (getenv("RG_NO_HISTO_SIDECAR") || !file_id(jhash, id)) return false;
  FILE* f = fopen((std::string(jhash) + ".rg_histo").c_str(), "rb");
  if (!f) return false;
  bool ok = fread(head, 8, 4, f) == 4 && head[0] == RG_HISTO_MAGIC && head[1] == id[0] && head[2] == id[1] &&
            head[3] == rg_histo_rows(1, 10000);
  if (ok) {
    rows.resize(head[3]);
    ok = fread(rows.data(), 8, rows.size(), f) == rows.size();
  }
  fclose(f);
  return ok;
}

// ---------------------------------------------------------------------------- count on several GPUs
// The devices `jellyfish count` uses: RG_DEVICES=0,1,... or "all"; default: the first GPU. One host thread parses
// the input (~1 GB/s of FASTQ), which cannot even feed one GPU, so more GPUs buy CAPACITY (a table too large for one
// device: 180 GB hold ~8 G distinct k-mers), not speed, and cost start-up time (measured on an 8-GPU box, C2-sized
// input: 16 s for three counts with 8 contexts + 2 GB of pinned buffers against 5 s with one GPU): opt-in.
static std::vector<int> count_devices() {
  const int n = rg_device_count();
  if (n <= 0) die("jellyfish count: no CUDA device (librufus_gpu has no CPU fallback)");
  std::vector<int> d;
  const char* e = getenv("RG_DEVICES");
  if (e && *e && strcmp(e, "all") != 0) {
    for (const char* p = e; *p;) {
      char* end = nullptr;
      const long v = strtol(p, &end, 10);
      if (end == p || v < 0 || v >= n) die(std::string("RG_DEVICES: bad device list '") + e + "'");
      d.push_back((int)v);
      p = *end == ',' ? end + 1 : end;
      if (*end != ',' && *end != 0) die(std::string("RG_DEVICES: bad device list '") + e + "'");
    }
  } else if (e && *e) {
    for (int i = 0; i < n && i < 8; ++i) d.push_back(i);
  }
  if (d.empty()) d.push_back(0);
  return d;
}

struct HostBarrier {
  std::mutex m;
  std::condition_variable cv;
  int n, waiting = 0, phase = 0;
  explicit HostBarrier(int n_) : n(n_) {}
  void wait() {
    std::unique_lock<std::mutex> lk(m);
    const int ph = phase;
    if (++waiting == n) {
      waiting = 0;
      ++phase;
      cv.notify_all();
    } else {
      cv.wait(lk, [&] { return phase != ph; });
    }
  }
};

// One process, one context + one hash-prefix shard per GPU (RunJellyForRUFUS.sh:29 with every GPU of the box
// behind it). The reader deals the sequence chunks round-robin; a ROUND = one chunk per GPU, counted in three
// lock-step phases by one host thread per GPU: (1) make room + zero my slab cursors, (2) upload my chunk and
// scatter its super-k-mer records straight into the owners' slabs over NVLink (peer stores, no collective),
// (3) my slabs -> tiles -> inserts. The next round is parsed while this one is counted. One merged .Jhash in
// (pos, key) order at the end.
struct MultiCount {
  std::vector<int> dev;
  std::vector<rg_ctx*> ctx;
  std::vector<rg_table*> tab;
  std::vector<char*> stage;              // device staging of one chunk per GPU
  size_t chunk;
  std::vector<ByteBuf> bufs[2];          // two rounds of pinned chunks: one being filled, one being counted
  int filling = 0, n_filled = 0;
  std::thread round;
  std::string err;

  MultiCount(const std::vector<int>& devices, int k, int canonical, size_t chunk_bytes) : dev(devices), chunk(chunk_bytes) {
    const int n = (int)dev.size();
    ctx.resize(n);
    tab.resize(n);
    stage.resize(n);
    const uint32_t groups = (uint32_t)std::max(8, 128 / n);
    for (int i = 0; i < n; ++i) {
      ctx[i] = rg_init(dev[i]);
      if (!ctx[i]) die(rg_last_error());
      tab[i] = rg_count_begin(ctx[i], k, canonical, 1ull << 22, RG_MODE_TILED);
      if (!tab[i]) die(rg_last_error());
      check(rg_table_set_shard(tab[i], (uint32_t)i, (uint32_t)n), "count");
      check(rg_table_p2p_prepare(tab[i], groups, chunk + chunk / 4 + (1u << 20)), "count");
      stage[i] = (char*)rg_dev_alloc(ctx[i], chunk + (16u << 20));
      if (!stage[i]) die(rg_last_error());
    }
    check(rg_tables_p2p_attach(tab.data(), n), "count");
    for (int r = 0; r < 2; ++r) {
      bufs[r] = std::vector<ByteBuf>(n);
      for (auto& b : bufs[r]) b.reserve(chunk + (8u << 20));
    }
  }
  void run_round(int which, int n_chunks) {
    const int n = (int)dev.size();
    size_t total = 0;
    for (int i = 0; i < n_chunks; ++i) total += bufs[which][i].size();
    HostBarrier bar(n);
    std::vector<std::thread> th;
    std::mutex em;
    auto fail = [&](const char* what) {
      std::lock_guard<std::mutex> g(em);
      if (err.empty()) err = std::string(what) + ": " + rg_last_error();
    };
    for (int i = 0; i < n; ++i)
      th.emplace_back([&, i] {
        ByteBuf& b = bufs[which][i];
        const size_t bytes = i < n_chunks ? b.size() : 0;
        // (1) room for my share of this round's k-mers, whatever their minimizers are; zero my cursors
        if (rg_table_reserve(tab[i], total / n + total / (2 * n) + 1024) || rg_count_p2p_begin(tab[i]) || rg_sync(ctx[i]))
          fail("count (reserve)");
        bar.wait();
        // (2) upload + scatter into every owner's slabs
        if (bytes && rg_h2d(ctx[i], stage[i], b.data(), bytes)) fail("count (upload)");
        if (rg_count_p2p_scatter_dev(tab[i], stage[i], bytes, 0, 0) || rg_sync(ctx[i])) fail("count (scatter)");
        bar.wait();
        // (3) my slabs -> tiles -> inserts; spills resolved
        if (rg_count_p2p_finish(tab[i]) || rg_count_end(tab[i])) fail("count (insert)");
        b.clear();
      });
    for (auto& t : th) t.join();
  }
  // the reader's flush: hand over one chunk; a full round starts counting in the background
  void add_chunk(ByteBuf& b) {
    bufs[filling][n_filled].swap(b);
    b.clear();
    if (++n_filled == (int)dev.size()) launch();
  }
  void launch() {
    if (round.joinable()) round.join();
    if (!err.empty()) die(err);
    const int which = filling, cnt = n_filled;
    round = std::thread([this, which, cnt] { run_round(which, cnt); });
    filling ^= 1;
    n_filled = 0;
  }
  void finish() {
    if (n_filled) launch();
    if (round.joinable()) round.join();
    if (!err.empty()) die(err);
  }
  ~MultiCount() {
    for (size_t i = 0; i < tab.size(); ++i) {
      rg_table_free(tab[i]);
      rg_dev_free(ctx[i], stage[i]);
      rg_shutdown(ctx[i]);
    }
  }
};

// ---------------------------------------------------------------------------- count
static int count_main(int argc, char** argv, int full_argc, char** full_argv) {
  int k = 0, canonical = 0;
  unsigned threads = 1;
  bool no_merge = false, no_unlink = false;
  uint64_t size = 0, lower = 0, upper = 0;
  uint32_t out_counter_len = 4;
  const char* out = "mer_counts.jf";
  std::vector<const char*> inputs;
  for (int i = 1; i < argc; ++i) {
    const char* a = argv[i];
    const char* v = nullptr;
    if (!strcmp(a, "-C") || !strcmp(a, "--canonical")) canonical = 1;
    else if (!strcmp(a, "--disk") || !strcmp(a, "--text") || !strcmp(a, "--no-write")) { if (a[2] == 't') die("--text output is not supported"); }
    else if (!strcmp(a, "--no-merge")) no_merge = true;
    else if (!strcmp(a, "--no-unlink")) no_unlink = true;
    else if (opt_value(a, "-m", "--mer-len", i, argc, argv, &v)) k = atoi(v);
    else if (opt_value(a, "-s", "--size", i, argc, argv, &v)) size = parse_size(v);
    else if (opt_value(a, "-t", "--threads", i, argc, argv, &v)) threads = (unsigned)std::max(1, std::min(64, atoi(v)));   // BAM inputs decode on this many threads
    else if (opt_value(a, "-o", "--output", i, argc, argv, &v)) out = v;
    else if (opt_value(a, "-L", "--lower-count", i, argc, argv, &v)) lower = parse_size(v);
    else if (opt_value(a, "-U", "--upper-count", i, argc, argv, &v)) upper = parse_size(v);
    else if (opt_value(a, "-c", "--counter-len", i, argc, argv, &v)) {}
    else if (opt_value(a, "-p", "--reprobes", i, argc, argv, &v)) { if (atoi(v) != 126) die("only the default --reprobes 126 is supported"); }
    else if (opt_value(a, "", "--out-counter-len", i, argc, argv, &v)) out_counter_len = (uint32_t)atoi(v);
    else if (opt_value(a, "-F", "--Files", i, argc, argv, &v) || opt_value(a, "-G", "--Generators", i, argc, argv, &v) ||
             opt_value(a, "-g", "--generator", i, argc, argv, &v) || opt_value(a, "-Q", "--min-qual-char", i, argc, argv, &v) ||
             opt_value(a, "", "--bf-size", i, argc, argv, &v) || opt_value(a, "", "--bc", i, argc, argv, &v) ||
             opt_value(a, "", "--if", i, argc, argv, &v))
      die(std::string("option not supported by the GPU count: ") + a);
    else if (a[0] == '-' && a[1] != 0) die(std::string("unknown option ") + a);
    else inputs.push_back(a);
  }
  if (k < 1 || k > 32 || (!canonical && k == 32)) die("mer length must be 1..32 (1..31 without -C)");
  if (size == 0) die("Missing -s SIZE");
  if (inputs.empty()) die("Missing input file");

  const std::vector<int> devices = count_devices();
  if (devices.size() > 1) {
    const char* cm = getenv("RG_CHUNK_MB");
    const size_t chunk = (size_t)(cm ? std::max(1, atoi(cm)) : 64) << 20;
    MultiCount mc(devices, k, canonical, chunk);
    ByteBuf fill;
    fill.reserve(chunk + (8u << 20));
    auto flush = [&](ByteBuf& b) {
      mc.add_chunk(b);
      b.reserve(chunk + (8u << 20));
    };
    for (const char* path : inputs) {
      FILE* f = fopen(path, "rb");
      if (!f) die(std::string("Failed to open input file '") + path + "'");
      read_input(f, path, fill, chunk, flush, threads);
      fclose(f);
    }
    if (!fill.empty()) flush(fill);
    mc.finish();
    check(rg_tables_write_jhash(mc.tab.data(), (int)mc.tab.size(), out, lower, upper, size, out_counter_len,
                                full_argv + 1, full_argc - 1), "count");
    if (upper == 0) {
      std::vector<uint64_t> rows(rg_histo_rows(1, 10000), 0), part(rows.size());
      for (rg_table* t : mc.tab) {
        check(rg_histo(t, lower ? lower : 1, 1, 10000, part.data()), "count");
        for (size_t i = 0; i < rows.size(); ++i) rows[i] += part[i];
      }
      write_histo_sidecar(out, rows);
    }
    return 0;
  }
  rg_ctx* ctx = rg_init(devices[0]);
  if (!ctx) die(rg_last_error());
  // the table grows with the input (rg_table_reserve); -s only fixes the geometry written to the header
  rg_table* t = rg_count_begin(ctx, k, canonical, 1ull << 22, RG_MODE_TILED);
  if (!t) die(rg_last_error());
  // two pinned buffers: the reader fills one while the other is copied and counted. The chunk starts small (a
  // small input should not pay for 2 x 128 MiB of pinned memory) and grows to 128 MiB.
  size_t chunk = 8u << 20;
  ByteBuf fill, inflight;
  fill.reserve(chunk + (1u << 20));
  // --disk (hash_counter.hpp:182-202): a table about to hold more than `spill_limit` distinct k-mers is written out as
  // the intermediate file <out><N> (every count, no -L/-U: they apply to the summed counts) and emptied; counting goes
  // on. The limit leaves room for the table (at most 60 % loaded and grown by half: up to 30 bytes per k-mer), the copy
  // it is re-hashed into while it grows, and the 48 bytes per record the .Jhash writer needs while it sorts: device
  // memory / 128 bytes (1.4 G distinct k-mers per part on 180 GB; a 30x human sample leaves ~9 parts).
  // No k-mer is looked at before the bytes pushed since the table was last empty could exceed the limit, so an input
  // that fits — everything below 1.4 Gbases, whatever it holds — runs exactly the path it ran without this.
  uint64_t spill_limit = 0, since_empty = 0;
  if (const char* e = getenv("RG_COUNT_SPILL_DISTINCT")) {
    spill_limit = parse_size(e);
  } else {
    uint64_t fr = 0, tot = 0;
    if (rg_device_memory(ctx, &fr, &tot) == 0) spill_limit = tot / 128;
  }
  std::vector<std::string> parts;
  auto spill = [&]() {
    if (out_counter_len != 4) die("jellyfish count: --out-counter-len with intermediate files is not supported");
    const std::string part = std::string(out) + std::to_string(parts.size());   // (count_main.cc: <output><N>)
    check(rg_table_write_jhash(t, part.c_str(), 0, 0, size, 4, full_argv + 1, full_argc - 1), "count (intermediate file)");
    check(rg_table_clear(t), "count");
    parts.push_back(part);
    since_empty = 0;
  };
  auto flush = [&](ByteBuf& b) {
    check(rg_count_end(t), "count");                 // the previous chunk (in `inflight`) is done; spills resolved
    if (spill_limit && since_empty && since_empty + b.size() > spill_limit) {
      uint64_t distinct = 0;
      check(rg_table_stats(t, &distinct, nullptr, nullptr), "count");
      if (distinct + b.size() > spill_limit) spill();
      else since_empty = distinct;                   // (what the table holds: a tighter bound than the bytes pushed)
    }
    check(rg_table_reserve(t, b.size()), "count");   // room for this chunk's k-mers, whatever they are
    inflight.swap(b);
    b.clear();
    check(rg_count_push(t, inflight.data(), inflight.size()), "count");   // asynchronous
    since_empty += inflight.size();
    chunk = std::min<size_t>(chunk * 4, 128u << 20);
    b.reserve(chunk + (1u << 20));
  };
  for (const char* path : inputs) {
    FILE* f = fopen(path, "rb");
    if (!f) die(std::string("Failed to open input file '") + path + "'");
    read_input(f, path, fill, chunk, flush, threads);
    fclose(f);
  }
  if (!fill.empty()) flush(fill);
  check(rg_count_end(t), "count");
  if (!parts.empty()) {
    // count_main.cc:322-337: the table's last state is one more part; the stock merge sums the parts, then -L/-U
    if (since_empty) spill();
    rg_table_free(t);
    rg_shutdown(ctx);
    if (no_merge) return 0;
    std::vector<uint64_t> rows(upper == 0 ? rg_histo_rows(1, 10000) : 0, 0);   // (the sidecar, as after an unspilled count)
    rgparts::merge(parts, out, lower, upper, &rows);
    if (upper == 0) write_histo_sidecar(out, rows);
    if (!no_unlink)
      for (const std::string& p : parts) unlink(p.c_str());
    return 0;
  }
  check(rg_table_write_jhash(t, out, lower, upper, size, out_counter_len, full_argv + 1, full_argc - 1), "count");
  if (upper == 0) {
    std::vector<uint64_t> rows(rg_histo_rows(1, 10000));
    check(rg_histo(t, lower ? lower : 1, 1, 10000, rows.data()), "count");
    write_histo_sidecar(out, rows);
  }
  rg_table_free(t);
  rg_shutdown(ctx);
  return 0;
}

// `jellyfish rg-merge-parts [-L min] [-U max] -o OUT PART...`: the merge `count --disk` runs over its intermediate files,
// on its own (parts kept with --no-merge; no GPU involved)
static int merge_parts_main(int argc, char** argv) {
  uint64_t lower = 0, upper = 0;
  const char* out = nullptr;
  std::vector<std::string> parts;
  for (int i = 1; i < argc; ++i) {
    const char* a = argv[i];
    const char* v = nullptr;
    if (opt_value(a, "-o", "--output", i, argc, argv, &v)) out = v;
    else if (opt_value(a, "-L", "--lower-count", i, argc, argv, &v)) lower = parse_size(v);
    else if (opt_value(a, "-U", "--upper-count", i, argc, argv, &v)) upper = parse_size(v);
    else if (a[0] == '-' && a[1] != 0) die(std::string("unknown option ") + a);
    else parts.push_back(a);
  }
  if (!out || parts.empty()) die("Usage: jellyfish rg-merge-parts [-L min] [-U max] -o OUT PART...");
  const uint64_t n = rgparts::merge(parts, out, lower, upper);
  fprintf(stderr, "%llu records\n", (unsigned long long)n);
  return 0;
}

// ---------------------------------------------------------------------------- helpers over a DB
struct Db {
  rg_ctx* ctx = nullptr;
  rg_table* t = nullptr;
  rg_jf_info info;
};
static Db open_db(rg_ctx* ctx, const char* path) {
  Db d;
  d.ctx = ctx;
  d.t = rg_table_read_jhash(ctx, path, &d.info);
  if (!d.t) die(rg_last_error());
  return d;
}

// ---------------------------------------------------------------------------- histo
static int histo_main(int argc, char** argv) {
  uint64_t low = 1, high = 10000, inc = 1;
  bool full = false;
  const char* out = nullptr;
  const char* db = nullptr;
  for (int i = 1; i < argc; ++i) {
    const char* a = argv[i];
    const char* v = nullptr;
    if (!strcmp(a, "-f") || !strcmp(a, "--full")) full = true;
    else if (!strcmp(a, "-v") || !strcmp(a, "--verbose")) {}
    else if (opt_value(a, "-l", "--low", i, argc, argv, &v)) low = parse_size(v);
    else if (opt_value(a, "-h", "--high", i, argc, argv, &v)) high = parse_size(v);
    else if (opt_value(a, "-i", "--increment", i, argc, argv, &v)) inc = parse_size(v);
    else if (opt_value(a, "-t", "--threads", i, argc, argv, &v)) {}
    else if (opt_value(a, "-o", "--output", i, argc, argv, &v)) out = v;
    else if (a[0] == '-' && a[1] != 0) die(std::string("unknown option ") + a);
    else db = a;
  }
  if (!db) die("Missing database file");
  if (high < low) die("High count value must be >= to low count value");
  if (inc != 1) die("only --increment 1 is supported");
  const uint64_t rows = rg_histo_rows(low, high);
  std::vector<uint64_t> h(rows);
  rg_ctx* ctx = nullptr;
  Db d;
  // the histogram the counting process left next to the file it wrote (computed on the GPU then), if it is about this file
  if (!(low == 1 && high == 10000 && read_histo_sidecar(db, h))) {
    ctx = rg_init(0);
    if (!ctx) die(rg_last_error());
    d = open_db(ctx, db);
    check(rg_histo(d.t, 1, low, high, h.data()), "histo");
  }
  FILE* f = open_out(out);
  const uint64_t base = 1 >= low ? 0 : low - 1;     // histo_main.cc:64
  for (uint64_t i = 0; i < rows; ++i)
    if (h[i] > 0 || full) fprintf(f, "%llu %llu\n", (unsigned long long)(base + i), (unsigned long long)h[i]);
  if (f != stdout) fclose(f);
  if (ctx) {
    rg_table_free(d.t);
    rg_shutdown(ctx);
  }
  return 0;
}

// records of a table in the file order of its .Jhash
static void ordered_records(Db& d, uint64_t lo, uint64_t hi, std::vector<uint64_t>& keys, std::vector<uint32_t>& counts) {
  uint64_t n = 0;
  check(rg_table_stats(d.t, &n, nullptr, nullptr), "stats");
  keys.resize(n ? n : 1);
  counts.resize(n ? n : 1);
  uint64_t need = 0;
  const int64_t got = rg_dump(d.t, lo, hi, keys.data(), counts.data(), keys.size(), &need);
  check(got, "dump");
  keys.resize((size_t)got);
  counts.resize((size_t)got);
  check(rg_jf_sort(d.ctx, keys.data(), counts.data(), keys.size(), d.info.size, d.info.k), "sort");
}

// ---------------------------------------------------------------------------- dump
static int dump_main(int argc, char** argv) {
  bool column = false, tab = false;
  uint64_t lo = 1, hi = 0xFFFFFFFFull;
  const char* out = nullptr;
  const char* db = nullptr;
  for (int i = 1; i < argc; ++i) {
    const char* a = argv[i];
    const char* v = nullptr;
    if (!strcmp(a, "-c") || !strcmp(a, "--column")) column = true;
    else if (!strcmp(a, "-t") || !strcmp(a, "--tab")) tab = true;
    else if (!strcmp(a, "-ct") || !strcmp(a, "-tc")) column = tab = true;
    else if (opt_value(a, "-L", "--lower-count", i, argc, argv, &v)) lo = parse_size(v);
    else if (opt_value(a, "-U", "--upper-count", i, argc, argv, &v)) hi = parse_size(v);
    else if (opt_value(a, "-o", "--output", i, argc, argv, &v)) out = v;
    else if (a[0] == '-' && a[1] != 0) die(std::string("unknown option ") + a);
    else db = a;
  }
  if (!db) die("Missing database file");
  rg_ctx* ctx = rg_init(0);
  if (!ctx) die(rg_last_error());
  Db d = open_db(ctx, db);
  std::vector<uint64_t> keys;
  std::vector<uint32_t> counts;
  ordered_records(d, lo ? lo : 1, hi, keys, counts);
  FILE* f = open_out(out);
  for (size_t i = 0; i < keys.size(); ++i) {
    const std::string m = kmer_text(keys[i], d.info.k);
    if (column) fprintf(f, "%s%c%u\n", m.c_str(), tab ? '\t' : ' ', counts[i]);
    else fprintf(f, ">%u\n%s\n", counts[i], m.c_str());
  }
  if (f != stdout) fclose(f);
  rg_table_free(d.t);
  rg_shutdown(ctx);
  return 0;
}

// ---------------------------------------------------------------------------- query
// the k-mers of a sequence stream in order (mer_iterator.hpp:59-89), canonical if asked
template <typename Bytes>
static void kmers_of(const Bytes& seq, int k, bool canonical, std::vector<uint64_t>& out) {
  const uint64_t mask = k == 32 ? ~0ull : ((1ull << (2 * k)) - 1);
  uint64_t fwd = 0, rc = 0;
  int filled = 0;
  for (char ch : seq) {
    int c;
    switch (ch) {
      case 'A': case 'a': c = 0; break;
      case 'C': case 'c': c = 1; break;
      case 'G': case 'g': c = 2; break;
      case 'T': case 't': c = 3; break;
      default: c = -1;
    }
    if (c < 0) { filled = 0; fwd = rc = 0; continue; }
    fwd = ((fwd << 2) | (uint64_t)c) & mask;
    rc = (rc >> 2) | ((uint64_t)(3 - c) << (2 * (k - 1)));
    if (++filled >= k) out.push_back(canonical && rc < fwd ? rc : fwd);
  }
}

static int query_main(int argc, char** argv) {
  std::vector<const char*> seq_files, pos;
  const char* out = nullptr;
  for (int i = 1; i < argc; ++i) {
    const char* a = argv[i];
    const char* v = nullptr;
    if (opt_value(a, "-s", "--sequence", i, argc, argv, &v)) seq_files.push_back(v);
    else if (opt_value(a, "-o", "--output", i, argc, argv, &v)) out = v;
    else if (!strcmp(a, "-l") || !strcmp(a, "--load") || !strcmp(a, "-L") || !strcmp(a, "--no-load")) {}
    else if (!strcmp(a, "-i") || !strcmp(a, "--interactive")) die("--interactive is not supported");
    else if (a[0] == '-' && a[1] != 0) die(std::string("unknown option ") + a);
    else pos.push_back(a);
  }
  if (pos.empty()) die("Missing database file");
  rg_ctx* ctx = rg_init(0);
  if (!ctx) die(rg_last_error());
  Db d = open_db(ctx, pos[0]);
  const int k = d.info.k;
  FILE* f = open_out(out);
  auto answer = [&](const std::vector<uint64_t>& keys) {
    std::vector<uint32_t> counts(keys.size() ? keys.size() : 1);
    check(rg_query(d.t, keys.data(), keys.size(), 1, counts.data()), "query");
    for (size_t i = 0; i < keys.size(); ++i) fprintf(f, "%s %u\n", kmer_text(keys[i], k).c_str(), counts[i]);
  };
  for (const char* path : seq_files) {
    FILE* in = fopen(path, "rb");
    if (!in) die(std::string("Failed to open input file '") + path + "'");
    HeapBuf seq;
    std::vector<uint64_t> keys;
    read_sequences(in, path, seq, 64u << 20, [&](HeapBuf& b) {
      keys.clear();
      kmers_of(b, k, d.info.canonical, keys);
      answer(keys);
      b.clear();
    });
    fclose(in);
    keys.clear();
    kmers_of(seq, k, d.info.canonical, keys);
    answer(keys);
  }
  std::vector<uint64_t> keys;
  for (size_t i = 1; i < pos.size(); ++i) {
    std::vector<char> m(pos[i], pos[i] + strlen(pos[i]));
    std::vector<uint64_t> one;
    if ((int)m.size() == k) kmers_of(m, k, d.info.canonical, one);
    if (one.size() != 1) fprintf(stderr, "Invalid mer '%s'\n", pos[i]);
    else keys.push_back(one[0]);
  }
  answer(keys);
  if (f != stdout) fclose(f);
  rg_table_free(d.t);
  rg_shutdown(ctx);
  return 0;
}

// ---------------------------------------------------------------------------- merge (modified)
static int merge_main(int argc, char** argv) {
  std::vector<const char*> files;
  for (int i = 1; i < argc; ++i) {
    const char* a = argv[i];
    const char* v = nullptr;
    if (opt_value(a, "-o", "--output", i, argc, argv, &v) || opt_value(a, "-L", "--lower-count", i, argc, argv, &v) ||
        opt_value(a, "-U", "--upper-count", i, argc, argv, &v)) {}   // the modified merge ignores them (merge_files.cc:68-155)
    else if (a[0] == '-' && a[1] != 0) die(std::string("unknown option ") + a);
    else files.push_back(a);
  }
  if (files.size() < 2) die("Missing input files (at least 2)");
  rg_ctx* ctx = rg_init(0);
  if (!ctx) die(rg_last_error());
  std::vector<Db> dbs;
  for (const char* p : files) dbs.push_back(open_db(ctx, p));
  for (size_t i = 1; i < dbs.size(); ++i) {          // merge_files.cc:193-202
    const rg_jf_info &a = dbs[0].info, &b = dbs[i].info;
    if (a.k != b.k) die(std::string("Can't merge hashes of different key lengths (") + files[i] + ")");
    if (a.max_reprobe_offset != b.max_reprobe_offset) die(std::string("Can't merge hashes with different reprobing strategies (") + files[i] + ")");
    if (a.size != b.size) die(std::string("Can't merge hash with different size (") + files[i] + ")");
    if (memcmp(a.columns, b.columns, sizeof(a.columns)) != 0 || a.matrix_r != b.matrix_r)
      die(std::string("Can't merge hash with different hash function (") + files[i] + ")");
  }
  std::vector<rg_table*> tabs;
  std::vector<uint64_t> mins(dbs.size(), 1);
  for (Db& d : dbs) tabs.push_back(d.t);
  uint64_t cap = 1 << 20, need = 0;
  std::vector<uint64_t> keys;
  std::vector<uint32_t> counts;
  int64_t n;
  for (;;) {
    keys.resize(cap);
    counts.resize(cap);
    n = rg_merge_unique(tabs.data(), mins.data(), (int)tabs.size(), 5, keys.data(), counts.data(), nullptr, cap, &need);
    if (n == RG_ERR_SMALL) { cap = need + 1024; continue; }
    check(n, "merge");
    break;
  }
  check(rg_jf_sort(ctx, keys.data(), counts.data(), (uint64_t)n, dbs[0].info.size, dbs[0].info.k), "sort");
  for (int64_t i = 0; i < n; ++i) printf("%s\t%u\n", kmer_text(keys[(size_t)i], dbs[0].info.k).c_str(), counts[(size_t)i]);
  for (Db& d : dbs) rg_table_free(d.t);
  rg_shutdown(ctx);
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 2) die("Usage: jellyfish <count|histo|dump|query|merge> [options] (GPU build: rufus_b200)");
  const std::string cmd = argv[1];
  if (cmd == "count") return count_main(argc - 1, argv + 1, argc, argv);
  if (cmd == "histo") return histo_main(argc - 1, argv + 1);
  if (cmd == "dump") return dump_main(argc - 1, argv + 1);
  if (cmd == "query") return query_main(argc - 1, argv + 1);
  if (cmd == "merge") return merge_main(argc - 1, argv + 1);
  if (cmd == "rg-merge-parts") return merge_parts_main(argc - 1, argv + 1);
  if (cmd == "--version" || cmd == "-V") { printf("jellyfish 2.2.5 (%s)\n", rg_version()); return 0; }
  die("Unknown command '" + cmd + "'");
}
