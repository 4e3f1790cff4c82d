"""TEST INFRASTRUCTURE — a recording stand-in for librufus_gpu.so, built into a temporary directory at test time and
injected with LD_LIBRARY_PATH into ONE subprocess: it lets the `-m "not gpu"` suite run the drop-in executables' own host
logic (argument parsing, input readers, chunking, batch building, output writing) without a GPU. It is NOT a CPU
implementation of anything: every entry point of include/rufus_gpu.h returns 0/NULL-ish success without computing; the
few that matter here RECORD their arguments —
  rg_count_push     appends the pushed bytes to $RG_RECORD_DIR/pushes.bin
  rg_filter_pairs   folds sequences, qualities and offsets of both mates into a running digest (printed on stderr) and
                    answers a fixed pattern: keep every pair whose index in the call is not 1 mod 3
  rg_filter_single  same, hits = index mod 4
  rg_table_write_jhash  prints its arguments and (with $RG_RECORD_PARTS) copies the next canned part file to the path
Nothing under rufus_b200/ or tools/ knows about it (tests/test_layout.py), and the real library is never replaced in
the tree."""
import os
import re
import subprocess

from conftest import ROOT

SPECIAL = r'''
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
extern "C" {
static unsigned long long dg = 1469598103934665603ull;
static void mix(const void* p, unsigned long long n) {
  const unsigned char* c = (const unsigned char*)p;
  for (unsigned long long i = 0; i < n; ++i) dg = (dg ^ c[i]) * 1099511628211ull;
}
void* rg_init(int) { return (void*)0x1000; }
int rg_device_count() { return 1; }
const char* rg_last_error() { return "recorder"; }
void* rg_host_alloc(uint64_t n) { return malloc(n ? n : 1); }
void rg_host_free(void* p) { free(p); }
void* rg_count_begin(void*, int, int, uint64_t, int) { return (void*)0x2000; }
uint64_t rg_histo_rows(uint64_t low, uint64_t high) { return (high + 1) + 1 - (low > 1 ? low - 1 : 0); }   // (histo_main.cc:64-68)
int rg_count_push(void*, const char* s, uint64_t n) {
  char path[4096];
  snprintf(path, sizeof path, "%s/pushes.bin", getenv("RG_RECORD_DIR"));
  FILE* f = fopen(path, "ab");
  fwrite(s, 1, n, f);
  fclose(f);
  fprintf(stderr, "push %llu\n", (unsigned long long)n);
  return 0;
}
// the --disk path of `jellyfish count`: the table "holds" $RG_RECORD_DISTINCT k-mers, and every .Jhash it is asked to write
// is a copy of the next canned file $RG_RECORD_PARTS/part<N> (made by the reference's binary in the test)
int rg_device_memory(void*, uint64_t* fr, uint64_t* tot) {
  if (fr) *fr = 1ull << 36;
  if (tot) *tot = 1ull << 37;
  return 0;
}
int rg_table_stats(void*, uint64_t* d, uint64_t* o, uint64_t* s) {
  const char* e = getenv("RG_RECORD_DISTINCT");
  if (d) *d = e ? strtoull(e, 0, 10) : 0;
  if (o) *o = 0;
  if (s) *s = 0;
  return 0;
}
int rg_table_clear(void*) {
  fprintf(stderr, "clear\n");
  return 0;
}
long rg_table_write_jhash(void*, const char* path, uint64_t lo, uint64_t hi, uint64_t size, uint32_t cl, const char* const*, int) {
  static int call = 0;
  fprintf(stderr, "write_jhash %s min=%llu max=%llu size=%llu counter_len=%u\n", path, (unsigned long long)lo,
          (unsigned long long)hi, (unsigned long long)size, cl);
  if (const char* dir = getenv("RG_RECORD_PARTS")) {
    char src[4096];
    snprintf(src, sizeof src, "%s/part%d", dir, call);
    FILE* a = fopen(src, "rb");
    FILE* b = fopen(path, "wb");
    if (!a || !b) return -4;
    char buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, a)) > 0) fwrite(buf, 1, n, b);
    fclose(a);
    fclose(b);
  }
  ++call;
  return 0;
}
void* rg_set_from_hashlist_text(void*, const char*, uint64_t, int, int) { return (void*)0x3000; }
int rg_filter_pairs(void*, const char* s1, const char* q1, const uint64_t* o1, const char* s2, const char* q2, const uint64_t* o2,
                    uint64_t n, int minq, int thresh, uint8_t* keep) {
  mix(s1, o1[n]); mix(q1, o1[n]); mix(o1, (n + 1) * 8); mix(s2, o2[n]); mix(q2, o2[n]); mix(o2, (n + 1) * 8);
  for (uint64_t i = 0; i < n; ++i) keep[i] = (i % 3) != 1;
  fprintf(stderr, "filter_pairs n=%llu minq=%d thresh=%d digest=%016llx\n", (unsigned long long)n, minq, thresh, dg);
  return 0;
}
int rg_filter_single(void*, const char* s1, const char* q1, const uint64_t* o1, uint64_t n, int minq, int32_t* hits) {
  mix(s1, o1[n]); mix(q1, o1[n]); mix(o1, (n + 1) * 8);
  for (uint64_t i = 0; i < n; ++i) hits[i] = (int)(i % 4);
  fprintf(stderr, "filter_single n=%llu minq=%d digest=%016llx\n", (unsigned long long)n, minq, dg);
  return 0;
}
'''


def build(tmp_dir):
    """compiles the recorder as <tmp_dir>/librufus_gpu.so; returns the environment that injects it"""
    header = open(os.path.join(ROOT, "include", "rufus_gpu.h")).read()
    names = sorted(set(re.findall(r"\b(rg_[a-z0-9_]+)\s*\(", header)))
    special = set(re.findall(r"\b(rg_[a-z0-9_]+)\s*\(", SPECIAL))
    src = SPECIAL + "".join(f"long {n}() {{ return 0; }}\n" for n in names if n not in special) + "}\n"
    cpp = os.path.join(tmp_dir, "recorder.cpp")
    open(cpp, "w").write(src)
    subprocess.run(["g++", "-O1", "-shared", "-fPIC", cpp, "-o", os.path.join(tmp_dir, "librufus_gpu.so")], check=True)
    os.makedirs(os.path.join(tmp_dir, "rec"), exist_ok=True)
    return dict(os.environ, LD_LIBRARY_PATH=str(tmp_dir), RG_RECORD_DIR=os.path.join(str(tmp_dir), "rec"))


def pushes(env):
    p = os.path.join(env["RG_RECORD_DIR"], "pushes.bin")
    data = open(p, "rb").read() if os.path.exists(p) else b""
    if os.path.exists(p):
        os.remove(p)
    return data
