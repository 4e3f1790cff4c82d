/* rufus_oracle.h — CPU restatement of the reference hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is linked, imported or
 * executed by the product (rufus_b200/, include/, tools/). Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may use it, and only as the checker.
 *
 * Parity pinning: this restatement is checked (tests/test_oracle.py) against
 * golden outputs produced by the reference's own binaries (oracle/_ref, built
 * from /root/reference sources by oracle/Makefile; generator script
 * tests/golden/make_golden.py) on the bundled testRun trio, two synthetic trios
 * (k=25, k=31) and a hand-made edge-case trio.
 *
 * All file:line citations are relative to /root/reference/.
 */
#ifndef RUFUS_ORACLE_H
#define RUFUS_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  uint64_t* keys;    /* Jellyfish packing: first base in the most significant 2 bits, A0 C1 G2 T3 */
  uint32_t* counts;  /* saturated at 2^32-1 like binary_dumper.hpp:44-48 */
  uint64_t  n;       /* entries, ascending key */
} ro_table;

/* jellyfish count -m k [-C] -L lower  (count_main.cc:154-157, mer_iterator.hpp:59-89,
 * mer_dna.hpp:46-63; lower bound applied at dump time, sorted_dumper.hpp:94).
 * seq = concatenated read bytes, read r = seq[offs[r] .. offs[r+1]). */
int  ro_count(const char* seq, const uint64_t* offs, uint64_t n_reads, int k, int canonical,
              uint64_t lower, ro_table* out);
void ro_table_free(ro_table* t);

/* jellyfish histo -l low -h high -i 1 (histo_main.cc:34-44,64-84). out must hold
 * ro_histo_buckets(low,high) entries; *base receives the count of row 0. */
uint64_t ro_histo_buckets(uint64_t low, uint64_t high);
void ro_histo(const ro_table* t, uint64_t low, uint64_t high, uint64_t* out, uint64_t* base);

/* modified `jellyfish merge` (src/modifiedJellyfish/jellyfish/merge_files.cc:86-154):
 * emit (key,count) for every key present in EXACTLY ONE input whose count there
 * is >= floor (5 in the reference). Output ascending key. Returns n emitted;
 * out arrays must hold sum(n_i) entries. */
uint64_t ro_merge_unique(const ro_table* const* tabs, int n_tabs, uint64_t floor,
                         uint64_t* out_keys, uint32_t* out_counts);

/* jellyfish query against one table (binary_dumper.hpp:156-197): 0 if absent. */
uint32_t ro_query(const ro_table* t, uint64_t key);

/* CheckJellyHashList.sh:12 = merge_unique | query subject | awk >=min | awk <=max. */
uint64_t ro_hashlist(const ro_table* subj, const ro_table* const* others, int n_others, uint64_t floor,
                     int64_t min_cov, int64_t max_cov, uint64_t* out_keys, uint32_t* out_counts);

/* Util::HashToLong (src/Util.cpp:51-84): base i at bits (2i,2i+1); A=0 C=2 G=1 T=3,
 * any other byte leaves the bits 0 (== 'A'). len <= 32. */
uint64_t ro_rufus_pack(const char* s, int len);
/* Util::RevComp (src/Util.cpp:187-210): ACGTN complemented, other bytes DROPPED. Returns new length. */
int ro_rufus_revcomp(const char* s, int len, char* out);

/* RUFUS.Filter per-read hit count (src/RUFUS.Filter.cpp:203-221; single-end
 * src/RUFUS.Filter.ss.cpp:176): set = sorted RUFUS-packed values (k-mer and its
 * RevComp, RUFUS.Filter.cpp:121-143). skip_last=1 reproduces the paired tool's
 * `i < length()-1` bound. */
int ro_filter_hits(const char* seq, const char* qual, int len, int k, int minq, int skip_last,
                   const uint64_t* set_sorted, uint64_t n_set);
/* keep[i] = hits(m1)>=thresh || hits(m2)>=thresh (RUFUS.Filter.cpp:222,237-262). */
void ro_filter_pairs(const char* s1, const char* q1, const uint64_t* o1,
                     const char* s2, const char* q2, const uint64_t* o2, uint64_t n_pairs,
                     int k, int minq, int thresh, const uint64_t* set_sorted, uint64_t n_set,
                     uint8_t* keep);
/* single-end: hits[i] per read, last base included. */
void ro_filter_single(const char* s, const char* q, const uint64_t* o, uint64_t n_reads,
                      int k, int minq, const uint64_t* set_sorted, uint64_t n_set, int32_t* hits);

/* Sliced counting for inputs too large to count whole on the CPU: the k-mer extraction of ro_count,
 * keeping only the keys with ro_slice_hash(key) % mod == rem, accumulated over any number of ro_slice_add calls
 * (fixed-length reads, back to back). ro_slice_finish: count >= lower, ascending key. */
typedef struct ro_slice ro_slice;
uint32_t ro_slice_hash(uint64_t key);
ro_slice* ro_slice_new(int k, int canonical, uint32_t mod, uint32_t rem);
void ro_slice_add(ro_slice* s, const char* seq, uint64_t n_reads, uint32_t read_len);
int ro_slice_finish(ro_slice* s, uint64_t lower, ro_table* out);
void ro_slice_free(ro_slice* s);

#ifdef __cplusplus
}
#endif
#endif
