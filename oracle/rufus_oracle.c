/* rufus_oracle.c — see rufus_oracle.h. TEST INFRASTRUCTURE ONLY. */
#include "rufus_oracle.h"
#include <stdlib.h>
#include <string.h>

/* mer_dna.hpp:46-63 — A/a 0, C/c 1, G/g 2, T/t 3, everything else negative. */
static inline int jf_code(unsigned char c) {
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return -1;
  }
}

static void radix_sort_u64(uint64_t* a, uint64_t* tmp, uint64_t n, int bits) {
  const int RB = 11;
  const uint64_t NB = 1u << RB;
  uint64_t* cnt = (uint64_t*)malloc(sizeof(uint64_t) * NB);
  uint64_t *src = a, *dst = tmp;
  int passes = 0;
  for (int sh = 0; sh < bits; sh += RB) {
    memset(cnt, 0, sizeof(uint64_t) * NB);
    for (uint64_t i = 0; i < n; i++) cnt[(src[i] >> sh) & (NB - 1)]++;
    uint64_t s = 0;
    for (uint64_t b = 0; b < NB; b++) { uint64_t c = cnt[b]; cnt[b] = s; s += c; }
    for (uint64_t i = 0; i < n; i++) dst[cnt[(src[i] >> sh) & (NB - 1)]++] = src[i];
    uint64_t* t = src; src = dst; dst = t;
    passes++;
  }
  if (passes & 1) memcpy(a, tmp, n * sizeof(uint64_t));
  free(cnt);
}

int ro_count(const char* seq, const uint64_t* offs, uint64_t n_reads, int k, int canonical,
             uint64_t lower, ro_table* out) {
  out->keys = NULL; out->counts = NULL; out->n = 0;
  if (k < 1 || k > 32) return -1;
  uint64_t cap = 0;
  for (uint64_t r = 0; r < n_reads; r++) {
    uint64_t len = offs[r + 1] - offs[r];
    if (len >= (uint64_t)k) cap += len - k + 1;
  }
  uint64_t* mers = (uint64_t*)malloc(sizeof(uint64_t) * (cap ? cap : 1));
  uint64_t* tmp = (uint64_t*)malloc(sizeof(uint64_t) * (cap ? cap : 1));
  if (!mers || !tmp) { free(mers); free(tmp); return -2; }
  const uint64_t mask = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1);
  const int hs = 2 * (k - 1);
  uint64_t n = 0;
  for (uint64_t r = 0; r < n_reads; r++) {
    /* reads are separated by 'N' in the parser (mer_overlap_sequence_parser.hpp:195):
     * the window restarts at every read boundary. */
    uint64_t m = 0, rc = 0; int filled = 0;
    for (uint64_t p = offs[r]; p < offs[r + 1]; p++) {
      int c = jf_code((unsigned char)seq[p]);
      if (c < 0) { filled = 0; continue; }
      m = ((m << 2) | (uint64_t)c) & mask;               /* shift_left, mer_dna.hpp:330-354 */
      rc = (rc >> 2) | ((uint64_t)(3 - c) << hs);        /* shift_right of the complement */
      if (filled < k) filled++;
      if (filled >= k) mers[n++] = (!canonical || m < rc) ? m : rc;   /* mer_iterator.hpp:59 */
    }
  }
  radix_sort_u64(mers, tmp, n, 2 * k);
  free(tmp);
  /* run-length encode */
  uint64_t d = 0;
  for (uint64_t i = 0; i < n;) {
    uint64_t j = i; while (j < n && mers[j] == mers[i]) j++;
    if (j - i >= lower) d++;
    i = j;
  }
  out->keys = (uint64_t*)malloc(sizeof(uint64_t) * (d ? d : 1));
  out->counts = (uint32_t*)malloc(sizeof(uint32_t) * (d ? d : 1));
  d = 0;
  for (uint64_t i = 0; i < n;) {
    uint64_t j = i; while (j < n && mers[j] == mers[i]) j++;
    uint64_t c = j - i;
    if (c >= lower) { out->keys[d] = mers[i]; out->counts[d] = c > 0xffffffffull ? 0xffffffffu : (uint32_t)c; d++; }
    i = j;
  }
  out->n = d;
  free(mers);
  return 0;
}

void ro_table_free(ro_table* t) { free(t->keys); free(t->counts); t->keys = NULL; t->counts = NULL; t->n = 0; }

uint64_t ro_histo_buckets(uint64_t low, uint64_t high) {
  const uint64_t inc = 1;
  uint64_t base = inc >= low ? 0 : low - inc, ceil = high + inc;
  return (ceil + inc - base) / inc;
}

void ro_histo(const ro_table* t, uint64_t low, uint64_t high, uint64_t* out, uint64_t* base_out) {
  const uint64_t inc = 1;
  uint64_t base = inc >= low ? 0 : low - inc, ceil = high + inc;
  uint64_t nb = (ceil + inc - base) / inc;
  memset(out, 0, nb * sizeof(uint64_t));
  for (uint64_t i = 0; i < t->n; i++) {
    uint64_t v = t->counts[i];
    if (v < base) out[0]++;
    else if (v > ceil) out[nb - 1]++;
    else out[(v - base) / inc]++;
  }
  if (base_out) *base_out = base;
}

uint64_t ro_merge_unique(const ro_table* const* tabs, int n_tabs, uint64_t floor,
                         uint64_t* out_keys, uint32_t* out_counts) {
  uint64_t* pos = (uint64_t*)calloc(n_tabs, sizeof(uint64_t));
  uint64_t n_out = 0;
  for (;;) {
    int have = 0; uint64_t key = 0;
    for (int i = 0; i < n_tabs; i++)
      if (pos[i] < tabs[i]->n && (!have || tabs[i]->keys[pos[i]] < key)) { key = tabs[i]->keys[pos[i]]; have = 1; }
    if (!have) break;
    int present = 0; uint32_t val = 0;
    for (int i = 0; i < n_tabs; i++)
      if (pos[i] < tabs[i]->n && tabs[i]->keys[pos[i]] == key) { present++; val = tabs[i]->counts[pos[i]]; pos[i]++; }
    if (present == 1 && val >= floor) { out_keys[n_out] = key; out_counts[n_out] = val; n_out++; }  /* merge_files.cc:147-152 */
  }
  free(pos);
  return n_out;
}

uint32_t ro_query(const ro_table* t, uint64_t key) {
  uint64_t lo = 0, hi = t->n;
  while (lo < hi) { uint64_t mid = (lo + hi) / 2; if (t->keys[mid] < key) lo = mid + 1; else hi = mid; }
  return (lo < t->n && t->keys[lo] == key) ? t->counts[lo] : 0;
}

uint64_t ro_hashlist(const ro_table* subj, const ro_table* const* others, int n_others, uint64_t floor,
                     int64_t min_cov, int64_t max_cov, uint64_t* out_keys, uint32_t* out_counts) {
  const ro_table** all = (const ro_table**)malloc(sizeof(void*) * (n_others + 1));
  uint64_t tot = subj->n;
  all[0] = subj;
  for (int i = 0; i < n_others; i++) { all[i + 1] = others[i]; tot += others[i]->n; }
  uint64_t* mk = (uint64_t*)malloc(sizeof(uint64_t) * (tot ? tot : 1));
  uint32_t* mc = (uint32_t*)malloc(sizeof(uint32_t) * (tot ? tot : 1));
  uint64_t nm = ro_merge_unique(all, n_others + 1, floor, mk, mc), n_out = 0;
  for (uint64_t i = 0; i < nm; i++) {
    int64_t c = (int64_t)ro_query(subj, mk[i]);      /* query returns the SUBJECT count, 0 if absent */
    if (c >= min_cov && c <= max_cov) { out_keys[n_out] = mk[i]; out_counts[n_out] = (uint32_t)c; n_out++; }
  }
  free(mk); free(mc); free(all);
  return n_out;
}

uint64_t ro_rufus_pack(const char* s, int len) {
  uint64_t v = 0;
  for (int i = 0; i < len && i < 32; i++) {
    uint64_t c;
    switch (s[i]) { case 'C': c = 2; break; case 'G': c = 1; break; case 'T': c = 3; break; default: c = 0; }
    v |= c << (2 * i);
  }
  return v;
}

int ro_rufus_revcomp(const char* s, int len, char* out) {
  int n = 0;
  for (int i = len - 1; i >= 0; i--) {
    switch (s[i]) {
      case 'A': out[n++] = 'T'; break; case 'C': out[n++] = 'G'; break;
      case 'G': out[n++] = 'C'; break; case 'T': out[n++] = 'A'; break;
      case 'N': out[n++] = 'N'; break; default: break;
    }
  }
  return n;
}

static int in_set(const uint64_t* a, uint64_t n, uint64_t v) {
  uint64_t lo = 0, hi = n;
  while (lo < hi) { uint64_t mid = (lo + hi) / 2; if (a[mid] < v) lo = mid + 1; else hi = mid; }
  return lo < n && a[lo] == v;
}

int ro_filter_hits(const char* seq, const char* qual, int len, int k, int minq, int skip_last,
                   const uint64_t* set_sorted, uint64_t n_set) {
  int hits = 0, streak = 0;
  int end = skip_last ? len - 1 : len;
  for (int i = 0; i < end; i++) {
    if (((int)qual[i] - 33) < minq || seq[i] == 'N') streak = 0; else streak++;
    if (streak >= k && in_set(set_sorted, n_set, ro_rufus_pack(seq + i - k + 1, k))) hits++;
  }
  return hits;
}

void ro_filter_pairs(const char* s1, const char* q1, const uint64_t* o1,
                     const char* s2, const char* q2, const uint64_t* o2, uint64_t n_pairs,
                     int k, int minq, int thresh, const uint64_t* set_sorted, uint64_t n_set,
                     uint8_t* keep) {
#pragma omp parallel for schedule(dynamic, 1024)
  for (uint64_t p = 0; p < n_pairs; p++) {
    int h = ro_filter_hits(s1 + o1[p], q1 + o1[p], (int)(o1[p + 1] - o1[p]), k, minq, 1, set_sorted, n_set);
    if (h < thresh)
      h = ro_filter_hits(s2 + o2[p], q2 + o2[p], (int)(o2[p + 1] - o2[p]), k, minq, 1, set_sorted, n_set);
    keep[p] = h >= thresh;
  }
}

void ro_filter_single(const char* s, const char* q, const uint64_t* o, uint64_t n_reads,
                      int k, int minq, const uint64_t* set_sorted, uint64_t n_set, int32_t* hits) {
#pragma omp parallel for schedule(dynamic, 1024)
  for (uint64_t r = 0; r < n_reads; r++)
    hits[r] = ro_filter_hits(s + o[r], q + o[r], (int)(o[r + 1] - o[r]), k, minq, 0, set_sorted, n_set);
}


/* ---------------------------------------------------------------------------
 * Sliced counting for inputs too large to count whole on the CPU (SURVEY §8d "parity at full scale via
 * hash-prefix slices"): the same k-mer extraction as ro_count (mer_iterator.hpp:59-89), but only the keys with
 * slice_hash(key) % mod == rem are kept, accumulated over any number of ro_slice_add calls in an
 * open-addressed map. The slice hash is a fixed function of the key (independent of any table geometry), so
 * the CUDA path's dump restricted to the same predicate must hold exactly these (key, count) pairs.
 * --------------------------------------------------------------------------- */
struct ro_slice {
  int k, canonical;
  uint32_t mod, rem;
  uint64_t* keys;
  uint64_t* counts;
  uint64_t cap, n;       /* cap: power of two */
};

uint32_t ro_slice_hash(uint64_t key) {
  uint32_t h = (uint32_t)key * 0xC2B2AE35u + (uint32_t)(key >> 32) * 0x27D4EB2Fu;
  h ^= h >> 15;
  return h * 0x2C1B3C6Du;
}

static void slice_put(ro_slice* s, uint64_t key, uint64_t add) {
  uint64_t i = (key * 0x9E3779B97F4A7C15ull) >> 20 & (s->cap - 1);
  for (;;) {
    if (s->counts[i] == 0) { s->keys[i] = key; s->counts[i] = add; s->n++; return; }
    if (s->keys[i] == key) { s->counts[i] += add; return; }
    i = (i + 1) & (s->cap - 1);
  }
}

static void slice_grow(ro_slice* s) {
  uint64_t ocap = s->cap, *ok = s->keys, *oc = s->counts;
  s->cap = ocap * 2;
  s->keys = (uint64_t*)calloc(s->cap, sizeof(uint64_t));
  s->counts = (uint64_t*)calloc(s->cap, sizeof(uint64_t));
  s->n = 0;
  for (uint64_t i = 0; i < ocap; i++) if (oc[i]) slice_put(s, ok[i], oc[i]);
  free(ok); free(oc);
}

ro_slice* ro_slice_new(int k, int canonical, uint32_t mod, uint32_t rem) {
  if (k < 1 || k > 32 || mod == 0 || rem >= mod) return NULL;
  ro_slice* s = (ro_slice*)calloc(1, sizeof(ro_slice));
  s->k = k; s->canonical = canonical; s->mod = mod; s->rem = rem;
  s->cap = 1u << 16;
  s->keys = (uint64_t*)calloc(s->cap, sizeof(uint64_t));
  s->counts = (uint64_t*)calloc(s->cap, sizeof(uint64_t));
  return s;
}

/* n_reads reads of read_len bytes each, back to back */
void ro_slice_add(ro_slice* s, const char* seq, uint64_t n_reads, uint32_t read_len) {
  const int k = s->k;
  const uint64_t mask = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1);
  const int hs = 2 * (k - 1);
  for (uint64_t r = 0; r < n_reads; r++) {
    uint64_t m = 0, rc = 0; int filled = 0;
    const char* p = seq + r * (uint64_t)read_len;
    for (uint32_t j = 0; j < read_len; j++) {
      int c = jf_code((unsigned char)p[j]);
      if (c < 0) { filled = 0; continue; }
      m = ((m << 2) | (uint64_t)c) & mask;
      rc = (rc >> 2) | ((uint64_t)(3 - c) << hs);
      if (filled < k) filled++;
      if (filled >= k) {
        const uint64_t key = (!s->canonical || m < rc) ? m : rc;
        if (ro_slice_hash(key) % s->mod == s->rem) {
          if (s->n * 2 >= s->cap) slice_grow(s);
          slice_put(s, key, 1);
        }
      }
    }
  }
}

static int cmp_u64(const void* a, const void* b) {
  const uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
  return x < y ? -1 : x > y;
}

/* entries with count >= lower, ascending key, counts saturated at 2^32-1 */
int ro_slice_finish(ro_slice* s, uint64_t lower, ro_table* out) {
  uint64_t d = 0;
  for (uint64_t i = 0; i < s->cap; i++) if (s->counts[i] >= lower && s->counts[i]) d++;
  uint64_t* pairs = (uint64_t*)malloc(sizeof(uint64_t) * 2 * (d ? d : 1));
  if (!pairs) return -2;
  uint64_t q = 0;
  for (uint64_t i = 0; i < s->cap; i++)
    if (s->counts[i] >= lower && s->counts[i]) { pairs[2 * q] = s->keys[i]; pairs[2 * q + 1] = s->counts[i]; q++; }
  qsort(pairs, d, 2 * sizeof(uint64_t), cmp_u64);
  out->keys = (uint64_t*)malloc(sizeof(uint64_t) * (d ? d : 1));
  out->counts = (uint32_t*)malloc(sizeof(uint32_t) * (d ? d : 1));
  out->n = d;
  for (uint64_t i = 0; i < d; i++) {
    out->keys[i] = pairs[2 * i];
    out->counts[i] = pairs[2 * i + 1] > 0xffffffffull ? 0xffffffffu : (uint32_t)pairs[2 * i + 1];
  }
  free(pairs);
  return 0;
}

void ro_slice_free(ro_slice* s) {
  if (!s) return;
  free(s->keys); free(s->counts); free(s);
}
