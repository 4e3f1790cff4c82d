/* rufus_gpu.h — C ABI of librufus_gpu.so: the B200 (sm_100a) implementation of
 * RUFUS's k-mer count -> subtract -> read-pull hot path.
 *
 * The reference has no linked API for this path: its boundary is the argv +
 * file contract of the binaries bound at runRufus.sh:748-759.
 * Each entry point below replaces the inner work of one of those binaries; the
 * CLI shims under tools/ re-create their argv on top of this ABI, and
 * INTEGRATION.md shows how runRufus.sh is pointed at them.
 *
 * Conventions: plain pointers and sizes only. Functions returning int give 0 on
 * success and a negative code on failure; functions returning a pointer give
 * NULL on failure; rg_last_error() (thread-local) holds the message. The caller
 * owns every host buffer; the library owns device memory behind the opaque
 * handles. There is NO CPU fallback: without a CUDA device rg_init fails.
 *
 * K-mer encodings (citations relative to /root/reference/):
 *  - "jf key"   : Jellyfish packing, first base in the most-significant 2 bits of
 *                 a 2k-bit integer, A0 C1 G2 T3 (mer_dna.hpp:46-63,330-354);
 *                 canonical = min(forward, reverse complement) (mer_iterator.hpp:59).
 *  - "rufus key": Util::HashToLong packing, base i at bits (2i,2i+1), A0 G1 C2 T3,
 *                 any other byte == A (src/Util.cpp:51-84). Used by the read filter only.
 */
#ifndef RUFUS_GPU_H
#define RUFUS_GPU_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rg_ctx rg_ctx;
typedef struct rg_table rg_table;
typedef struct rg_set rg_set;

/* insert strategies for rg_count_begin */
#define RG_MODE_AUTO 0
#define RG_MODE_DIRECT 1 /* global-memory CAS/atomic insert straight from the extraction kernel */
#define RG_MODE_TILED 2  /* radix partition by hash prefix -> shared-memory tile tables (TMA load/store) */

#define RG_ERR_ARG (-1)
#define RG_ERR_CUDA (-2)
#define RG_ERR_FULL (-3)
#define RG_ERR_IO (-4)
#define RG_ERR_SMALL (-5) /* caller's output buffer too small; needed size is reported */

const char* rg_last_error(void);
const char* rg_version(void);

/* ---- context: one per GPU / process ---------------------------------- */
rg_ctx* rg_init(int device);
void rg_shutdown(rg_ctx* ctx);
/* Run on the caller's CUDA stream (a cudaStream_t passed as void*), e.g. torch's
 * current stream, so the caller's events time our kernels and its other work (NCCL
 * collectives, reductions) is ordered with ours. NULL = the library's own stream; pass
 * cudaStreamLegacy ((void*)0x1) for the legacy default stream. */
int rg_set_stream(rg_ctx* ctx, void* cuda_stream);
/* rg_set_stream without waiting for the work enqueued on the previous stream (the caller orders the
 * streams with events): one context can feed two streams that run concurrently. */
int rg_use_stream(rg_ctx* ctx, void* cuda_stream);
int rg_sync(rg_ctx* ctx);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t rg_launch_count(rg_ctx* ctx);

/* Per-kernel device timing (CUDA events on the launch stream around each of our
 * kernels). Off by default. rg_profile_get: total ms and launch count of the
 * kernel called `name` since the last reset; returns 0 if known. */
int rg_profile_enable(rg_ctx* ctx, int on);
int rg_profile_reset(rg_ctx* ctx);
int rg_profile_get(rg_ctx* ctx, const char* name, double* total_ms, uint64_t* n_launches);
/* names of all profiled kernels, '\n'-separated, into buf (truncated to cap) */
int rg_profile_names(rg_ctx* ctx, char* buf, uint64_t cap);

/* plain device-memory helpers for callers that keep inputs resident in HBM */
void* rg_dev_alloc(rg_ctx* ctx, uint64_t bytes);
void rg_dev_free(rg_ctx* ctx, void* dptr);
/* pinned host memory: host buffers allocated here make the H2D copies of the host entry points
 * asynchronous (they overlap the kernels); any other host memory works too, only slower */
void* rg_host_alloc(uint64_t bytes);
void rg_host_free(void* p);
int rg_h2d(rg_ctx* ctx, void* dst_dev, const void* src_host, uint64_t bytes);
int rg_d2h(rg_ctx* ctx, void* dst_host, const void* src_dev, uint64_t bytes);

/* ---- K2/K3: k-mer extraction + counting -------------------------------
 * replaces `jellyfish count -m k [-C] -s est ...` (count_main.cc:148-157,214-354;
 * mer_iterator.hpp:61-89; large_hash_array.hpp:298-302,513-601). Counting is per
 * occurrence; any byte outside ACGTacgt restarts the window. */
rg_table* rg_count_begin(rg_ctx* ctx, int k, int canonical, uint64_t est_distinct, int mode);
/* HOST stream of sequence bytes, reads delimited by any non-ACGT byte (e.g. the
 * sequence lines of a FASTQ joined by '\n' — what mer_overlap_sequence_parser.hpp:179-207
 * builds with 'N' separators). */
int rg_count_push(rg_table* t, const char* host_stream, uint64_t n_bytes);
/* HOST concatenated reads + offsets (n_reads+1 entries); no separators needed. */
int rg_count_push_reads(rg_table* t, const char* host_seq, const uint64_t* host_offs, uint64_t n_reads);
/* DEVICE-resident stream (16-byte aligned). read_stride > 0: a read starts every
 * read_stride bytes (fixed-length reads without separators); 0: separators only. */
int rg_count_push_dev(rg_table* t, const char* dev_stream, uint64_t n_bytes, uint32_t read_stride);
/* Make room for up to extra_distinct more distinct keys (grows and re-hashes the table if the
 * load would pass ~60 %): for streaming callers that learn the input size as they read it. */
int rg_table_reserve(rg_table* t, uint64_t extra_distinct);
/* Finish counting: wait for the enqueued pushes, check them (RG_ERR_FULL if k-mers found no slot), grow the
 * table when its overflow region is filling up. */
int rg_count_end(rg_table* t);
void rg_table_free(rg_table* t);
/* Empty the table for re-use (keeps its allocation and geometry). */
int rg_table_clear(rg_table* t);
/* Release the counting-time staging buffers (bucket slabs, H2D staging); the table stays usable. */
int rg_table_trim(rg_table* t);
/* n_distinct (count>=1), total occurrences, slots allocated */
int rg_table_stats(rg_table* t, uint64_t* n_distinct, uint64_t* n_occurrences, uint64_t* n_slots);
int rg_table_k(rg_table* t);
int rg_table_canonical(rg_table* t);

/* Multi-GPU building blocks (hash-prefix sharding). The hash prefix of a k-mer is the mixed
 * order value of its MINIMIZER (the smallest, under a strand-symmetric hash, of its k-m+1 m-mers): the
 * consecutive k-mers of a read that share it travel to their shard and tile as one 16-byte record.
 * rg_extract_partition_dev: extract + partition the k-mers of a read stream into the table's shard_parts
 * contiguous buckets (one per owning shard), for an NCCL all-to-all between ranks; the received packed
 * keys are inserted with rg_count_push_keys_dev. All DEVICE pointers. part_offsets_dev: n_parts+1 uint64
 * written by the library. keys_out_dev needs room for one uint64 per window (n_bytes). */
int rg_extract_partition_dev(rg_table* t, const char* dev_stream, uint64_t n_bytes, uint32_t read_stride,
                             uint64_t* keys_out_dev, uint64_t* part_offsets_dev);
/* which shard (0..n_parts-1) owns a jf key of a table with `groups` hash-prefix groups over all shards and
 * minimizer length m — the function the kernels use (host arithmetic, no device needed) */
uint32_t rg_key_shard(uint64_t jf_key, int k, int m, uint32_t groups, uint32_t n_parts);
/* minimizer length the library picks for k-mers of length k in a table of n_tiles_global tiles */
int rg_minimizer_len(int k, uint64_t n_tiles_global);
/* geometry of a table: tiles of this shard, hash-prefix groups over all shards, minimizer length, slots of
 * the overflow region (any pointer may be NULL) */
int rg_table_geometry(rg_table* t, uint32_t* n_tiles, uint32_t* groups_global, int* m, uint64_t* ovf_slots);
int rg_count_push_keys_dev(rg_table* t, const uint64_t* dev_keys, uint64_t n_keys);
/* restrict a table to the keys of shard `part` of `n_parts` (tile numbering is global) */
int rg_table_set_shard(rg_table* t, uint32_t part, uint32_t n_parts);

/* Peer-to-peer multi-GPU counting (one process per GPU, NVLink peer memory; no collective in the
 * data path): every GPU's partition pass sorts its reads into the hash-prefix groups of ALL shards
 * and writes each run straight into the owner's slab; then every GPU finishes its own groups.
 *   setup (once, after rg_table_set_shard): rg_table_p2p_prepare on every rank, all-gather the
 *     64-byte handles of rg_table_p2p_export, rg_table_p2p_import each of them;
 *   per sample: rg_count_p2p_begin -> barrier -> rg_count_p2p_scatter_dev -> barrier ->
 *     rg_count_p2p_finish (the barriers are the caller's, e.g. a tiny all-reduce on the stream).
 * groups_per_shard x n_parts must fit one partition pass (<= ~2900); max_keys_per_push bounds the
 * keys a shard receives per sample (sizes the exported slabs; they cannot grow). */
int rg_table_p2p_prepare(rg_table* t, uint32_t groups_per_shard, uint64_t max_keys_per_push);
int rg_table_p2p_export(rg_table* t, void* handle64_out);
int rg_table_p2p_import(rg_table* t, uint32_t rank, const void* handle64);
/* ONE process driving several GPUs (one rg_ctx + one prepared shard table per device; how `jellyfish count` uses
 * every visible GPU): instead of export / import, attach the shards to one another directly (enables peer access
 * between their devices). tabs[r] = shard r of n. */
int rg_tables_p2p_attach(rg_table* const* tabs, int n);
/* number of CUDA devices visible to the process (0 without a driver) */
int rg_device_count(void);
/* free and total bytes of the context's device (cudaMemGetInfo): what a streaming caller sizes its table budget
 * from — `jellyfish count --disk` writes the table out as an intermediate file before it would outgrow the device
 * (hash_counter.hpp:182-202) */
int rg_device_memory(rg_ctx* ctx, uint64_t* free_bytes, uint64_t* total_bytes);
int rg_count_p2p_begin(rg_table* t);
/* max_ctas: CTAs of the scatter kernel (0 = one per SM). It is NVLink-bound: with fewer CTAs it leaves
 * SMs to the finish kernels of the previous sample running on another stream. */
int rg_count_p2p_scatter_dev(rg_table* t, const char* dev_stream, uint64_t n_bytes, uint32_t read_stride,
                             uint32_t max_ctas);
int rg_count_p2p_finish(rg_table* t);

/* ---- H1: histogram ------------------------------------------------------
 * replaces `jellyfish histo -l low -h high -i 1` (histo_main.cc:34-44,64-84) over
 * the entries a `-L min_count` dump would hold. out (HOST) has high+2-base rows
 * where base = (low<=1 ? 0 : low-1); see rg_histo_rows. */
uint64_t rg_histo_rows(uint64_t low, uint64_t high);
int rg_histo(rg_table* t, uint64_t min_count, uint64_t low, uint64_t high, uint64_t* out_host);
/* same, rows left on the DEVICE (rg_histo_rows(low, high) uint64); enqueued only, no host round trip */
int rg_histo_dev(rg_table* t, uint64_t min_count, uint64_t low, uint64_t high, uint64_t* out_dev);

/* A whole pass (count -> histo -> subtract -> read pull) enqueued without host round trips: with
 * rg_defer_checks(ctx, 1) the readers of a table no longer wait for the status words of the enqueued pushes;
 * the caller fetches them with rg_table_status_async (9 uint64 of pinned host memory: [1] foreign keys,
 * [2] k-mers that found no slot, [8] records spilled by peer GPUs; all zero = clean), synchronises once at the
 * end of the pass and, when a table is not clean, turns the deferral off, calls rg_count_end (which resolves
 * it) and repeats the readers. */
int rg_defer_checks(rg_ctx* ctx, int on);
int rg_table_status_async(rg_table* t, uint64_t* host_pinned9);
/* the same 9 words copied into DEVICE memory (callers that reduce them on the device, e.g. across ranks) */
int rg_table_status_dev(rg_table* t, uint64_t* dev_out9);

/* ---- K5: dump -----------------------------------------------------------
 * entries with min_count <= count <= max_count (sorted_dumper.hpp:94) as jf keys +
 * uint32 counts (binary_dumper.hpp:44-48), unordered. Returns n written, or
 * RG_ERR_SMALL with *needed set when cap is too small. */
int64_t rg_dump(rg_table* t, uint64_t min_count, uint64_t max_count, uint64_t* keys_host, uint32_t* counts_host,
                uint64_t cap, uint64_t* needed);
/* load (key,count) records — e.g. a parsed .Jhash — into a device table */
rg_table* rg_table_from_pairs(rg_ctx* ctx, int k, int canonical, const uint64_t* keys_host,
                              const uint32_t* counts_host, uint64_t n);
/* `jellyfish query` (query_main.cc:44-51; binary_dumper.hpp:156-197): count of each jf key,
 * 0 when absent or below min_count. Host in/out. */
int rg_query(rg_table* t, const uint64_t* keys_host, uint64_t n, uint64_t min_count, uint32_t* counts_host);

/* ---- K5 on disk: the `.Jhash` file -------------------
 * Byte-compatible with Jellyfish 2.2.5 binary/sorted files: header of
 * generic_file_header.hpp:96-119 + file_header.hpp:34-41, records of
 * binary_dumper.hpp:44-48 in the (pos, key) order of sorted_dumper.hpp:80-111, hash
 * matrix regenerated exactly as large_hash_array.hpp:948 does (unseeded glibc random()).
 * Files written here load in real `jellyfish dump/histo/query/merge`, and real files
 * (e.g. `-e` excludes) load here. */
typedef struct rg_jf_info {
  int k, canonical;
  uint64_t size;            /* header "size" (power of two) */
  uint32_t counter_len;     /* bytes per count */
  uint64_t n_records;
  uint32_t matrix_r;
  uint64_t max_reprobe_offset;
  uint64_t columns[64];     /* matrix1.columns, 2k used */
} rg_jf_info;
/* matrix of a `jellyfish count -s size -m k` table: 2k columns, r = ceil(log2(size)) rows */
int rg_jf_matrix(uint64_t size, int k, uint64_t* columns_out, uint32_t* r_out, uint64_t* size_out);
/* position of a key in that table (matrix * key & (size-1)); ~0 on bad arguments */
uint64_t rg_jf_pos(uint64_t jf_key, uint64_t size, int k);
/* reorder HOST (key,count) pairs into the file order of such a table, on the device. Returns n. */
int64_t rg_jf_sort(rg_ctx* ctx, uint64_t* keys_host, uint32_t* counts_host, uint64_t n, uint64_t size, int k);
/* `jellyfish count -o path -L min_count [-U max_count] -s size --out-counter-len counter_len`:
 * write the table as a .Jhash; cmdline goes into the header. Returns records written. 0 = default
 * for max_count (none) and counter_len (4). */
int64_t rg_table_write_jhash(rg_table* t, const char* path, uint64_t min_count, uint64_t max_count, uint64_t size,
                             uint32_t counter_len, const char* const* cmdline, int n_cmdline);
/* the same from the n hash-prefix shards of one table (one per GPU, tabs[r] = shard r): ONE file, records in
 * the (pos, key) order of a single table */
int64_t rg_tables_write_jhash(rg_table* const* tabs, int n_tabs, const char* path, uint64_t min_count, uint64_t max_count,
                              uint64_t size, uint32_t counter_len, const char* const* cmdline, int n_cmdline);
/* load a .Jhash (ours or real Jellyfish's) into a device table; info (optional) gets the header values */
rg_table* rg_table_read_jhash(rg_ctx* ctx, const char* path, rg_jf_info* info);

/* the entries with count >= min_count whose key lies in the slice {key : h(key) % slice_mod == slice_rem} (h = a
 * fixed 32-bit hash of the key, independent of the table geometry; slice_mod 0 = every key), into DEVICE arrays;
 * returns their number (RG_ERR_SMALL / *needed when cap is too small). For sliced parity of tables too large to
 * check whole: a CPU count restricted to the same slice must give exactly these pairs. */
int64_t rg_dump_slice_dev(rg_table* t, uint64_t min_count, uint32_t slice_mod, uint32_t slice_rem, uint64_t* keys_dev,
                          uint32_t* counts_dev, uint64_t cap, uint64_t* needed);

/* ---- synthetic inputs ("synth-trio v2"): counter-based read generator ----------------
 * Pure function of (seed, coordinates): the device and the host entry points write the SAME bytes, so the large
 * benchmark configs are built in HBM while the CPU restatement reads identical inputs (sliced parity). Not part
 * of the counting path. Pair i of a sample: 2 x read_len bases + qualities; buffers hold n_pairs * read_len bytes
 * each, pairs first_pair .. first_pair + n_pairs - 1 (a rank's slice of the sample). */
#define RG_SYNTH_TRIO 0 /* samples: 0 child/proband, 1 mother, 2 father, 3 sibling (quad) */
#define RG_SYNTH_PAIR 1 /* samples: 0 tumor, 1 normal */
typedef struct {
  uint64_t seed, genome_len, first_pair, n_pairs;
  uint32_t read_len, sample, kind, repeats; /* repeats != 0: 5 % of the genome in 100 repeat families, 2 % divergence */
} rg_synth_spec;
int rg_synth_pairs_dev(rg_ctx* ctx, const rg_synth_spec* spec, char* s1_dev, char* q1_dev, char* s2_dev, char* q2_dev);
int rg_synth_pairs_host(const rg_synth_spec* spec, char* s1, char* q1, char* s2, char* q2);

/* ---- S1: subtraction ----------------------------------------------------
 * modified `jellyfish merge` (src/modifiedJellyfish/jellyfish/merge_files.cc:86-154):
 * every key present (count >= mins[i]) in EXACTLY ONE table i with count >= floor.
 * which_host (optional) receives the table index of each emitted key. */
int64_t rg_merge_unique(rg_table* const* tabs, const uint64_t* mins, int n_tabs, uint64_t floor,
                        uint64_t* keys_host, uint32_t* counts_host, int32_t* which_host, uint64_t cap,
                        uint64_t* needed);
/* fused merge + CheckJellyHashList.sh:12 (query subject | awk >=min | awk <=max):
 * { (x, c_S(x)) : c_S(x) >= max(subj_min, floor), x absent (count < other_min[i]) from
 * every other table, min_cov <= c_S(x) <= max_cov }. One scan of the subject table. */
int64_t rg_subtract(rg_table* subj, uint64_t subj_min, rg_table* const* others, const uint64_t* other_min,
                    int n_others, uint64_t floor, int64_t min_cov, int64_t max_cov, uint64_t* keys_host,
                    uint32_t* counts_host, uint64_t cap, uint64_t* needed);
/* same, result left on the device as an rg_set-ready list; returns n (device-timed path) */
int64_t rg_subtract_dev(rg_table* subj, uint64_t subj_min, rg_table* const* others, const uint64_t* other_min,
                        int n_others, uint64_t floor, int64_t min_cov, int64_t max_cov, uint64_t* keys_dev,
                        uint32_t* counts_dev, uint64_t cap);
/* same, enqueued only: the number of survivors stays on the device (*count_dev; it may exceed cap, the excess
 * is dropped) and the unused tail of keys_dev is filled with ~0 (never a canonical key) */
int rg_subtract_enqueue(rg_table* subj, uint64_t subj_min, rg_table* const* others, const uint64_t* other_min,
                        int n_others, uint64_t floor, int64_t min_cov, int64_t max_cov, uint64_t* keys_dev,
                        uint32_t* counts_dev, uint64_t cap, uint64_t* count_dev);

/* ---- S2: RUFUS.Build, the legacy selection (src/RUFUS.Build.cpp:224-283) -------------
 * { (x, c_S(x), sum_i c_i(x)) : min_subject <= c_S(x) <= max_subject, sum over the control tables
 * <= max_control_sum }. Unordered; RG_ERR_SMALL / *needed as rg_dump. */
int64_t rg_build_select(rg_table* subj, int64_t min_subject, int64_t max_subject, rg_table* const* controls,
                        int n_controls, int64_t max_control_sum, uint64_t* keys_host, uint32_t* counts_host,
                        uint32_t* control_sums_host, uint64_t cap, uint64_t* needed);

/* ---- F1/F2: read pull ---------------------------------------------------
 * RUFUS.Filter (src/RUFUS.Filter.cpp:121-143,196-277) and RUFUS.Filter.single
 * (src/RUFUS.Filter.ss.cpp:100-121,164-201). */
uint64_t rg_rufus_pack(const char* s, int len);            /* Util.cpp:51-84 */
int rg_rufus_revcomp(const char* s, int len, char* out);   /* Util.cpp:187-210; returns new length */
/* the Mutations set from HashList text lines; paired=1 uses RUFUS.Filter.cpp's
 * column rules (' ' first), 0 RUFUS.Filter.ss.cpp's ('\t' first) */
rg_set* rg_set_from_hashlist_text(rg_ctx* ctx, const char* text, uint64_t n_bytes, int k, int paired);
rg_set* rg_set_from_rufus_keys(rg_ctx* ctx, const uint64_t* rufus_keys_host, uint64_t n, int k);
/* from canonical jf keys (e.g. rg_subtract output): inserts k-mer and reverse complement */
rg_set* rg_set_from_jf_keys(rg_ctx* ctx, const uint64_t* jf_keys_host, uint64_t n, int k);
rg_set* rg_set_from_jf_keys_dev(rg_ctx* ctx, const uint64_t* jf_keys_dev, uint64_t n, int k);
/* a set rebuilt in place every pass (no allocation, no host round trip): rg_set_create(k < 32, room for max_keys
 * jf keys), then rg_set_rebuild_dev(jf_keys_dev, n_max <= max_keys): entries equal to ~0 are padding */
rg_set* rg_set_create(rg_ctx* ctx, int k, uint64_t max_keys);
int rg_set_rebuild_dev(rg_set* s, const uint64_t* jf_keys_dev, uint64_t n_max);
uint64_t rg_set_size(rg_set* s);
void rg_set_free(rg_set* s);
/* per-read hit counts; all DEVICE pointers. skip_last=1: paired tool's `i < length()-1` bound */
int rg_filter_hits_dev(rg_set* s, const char* seq_dev, const char* qual_dev, const uint64_t* offs_dev,
                       uint64_t n_reads, uint64_t n_bytes, int minq, int skip_last, int32_t* hits_dev);
/* fixed-length reads (every read read_len bytes, no offsets array needed) */
int rg_filter_hits_fixed_dev(rg_set* s, const char* seq_dev, const char* qual_dev, uint32_t read_len,
                             uint64_t n_reads, int minq, int skip_last, int32_t* hits_dev);
/* paired pull on DEVICE-resident fixed-length mates: hits per mate + keep mask, all device */
int rg_filter_pairs_fixed_dev(rg_set* s, const char* s1_dev, const char* q1_dev, const char* s2_dev,
                              const char* q2_dev, uint32_t read_len, uint64_t n_pairs, int minq, int thresh,
                              int32_t* hits1_dev, int32_t* hits2_dev, uint8_t* keep_dev);
/* keep[i] = hits(mate1)>=thresh || hits(mate2)>=thresh; HOST buffers */
int rg_filter_pairs(rg_set* s, const char* s1, const char* q1, const uint64_t* o1, const char* s2, const char* q2,
                    const uint64_t* o2, uint64_t n_pairs, int minq, int thresh, uint8_t* keep_host);
/* AnnotateOverlap (src/AnnotateOverlap.cpp:95-146), the `.hashcount.fastq` depth line: per BASE, the
 * number of mutant k-mer windows covering it (capped at 255; the tool prints min(depth,93)+33). A
 * window counts when it holds no 'N', every quality-33 >= minq (the tool: 3), it or its reverse
 * complement is in the set, and it is not the read's last window (`i < length-K`). HOST buffers. */
int rg_annotate(rg_set* s, const char* seq, const char* qual, const uint64_t* offs, uint64_t n_reads, int minq,
                uint8_t* depth_host);
/* single-end: hits per read, last base included; HOST buffers */
int rg_filter_single(rg_set* s, const char* seq, const char* qual, const uint64_t* offs, uint64_t n_reads,
                     int minq, int32_t* hits_host);

/* ---- O2-O4: the device side of the greedy overlap assembly (src/Overlap.cpp) -----------
 * The greedy loop is serial and stays with the caller (tools/overlap.cpp); the seed index, the
 * seed lookups and the Align3 placement search run on the GPU.
 *  rg_overlap_build     O2, RebuildHashTable (:34-76): load the reads, index the SearchHash-mers
 *                       (and their reverse complements) of reads >= first
 *  rg_overlap_update    a merge replaced read idx by a contig (:1066-1070); the index is not
 *                       updated, like the reference's (:1083,:1101)
 *  rg_overlap_seed_hits O3 lookups of PrepairSearchList (:78-103): per query, the read indices
 *                       > q_ai[q] that share a seed, one entry per shared occurrence
 *  rg_overlap_align     O4, Align3 (:169-360): A against the current sequence of every candidate.
 *                       Per candidate: best of the "shorter inside longer" placements (score1,
 *                       overlap1, perfect1) and best over all three placement families
 *                       (score_all, overlap_all); -1 = nothing accepted. The caller applies the
 *                       PerfectMatch short-circuit across candidates (see tools/overlap.cpp). */
typedef struct rg_overlap rg_overlap;
typedef struct rg_align_result {
  int32_t score1, overlap1, perfect1, score_all, overlap_all;
} rg_align_result;
rg_overlap* rg_overlap_create(rg_ctx* ctx, int search_hash);
void rg_overlap_free(rg_overlap* ov);
int rg_overlap_build(rg_overlap* ov, const char* seq, const uint64_t* offs, uint64_t n_reads, uint64_t first);
int rg_overlap_update(rg_overlap* ov, uint64_t idx, const char* seq, uint64_t len);
int64_t rg_overlap_seed_hits(rg_overlap* ov, const char* qseq, const uint64_t* qoffs, const uint64_t* q_ai, uint64_t nq,
                             uint32_t* hits_host, uint64_t cap, uint64_t* hit_offs_host);
int rg_overlap_align(rg_overlap* ov, const char* a, uint64_t a_len, const uint32_t* cand, uint64_t n_cand,
                     float min_percent, int min_overlap, rg_align_result* out_host);
/* the read and its reverse complement (same length) against their two candidate lists in one launch */
int rg_overlap_align_both(rg_overlap* ov, const char* a, const char* a_rev, uint64_t a_len, const uint32_t* cand,
                          uint64_t n_cand, const uint32_t* cand_rev, uint64_t n_cand_rev, float min_percent,
                          int min_overlap, rg_align_result* out_host, rg_align_result* out_rev_host);

#ifdef __cplusplus
}
#endif
#endif
