// rg_internal.cuh — shared declarations of librufus_gpu.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "rufus_gpu.h"

// ---------------------------------------------------------------------------
// Table geometry. A table is an array of TILES; a tile is an independent
// open-addressed (linear probing, wrap inside the tile) hash table of
// RG_TILE_SLOTS slots, stored SoA: keys[] (u64, EMPTY = ~0) and counts[] (u32).
// One tile = 128 KiB of keys + 64 KiB of counts = 192 KiB: exactly what one CTA
// stages in shared memory (227 KiB/SM on B200) in the tiled insert.
//   h      = hash32(key)                  (lo * C1 + hi * C2 mod 2^32)
//   p      = h * n_tiles_global           (64-bit product)
//   tile   = p >> 32                      ("hash prefix": also the shard id of the multi-GPU partition)
//   slot0  = low32(p) >> (32 - RG_TILE_BITS)   (the position of h inside its tile's hash range)
// ---------------------------------------------------------------------------
#define RG_TILE_BITS 14
#define RG_TILE_SLOTS (1u << RG_TILE_BITS)
#define RG_EMPTY 0xFFFFFFFFFFFFFFFFull
#define RG_TILE_KEY_BYTES (RG_TILE_SLOTS * 8u)
#define RG_TILE_CNT_BYTES (RG_TILE_SLOTS * 4u)

struct rg_prof_entry {
  double ms = 0;
  uint64_t n = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
};

struct rg_ctx {
  bool profiling = false;
  std::map<std::string, rg_prof_entry> prof;
  std::vector<cudaEvent_t> event_pool;
  int device = 0;
  int sm_count = 148;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  uint64_t launches = 0;
  // small pinned scratch for scalar read-backs
  uint64_t* h_scratch = nullptr;  // pinned, 64 u64
  uint64_t* d_scratch = nullptr;  // device, 64 u64
  // host -> device pipeline of the host-buffer entry points: two staging slots filled on a copy
  // stream while the kernels of the previous slot run on `stream`
  cudaStream_t copy_stream = nullptr;
  char* slot[2] = {nullptr, nullptr};
  uint64_t slot_bytes = 0;
  cudaEvent_t slot_copied[2] = {nullptr, nullptr};   // recorded on copy_stream after the slot's H2D
  cudaEvent_t slot_free[2] = {nullptr, nullptr};     // recorded on stream after the last kernel that reads the slot
  uint32_t slot_next = 0;
  // grow-only device scratch for the read filter's host entry points (offsets, hit counters, keep mask)
  void* pool[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  uint64_t pool_bytes[6] = {0, 0, 0, 0, 0, 0};
};

struct rg_table {
  rg_ctx* ctx = nullptr;
  int k = 0;
  int canonical = 1;
  int mode = RG_MODE_DIRECT;
  // geometry
  uint32_t n_tiles = 0;         // tiles held by this table (this shard)
  uint32_t n_tiles_global = 0;  // n_tiles * n_parts
  uint32_t tile_first = 0;      // global id of local tile 0
  uint32_t shard_part = 0, shard_parts = 1;
  uint64_t* keys = nullptr;     // [n_tiles * RG_TILE_SLOTS]
  uint32_t* counts = nullptr;
  // device-side status words: [0]=spill count, [1]=foreign-key count (wrong shard)
  uint64_t* d_status = nullptr;
  uint64_t* spill = nullptr;    // keys that found their tile full
  uint64_t spill_cap = 0;
  bool tiles_dirty = false;     // table holds data (tiled insert must load tiles, not zero them)
  bool unchecked = false;       // pushes were enqueued whose status words (spill / foreign) are not read back yet
  // tiled-mode staging
  uint64_t* part_keys = nullptr;     // bucketed keys
  uint64_t part_cap = 0;             // total key capacity
  uint32_t* part_cursor = nullptr;   // [n_tiles] fill of each bucket
  uint32_t bucket_cap = 0;
  uint64_t* stage = nullptr;         // host->device staging for rg_count_push
  uint64_t stage_bytes = 0;
  uint32_t* brk = nullptr;           // read-start bitmap for offsets pushes
  uint64_t brk_words = 0;
  uint64_t* d_offs = nullptr;
  uint64_t d_offs_cap = 0;
};

struct rg_set {
  rg_ctx* ctx = nullptr;
  int k = 0;
  uint64_t n = 0;           // distinct rufus keys
  uint32_t slots = 0;       // power of two
  uint64_t* table = nullptr;  // open addressing, EMPTY = ~0
  uint32_t* bitmap = nullptr;  // bitmap_bits-bit prefilter (allocated at RG_SET_BITMAP_BITS)
  uint32_t bitmap_bits = 1u << 15;  // power of two in [2^15, RG_SET_BITMAP_BITS], ~64 bits per member
  int has_all_ones = 0;     // the value ~0 itself is a member (k = 32, all T)
};
#define RG_SET_BITMAP_BITS (1u << 19)

void rg_set_error(const std::string& msg);
#define RG_CUDA_OK(expr)                                                                    \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      rg_set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                     \
      return RG_ERR_CUDA;                                                                   \
    }                                                                                       \
  } while (0)
#define RG_CUDA_OKP(expr)                                                                   \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      rg_set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                     \
      return nullptr;                                                                       \
    }                                                                                       \
  } while (0)

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t rg_mix64(uint64_t h) {
  // murmur3 fmix64 — a bijection on 64 bits
  h ^= h >> 33;
  h *= 0xff51afd7ed558ccdull;
  h ^= h >> 33;
  h *= 0xc4ceb9fe1a85ec53ull;
  h ^= h >> 33;
  return h;
}

// table hash: two 32-bit multiply-adds (both on the FMA pipe, nothing on the integer ALU pipe that the
// extraction saturates). Checked against the previous xor-shift-multiply hash on real k-mer sets: same
// tile-load variance (Poisson) and the same mean probe length (profiles/README.md).
__host__ __device__ __forceinline__ uint32_t rg_hash32(uint64_t key) {
  return (uint32_t)key * 0x9E3779B1u + (uint32_t)(key >> 32) * 0x85EBCA77u;
}

// tile (global numbering) and home slot of a key
__host__ __device__ __forceinline__ void rg_locate(uint64_t key, uint32_t n_tiles_global, uint32_t& tile,
                                                   uint32_t& slot) {
  const uint64_t p = (uint64_t)rg_hash32(key) * n_tiles_global;
  tile = (uint32_t)(p >> 32);
  slot = (uint32_t)p >> (32 - RG_TILE_BITS);
}

__host__ __device__ __forceinline__ uint32_t rg_tile_of_key(uint64_t key, uint32_t n_tiles_global) {
  return (uint32_t)(((uint64_t)rg_hash32(key) * n_tiles_global) >> 32);
}

// reverse complement of a jf key (first base in the MSBs of 2k bits)
__host__ __device__ __forceinline__ uint64_t rg_revcomp_jf(uint64_t x, int k) {
  // reverse the order of the 2-bit groups of the 64-bit word, complement, shift down
  x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
  x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
#ifdef __CUDA_ARCH__
  uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
  lo = __byte_perm(lo, 0, 0x0123);
  hi = __byte_perm(hi, 0, 0x0123);
  x = ((uint64_t)lo << 32) | hi;
#else
  x = __builtin_bswap64(x);
#endif
  return (~x) >> (64 - 2 * k);
}

struct TableView {
  uint64_t* keys;
  uint32_t* counts;
  uint32_t n_tiles;
  uint32_t n_tiles_global;
  uint32_t tile_first;
};

static inline TableView rg_view(const rg_table* t) {
  TableView v;
  v.keys = t->keys;
  v.counts = t->counts;
  v.n_tiles = t->n_tiles;
  v.n_tiles_global = t->n_tiles_global;
  v.tile_first = t->tile_first;
  return v;
}

// add `add` to key's counter in a global-memory table. Returns 0 ok, 1 tile full, 2 foreign shard.
__device__ __forceinline__ int rg_table_add(const TableView& tv, uint64_t key, uint32_t add) {
  uint32_t tile, slot;
  rg_locate(key, tv.n_tiles_global, tile, slot);
  tile -= tv.tile_first;
  if (tile >= tv.n_tiles) return 2;
  const uint64_t base = (uint64_t)tile << RG_TILE_BITS;
  for (uint32_t probe = 0; probe < RG_TILE_SLOTS; ++probe) {
    unsigned long long* kp = (unsigned long long*)(tv.keys + base + slot);
    unsigned long long cur = *(volatile unsigned long long*)kp;
    if (cur == RG_EMPTY) cur = atomicCAS(kp, (unsigned long long)RG_EMPTY, (unsigned long long)key);
    if (cur == RG_EMPTY || cur == key) {
      atomicAdd(tv.counts + base + slot, add);
      return 0;
    }
    slot = (slot + 1) & (RG_TILE_SLOTS - 1);
  }
  return 1;
}

// count of key (0 if absent)
__device__ __forceinline__ uint32_t rg_table_get(const TableView& tv, uint64_t key) {
  uint32_t tile, slot;
  rg_locate(key, tv.n_tiles_global, tile, slot);
  tile -= tv.tile_first;
  if (tile >= tv.n_tiles) return 0;
  const uint64_t base = (uint64_t)tile << RG_TILE_BITS;
  for (uint32_t probe = 0; probe < RG_TILE_SLOTS; ++probe) {
    const uint64_t cur = __ldg(tv.keys + base + slot);
    if (cur == key) return __ldg(tv.counts + base + slot);
    if (cur == RG_EMPTY) return 0;
    slot = (slot + 1) & (RG_TILE_SLOTS - 1);
  }
  return 0;
}

// ---------------------------------------------------------------------------
// host-side launchers implemented in the .cu files
// ---------------------------------------------------------------------------
struct ExtractParams {
  const char* stream;      // device bytes
  uint64_t n_bytes;
  const uint32_t* brk;     // optional read-start bitmap (bit (p&31) of word p>>5)
  uint32_t read_stride;    // optional fixed read stride
  int k;
  int canonical;
};

// read back the status words of the enqueued pushes and resolve spills (every reader calls it first)
int rg_table_finish(rg_table* t);

int rg_launch_count_direct(rg_table* t, const ExtractParams& ep);
int rg_launch_partition(rg_table* t, const ExtractParams& ep);   // into t->part_keys buckets
int rg_launch_partition_keys(rg_table* t, const uint64_t* dev_keys, uint64_t n);
int rg_launch_tile_insert(rg_table* t);
uint64_t rg_partition_pad_keys(uint32_t n_tiles, uint64_t n_units);                           // consume buckets
int rg_launch_insert_keys(rg_table* t, const uint64_t* dev_keys, uint64_t n, const uint32_t* dev_counts);
int rg_launch_mark_breaks(rg_ctx* ctx, const uint64_t* d_offs, uint64_t n_reads, uint32_t* brk);
int rg_launch_extract_partition(rg_ctx* ctx, const ExtractParams& ep, uint32_t n_parts, uint64_t* keys_out,
                                uint64_t* part_offsets, uint64_t* d_tmp_counts);

int rg_launch_histo(rg_table* t, uint64_t min_count, uint64_t base, uint64_t nb, uint64_t* d_out);
int rg_launch_stats(rg_table* t, uint64_t* d_out2);
int rg_launch_dump(rg_table* t, uint64_t min_count, uint64_t max_count, uint64_t* d_keys, uint32_t* d_counts,
                   uint64_t cap, uint64_t* d_cursor);
int rg_launch_query(rg_table* t, const uint64_t* d_keys, uint64_t n, uint64_t min_count, uint32_t* d_out);
struct OtherTables {
  TableView tv[16];
  uint32_t min_count[16];
  int n;
};
int rg_launch_subtract(rg_table* subj, uint64_t lo, uint64_t hi, const OtherTables& others, uint64_t* d_keys,
                       uint32_t* d_counts, uint64_t cap, uint64_t* d_cursor);
int rg_launch_rehash(rg_table* dst, const uint64_t* old_keys, const uint32_t* old_counts, uint64_t n_slots);

int rg_launch_set_build(rg_set* s, const uint64_t* d_vals, uint64_t n);
int rg_launch_jf_to_rufus(rg_ctx* ctx, const uint64_t* d_jf, uint64_t n, int k, uint64_t* d_out2n);
// pos_base: stream position of seq[0] (offs / read_len are absolute); zero_hits: clear hits[0..n_reads) first
int rg_launch_filter_hits(rg_set* s, const char* seq, const char* qual, const uint64_t* offs, uint32_t read_len,
                          uint64_t n_reads, uint64_t n_bytes, uint64_t pos_base, bool zero_hits, int minq, int skip_last,
                          int32_t* hits);
int rg_launch_keep(rg_ctx* ctx, const int32_t* h1, const int32_t* h2, uint64_t n, int thresh, uint8_t* keep);

// RAII: times one kernel launch when profiling is on, and counts the launch
struct rg_launch_scope {
  rg_ctx* ctx;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  const char* name;
  rg_launch_scope(rg_ctx* c, const char* nm) : ctx(c), name(nm) {
    ctx->launches++;
    if (ctx->profiling) {
      e0 = take();
      e1 = take();
      cudaEventRecord(e0, ctx->stream);
    }
  }
  ~rg_launch_scope() {
    if (e0) {
      cudaEventRecord(e1, ctx->stream);
      ctx->prof[name].pending.emplace_back(e0, e1);
    }
  }
  cudaEvent_t take() {
    if (!ctx->event_pool.empty()) {
      cudaEvent_t e = ctx->event_pool.back();
      ctx->event_pool.pop_back();
      return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
  }
};

static inline unsigned rg_grid(const rg_ctx* ctx, int ctas_per_sm) { return (unsigned)(ctx->sm_count * ctas_per_sm); }
