// rg_internal.cuh — shared declarations of librufus_gpu.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "rufus_gpu.h"

// ---------------------------------------------------------------------------
// Table geometry. A table is an array of TILES plus an OVERFLOW region; a tile is an
// open-addressed (linear probing, wrap inside the tile) hash table of RG_TILE_SLOTS slots,
// stored SoA: keys[] (u64, EMPTY = ~0) and counts[] (u32). One tile = 128 KiB of keys +
// 64 KiB of counts = 192 KiB: what one CTA stages in shared memory in the tiled insert.
//
// The tile of a k-mer is chosen by its MINIMIZER, so that the consecutive k-mers of a read that
// share a minimizer (a "super-k-mer") travel to their tile as ONE 16-byte record (rg_sk.cuh):
//   y_j    = (((f_j * C1) ^ (r_j * C1)) + (f_j & r_j) * C2) << (32 - 2m)   j = 0..k-m; f_j / r_j = the m-mer at base j
//            of the k-mer and its reverse complement: strand-symmetric without a compare. (Not (f + r) * C: f + r only
//            depends on the differences of mirrored bases, 7^(m/2) classes instead of 4^m / 2; and the xor alone is 0
//            for every palindromic m-mer, which would all share one tile: the (f & r) term separates them.)
//   ymin   = min_j y_j ;  mh = mix32(ymin) & ~15     (the hash prefix every level of the partition uses)
//   p      = mh * groups                  (64-bit product; groups = hash-prefix GROUPS over all shards)
//   group  = p >> 32                      (the level-1 partition bucket; group / groups_per_shard = owning GPU)
//   q      = low32(p) * group_tiles       (64-bit product; group_tiles = tiles per group, 1 for small tables)
//   tile   = group * group_tiles + (q >> 32)
//   slot0  = hash32(key) >> (32 - RG_TILE_BITS)
// A key is probed over at most RG_PROBE_MAX consecutive slots of its tile; a key that finds that window
// full lives in the overflow region (slots [n_tiles << RG_TILE_BITS, ... + ovf_slots), one open-addressed
// table addressed by a second hash of the key, global-memory atomics only). Minimizer-owned tiles are
// loaded unevenly (one hot minimizer can exceed a tile whatever the tile count); the overflow region
// takes what does not fit, and the host grows it (or the tile count) between pushes.
// With group_tiles == 1 a group IS a tile and one partition pass feeds the tile inserts; larger
// tables (more tiles than the partition kernel can histogram in shared memory) use two passes:
// stream -> groups, then group -> tiles.
// ---------------------------------------------------------------------------
#define RG_TILE_BITS 14
#define RG_PROBE_MAX 256u          // probe window inside a tile
#define RG_OVF_PROBE_MAX 128u      // probe window inside the overflow region (beyond it: the spill list)
#define RG_MM_C 0x9E3779B1u        // m-mer order hash multipliers (odd)
#define RG_MM_C2 0x85EBCA77u
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
  bool defer_checks = false;   // rg_defer_checks: readers do not wait for the status words of enqueued pushes
  uint32_t force_groups = 0;   // RG_GROUPS_LOCAL (test knob): tiled tables use this many groups, i.e. the two-level partition
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
  uint32_t group_tiles = 1;     // tiles per hash-prefix group (level-2 fanout); n_tiles % group_tiles == 0
  uint32_t shard_part = 0, shard_parts = 1;
  uint64_t* keys = nullptr;     // [n_tiles * RG_TILE_SLOTS]
  uint32_t* counts = nullptr;
  int m = 0;                    // minimizer length (<= min(k, 16)); k - m + 1 = window of the partition kernel
  uint64_t ovf_slots = 0;       // overflow region (power of two, multiple of RG_TILE_SLOTS), after the tiles
  uint64_t ovf_used = 0;        // keys placed in the overflow region so far (host copy, updated by rg_table_finish)
  // device-side status words (RG_ST_*)
  uint64_t* d_status = nullptr;
  uint64_t* spill_keys = nullptr;   // (key, count) pairs that found tiles and overflow region full (RG_ST_LOST)
  uint32_t* spill_counts = nullptr;
  uint64_t spill_cap = 0;
  bool tiles_dirty = false;     // table holds data (tiled insert must load tiles, not zero them)
  uint32_t p2p_groups = 0;      // > 0: peer-to-peer multi-GPU counting with this many groups per shard
  bool unchecked = false;       // pushes were enqueued whose status words (spill / foreign) are not read back yet
  // tiled-mode staging
  uint4* part_keys = nullptr;        // bucketed super-k-mer records
  uint64_t part_cap = 0;             // total record capacity
  uint32_t* part_cursor = nullptr;   // [n_tiles] fill of each bucket
  uint32_t bucket_cap = 0;
  // two-level partition (group_tiles > 1) and peer-to-peer multi-GPU counting: slabs of the groups
  uint4* grp_keys = nullptr;         // [n_groups_local][grp_cap] records
  uint32_t* grp_cursor = nullptr;    // [n_groups_local]
  uint32_t grp_cap = 0;
  uint64_t grp_alloc_keys = 0;       // records allocated behind grp_keys
  // peer-to-peer: one allocation [cursors | status | spill | slabs], exported with CUDA IPC
  void* p2p_block = nullptr;
  uint64_t p2p_block_bytes = 0;
  uint64_t p2p_spill_cap = 0;
  bool p2p_pending = false;          // a finished peer-to-peer push whose exported spill list is not collected yet
  uint32_t p2p_ctas = 0;             // CTAs of the scatter kernel (0 = one per SM); fewer leave SMs to a concurrent stream
  uint64_t p2p_max_keys = 0;         // keys one shard may receive per push (sizes the slabs)
  uint32_t* p2p_src_cursor = nullptr;  // LOCAL fill of my sub-slab in every group of every shard [groups_global]
  uint64_t* p2p_status = nullptr;    // [0] = number of records other GPUs could not place in a slab (spilled)
  uint4* p2p_spill = nullptr;
  uint4* peer_keys[8] = {};
  uint32_t* peer_cursor[8] = {};
  uint64_t* peer_status[8] = {};
  uint4* peer_spill[8] = {};
  void* peer_mapped[8] = {};         // bases returned by cudaIpcOpenMemHandle (to close them)
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
  uint64_t* scratch = nullptr;   // rg_set_create: room for the 2 * max_keys rufus keys of a rebuild
  uint64_t max_keys = 0;
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
// status words of a table (device, uint64 each)
#define RG_ST_FOREIGN 1   // keys / records that do not belong to this table's shard
#define RG_ST_LOST 2      // (key, count) pairs that found the overflow region full: appended to the spill list, re-inserted
                          // by the host after it has grown the table (beyond the list's capacity they are lost: error)
#define RG_ST_OVF_NEW 3   // keys newly placed in the overflow region
#define RG_ST_DIRECT 4    // != 0: global-memory inserts happened since the table was last empty (the tiled insert must load tiles)
#define RG_ST_WORDS 8

// slot hash inside a tile: two 32-bit multiply-adds (FMA pipe)
__host__ __device__ __forceinline__ uint32_t rg_hash32(uint64_t key) {
  return (uint32_t)key * 0x9E3779B1u + (uint32_t)(key >> 32) * 0x85EBCA77u;
}
// second, independent hash: slot in the overflow region
__host__ __device__ __forceinline__ uint32_t rg_hash32b(uint64_t key) {
  uint32_t h = (uint32_t)key * 0xC2B2AE35u + (uint32_t)(key >> 32) * 0x27D4EB2Fu;
  h ^= h >> 15;
  return h * 0x2C1B3C6Du;
}

// table geometry as the kernels see it
struct RgGeo {
  uint32_t groups;        // hash-prefix groups over ALL shards (= n_tiles_global / group_tiles)
  uint32_t group_tiles;   // tiles per group
  uint32_t k, m;          // k-mer and minimizer length
};

// reverse complement of a jf key (first base in the MSBs of 2k bits)
__host__ __device__ __forceinline__ uint64_t rg_revcomp_jf(uint64_t x, int k) {
#ifdef __CUDA_ARCH__
  // bit-reverse the 64-bit word (two BREV), swap the bits inside every pair, complement, shift down
  // (one LOP3 per word swaps the pair bits AND complements: ~(((r >> 1) & 0x55..) | ((r << 1) & 0xAA..)), LUT 0x1B)
  const uint32_t lo = __brev((uint32_t)(x >> 32)), hi = __brev((uint32_t)x);
  uint32_t l2, h2;
  asm("lop3.b32 %0, %1, %2, 0x55555555, 0x1B;" : "=r"(l2) : "r"(lo >> 1), "r"(lo << 1));
  asm("lop3.b32 %0, %1, %2, 0x55555555, 0x1B;" : "=r"(h2) : "r"(hi >> 1), "r"(hi << 1));
  return (((uint64_t)h2 << 32) | l2) >> (64 - 2 * k);
#else
  x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
  x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
  x = __builtin_bswap64(x);
#endif
  return (~x) >> (64 - 2 * k);
}

// final mix of the minimizer order value -> the hash prefix (low 4 bits are free for the record's length)
__host__ __device__ __forceinline__ uint32_t rg_mix_min(uint32_t ymin) {
  uint32_t h = ymin * 0x9E3779B1u;
  h = (h ^ (h >> 15)) * 0x85EBCA77u;
  h ^= h >> 13;
  return h & ~15u;
}

// hash prefix of a k-mer from its packed key (slow path: k-m+1 m-mers; the partition kernel gets the same
// value incrementally from the read). Strand-symmetric: key and its reverse complement give the same value.
__host__ __device__ __forceinline__ uint32_t rg_min_hash(uint64_t key, uint32_t k, uint32_t m) {
  const uint64_t rc = rg_revcomp_jf(key, (int)k);
  const uint32_t sh = 32u - 2u * m;
  const uint32_t cm = RG_MM_C << sh, cm2 = RG_MM_C2 << sh;   // (x * cm) == ((x mod 4^m) * RG_MM_C) << sh  (mod 2^32)
  uint32_t ymin = 0xFFFFFFFFu;
  for (uint32_t j = 0; j + m <= k; ++j) {
    const uint32_t f = (uint32_t)(key >> (2u * (k - m - j)));   // m-mer at base j (garbage above 2m bits: multiplied away)
    const uint32_t r = (uint32_t)(rc >> (2u * j));              // its reverse complement
    const uint32_t y = ((f * cm) ^ (r * cm)) + (f & r) * cm2;
    ymin = y < ymin ? y : ymin;
  }
  return rg_mix_min(ymin);
}

// level-1 bucket (hash-prefix group, global numbering) of a hash prefix
__host__ __device__ __forceinline__ uint32_t rg_group_of_mh(uint32_t mh, uint32_t groups) {
  return (uint32_t)(((uint64_t)mh * groups) >> 32);
}
// tile (global numbering) of a hash prefix
__host__ __device__ __forceinline__ uint32_t rg_tile_of_mh(uint32_t mh, RgGeo geo) {
  const uint64_t p = (uint64_t)mh * geo.groups;
  const uint64_t q = (uint64_t)(uint32_t)p * geo.group_tiles;
  return (uint32_t)(p >> 32) * geo.group_tiles + (uint32_t)(q >> 32);
}
// tile inside its group (level-2 partition)
__host__ __device__ __forceinline__ uint32_t rg_tile_in_group_of_mh(uint32_t mh, RgGeo geo) {
  const uint32_t r = mh * geo.groups;             // low word of mh * groups
  return (uint32_t)(((uint64_t)r * geo.group_tiles) >> 32);
}
__host__ __device__ __forceinline__ uint32_t rg_slot_of_key(uint64_t key) { return rg_hash32(key) >> (32 - RG_TILE_BITS); }

struct TableView {
  uint64_t* keys;
  uint32_t* counts;
  uint32_t n_tiles;
  uint32_t tile_first;
  RgGeo geo;
  uint64_t ovf_base;      // first slot of the overflow region (= n_tiles << RG_TILE_BITS)
  uint64_t ovf_mask;      // ovf_slots - 1
  uint64_t* status;       // RG_ST_* words of the table
  uint64_t* spill_keys;   // last resort: what found the overflow region full
  uint32_t* spill_counts;
  uint64_t spill_cap;
};

static inline TableView rg_view(const rg_table* t) {
  TableView v;
  v.keys = t->keys;
  v.counts = t->counts;
  v.n_tiles = t->n_tiles;
  v.tile_first = t->tile_first;
  v.geo.groups = t->n_tiles_global / t->group_tiles;
  v.geo.group_tiles = t->group_tiles;
  v.geo.k = (uint32_t)t->k;
  v.geo.m = (uint32_t)t->m;
  v.ovf_base = (uint64_t)t->n_tiles << RG_TILE_BITS;
  v.ovf_mask = t->ovf_slots - 1;
  v.status = t->d_status;
  v.spill_keys = t->spill_keys;
  v.spill_counts = t->spill_counts;
  v.spill_cap = t->spill_cap;
  return v;
}
static inline uint64_t rg_table_slots(const rg_table* t) { return ((uint64_t)t->n_tiles << RG_TILE_BITS) + t->ovf_slots; }

#ifdef __CUDACC__
// add `add` to key's counter in the overflow region. Returns 0 ok, 1 probe window full (the pair goes to the spill list).
__device__ __forceinline__ int rg_ovf_add(const TableView& tv, uint64_t key, uint32_t add) {
  uint64_t slot = rg_hash32b(key) & tv.ovf_mask;
  for (uint32_t probe = 0; probe < RG_OVF_PROBE_MAX; ++probe) {
    unsigned long long* kp = (unsigned long long*)(tv.keys + tv.ovf_base + slot);
    unsigned long long cur = *(volatile unsigned long long*)kp;
    if (cur == RG_EMPTY) {
      cur = atomicCAS(kp, (unsigned long long)RG_EMPTY, (unsigned long long)key);
      if (cur == RG_EMPTY) atomicAdd((unsigned long long*)&tv.status[RG_ST_OVF_NEW], 1ull);
    }
    if (cur == RG_EMPTY || cur == key) {
      atomicAdd(tv.counts + tv.ovf_base + slot, add);
      return 0;
    }
    slot = (slot + 1) & tv.ovf_mask;
  }
  const unsigned long long at = atomicAdd((unsigned long long*)&tv.status[RG_ST_LOST], 1ull);
  if (at < tv.spill_cap) {
    tv.spill_keys[at] = key;
    tv.spill_counts[at] = add;
  }
  return 1;
}

// add `add` to key's counter in a global-memory table, tile already known (local numbering).
// Returns 0 ok, 1 lost (overflow region full).
__device__ __forceinline__ int rg_table_add_at(const TableView& tv, uint32_t tile, uint64_t key, uint32_t add) {
  const uint64_t base = (uint64_t)tile << RG_TILE_BITS;
  uint32_t slot = rg_slot_of_key(key);
  for (uint32_t probe = 0; probe < RG_PROBE_MAX; ++probe) {
    unsigned long long* kp = (unsigned long long*)(tv.keys + base + slot);
    unsigned long long cur = *(volatile unsigned long long*)kp;
    if (cur == RG_EMPTY) cur = atomicCAS(kp, (unsigned long long)RG_EMPTY, (unsigned long long)key);
    if (cur == RG_EMPTY || cur == key) {
      atomicAdd(tv.counts + base + slot, add);
      return 0;
    }
    slot = (slot + 1) & (RG_TILE_SLOTS - 1);
  }
  return rg_ovf_add(tv, key, add);
}

// same, tile from the key's minimizer. Returns 0 ok, 1 lost, 2 foreign shard.
__device__ __forceinline__ int rg_table_add(const TableView& tv, uint64_t key, uint32_t add) {
  const uint32_t tile = rg_tile_of_mh(rg_min_hash(key, tv.geo.k, tv.geo.m), tv.geo) - tv.tile_first;
  if (tile >= tv.n_tiles) {
    atomicAdd((unsigned long long*)&tv.status[RG_ST_FOREIGN], 1ull);
    return 2;
  }
  return rg_table_add_at(tv, tile, key, add);
}

// count of key (0 if absent), tile known (local numbering)
__device__ __forceinline__ uint32_t rg_table_get_at(const TableView& tv, uint32_t tile, uint64_t key) {
  const uint64_t base = (uint64_t)tile << RG_TILE_BITS;
  uint32_t slot = rg_slot_of_key(key);
  for (uint32_t probe = 0; probe < RG_PROBE_MAX; ++probe) {
    const uint64_t cur = __ldg(tv.keys + base + slot);
    if (cur == key) return __ldg(tv.counts + base + slot);
    if (cur == RG_EMPTY) return 0;
    slot = (slot + 1) & (RG_TILE_SLOTS - 1);
  }
  uint64_t os = rg_hash32b(key) & tv.ovf_mask;
  for (uint32_t probe = 0; probe < RG_OVF_PROBE_MAX; ++probe) {
    const uint64_t cur = __ldg(tv.keys + tv.ovf_base + os);
    if (cur == key) return __ldg(tv.counts + tv.ovf_base + os);
    if (cur == RG_EMPTY) return 0;
    os = (os + 1) & tv.ovf_mask;
  }
  return 0;
}
__device__ __forceinline__ uint32_t rg_table_get(const TableView& tv, uint64_t key) {
  const uint32_t tile = rg_tile_of_mh(rg_min_hash(key, tv.geo.k, tv.geo.m), tv.geo) - tv.tile_first;
  if (tile >= tv.n_tiles) return 0;
  return rg_table_get_at(tv, tile, key);
}
#endif

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
  uint32_t skip = 0;       // windows starting before this position are not emitted (the halo of a split push)
};

// read back the status words of the enqueued pushes, grow the table when its overflow region fills (every reader calls it first)
int rg_table_finish(rg_table* t);

int rg_launch_count_direct(rg_table* t, const ExtractParams& ep);
int rg_launch_partition(rg_table* t, const ExtractParams& ep);   // into t->part_keys buckets (or the group slabs)
int rg_launch_partition_groups(rg_table* t);                      // level 2: group slabs -> tile buckets
int rg_launch_partition_multi(rg_table* t, const ExtractParams& ep);   // level 1 into every GPU's group slabs
uint32_t rg_partition_max_buckets();
double rg_partition_record_density(int k, int m);
int rg_launch_tile_insert(rg_table* t);                           // consume buckets
int rg_launch_insert_keys(rg_table* t, const uint64_t* dev_keys, uint64_t n, const uint32_t* dev_counts);
int rg_launch_insert_records(rg_table* t, const uint4* dev_recs, uint64_t n);
int rg_launch_mark_breaks(rg_ctx* ctx, const uint64_t* d_offs, uint64_t n_reads, uint32_t* brk);
int rg_launch_publish_cursors(rg_table* t);                       // peer-to-peer: my sub-slab fills -> every owner's cursor array
int rg_launch_extract_partition(rg_ctx* ctx, const ExtractParams& ep, RgGeo geo, uint32_t n_parts, uint64_t* keys_out,
                                uint64_t* part_offsets, uint64_t* d_tmp_counts);

int rg_launch_histo(rg_table* t, uint64_t min_count, uint64_t base, uint64_t nb, uint64_t* d_out);
int rg_launch_stats(rg_table* t, uint64_t* d_out2);
int rg_launch_dump(rg_table* t, uint64_t min_count, uint64_t max_count, uint64_t* d_keys, uint32_t* d_counts,
                   uint64_t cap, uint64_t* d_cursor, uint32_t slice_mod = 0, uint32_t slice_rem = 0);
int rg_launch_query(rg_table* t, const uint64_t* d_keys, uint64_t n, uint64_t min_count, uint32_t* d_out);
struct OtherTables {
  TableView tv[16];
  uint32_t min_count[16];
  uint32_t same_geo;     // bit o: table o has the subject's tile geometry (a key sits in the same tile in both)
  int n;
};
int rg_launch_subtract(rg_table* subj, uint64_t lo, uint64_t hi, const OtherTables& others, uint64_t* d_keys,
                       uint32_t* d_counts, uint64_t cap, uint64_t* d_cursor);
int rg_launch_build_select(rg_table* subj, uint64_t lo, uint64_t hi, const OtherTables& others, uint64_t max_ctl_sum,
                           uint64_t* d_keys, uint32_t* d_counts, uint32_t* d_sums, uint64_t cap, uint64_t* d_cursor);
int rg_launch_rehash(rg_table* dst, const uint64_t* old_keys, const uint32_t* old_counts, uint64_t n_slots);

int rg_launch_set_build(rg_set* s, const uint64_t* d_vals, uint64_t n);
int rg_launch_jf_to_rufus(rg_ctx* ctx, const uint64_t* d_jf, uint64_t n, int k, uint64_t* d_out2n);
// pos_base: stream position of seq[0] (offs / read_len are absolute); zero_hits: clear hits[0..n_reads) first
int rg_launch_filter_hits(rg_set* s, const char* seq, const char* qual, const uint64_t* offs, uint32_t read_len,
                          uint64_t n_reads, uint64_t n_bytes, uint64_t pos_base, bool zero_hits, int minq, int skip_last,
                          int32_t* hits, uint8_t* hit_flags = nullptr);
int rg_launch_coverage(rg_ctx* ctx, const uint8_t* hit_flags, const uint64_t* offs, uint64_t n_reads, uint64_t n_bytes,
                       int k, uint8_t* cov);
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
