// rg_api.cu — the C ABI of librufus_gpu.so (include/rufus_gpu.h): context,
// table lifetime/growth, host<->device staging. No CPU compute path exists here:
// every entry point ends in a kernel launch from rg_count.cu / rg_ops.cu / rg_filter.cu.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "rg_internal.cuh"
#include "rg_sk.cuh"

static thread_local std::string g_err;
void rg_set_error(const std::string& msg) { g_err = msg; }

extern "C" {

const char* rg_last_error(void) { return g_err.c_str(); }
const char* rg_version(void) { return "rufus_b200 0.1 (sm_100a)"; }

// ---------------------------------------------------------------------------
rg_ctx* rg_init(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    rg_set_error(std::string("rg_init: no CUDA device (") + cudaGetErrorString(e) +
                 "); librufus_gpu has no CPU fallback");
    return nullptr;
  }
  if (device < 0 || device >= n) {
    rg_set_error("rg_init: bad device index");
    return nullptr;
  }
  RG_CUDA_OKP(cudaSetDevice(device));
  rg_ctx* ctx = new rg_ctx();
  ctx->device = device;
  cudaDeviceProp prop;
  RG_CUDA_OKP(cudaGetDeviceProperties(&prop, device));
  ctx->sm_count = prop.multiProcessorCount;
  RG_CUDA_OKP(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
  ctx->stream = ctx->own_stream;
  RG_CUDA_OKP(cudaMallocHost(&ctx->h_scratch, 256 * sizeof(uint64_t)));
  RG_CUDA_OKP(cudaMalloc(&ctx->d_scratch, 256 * sizeof(uint64_t)));
  if (const char* fg = getenv("RG_GROUPS_LOCAL")) ctx->force_groups = (uint32_t)atoi(fg);
  RG_CUDA_OKP(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    RG_CUDA_OKP(cudaEventCreateWithFlags(&ctx->slot_copied[i], cudaEventDisableTiming));
    RG_CUDA_OKP(cudaEventCreateWithFlags(&ctx->slot_free[i], cudaEventDisableTiming));
  }
  return ctx;
}

void rg_shutdown(rg_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  cudaStreamSynchronize(ctx->copy_stream);
  cudaFreeHost(ctx->h_scratch);
  cudaFree(ctx->d_scratch);
  for (int i = 0; i < 2; ++i) {
    cudaFree(ctx->slot[i]);
    cudaEventDestroy(ctx->slot_copied[i]);
    cudaEventDestroy(ctx->slot_free[i]);
  }
  for (int i = 0; i < 6; ++i) cudaFree(ctx->pool[i]);
  for (cudaEvent_t e : ctx->event_pool) cudaEventDestroy(e);
  cudaStreamDestroy(ctx->copy_stream);
  cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

int rg_set_stream(rg_ctx* ctx, void* cuda_stream) {
  if (!ctx) return RG_ERR_ARG;
  RG_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return 0;
}

// same without waiting for the work already enqueued on the previous stream: the caller orders the
// two streams itself (events); lets one context feed two streams that run concurrently
int rg_use_stream(rg_ctx* ctx, void* cuda_stream) {
  if (!ctx) return RG_ERR_ARG;
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return 0;
}

int rg_sync(rg_ctx* ctx) {
  cudaSetDevice(ctx->device);
  RG_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

uint64_t rg_launch_count(rg_ctx* ctx) { return ctx ? ctx->launches : 0; }

static void prof_collect(rg_ctx* ctx) {
  cudaStreamSynchronize(ctx->stream);
  for (auto& kv : ctx->prof) {
    for (auto& ev : kv.second.pending) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, ev.first, ev.second) == cudaSuccess) {
        kv.second.ms += ms;
        kv.second.n++;
      }
      ctx->event_pool.push_back(ev.first);
      ctx->event_pool.push_back(ev.second);
    }
    kv.second.pending.clear();
  }
}

int rg_profile_enable(rg_ctx* ctx, int on) {
  if (!ctx) return RG_ERR_ARG;
  prof_collect(ctx);
  ctx->profiling = on != 0;
  return 0;
}

int rg_profile_reset(rg_ctx* ctx) {
  if (!ctx) return RG_ERR_ARG;
  prof_collect(ctx);
  ctx->prof.clear();
  return 0;
}

int rg_profile_get(rg_ctx* ctx, const char* name, double* total_ms, uint64_t* n_launches) {
  if (!ctx || !name) return RG_ERR_ARG;
  prof_collect(ctx);
  auto it = ctx->prof.find(name);
  if (it == ctx->prof.end()) return RG_ERR_ARG;
  if (total_ms) *total_ms = it->second.ms;
  if (n_launches) *n_launches = it->second.n;
  return 0;
}

int rg_profile_names(rg_ctx* ctx, char* buf, uint64_t cap) {
  if (!ctx || !buf || cap == 0) return RG_ERR_ARG;
  std::string s;
  for (auto& kv : ctx->prof) s += kv.first + "\n";
  const uint64_t n = std::min<uint64_t>(cap - 1, s.size());
  memcpy(buf, s.data(), n);
  buf[n] = 0;
  return 0;
}

void* rg_dev_alloc(rg_ctx* ctx, uint64_t bytes) {
  void* p = nullptr;
  cudaSetDevice(ctx->device);
  RG_CUDA_OKP(cudaMalloc(&p, bytes ? bytes : 16));
  return p;
}
void rg_dev_free(rg_ctx* ctx, void* p) {
  if (p) {
    cudaStreamSynchronize(ctx->stream);
    cudaFree(p);
  }
}
void* rg_host_alloc(uint64_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 16) != cudaSuccess) {
    rg_set_error("rg_host_alloc: cudaMallocHost failed");
    return nullptr;
  }
  return p;
}
void rg_host_free(void* p) {
  if (p) cudaFreeHost(p);
}
int rg_h2d(rg_ctx* ctx, void* dst, const void* src, uint64_t bytes) {
  cudaSetDevice(ctx->device);
  RG_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}
int rg_d2h(rg_ctx* ctx, void* dst, const void* src, uint64_t bytes) {
  RG_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  RG_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// table internals
// ---------------------------------------------------------------------------
static int read_scalars(rg_ctx* ctx, const uint64_t* d, int n, uint64_t* out) {
  RG_CUDA_OK(cudaMemcpyAsync(ctx->h_scratch, d, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  RG_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < n; ++i) out[i] = ctx->h_scratch[i];
  return 0;
}

// ---------------------------------------------------------------------------
// host -> device pipeline: two staging slots. slot_acquire() makes the copy stream wait until the
// kernels that last read the slot are done; slot_publish() makes `stream` wait for the slot's copy;
// slot_release() marks the point on `stream` after which the slot may be overwritten.
// ---------------------------------------------------------------------------
static const uint64_t RG_SLOT_BYTES = 32ull << 20;

static int slots_ensure(rg_ctx* ctx) {
  if (ctx->slot_bytes) return 0;
  for (int i = 0; i < 2; ++i) {
    RG_CUDA_OK(cudaMalloc(&ctx->slot[i], RG_SLOT_BYTES + 256));
    RG_CUDA_OK(cudaEventRecord(ctx->slot_free[i], ctx->stream));
  }
  ctx->slot_bytes = RG_SLOT_BYTES;
  return 0;
}
static int slot_acquire(rg_ctx* ctx, uint32_t* which) {
  const uint32_t i = ctx->slot_next++ & 1u;
  RG_CUDA_OK(cudaStreamWaitEvent(ctx->copy_stream, ctx->slot_free[i], 0));
  *which = i;
  return 0;
}
static int slot_publish(rg_ctx* ctx, uint32_t i) {
  RG_CUDA_OK(cudaEventRecord(ctx->slot_copied[i], ctx->copy_stream));
  RG_CUDA_OK(cudaStreamWaitEvent(ctx->stream, ctx->slot_copied[i], 0));
  return 0;
}
static int slot_release(rg_ctx* ctx, uint32_t i) {
  RG_CUDA_OK(cudaEventRecord(ctx->slot_free[i], ctx->stream));
  return 0;
}
// grow-only device scratch, index 0..5
static int pool_ensure(rg_ctx* ctx, int i, uint64_t bytes) {
  if (ctx->pool_bytes[i] >= bytes) return 0;
  if (ctx->pool[i]) {
    RG_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->pool[i]);
    ctx->pool[i] = nullptr;
    ctx->pool_bytes[i] = 0;
  }
  const uint64_t want = (bytes + (bytes >> 2) + 4096) & ~255ull;
  RG_CUDA_OK(cudaMalloc(&ctx->pool[i], want));
  ctx->pool_bytes[i] = want;
  return 0;
}

// Tile count and group geometry for at least `wanted` tiles (header of rg_internal.cuh):
//  * direct mode: no constraint;
//  * tiled mode, small table: group_tiles = 1 (one partition pass), whole waves of tile-insert CTAs;
//  * tiled mode, more tiles than one sorted pass can histogram in shared memory: one group per SM
//    (two passes: stream -> groups -> tiles);
//  * peer-to-peer multi-GPU counting: p2p_groups groups per shard (the level-1 pass sorts into the
//    groups of ALL shards at once).
static void choose_geometry(const rg_table* t, uint32_t wanted, uint32_t* n_tiles, uint32_t* group_tiles) {
  const rg_ctx* ctx = t->ctx;
  const uint32_t sm = (uint32_t)ctx->sm_count;
  *group_tiles = 1;
  *n_tiles = wanted ? wanted : 1;
  if (t->mode != RG_MODE_TILED) return;
  const uint32_t max1 = rg_partition_max_buckets();
  uint32_t groups = 0;
  if (t->p2p_groups) groups = t->p2p_groups;
  else if (ctx->force_groups) groups = ctx->force_groups;
  else if (wanted > max1) groups = std::min(sm, max1);   // measured on 9472 tiles: 148 groups 5.3+4.9 ms, 1184 groups 6.0+5.1 ms
  if (groups) {
    if (groups > wanted) groups = wanted ? wanted : 1;
    *group_tiles = (wanted + groups - 1) / groups;
    *n_tiles = groups * *group_tiles;
    return;
  }
  if (wanted > sm) {
    const uint32_t rounded = (wanted + sm - 1) / sm * sm;   // tiled insert runs one CTA per tile: whole waves
    *n_tiles = rounded <= max1 ? rounded : wanted;
  }
}

static void drop_buckets(rg_table* t) {
  if (t->part_cursor) { cudaFree(t->part_cursor); t->part_cursor = nullptr; }
  if (t->part_keys) { cudaFree(t->part_keys); t->part_keys = nullptr; t->part_cap = 0; }
  if (!t->p2p_block) {   // (the peer-to-peer block is exported: it lives as long as the table)
    if (t->grp_cursor) { cudaFree(t->grp_cursor); t->grp_cursor = nullptr; }
    if (t->grp_keys) { cudaFree(t->grp_keys); t->grp_keys = nullptr; t->grp_alloc_keys = 0; }
  }
}

// overflow region of a table with main_slots tile slots: 1/16 of them, at least one tile's worth (power of two)
static uint64_t ovf_for(uint64_t main_slots) {
  uint64_t o = RG_TILE_SLOTS;
  while (o < main_slots / 16) o <<= 1;
  return o;
}

static int table_alloc_tiles(rg_table* t, uint32_t wanted, uint64_t ovf_slots = 0) {
  rg_ctx* ctx = t->ctx;
  cudaSetDevice(ctx->device);
  uint32_t n_tiles, group_tiles;
  choose_geometry(t, wanted, &n_tiles, &group_tiles);
  const uint64_t main_slots = (uint64_t)n_tiles << RG_TILE_BITS;
  if (ovf_slots < ovf_for(main_slots)) ovf_slots = ovf_for(main_slots);
  const uint64_t slots = main_slots + ovf_slots;
  RG_CUDA_OK(cudaMalloc(&t->keys, slots * 8));
  RG_CUDA_OK(cudaMalloc(&t->counts, slots * 4));
  RG_CUDA_OK(cudaMemsetAsync(t->keys, 0xFF, slots * 8, ctx->stream));
  RG_CUDA_OK(cudaMemsetAsync(t->counts, 0, slots * 4, ctx->stream));
  t->n_tiles = n_tiles;
  t->group_tiles = group_tiles;
  t->n_tiles_global = n_tiles * t->shard_parts;
  t->tile_first = n_tiles * t->shard_part;
  t->ovf_slots = ovf_slots;
  t->ovf_used = 0;
  // (a peer-to-peer table keeps its minimizer length for life: the other GPUs route records with it)
  if (!t->p2p_block) t->m = sk_choose_m(t->k, t->n_tiles_global);
  return 0;
}

static uint32_t tiles_for(uint64_t est_distinct) {
  // aim for <= 60 % load
  uint64_t n = (est_distinct * 5 / 3 + RG_TILE_SLOTS - 1) >> RG_TILE_BITS;
  if (n < 1) n = 1;
  if (n > (1u << 22)) n = 1u << 22;
  return (uint32_t)n;
}

// rebuild with new_tiles tiles (and at least new_ovf overflow slots) and re-insert every (key,count)
static int table_grow(rg_table* t, uint32_t new_tiles, uint64_t new_ovf = 0) {
  rg_ctx* ctx = t->ctx;
  cudaSetDevice(ctx->device);
  uint64_t* old_keys = t->keys;
  uint32_t* old_counts = t->counts;
  const uint64_t old_slots = rg_table_slots(t);
  // (a peer-to-peer table grows like any other: the number of hash-prefix groups, hence the exported slabs
  // and what the other GPUs send where, stays; only the tiles per group change, which is local)
  t->keys = nullptr;
  t->counts = nullptr;
  int rc = table_alloc_tiles(t, new_tiles, new_ovf);
  if (rc) {
    // no room for the larger table next to the old one: the table stays exactly what it was (table_alloc_tiles sets
    // the geometry only after both allocations), the caller gets a message that says what was being attempted
    const std::string why = rg_last_error();
    cudaFree(t->keys);
    cudaFree(t->counts);
    t->keys = old_keys;
    t->counts = old_counts;
    (void)cudaGetLastError();       // (a failed cudaMalloc leaves its error behind for the next launch check)
    rg_set_error("count: cannot grow the table to " + std::to_string(new_tiles) + " tiles (" +
                 std::to_string(((((uint64_t)new_tiles << RG_TILE_BITS) + new_ovf) * 12) >> 20) +
                 " MiB next to the " + std::to_string((old_slots * 12) >> 20) + " MiB it holds): " + why);
    return rc;
  }
  // bucket staging is sized per tile count: drop it, it is rebuilt on the next push
  drop_buckets(t);
  RG_CUDA_OK(cudaMemsetAsync(t->d_status, 0, RG_ST_WORDS * sizeof(uint64_t), ctx->stream));
  rc = rg_launch_rehash(t, old_keys, old_counts, old_slots);
  if (rc) return rc;
  uint64_t st[RG_ST_WORDS];
  rc = read_scalars(ctx, t->d_status, RG_ST_WORDS, st);
  cudaFree(old_keys);
  cudaFree(old_counts);
  if (rc) return rc;
  if (st[RG_ST_LOST]) {      // (cannot happen: the new table is larger than what it has to hold)
    rg_set_error("count: rehash overflow");
    return RG_ERR_FULL;
  }
  t->ovf_used = st[RG_ST_OVF_NEW];
  RG_CUDA_OK(cudaMemsetAsync(t->d_status + RG_ST_OVF_NEW, 0, sizeof(uint64_t), ctx->stream));
  t->tiles_dirty = true;
  return 0;
}

// the spill list (last resort of a push that found tiles AND overflow region full) holds at least
// min(n_keys / 4, 2^27) pairs, 4 Mi at the least
static int ensure_spill(rg_table* t, uint64_t n_keys) {
  uint64_t want = std::max<uint64_t>(1ull << 22, std::min<uint64_t>(n_keys / 4, 1ull << 27));
  if (t->spill_cap >= want) return 0;
  rg_ctx* ctx = t->ctx;
  cudaSetDevice(ctx->device);
  if (t->spill_keys) {
    int rc = rg_table_finish(t);          // (entries of earlier pushes are resolved first)
    if (rc) return rc;
    RG_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    cudaFree(t->spill_keys);
    cudaFree(t->spill_counts);
    t->spill_keys = nullptr;
    t->spill_counts = nullptr;
    t->spill_cap = 0;
  }
  RG_CUDA_OK(cudaMalloc(&t->spill_keys, want * 8));
  RG_CUDA_OK(cudaMalloc(&t->spill_counts, want * 4));
  t->spill_cap = want;
  return 0;
}

// after a push: read the status words; re-insert what found no slot after growing the table; keep the
// overflow region at most half full
static int table_check_status(rg_table* t) {
  rg_ctx* ctx = t->ctx;
  cudaSetDevice(ctx->device);
  for (int round = 0; round < 10; ++round) {
    uint64_t st[RG_ST_WORDS];
    int rc = read_scalars(ctx, t->d_status, RG_ST_WORDS, st);
    if (rc) return rc;
    if (st[RG_ST_FOREIGN] != 0) {
      rg_set_error("count: " + std::to_string(st[RG_ST_FOREIGN]) + " keys do not belong to this table's shard");
      return RG_ERR_ARG;
    }
    if (st[RG_ST_LOST] > t->spill_cap) {
      rg_set_error("count: table full, " + std::to_string(st[RG_ST_LOST]) +
                   " k-mer occurrences found no slot and the spill list overflowed; raise est_distinct or "
                   "call rg_table_reserve before the push");
      return RG_ERR_FULL;
    }
    if (st[RG_ST_OVF_NEW]) {
      t->ovf_used += st[RG_ST_OVF_NEW];
      RG_CUDA_OK(cudaMemsetAsync(t->d_status + RG_ST_OVF_NEW, 0, sizeof(uint64_t), ctx->stream));
    }
    const uint64_t n_spill = st[RG_ST_LOST];
    if (n_spill == 0 && t->ovf_used * 2 <= t->ovf_slots) return 0;
    // Either the tiles are too loaded (more tiles), or a few minimizers own more k-mers than a tile holds
    // (a larger overflow region). The spilled pairs are moved aside and re-inserted into the rebuilt table.
    uint64_t* tmp_k = nullptr;
    uint32_t* tmp_c = nullptr;
    if (n_spill) {
      RG_CUDA_OK(cudaMalloc(&tmp_k, n_spill * 8));
      RG_CUDA_OK(cudaMalloc(&tmp_c, n_spill * 4));
      RG_CUDA_OK(cudaMemcpyAsync(tmp_k, t->spill_keys, n_spill * 8, cudaMemcpyDeviceToDevice, ctx->stream));
      RG_CUDA_OK(cudaMemcpyAsync(tmp_c, t->spill_counts, n_spill * 4, cudaMemcpyDeviceToDevice, ctx->stream));
      RG_CUDA_OK(cudaMemsetAsync(t->d_status + RG_ST_LOST, 0, sizeof(uint64_t), ctx->stream));
    }
    uint64_t distinct = 0;
    rc = rg_launch_stats(t, ctx->d_scratch);
    if (rc == 0) rc = read_scalars(ctx, ctx->d_scratch, 1, &distinct);
    if (rc == 0) {
      const uint64_t main_slots = (uint64_t)t->n_tiles << RG_TILE_BITS;
      uint32_t want_tiles = t->n_tiles;
      // (every spilled pair may be a new key)
      if ((distinct + n_spill) * 2 > main_slots)
        want_tiles = std::max<uint32_t>(t->n_tiles * 2, tiles_for((distinct + n_spill) * 2));
      rc = table_grow(t, want_tiles, t->ovf_slots * 2);
    }
    if (rc == 0 && n_spill) rc = rg_launch_insert_keys(t, tmp_k, n_spill, tmp_c);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(tmp_k);
    cudaFree(tmp_c);
    if (rc) return rc;
  }
  rg_set_error("count: could not make room after growing 10 times");
  return RG_ERR_FULL;
}

static int ensure_stage(rg_table* t, uint64_t bytes) {
  if (t->stage_bytes >= bytes) return 0;
  if (t->stage) {
    RG_CUDA_OK(cudaStreamSynchronize(t->ctx->stream));
    cudaFree(t->stage);
    t->stage = nullptr;
  }
  const uint64_t want = (bytes + (1 << 20)) & ~((1ull << 20) - 1);
  RG_CUDA_OK(cudaMalloc(&t->stage, want));
  t->stage_bytes = want;
  return 0;
}

// (re)size the per-tile bucket slabs (super-k-mer records) for a push of n_bytes stream bytes and zero the
// cursors. Records that find their slab full are inserted straight into the table by the partition pass, so
// the sizing is a speed matter only.
static int prepare_buckets(rg_table* t, uint64_t n_bytes) {
  rg_ctx* ctx = t->ctx;
  cudaSetDevice(ctx->device);
  const uint64_t n_rec = (uint64_t)((double)n_bytes * rg_partition_record_density(t->k, t->m)) + 1;
  // bucket capacity: mean records per tile + 30 % + slack
  const uint64_t mean = n_rec / t->n_tiles + 1;
  uint64_t cap = mean + mean * 3 / 10 + 128;
  cap = (cap + 1) & ~1ull;  // slabs start on a 32-byte sector boundary
  if (!t->part_cursor) RG_CUDA_OK(cudaMalloc(&t->part_cursor, (uint64_t)t->n_tiles * sizeof(uint32_t)));
  if (t->part_cap < cap * t->n_tiles) {
    if (t->part_keys) {
      RG_CUDA_OK(cudaStreamSynchronize(ctx->stream));
      cudaFree(t->part_keys);
      t->part_keys = nullptr;
    }
    RG_CUDA_OK(cudaMalloc(&t->part_keys, cap * t->n_tiles * 16 + 64));
    t->part_cap = cap * t->n_tiles;
  }
  t->bucket_cap = (uint32_t)(t->part_cap / t->n_tiles);
  RG_CUDA_OK(cudaMemsetAsync(t->part_cursor, 0, (uint64_t)t->n_tiles * sizeof(uint32_t), ctx->stream));
  // two-level partition: the slabs of the groups (a peer-to-peer table owns fixed ones)
  if (t->group_tiles > 1 && !t->p2p_block) {
    const uint32_t n_groups = t->n_tiles / t->group_tiles;
    const uint64_t gmean = n_rec / n_groups + 1;
    uint64_t gcap = gmean + gmean / 4 + 1024;
    gcap = (gcap + 1) & ~1ull;
    if (!t->grp_cursor) RG_CUDA_OK(cudaMalloc(&t->grp_cursor, (uint64_t)n_groups * sizeof(uint32_t)));
    if (t->grp_alloc_keys < gcap * n_groups) {
      if (t->grp_keys) {
        RG_CUDA_OK(cudaStreamSynchronize(ctx->stream));
        cudaFree(t->grp_keys);
        t->grp_keys = nullptr;
      }
      RG_CUDA_OK(cudaMalloc(&t->grp_keys, gcap * n_groups * 16 + 64));
      t->grp_alloc_keys = gcap * n_groups;
    }
    t->grp_cap = (uint32_t)(t->grp_alloc_keys / n_groups);
    RG_CUDA_OK(cudaMemsetAsync(t->grp_cursor, 0, (uint64_t)n_groups * sizeof(uint32_t), ctx->stream));
  }
  return 0;
}

// after the level-1 launches of a push: (group slabs -> tile buckets,) then the tile inserts
static int finish_tiled_push(rg_table* t) {
  int rc = 0;
  if (t->group_tiles > 1) rc = rg_launch_partition_groups(t);
  if (rc == 0) rc = rg_launch_tile_insert(t);
  return rc;
}

static int push_dev_piece(rg_table* t, const char* dev_stream, uint64_t n_bytes, uint32_t read_stride,
                          const uint32_t* brk, uint32_t skip) {
  ExtractParams ep;
  ep.stream = dev_stream;
  ep.n_bytes = n_bytes;
  ep.brk = brk;
  ep.read_stride = read_stride;
  ep.k = t->k;
  ep.canonical = t->canonical;
  ep.skip = skip;
  int rc = ensure_spill(t, n_bytes);
  if (rc) return rc;
  if (t->mode == RG_MODE_TILED) {
    rc = prepare_buckets(t, n_bytes);
    if (rc) return rc;
    rc = rg_launch_partition(t, ep);
    if (rc) return rc;
    rc = finish_tiled_push(t);
    if (rc) return rc;
  } else {
    rc = rg_launch_count_direct(t, ep);
    if (rc) return rc;
  }
  t->tiles_dirty = true;
  t->unchecked = true;      // status words are read back by rg_table_finish (rg_count_end or the first reader)
  return 0;
}

// A push larger than RG_PUSH_SPLIT bytes is counted in pieces (the bucket slabs are indexed with 32
// bits). Fixed-stride input is cut at read boundaries; otherwise a piece starts 32 bytes early (the
// stream pointer and the read-start bitmap stay aligned) and skips the windows that start inside that
// halo and were counted with the previous piece — except those reaching into the new bytes.
static uint64_t g_push_split = 1ull << 31;
static int push_dev_impl(rg_table* t, const char* dev_stream, uint64_t n_bytes, uint32_t read_stride,
                         const uint32_t* brk) {
  if (n_bytes == 0) return 0;
  if (((uintptr_t)dev_stream & 15) != 0 || read_stride >= (1u << 31)) {
    rg_set_error("count: device stream must be 16-byte aligned (and read_stride < 2^31)");
    return RG_ERR_ARG;
  }
  if (const char* e = getenv("RG_PUSH_SPLIT")) g_push_split = std::max<uint64_t>(4096, strtoull(e, nullptr, 10));
  if (n_bytes <= g_push_split) return push_dev_piece(t, dev_stream, n_bytes, read_stride, brk, 0);
  if (read_stride) {
    // pieces of whole reads whose starts stay 16-byte aligned: multiples of lcm(stride, 16)
    uint64_t unit = read_stride;
    while (unit % 16) unit += read_stride;
    const uint64_t piece = std::max<uint64_t>(unit, g_push_split / unit * unit);
    for (uint64_t off = 0; off < n_bytes; off += piece) {
      const int rc = push_dev_piece(t, dev_stream + off, std::min(piece, n_bytes - off), read_stride, nullptr, 0);
      if (rc) return rc;
    }
    return 0;
  }
  const uint64_t piece = g_push_split & ~31ull;
  for (uint64_t off = 0; off < n_bytes; off += piece) {
    const uint64_t halo = off ? 32 : 0;
    // windows starting at halo positions < 32-(k-1) lie entirely in the previous piece
    const uint32_t skip = off ? (uint32_t)(32 - (t->k - 1)) : 0u;
    const int rc = push_dev_piece(t, dev_stream + off - halo, std::min(piece, n_bytes - off) + halo, 0,
                                  brk ? brk + (off - halo) / 32 : nullptr, t->k - 1 >= 32 ? 0u : skip);
    if (rc) return rc;
  }
  return 0;
}

// every reader of a table goes through this: resolve what the enqueued pushes left behind
static int p2p_collect_spill(rg_table* t) {
  rg_ctx* ctx = t->ctx;
  cudaSetDevice(ctx->device);
  t->p2p_pending = false;
  uint64_t n_spill = 0;
  int rc = read_scalars(ctx, t->p2p_status, 1, &n_spill);
  if (rc) return rc;
  if (n_spill > t->p2p_spill_cap) {
    rg_set_error("count: peer-to-peer spill list overflowed; raise max_keys_per_push");
    return RG_ERR_FULL;
  }
  if (n_spill) return rg_launch_insert_records(t, t->p2p_spill, n_spill);
  return 0;
}

int rg_table_finish(rg_table* t) {
  if (!t->unchecked || t->ctx->defer_checks) return 0;
  if (t->p2p_pending) {
    const int prc = p2p_collect_spill(t);
    if (prc) return prc;
  }
  const int rc = table_check_status(t);
  if (rc == 0) t->unchecked = false;
  return rc;
}

// enqueue partition launches for one staged piece of a host stream (tiled mode) or the fused
// count (direct mode); the tile insert of the tiled mode is launched once per push by the caller
static int push_piece(rg_table* t, const char* dev, uint64_t n_bytes) {
  ExtractParams ep;
  ep.stream = dev;
  ep.n_bytes = n_bytes;
  ep.brk = nullptr;
  ep.read_stride = 0;
  ep.k = t->k;
  ep.canonical = t->canonical;
  return t->mode == RG_MODE_TILED ? rg_launch_partition(t, ep) : rg_launch_count_direct(t, ep);
}

// device temporaries of one entry point: freed on every way out (an RG_CUDA_OK that returns early included)
struct DevTemps {
  void* p[4] = {nullptr, nullptr, nullptr, nullptr};
  int n = 0;
  DevTemps() = default;
  DevTemps(const DevTemps&) = delete;
  DevTemps& operator=(const DevTemps&) = delete;
  ~DevTemps() {
    for (int i = 0; i < n; ++i) cudaFree(p[i]);
  }
  template <typename T>
  cudaError_t alloc(T** out, size_t bytes) {
    const cudaError_t e = cudaMalloc(out, bytes);
    if (e == cudaSuccess && n < 4) p[n++] = *out;
    return e;
  }
};

extern "C" {

rg_table* rg_count_begin(rg_ctx* ctx, int k, int canonical, uint64_t est_distinct, int mode) {
  if (!ctx) {
    rg_set_error("rg_count_begin: null context");
    return nullptr;
  }
  if (k < 1 || k > 32 || (!canonical && k == 32)) {
    rg_set_error("rg_count_begin: k must be 1..32 (1..31 when not canonical)");
    return nullptr;
  }
  cudaSetDevice(ctx->device);
  rg_table* t = new rg_table();
  t->ctx = ctx;
  t->k = k;
  t->canonical = canonical ? 1 : 0;
  t->mode = (mode == RG_MODE_AUTO) ? RG_MODE_TILED : mode;
  if (est_distinct == 0) est_distinct = 1ull << 24;
  if (table_alloc_tiles(t, tiles_for(est_distinct)) != 0) {
    delete t;
    return nullptr;
  }
  if (cudaMalloc(&t->d_status, RG_ST_WORDS * sizeof(uint64_t)) != cudaSuccess || ensure_spill(t, 0) != 0) {
    rg_set_error("rg_count_begin: out of device memory");
    rg_table_free(t);
    return nullptr;
  }
  cudaMemsetAsync(t->d_status, 0, RG_ST_WORDS * sizeof(uint64_t), ctx->stream);
  return t;
}

int rg_table_set_shard(rg_table* t, uint32_t part, uint32_t n_parts) {
  if (!t || n_parts == 0 || part >= n_parts || t->tiles_dirty) {
    rg_set_error("rg_table_set_shard: bad arguments (must be called before the first push)");
    return RG_ERR_ARG;
  }
  t->shard_part = part;
  t->shard_parts = n_parts;
  t->n_tiles_global = t->n_tiles * n_parts;
  t->tile_first = t->n_tiles * part;
  t->m = sk_choose_m(t->k, t->n_tiles_global);     // (the same on every shard: it depends on the global tile count)
  return 0;
}

// ---------------------------------------------------------------------------
// Peer-to-peer multi-GPU counting (one process per GPU, NVLink peer memory, no collective in the
// data path): every GPU's level-1 partition pass sorts its reads into the groups of ALL shards and
// writes each run straight into the OWNER's group slab (system-scope atomics reserve it), then
// every GPU finishes its own groups locally (groups -> tiles -> tile inserts).
//   rg_table_p2p_prepare : after rg_table_set_shard, before the first push; fixes the geometry and
//                          allocates the exported block [cursors | status | spill | slabs]
//   rg_table_p2p_export  : 64-byte CUDA IPC handle of that block (all-gather it between the ranks)
//   rg_table_p2p_import  : attach rank r's block (r == own rank: local pointers)
//   per sample: rg_count_p2p_begin (zero my cursors) -> barrier -> rg_count_p2p_scatter_dev ->
//               barrier -> rg_count_p2p_finish
// ---------------------------------------------------------------------------
static void p2p_layout(const rg_table* t, uint64_t* off_status, uint64_t* off_spill, uint64_t* off_slabs,
                       uint64_t* total) {
  const uint64_t n_groups = (uint64_t)(t->n_tiles / t->group_tiles) * t->shard_parts;   // (group, source GPU) sub-slabs
  uint64_t o = ((n_groups * 4 + 255) / 256) * 256;
  *off_status = o;
  o += 256;
  *off_spill = o;
  o += ((t->p2p_spill_cap * 16 + 255) / 256) * 256;
  *off_slabs = o;
  o += n_groups * (uint64_t)t->grp_cap * 16 + 256;
  *total = o;
}
static void p2p_pointers(const rg_table* t, void* base, uint32_t r, rg_table* into) {
  uint64_t os, osp, osl, tot;
  p2p_layout(t, &os, &osp, &osl, &tot);
  char* b = static_cast<char*>(base);
  into->peer_cursor[r] = reinterpret_cast<uint32_t*>(b);
  into->peer_status[r] = reinterpret_cast<uint64_t*>(b + os);
  into->peer_spill[r] = reinterpret_cast<uint4*>(b + osp);
  into->peer_keys[r] = reinterpret_cast<uint4*>(b + osl);
}

int rg_table_p2p_prepare(rg_table* t, uint32_t groups_per_shard, uint64_t max_keys_per_push) {
  if (!t || t->tiles_dirty || t->mode != RG_MODE_TILED || t->shard_parts > 8 || groups_per_shard == 0 ||
      max_keys_per_push == 0 || t->p2p_block) {
    rg_set_error("rg_table_p2p_prepare: needs an empty tiled table of at most 8 shards, once");
    return RG_ERR_ARG;
  }
  rg_ctx* ctx = t->ctx;
  cudaSetDevice(ctx->device);
  if ((uint64_t)groups_per_shard * t->shard_parts > rg_partition_max_buckets()) {
    rg_set_error("rg_table_p2p_prepare: too many groups for one partition pass");
    return RG_ERR_ARG;
  }
  // re-shape the (empty) table for the group geometry
  RG_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  const uint32_t wanted = t->n_tiles;
  cudaFree(t->keys);
  cudaFree(t->counts);
  t->keys = nullptr;
  t->counts = nullptr;
  drop_buckets(t);
  // at least two tiles per group: the sub-slabs of a group (one per source GPU) are merged by the
  // groups -> tiles pass
  t->p2p_groups = std::max<uint32_t>(1, std::min<uint32_t>(groups_per_shard, std::max<uint32_t>(wanted, 2) / 2));
  int rc = table_alloc_tiles(t, std::max<uint32_t>(wanted, 2));
  if (rc) return rc;
  // the minimizer length is fixed from here on (every GPU computes the same value: same k, same global tile
  // count); leave room for the table to grow 16x
  t->m = sk_choose_m(t->k, (uint64_t)t->n_tiles_global * 16);
  const uint64_t n_groups = (uint64_t)(t->n_tiles / t->group_tiles) * t->shard_parts;      // sub-slabs I own
  const uint64_t n_rec = (uint64_t)((double)max_keys_per_push * rg_partition_record_density(t->k, t->m)) + 1;
  const uint64_t gmean = n_rec / n_groups + 1;
  t->grp_cap = (uint32_t)((gmean + gmean / 2 + 1024 + 3) & ~3ull);
  t->p2p_spill_cap = std::max<uint64_t>(1ull << 20, max_keys_per_push / 16);
  t->p2p_max_keys = max_keys_per_push;
  uint64_t os, osp, osl, tot;
  p2p_layout(t, &os, &osp, &osl, &tot);
  RG_CUDA_OK(cudaMalloc(&t->p2p_block, tot));
  RG_CUDA_OK(cudaMemsetAsync(t->p2p_block, 0, osl, ctx->stream));
  t->p2p_block_bytes = tot;
  p2p_pointers(t, t->p2p_block, t->shard_part, t);
  t->grp_cursor = t->peer_cursor[t->shard_part];
  t->grp_keys = t->peer_keys[t->shard_part];
  t->p2p_status = t->peer_status[t->shard_part];
  t->p2p_spill = t->peer_spill[t->shard_part];
  t->grp_alloc_keys = n_groups * (uint64_t)t->grp_cap;
  RG_CUDA_OK(cudaMalloc(&t->p2p_src_cursor, (uint64_t)t->n_tiles_global / t->group_tiles * 4));
  RG_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int rg_table_p2p_export(rg_table* t, void* handle64) {
  if (!t || !t->p2p_block || !handle64) return RG_ERR_ARG;
  cudaIpcMemHandle_t h;
  RG_CUDA_OK(cudaIpcGetMemHandle(&h, t->p2p_block));
  static_assert(sizeof(h) == 64, "CUDA IPC handle size");
  memcpy(handle64, &h, 64);
  return 0;
}

int rg_table_p2p_import(rg_table* t, uint32_t rank, const void* handle64) {
  if (!t || !t->p2p_block || rank >= t->shard_parts || !handle64) return RG_ERR_ARG;
  if (rank == t->shard_part) return 0;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* base = nullptr;
  RG_CUDA_OK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
  t->peer_mapped[rank] = base;
  p2p_pointers(t, base, rank, t);
  return 0;
}

// One process driving several GPUs (no IPC): attach every table's exported block to every other table
// directly, after enabling peer access between their devices. tabs[r] must be shard r of n.
int rg_tables_p2p_attach(rg_table* const* tabs, int n) {
  if (!tabs || n < 1 || n > 8) return RG_ERR_ARG;
  for (int i = 0; i < n; ++i)
    if (!tabs[i] || !tabs[i]->p2p_block || tabs[i]->shard_parts != (uint32_t)n || tabs[i]->shard_part != (uint32_t)i) {
      rg_set_error("rg_tables_p2p_attach: tabs[r] must be the prepared shard r of n");
      return RG_ERR_ARG;
    }
  for (int i = 0; i < n; ++i) {
    RG_CUDA_OK(cudaSetDevice(tabs[i]->ctx->device));
    for (int j = 0; j < n; ++j) {
      if (i == j) continue;
      const int dj = tabs[j]->ctx->device;
      if (dj != tabs[i]->ctx->device) {
        int can = 0;
        RG_CUDA_OK(cudaDeviceCanAccessPeer(&can, tabs[i]->ctx->device, dj));
        if (!can) {
          rg_set_error("rg_tables_p2p_attach: no peer access between the devices");
          return RG_ERR_CUDA;
        }
        const cudaError_t e = cudaDeviceEnablePeerAccess(dj, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
          rg_set_error(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
          return RG_ERR_CUDA;
        }
        cudaGetLastError();
      }
      p2p_pointers(tabs[j], tabs[j]->p2p_block, (uint32_t)j, tabs[i]);
    }
  }
  return 0;
}

int rg_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int rg_device_memory(rg_ctx* ctx, uint64_t* free_bytes, uint64_t* total_bytes) {
  if (!ctx) return RG_ERR_ARG;
  cudaSetDevice(ctx->device);
  size_t fr = 0, tot = 0;
  RG_CUDA_OK(cudaMemGetInfo(&fr, &tot));
  if (free_bytes) *free_bytes = (uint64_t)fr;
  if (total_bytes) *total_bytes = (uint64_t)tot;
  return 0;
}

int rg_count_p2p_begin(rg_table* t) {
  if (!t || !t->p2p_block) return RG_ERR_ARG;
  cudaSetDevice(t->ctx->device);
  RG_CUDA_OK(cudaMemsetAsync(t->p2p_src_cursor, 0, (uint64_t)t->n_tiles_global / t->group_tiles * 4, t->ctx->stream));
  RG_CUDA_OK(cudaMemsetAsync(t->p2p_status, 0, 64, t->ctx->stream));
  return 0;
}

int rg_count_p2p_scatter_dev(rg_table* t, const char* dev_stream, uint64_t n_bytes, uint32_t read_stride,
                             uint32_t max_ctas) {
  if (!t || !t->p2p_block) return RG_ERR_ARG;
  cudaSetDevice(t->ctx->device);
  t->p2p_ctas = max_ctas;
  for (uint32_t r = 0; r < t->shard_parts; ++r)
    if (!t->peer_keys[r]) {
      rg_set_error("rg_count_p2p_scatter_dev: not every peer is imported");
      return RG_ERR_ARG;
    }
  if (((uintptr_t)dev_stream & 15) != 0 || read_stride >= (1u << 31)) {
    rg_set_error("count: device stream must be 16-byte aligned, read_stride < 2^31");
    return RG_ERR_ARG;
  }
  ExtractParams ep;
  ep.stream = dev_stream;
  ep.n_bytes = n_bytes;
  ep.brk = nullptr;
  ep.read_stride = read_stride;
  ep.k = t->k;
  ep.canonical = t->canonical;
  int rc = rg_launch_partition_multi(t, ep);
  if (rc) return rc;
  // publish the fills of my sub-slabs: owner o's cursor[g][me] = my local cursor of its group g
  return rg_launch_publish_cursors(t);
}

// after every rank's scatter has completed (caller's barrier): my groups -> tiles -> inserts.
// Asynchronous: what the other GPUs could not place in my slabs (their overflow goes to my exported
// spill list) is picked up by rg_count_end / the first reader (rg_table_finish).
int rg_count_p2p_finish(rg_table* t) {
  if (!t || !t->p2p_block) return RG_ERR_ARG;
  cudaSetDevice(t->ctx->device);
  // bucket slabs of my tiles, then (group, source) sub-slabs -> tiles, then the tile inserts
  int rc = prepare_buckets(t, t->p2p_max_keys);
  if (rc == 0) rc = rg_launch_partition_groups(t);
  if (rc == 0) rc = rg_launch_tile_insert(t);
  if (rc) return rc;
  t->tiles_dirty = true;
  t->unchecked = true;
  t->p2p_pending = true;
  return 0;
}

int rg_count_push_dev(rg_table* t, const char* dev_stream, uint64_t n_bytes, uint32_t read_stride) {
  if (!t) return RG_ERR_ARG;
  return push_dev_impl(t, dev_stream, n_bytes, read_stride, nullptr);
}

// Asynchronous for pinned host memory: the pieces are copied on the context's copy stream into two
// staging slots while the partition kernel of the previous piece runs, so the H2D copy and the
// kernels overlap; the caller's buffer must stay valid until rg_count_end / rg_sync. A piece
// carries the k-1 bytes before it, so that the windows spanning a piece boundary are counted
// (exactly once: a window is only taken where it fits entirely).
int rg_count_push(rg_table* t, const char* host_stream, uint64_t n_bytes) {
  if (!t) return RG_ERR_ARG;
  if (n_bytes == 0) return 0;
  rg_ctx* ctx = t->ctx;
  cudaSetDevice(ctx->device);
  int rc = slots_ensure(ctx);
  if (rc == 0) rc = ensure_spill(t, n_bytes);
  if (rc) return rc;
  if (t->mode == RG_MODE_TILED) {
    rc = prepare_buckets(t, n_bytes);
    if (rc) return rc;
  }
  const uint64_t halo = (uint64_t)t->k - 1;
  const uint64_t piece = ctx->slot_bytes - 64;          // payload bytes per slot (halo fits in the slack)
  for (uint64_t off = 0; off < n_bytes; off += piece) {
    const uint64_t len = std::min(piece, n_bytes - off);
    const uint64_t pre = off ? halo : 0;
    uint32_t i;
    rc = slot_acquire(ctx, &i);
    if (rc) return rc;
    RG_CUDA_OK(cudaMemcpyAsync(ctx->slot[i], host_stream + off - pre, len + pre, cudaMemcpyHostToDevice,
                               ctx->copy_stream));
    rc = slot_publish(ctx, i);
    if (rc) return rc;
    rc = push_piece(t, ctx->slot[i], len + pre);
    if (rc) return rc;
    rc = slot_release(ctx, i);
    if (rc) return rc;
  }
  if (t->mode == RG_MODE_TILED) {
    rc = finish_tiled_push(t);
    if (rc) return rc;
  }
  t->tiles_dirty = true;
  t->unchecked = true;
  return 0;
}

int rg_count_push_reads(rg_table* t, const char* host_seq, const uint64_t* host_offs, uint64_t n_reads) {
  if (!t) return RG_ERR_ARG;
  if (n_reads == 0) return 0;
  rg_ctx* ctx = t->ctx;
  cudaSetDevice(ctx->device);
  const uint64_t base = host_offs[0];
  const uint64_t n_bytes = host_offs[n_reads] - base;
  if (n_bytes == 0) return 0;
  int rc = ensure_stage(t, n_bytes);
  if (rc) return rc;
  RG_CUDA_OK(cudaMemcpyAsync(t->stage, host_seq + base, n_bytes, cudaMemcpyHostToDevice, ctx->stream));
  if (t->d_offs_cap < n_reads + 1) {
    if (t->d_offs) { RG_CUDA_OK(cudaStreamSynchronize(ctx->stream)); cudaFree(t->d_offs); t->d_offs = nullptr; }
    RG_CUDA_OK(cudaMalloc(&t->d_offs, (n_reads + 1) * 8));
    t->d_offs_cap = n_reads + 1;
  }
  const uint64_t* offs = host_offs;
  std::vector<uint64_t> rebased;
  if (base != 0) {
    rebased.resize(n_reads + 1);
    for (uint64_t i = 0; i <= n_reads; ++i) rebased[i] = host_offs[i] - base;
    offs = rebased.data();
  }
  RG_CUDA_OK(cudaMemcpyAsync(t->d_offs, offs, (n_reads + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
  RG_CUDA_OK(cudaStreamSynchronize(ctx->stream));   // `rebased` must outlive the copy
  const uint64_t words = n_bytes / 32 + 2;
  if (t->brk_words < words) {
    if (t->brk) cudaFree(t->brk);
    RG_CUDA_OK(cudaMalloc(&t->brk, words * 4));
    t->brk_words = words;
  }
  RG_CUDA_OK(cudaMemsetAsync(t->brk, 0, words * 4, ctx->stream));
  // empty reads at the very end would mark position n_bytes: harmless (bitmap has slack)
  rc = rg_launch_mark_breaks(ctx, t->d_offs, n_reads, t->brk);
  if (rc) return rc;
  return push_dev_impl(t, reinterpret_cast<const char*>(t->stage), n_bytes, 0, t->brk);
}

int rg_count_push_keys_dev(rg_table* t, const uint64_t* dev_keys, uint64_t n_keys) {
  if (!t) return RG_ERR_ARG;
  // packed keys carry no read context: every key finds its tile through its own minimizer (global atomics)
  int rc = ensure_spill(t, n_keys);
  if (rc == 0) rc = rg_launch_insert_keys(t, dev_keys, n_keys, nullptr);
  if (rc) return rc;
  t->tiles_dirty = true;
  t->unchecked = true;
  return 0;
}

int rg_count_end(rg_table* t) {
  if (!t) return RG_ERR_ARG;
  cudaSetDevice(t->ctx->device);
  t->unchecked = true;
  return rg_table_finish(t);
}

// make room for up to extra_distinct new keys before a push (streaming callers do not know the
// number of distinct k-mers up front): grows (re-hashes) when the load would exceed ~60 %, and keeps
// the overflow region large enough to take all of them (whatever their minimizers, the push cannot
// lose a key)
int rg_table_reserve(rg_table* t, uint64_t extra_distinct) {
  if (!t) return RG_ERR_ARG;
  cudaSetDevice(t->ctx->device);
  int rc = rg_table_finish(t);
  if (rc) return rc;
  uint64_t distinct = 0;
  rc = rg_table_stats(t, &distinct, nullptr, nullptr);
  if (rc) return rc;
  uint32_t want = tiles_for(distinct + extra_distinct);
  uint64_t want_ovf = t->ovf_slots;
  while (want_ovf < 2 * (t->ovf_used + extra_distinct)) want_ovf <<= 1;
  if (want <= t->n_tiles && want_ovf == t->ovf_slots) return 0;
  if (want > t->n_tiles) want = std::max<uint32_t>(want, t->n_tiles + t->n_tiles / 2);
  else want = t->n_tiles;
  return table_grow(t, want, want_ovf);
}

int rg_table_trim(rg_table* t) {
  if (!t) return RG_ERR_ARG;
  RG_CUDA_OK(cudaStreamSynchronize(t->ctx->stream));
  if (t->part_keys) { cudaFree(t->part_keys); t->part_keys = nullptr; t->part_cap = 0; }
  if (!t->p2p_block && t->grp_keys) { cudaFree(t->grp_keys); t->grp_keys = nullptr; t->grp_alloc_keys = 0; }
  if (t->stage) { cudaFree(t->stage); t->stage = nullptr; t->stage_bytes = 0; }
  if (t->brk) { cudaFree(t->brk); t->brk = nullptr; t->brk_words = 0; }
  if (t->d_offs) { cudaFree(t->d_offs); t->d_offs = nullptr; t->d_offs_cap = 0; }
  return 0;
}

int rg_table_clear(rg_table* t) {
  if (!t) return RG_ERR_ARG;
  rg_ctx* ctx = t->ctx;
  cudaSetDevice(ctx->device);
  const uint64_t slots = rg_table_slots(t);
  RG_CUDA_OK(cudaMemsetAsync(t->keys, 0xFF, slots * 8, ctx->stream));
  RG_CUDA_OK(cudaMemsetAsync(t->counts, 0, slots * 4, ctx->stream));
  RG_CUDA_OK(cudaMemsetAsync(t->d_status, 0, RG_ST_WORDS * sizeof(uint64_t), ctx->stream));
  t->tiles_dirty = false;
  t->ovf_used = 0;
  return 0;
}

void rg_table_free(rg_table* t) {
  if (!t) return;
  cudaSetDevice(t->ctx->device);
  cudaStreamSynchronize(t->ctx->stream);
  cudaFree(t->keys);
  cudaFree(t->counts);
  cudaFree(t->d_status);
  cudaFree(t->spill_keys);
  cudaFree(t->spill_counts);
  cudaFree(t->part_keys);
  cudaFree(t->part_cursor);
  if (t->p2p_block) {
    for (int i = 0; i < 8; ++i)
      if (t->peer_mapped[i]) cudaIpcCloseMemHandle(t->peer_mapped[i]);
    cudaFree(t->p2p_block);
    cudaFree(t->p2p_src_cursor);
  } else {
    cudaFree(t->grp_keys);
    cudaFree(t->grp_cursor);
  }
  cudaFree(t->stage);
  cudaFree(t->brk);
  cudaFree(t->d_offs);
  delete t;
}

int rg_table_stats(rg_table* t, uint64_t* n_distinct, uint64_t* n_occ, uint64_t* n_slots) {
  if (!t) return RG_ERR_ARG;
  cudaSetDevice(t->ctx->device);
  int rc = rg_table_finish(t);
  if (rc) return rc;
  rc = rg_launch_stats(t, t->ctx->d_scratch);
  if (rc) return rc;
  uint64_t v[2];
  rc = read_scalars(t->ctx, t->ctx->d_scratch, 2, v);
  if (rc) return rc;
  if (n_distinct) *n_distinct = v[0];
  if (n_occ) *n_occ = v[1];
  if (n_slots) *n_slots = rg_table_slots(t);
  return 0;
}

int rg_table_k(rg_table* t) { return t ? t->k : 0; }
int rg_table_canonical(rg_table* t) { return t ? t->canonical : 0; }

uint32_t rg_key_shard(uint64_t jf_key, int k, int m, uint32_t groups, uint32_t n_parts) {
  if (k < 1 || k > 32 || m < 1 || m > 16 || m > k || n_parts == 0 || groups < n_parts) return 0;
  return rg_group_of_mh(rg_min_hash(jf_key, (uint32_t)k, (uint32_t)m), groups) / (groups / n_parts);
}
int rg_minimizer_len(int k, uint64_t n_tiles_global) { return sk_choose_m(k, n_tiles_global); }
int rg_table_geometry(rg_table* t, uint32_t* n_tiles, uint32_t* groups_global, int* m, uint64_t* ovf_slots) {
  if (!t) return RG_ERR_ARG;
  if (n_tiles) *n_tiles = t->n_tiles;
  if (groups_global) *groups_global = t->n_tiles_global / t->group_tiles;
  if (m) *m = t->m;
  if (ovf_slots) *ovf_slots = t->ovf_slots;
  return 0;
}

int rg_extract_partition_dev(rg_table* t, const char* dev_stream, uint64_t n_bytes, uint32_t read_stride,
                             uint64_t* keys_out_dev, uint64_t* part_offsets_dev) {
  if (!t) return RG_ERR_ARG;
  rg_ctx* ctx = t->ctx;
  cudaSetDevice(ctx->device);
  ExtractParams ep;
  ep.stream = dev_stream;
  ep.n_bytes = n_bytes;
  ep.brk = nullptr;
  ep.read_stride = read_stride;
  ep.k = t->k;
  ep.canonical = t->canonical;
  return rg_launch_extract_partition(ctx, ep, rg_view(t).geo, t->shard_parts, keys_out_dev, part_offsets_dev,
                                     ctx->d_scratch + 64);
}

// ---- deferred checks: a whole pass enqueued without a host round trip ---------
// With rg_defer_checks(ctx, 1) the readers of a table (histo, subtract, dump ...) no longer wait for the
// status words of the enqueued pushes; the caller fetches them with rg_table_status_async into pinned
// memory, synchronises ONCE at the end of its pass and, if a table reports spilled / foreign keys, turns
// the deferral off, calls rg_count_end (which resolves them) and repeats the readers.
int rg_defer_checks(rg_ctx* ctx, int on) {
  if (!ctx) return RG_ERR_ARG;
  ctx->defer_checks = on != 0;
  return 0;
}

// host_pinned[0..7] = the table's RG_ST_* words, [8] = records other GPUs spilled into my exported list
int rg_table_status_async(rg_table* t, uint64_t* host_pinned) {
  if (!t || !host_pinned) return RG_ERR_ARG;
  rg_ctx* ctx = t->ctx;
  cudaSetDevice(ctx->device);
  RG_CUDA_OK(cudaMemcpyAsync(host_pinned, t->d_status, RG_ST_WORDS * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  host_pinned[8] = 0;
  if (t->p2p_block)
    RG_CUDA_OK(cudaMemcpyAsync(host_pinned + 8, t->p2p_status, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  return 0;
}

// same words into DEVICE memory (9 uint64), for callers that reduce them on the device
int rg_table_status_dev(rg_table* t, uint64_t* dev_out9) {
  if (!t || !dev_out9) return RG_ERR_ARG;
  rg_ctx* ctx = t->ctx;
  cudaSetDevice(ctx->device);
  RG_CUDA_OK(cudaMemcpyAsync(dev_out9, t->d_status, RG_ST_WORDS * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
  if (t->p2p_block)
    RG_CUDA_OK(cudaMemcpyAsync(dev_out9 + 8, t->p2p_status, sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
  else
    RG_CUDA_OK(cudaMemsetAsync(dev_out9 + 8, 0, sizeof(uint64_t), ctx->stream));
  return 0;
}

// ---- histo ----------------------------------------------------------------
static inline uint64_t histo_base(uint64_t low) { return 1 >= low ? 0 : low - 1; }   // histo_main.cc:64
uint64_t rg_histo_rows(uint64_t low, uint64_t high) { return (high + 1) + 1 - histo_base(low); }

int rg_histo(rg_table* t, uint64_t min_count, uint64_t low, uint64_t high, uint64_t* out_host) {
  if (!t || high < low) {
    rg_set_error("rg_histo: High count value must be >= to low count value");
    return RG_ERR_ARG;
  }
  rg_ctx* ctx = t->ctx;
  cudaSetDevice(ctx->device);
  const uint64_t nb = rg_histo_rows(low, high);
  int rc = rg_table_finish(t);
  if (rc == 0) rc = pool_ensure(ctx, 5, nb * 8);
  if (rc) return rc;
  uint64_t* d_out = static_cast<uint64_t*>(ctx->pool[5]);
  rc = rg_launch_histo(t, min_count, histo_base(low), nb, d_out);
  if (rc == 0) {
    cudaError_t e = cudaMemcpyAsync(out_host, d_out, nb * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { rg_set_error(cudaGetErrorString(e)); rc = RG_ERR_CUDA; }
  }
  return rc;
}

// histogram left on the device (rg_histo_rows(low, high) uint64): no host round trip
int rg_histo_dev(rg_table* t, uint64_t min_count, uint64_t low, uint64_t high, uint64_t* out_dev) {
  if (!t || high < low || !out_dev) return RG_ERR_ARG;
  cudaSetDevice(t->ctx->device);
  int rc = rg_table_finish(t);
  if (rc) return rc;
  return rg_launch_histo(t, min_count, histo_base(low), rg_histo_rows(low, high), out_dev);
}

// ---- dump / query -----------------------------------------------------------
static int64_t copy_out(rg_ctx* ctx, uint64_t n, const uint64_t* d_keys, const uint32_t* d_counts, uint64_t* keys_host,
                        uint32_t* counts_host) {
  if (n) {
    RG_CUDA_OK(cudaMemcpyAsync(keys_host, d_keys, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    RG_CUDA_OK(cudaMemcpyAsync(counts_host, d_counts, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    RG_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  }
  return (int64_t)n;
}

int64_t rg_dump(rg_table* t, uint64_t min_count, uint64_t max_count, uint64_t* keys_host, uint32_t* counts_host,
                uint64_t cap, uint64_t* needed) {
  if (!t) return RG_ERR_ARG;
  rg_ctx* ctx = t->ctx;
  cudaSetDevice(ctx->device);
  uint64_t* d_keys = nullptr;
  uint32_t* d_counts = nullptr;
  DevTemps tmp;
  RG_CUDA_OK(tmp.alloc(&d_keys, (cap ? cap : 1) * 8));
  RG_CUDA_OK(tmp.alloc(&d_counts, (cap ? cap : 1) * 4));
  int64_t rc = rg_table_finish(t);
  if (rc == 0) rc = rg_launch_dump(t, min_count ? min_count : 1, max_count, d_keys, d_counts, cap, ctx->d_scratch);
  uint64_t n = 0;
  if (rc == 0) rc = read_scalars(ctx, ctx->d_scratch, 1, &n);
  if (rc == 0) {
    if (needed) *needed = n;
    if (n > cap) {
      rg_set_error("rg_dump: output buffer too small");
      rc = RG_ERR_SMALL;
    } else {
      rc = copy_out(ctx, n, d_keys, d_counts, keys_host, counts_host);
    }
  }
  return rc;
}

// the entries with count >= min_count whose key lies in the slice {key : rg_hash32b(key) % slice_mod == slice_rem}
// (slice_mod 0: every key), into DEVICE arrays; *needed (host) = their number. A slice is a fixed function of the
// key, so shards / tables of any geometry can be compared slice by slice with a CPU count of the same slice.
int64_t rg_dump_slice_dev(rg_table* t, uint64_t min_count, uint32_t slice_mod, uint32_t slice_rem, uint64_t* keys_dev,
                          uint32_t* counts_dev, uint64_t cap, uint64_t* needed) {
  if (!t || (slice_mod && slice_rem >= slice_mod)) return RG_ERR_ARG;
  rg_ctx* ctx = t->ctx;
  cudaSetDevice(ctx->device);
  int64_t rc = rg_table_finish(t);
  if (rc == 0)
    rc = rg_launch_dump(t, min_count ? min_count : 1, ~0ull, keys_dev, counts_dev, cap, ctx->d_scratch, slice_mod, slice_rem);
  uint64_t n = 0;
  if (rc == 0) rc = read_scalars(ctx, ctx->d_scratch, 1, &n);
  if (rc) return rc;
  if (needed) *needed = n;
  if (n > cap) {
    rg_set_error("rg_dump_slice_dev: output buffer too small");
    return RG_ERR_SMALL;
  }
  return (int64_t)n;
}

rg_table* rg_table_from_pairs(rg_ctx* ctx, int k, int canonical, const uint64_t* keys_host,
                              const uint32_t* counts_host, uint64_t n) {
  rg_table* t = rg_count_begin(ctx, k, canonical, n ? n : 1, RG_MODE_DIRECT);
  if (!t) return nullptr;
  if (n) {
    uint64_t* dk = nullptr;
    uint32_t* dc = nullptr;
    if (cudaMalloc(&dk, n * 8) != cudaSuccess || cudaMalloc(&dc, n * 4) != cudaSuccess) {
      rg_set_error("rg_table_from_pairs: out of device memory");
      cudaFree(dk);
      rg_table_free(t);
      return nullptr;
    }
    cudaMemcpyAsync(dk, keys_host, n * 8, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(dc, counts_host, n * 4, cudaMemcpyHostToDevice, ctx->stream);
    int rc = rg_launch_insert_keys(t, dk, n, dc);
    t->tiles_dirty = true;
    if (rc == 0) rc = rg_count_end(t);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(dk);
    cudaFree(dc);
    if (rc) {
      rg_table_free(t);
      return nullptr;
    }
  }
  return t;
}

int rg_query(rg_table* t, const uint64_t* keys_host, uint64_t n, uint64_t min_count, uint32_t* counts_host) {
  if (!t) return RG_ERR_ARG;
  if (n == 0) return 0;
  rg_ctx* ctx = t->ctx;
  cudaSetDevice(ctx->device);
  uint64_t* dk = nullptr;
  uint32_t* dc = nullptr;
  DevTemps tmp;
  RG_CUDA_OK(tmp.alloc(&dk, n * 8));
  RG_CUDA_OK(tmp.alloc(&dc, n * 4));
  int rc = rg_table_finish(t);
  if (rc == 0) rc = RG_ERR_CUDA;
  if (rc == RG_ERR_CUDA && cudaMemcpyAsync(dk, keys_host, n * 8, cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess) {
    rc = rg_launch_query(t, dk, n, min_count, dc);
    if (rc == 0 && (cudaMemcpyAsync(counts_host, dc, n * 4, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
                    cudaStreamSynchronize(ctx->stream) != cudaSuccess)) {
      rg_set_error("rg_query: copy failed");
      rc = RG_ERR_CUDA;
    }
  }
  return rc;
}

// ---- subtraction ------------------------------------------------------------
static int fill_others(OtherTables& ot, rg_table* const* others, const uint64_t* other_min, int n_others, int k,
                       const rg_table* subj) {
  if (n_others < 0 || n_others > 16) {
    rg_set_error("subtract: at most 16 control/exclude tables");
    return RG_ERR_ARG;
  }
  ot.n = n_others;
  ot.same_geo = 0;
  for (int i = 0; i < n_others; ++i) {
    if (!others[i] || others[i]->k != k) {
      rg_set_error("subtract: tables must share k");
      return RG_ERR_ARG;
    }
    const int frc = rg_table_finish(others[i]);
    if (frc) return frc;
    ot.tv[i] = rg_view(others[i]);
    if (subj && others[i]->n_tiles == subj->n_tiles && others[i]->group_tiles == subj->group_tiles &&
        others[i]->n_tiles_global == subj->n_tiles_global && others[i]->tile_first == subj->tile_first &&
        others[i]->m == subj->m)
      ot.same_geo |= 1u << i;
    uint64_t m = other_min ? other_min[i] : 1;
    ot.min_count[i] = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(m, 1), 0xFFFFFFFFull);
  }
  return 0;
}

// the subtraction enqueued only: survivors in keys_dev / counts_dev (the unused tail of keys_dev is filled with
// ~0, never a canonical key, so a fixed-capacity buffer can travel as it is), their number (it may exceed cap:
// the excess is dropped) in *count_dev. No host round trip.
int rg_subtract_enqueue(rg_table* subj, uint64_t subj_min, rg_table* const* others, const uint64_t* other_min,
                        int n_others, uint64_t floor, int64_t min_cov, int64_t max_cov, uint64_t* keys_dev,
                        uint32_t* counts_dev, uint64_t cap, uint64_t* count_dev) {
  if (!subj || !count_dev) return RG_ERR_ARG;
  rg_ctx* ctx = subj->ctx;
  cudaSetDevice(ctx->device);
  OtherTables ot;
  int rc = rg_table_finish(subj);
  if (rc) return rc;
  rc = fill_others(ot, others, other_min, n_others, subj->k, subj);
  if (rc) return rc;
  RG_CUDA_OK(cudaMemsetAsync(keys_dev, 0xFF, cap * 8, ctx->stream));
  uint64_t lo = std::max<uint64_t>(std::max<uint64_t>(subj_min, floor), 1);
  if (min_cov > 0) lo = std::max<uint64_t>(lo, (uint64_t)min_cov);
  if (max_cov < 0) {   // nothing can pass `count <= max`
    RG_CUDA_OK(cudaMemsetAsync(count_dev, 0, sizeof(uint64_t), ctx->stream));
    return 0;
  }
  return rg_launch_subtract(subj, lo, (uint64_t)max_cov, ot, keys_dev, counts_dev, cap, count_dev);
}

int64_t rg_subtract_dev(rg_table* subj, uint64_t subj_min, rg_table* const* others, const uint64_t* other_min,
                        int n_others, uint64_t floor, int64_t min_cov, int64_t max_cov, uint64_t* keys_dev,
                        uint32_t* counts_dev, uint64_t cap) {
  if (!subj) return RG_ERR_ARG;
  rg_ctx* ctx = subj->ctx;
  cudaSetDevice(ctx->device);
  OtherTables ot;
  int rc = rg_table_finish(subj);
  if (rc) return rc;
  rc = fill_others(ot, others, other_min, n_others, subj->k, subj);
  if (rc) return rc;
  // candidates: c >= subj_min (in the subject .Jhash), c >= floor (merge_files.cc:149), min_cov <= c <= max_cov
  uint64_t lo = std::max<uint64_t>(std::max<uint64_t>(subj_min, floor), 1);
  if (min_cov > 0) lo = std::max<uint64_t>(lo, (uint64_t)min_cov);
  if (max_cov < 0) {   // nothing can pass `count <= max`
    return 0;
  }
  const uint64_t hi = (uint64_t)max_cov;
  rc = rg_launch_subtract(subj, lo, hi, ot, keys_dev, counts_dev, cap, ctx->d_scratch);
  if (rc) return rc;
  uint64_t n = 0;
  rc = read_scalars(ctx, ctx->d_scratch, 1, &n);
  if (rc) return rc;
  if (n > cap) {
    rg_set_error("rg_subtract: output buffer too small, need " + std::to_string(n));
    ctx->h_scratch[8] = n;
    return RG_ERR_SMALL;
  }
  return (int64_t)n;
}

int64_t rg_subtract(rg_table* subj, uint64_t subj_min, rg_table* const* others, const uint64_t* other_min,
                    int n_others, uint64_t floor, int64_t min_cov, int64_t max_cov, uint64_t* keys_host,
                    uint32_t* counts_host, uint64_t cap, uint64_t* needed) {
  if (!subj) return RG_ERR_ARG;
  rg_ctx* ctx = subj->ctx;
  cudaSetDevice(ctx->device);
  uint64_t* dk = nullptr;
  uint32_t* dc = nullptr;
  DevTemps tmp;
  RG_CUDA_OK(tmp.alloc(&dk, (cap ? cap : 1) * 8));
  RG_CUDA_OK(tmp.alloc(&dc, (cap ? cap : 1) * 4));
  int64_t n = rg_subtract_dev(subj, subj_min, others, other_min, n_others, floor, min_cov, max_cov, dk, dc, cap);
  if (n == RG_ERR_SMALL && needed) *needed = ctx->h_scratch[8];
  if (n >= 0) {
    if (needed) *needed = (uint64_t)n;
    n = copy_out(ctx, (uint64_t)n, dk, dc, keys_host, counts_host);
  }
  return n;
}

// RUFUS.Build (legacy selection, src/RUFUS.Build.cpp:224-283)
int64_t rg_build_select(rg_table* subj, int64_t min_subject, int64_t max_subject, rg_table* const* controls,
                        int n_controls, int64_t max_control_sum, uint64_t* keys_host, uint32_t* counts_host,
                        uint32_t* control_sums_host, uint64_t cap, uint64_t* needed) {
  if (!subj) return RG_ERR_ARG;
  rg_ctx* ctx = subj->ctx;
  cudaSetDevice(ctx->device);
  OtherTables ot;
  int rc = rg_table_finish(subj);
  if (rc) return rc;
  rc = fill_others(ot, controls, nullptr, n_controls, subj->k, subj);
  if (rc) return rc;
  if (needed) *needed = 0;
  if (max_subject < 0 || max_control_sum < 0) return 0;     // nothing can pass
  const uint64_t lo = min_subject > 1 ? (uint64_t)min_subject : 1, hi = (uint64_t)max_subject;
  uint64_t* dk = nullptr;
  uint32_t *dc = nullptr, *ds = nullptr;
  DevTemps tmp;
  RG_CUDA_OK(tmp.alloc(&dk, (cap ? cap : 1) * 8));
  RG_CUDA_OK(tmp.alloc(&dc, (cap ? cap : 1) * 4));
  RG_CUDA_OK(tmp.alloc(&ds, (cap ? cap : 1) * 4));
  int64_t n = rg_launch_build_select(subj, lo, hi, ot, (uint64_t)max_control_sum, dk, dc, ds, cap, ctx->d_scratch);
  uint64_t got = 0;
  if (n == 0) n = read_scalars(ctx, ctx->d_scratch, 1, &got);
  if (n == 0) {
    if (needed) *needed = got;
    if (got > cap) {
      rg_set_error("rg_build_select: output buffer too small");
      n = RG_ERR_SMALL;
    } else {
      n = copy_out(ctx, got, dk, dc, keys_host, counts_host);
      if (n >= 0 && got) {
        cudaError_t e = cudaMemcpyAsync(control_sums_host, ds, got * 4, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { rg_set_error(cudaGetErrorString(e)); n = RG_ERR_CUDA; }
      }
    }
  }
  return n;
}

int64_t rg_merge_unique(rg_table* const* tabs, const uint64_t* mins, int n_tabs, uint64_t floor, uint64_t* keys_host,
                        uint32_t* counts_host, int32_t* which_host, uint64_t cap, uint64_t* needed) {
  if (!tabs || n_tabs < 1 || n_tabs > 17) {
    rg_set_error("rg_merge_unique: 1..17 tables");
    return RG_ERR_ARG;
  }
  uint64_t total = 0;
  bool small = false;
  for (int i = 0; i < n_tabs; ++i) {
    std::vector<rg_table*> others;
    std::vector<uint64_t> omin;
    for (int j = 0; j < n_tabs; ++j)
      if (j != i) {
        others.push_back(tabs[j]);
        omin.push_back(mins ? mins[j] : 1);
      }
    uint64_t need_i = 0;
    const uint64_t room = total < cap ? cap - total : 0;
    int64_t n = rg_subtract(tabs[i], mins ? mins[i] : 1, others.data(), omin.data(), (int)others.size(), floor, 0,
                            (int64_t)0xFFFFFFFFll, keys_host + std::min(total, cap),
                            counts_host + std::min(total, cap), room, &need_i);
    if (n == RG_ERR_SMALL) {
      small = true;
      total += need_i;
      continue;
    }
    if (n < 0) return n;
    if (which_host && !small)
      for (int64_t q = 0; q < n; ++q) which_host[total + q] = i;
    total += (uint64_t)n;
  }
  if (needed) *needed = total;
  if (small) {
    rg_set_error("rg_merge_unique: output buffer too small");
    return RG_ERR_SMALL;
  }
  return (int64_t)total;
}

// ---- filter -----------------------------------------------------------------
uint64_t rg_rufus_pack(const char* s, int len) {
  uint64_t v = 0;
  for (int i = 0; i < len && i < 32; ++i) {
    uint64_t c = 0;
    if (s[i] == 'C') c = 2;
    else if (s[i] == 'G') c = 1;
    else if (s[i] == 'T') c = 3;
    v |= c << (2 * i);
  }
  return v;
}

int rg_rufus_revcomp(const char* s, int len, char* out) {
  int n = 0;
  for (int i = len - 1; i >= 0; --i) {
    const char c = s[i];
    if (c == 'A') out[n++] = 'T';
    else if (c == 'C') out[n++] = 'G';
    else if (c == 'G') out[n++] = 'C';
    else if (c == 'T') out[n++] = 'A';
    else if (c == 'N') out[n++] = 'N';
  }
  return n;
}

static rg_set* set_alloc(rg_ctx* ctx, uint64_t n_vals, int k) {
  if (!ctx || k < 1 || k > 32) {
    rg_set_error("rg_set: k must be 1..32");
    return nullptr;
  }
  cudaSetDevice(ctx->device);
  rg_set* s = new rg_set();
  s->ctx = ctx;
  s->k = k;
  uint64_t slots = 1024;
  while (slots < n_vals * 2 + 2) slots <<= 1;
  if (slots > (1ull << 31)) {
    rg_set_error("rg_set: too many k-mers");
    delete s;
    return nullptr;
  }
  s->slots = (uint32_t)slots;
  s->bitmap_bits = 1u << 15;
  while (s->bitmap_bits < RG_SET_BITMAP_BITS && (uint64_t)s->bitmap_bits < n_vals * 64) s->bitmap_bits <<= 1;
  if (cudaMalloc(&s->table, slots * 8) != cudaSuccess || cudaMalloc(&s->bitmap, RG_SET_BITMAP_BITS / 8) != cudaSuccess) {
    rg_set_error("rg_set: out of device memory");
    cudaFree(s->table);
    delete s;
    return nullptr;
  }
  cudaMemsetAsync(s->table, 0xFF, slots * 8, ctx->stream);
  cudaMemsetAsync(s->bitmap, 0, RG_SET_BITMAP_BITS / 8, ctx->stream);
  return s;
}

static int set_finish(rg_set* s, const uint64_t* d_vals, uint64_t n) {
  rg_ctx* ctx = s->ctx;
  cudaSetDevice(ctx->device);
  int rc = rg_launch_set_build(s, d_vals, n);
  if (rc) return rc;
  uint64_t d = 0;
  if (n) {
    rc = read_scalars(ctx, ctx->d_scratch, 1, &d);
    if (rc) return rc;
  }
  s->n = d + (s->has_all_ones ? 1 : 0);
  return 0;
}

rg_set* rg_set_from_rufus_keys(rg_ctx* ctx, const uint64_t* vals_host, uint64_t n, int k) {
  rg_set* s = set_alloc(ctx, n, k);
  if (!s) return nullptr;
  for (uint64_t i = 0; i < n; ++i)
    if (vals_host[i] == RG_EMPTY) s->has_all_ones = 1;
  uint64_t* d = nullptr;
  if (n) {
    if (cudaMalloc(&d, n * 8) != cudaSuccess) {
      rg_set_error("rg_set: out of device memory");
      rg_set_free(s);
      return nullptr;
    }
    cudaMemcpyAsync(d, vals_host, n * 8, cudaMemcpyHostToDevice, ctx->stream);
  }
  int rc = set_finish(s, d, n);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(d);
  if (rc) {
    rg_set_free(s);
    return nullptr;
  }
  return s;
}

rg_set* rg_set_from_jf_keys_dev(rg_ctx* ctx, const uint64_t* jf_keys_dev, uint64_t n, int k) {
  rg_set* s = set_alloc(ctx, 2 * n, k);
  if (!s) return nullptr;
  uint64_t* d = nullptr;
  int rc = 0;
  if (n) {
    rc = pool_ensure(ctx, 5, 2 * n * 8);      // scratch: the 2n rufus keys (k-mer + reverse complement)
    if (rc) {
      rg_set_free(s);
      return nullptr;
    }
    d = static_cast<uint64_t*>(ctx->pool[5]);
    rc = rg_launch_jf_to_rufus(ctx, jf_keys_dev, n, k, d);
  }
  // k == 32: poly-A's reverse complement is the all-ones rufus key; canonical jf key 0 marks it
  if (rc == 0 && k == 32 && n) {
    std::vector<uint64_t> h(n);
    cudaMemcpyAsync(h.data(), jf_keys_dev, n * 8, cudaMemcpyDeviceToHost, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    for (uint64_t i = 0; i < n; ++i)
      if (h[i] == 0) s->has_all_ones = 1;
  }
  if (rc == 0) rc = set_finish(s, d, 2 * n);
  if (rc) {
    rg_set_free(s);
    return nullptr;
  }
  return s;
}

// A set that is rebuilt in place every pass (no allocation, no host round trip): room for max_keys jf keys.
rg_set* rg_set_create(rg_ctx* ctx, int k, uint64_t max_keys) {
  if (k >= 32) {
    rg_set_error("rg_set_create: k < 32 (the all-ones rufus key of k = 32 needs the host's attention)");
    return nullptr;
  }
  rg_set* s = set_alloc(ctx, 2 * max_keys, k);
  if (!s) return nullptr;
  if (cudaMalloc(&s->scratch, (max_keys ? max_keys : 1) * 16) != cudaSuccess) {
    rg_set_error("rg_set_create: out of device memory");
    rg_set_free(s);
    return nullptr;
  }
  s->max_keys = max_keys;
  return s;
}

// members = the n_max canonical jf keys at jf_keys_dev (and their reverse complements); entries equal to ~0 are
// padding. Enqueued only.
int rg_set_rebuild_dev(rg_set* s, const uint64_t* jf_keys_dev, uint64_t n_max) {
  if (!s || !s->scratch || n_max > s->max_keys) {
    rg_set_error("rg_set_rebuild_dev: needs a set from rg_set_create with room for n_max keys");
    return RG_ERR_ARG;
  }
  rg_ctx* ctx = s->ctx;
  cudaSetDevice(ctx->device);
  RG_CUDA_OK(cudaMemsetAsync(s->table, 0xFF, (uint64_t)s->slots * 8, ctx->stream));
  RG_CUDA_OK(cudaMemsetAsync(s->bitmap, 0, RG_SET_BITMAP_BITS / 8, ctx->stream));
  int rc = rg_launch_jf_to_rufus(ctx, jf_keys_dev, n_max, s->k, s->scratch);
  if (rc == 0) rc = rg_launch_set_build(s, s->scratch, 2 * n_max);
  return rc;
}

rg_set* rg_set_from_jf_keys(rg_ctx* ctx, const uint64_t* jf_keys_host, uint64_t n, int k) {
  uint64_t* d = nullptr;
  if (n) {
    RG_CUDA_OKP(cudaMalloc(&d, n * 8));
    cudaMemcpyAsync(d, jf_keys_host, n * 8, cudaMemcpyHostToDevice, ctx->stream);
  }
  rg_set* s = rg_set_from_jf_keys_dev(ctx, d, n, k);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(d);
  return s;
}

// HashList text -> Mutations set. paired=1: RUFUS.Filter.cpp:121-143 (split on ' ' first),
// paired=0: RUFUS.Filter.ss.cpp:100-121 (split on '\t' first). Split = Util.cpp:24-33.
static void split(const std::string& line, char delim, std::vector<std::string>& out) {
  out.clear();
  if (line.empty()) return;
  size_t start = 0;
  for (;;) {
    const size_t pos = line.find(delim, start);
    if (pos == std::string::npos) {
      if (start < line.size()) out.push_back(line.substr(start));   // getline drops a trailing empty token
      break;
    }
    out.push_back(line.substr(start, pos - start));
    start = pos + 1;
  }
}

rg_set* rg_set_from_hashlist_text(rg_ctx* ctx, const char* text, uint64_t n_bytes, int k, int paired) {
  std::vector<uint64_t> vals;
  std::vector<std::string> f;
  std::vector<char> rc;
  auto add = [&](const std::string& kmer) {
    vals.push_back(rg_rufus_pack(kmer.data(), (int)kmer.size()));
    rc.resize(kmer.size() + 1);
    const int n = rg_rufus_revcomp(kmer.data(), (int)kmer.size(), rc.data());
    vals.push_back(rg_rufus_pack(rc.data(), n));
  };
  const char first = paired ? ' ' : '\t', second = paired ? '\t' : ' ';
  uint64_t pos = 0;
  while (pos < n_bytes) {
    const char* nl = (const char*)memchr(text + pos, '\n', n_bytes - pos);
    const uint64_t end = nl ? (uint64_t)(nl - text) : n_bytes;
    std::string line(text + pos, end - pos);
    pos = end + 1;
    split(line, first, f);
    if (f.size() == 2) add(f[0]);
    else if (f.size() == 4) add(f[3]);
    if (f.size() == 1) {
      split(line, second, f);
      if (!f.empty()) add(f[0]);
    }
  }
  return rg_set_from_rufus_keys(ctx, vals.data(), vals.size(), k);
}

uint64_t rg_set_size(rg_set* s) { return s ? s->n : 0; }

void rg_set_free(rg_set* s) {
  if (!s) return;
  cudaStreamSynchronize(s->ctx->stream);
  cudaFree(s->table);
  cudaFree(s->bitmap);
  cudaFree(s->scratch);
  delete s;
}

int rg_filter_hits_dev(rg_set* s, const char* seq_dev, const char* qual_dev, const uint64_t* offs_dev,
                       uint64_t n_reads, uint64_t n_bytes, int minq, int skip_last, int32_t* hits_dev) {
  if (!s || (((uintptr_t)seq_dev | (uintptr_t)qual_dev) & 15)) {
    rg_set_error("rg_filter_hits_dev: null set or unaligned (16 B) buffers");
    return RG_ERR_ARG;
  }
  return rg_launch_filter_hits(s, seq_dev, qual_dev, offs_dev, 0, n_reads, n_bytes, 0, true, minq, skip_last, hits_dev);
}

int rg_filter_hits_fixed_dev(rg_set* s, const char* seq_dev, const char* qual_dev, uint32_t read_len,
                             uint64_t n_reads, int minq, int skip_last, int32_t* hits_dev) {
  if (!s || read_len == 0 || (((uintptr_t)seq_dev | (uintptr_t)qual_dev) & 15)) {
    rg_set_error("rg_filter_hits_fixed_dev: bad arguments");
    return RG_ERR_ARG;
  }
  return rg_launch_filter_hits(s, seq_dev, qual_dev, nullptr, read_len, n_reads, (uint64_t)read_len * n_reads, 0, true,
                               minq, skip_last, hits_dev);
}

int rg_filter_pairs_fixed_dev(rg_set* s, const char* s1, const char* q1, const char* s2, const char* q2,
                              uint32_t read_len, uint64_t n_pairs, int minq, int thresh, int32_t* hits1_dev,
                              int32_t* hits2_dev, uint8_t* keep_dev) {
  int rc = rg_filter_hits_fixed_dev(s, s1, q1, read_len, n_pairs, minq, 1, hits1_dev);
  if (rc == 0) rc = rg_filter_hits_fixed_dev(s, s2, q2, read_len, n_pairs, minq, 1, hits2_dev);
  if (rc == 0) rc = rg_launch_keep(s->ctx, hits1_dev, hits2_dev, n_pairs, thresh, keep_dev);
  return rc;
}

// Host-buffer read pull: the offsets go up once, the sequence + quality bytes are streamed through
// the two staging slots in pieces that end on read boundaries (copy of piece i+1 overlaps the
// filter kernel of piece i); device scratch comes from the context's grow-only pool.
static int filter_host_mate(rg_set* s, const char* seq, const char* qual, const uint64_t* offs_host,
                            const uint64_t* offs_dev, uint64_t n_reads, int minq, int skip_last, int32_t* hits_dev,
                            uint8_t* hit_flags_dev = nullptr) {
  rg_ctx* ctx = s->ctx;
  cudaSetDevice(ctx->device);
  RG_CUDA_OK(cudaMemsetAsync(hits_dev, 0, n_reads * sizeof(int32_t), ctx->stream));
  const uint64_t half = (ctx->slot_bytes / 2) & ~255ull;     // sequence in the first half of a slot, quality in the second
  uint64_t r0 = 0;
  while (r0 < n_reads) {
    // largest r1 with offs[r1] - offs[r0] <= half (at least one read; a longer read is rejected)
    uint64_t lo = r0 + 1, hi = n_reads;
    if (offs_host[lo] - offs_host[r0] > half) {
      rg_set_error("filter: a read is longer than the staging slot");
      return RG_ERR_ARG;
    }
    while (lo < hi) {
      const uint64_t mid = (lo + hi + 1) >> 1;
      if (offs_host[mid] - offs_host[r0] <= half) lo = mid; else hi = mid - 1;
    }
    const uint64_t r1 = lo, b0 = offs_host[r0], len = offs_host[r1] - b0;
    if (len) {
      uint32_t i;
      int rc = slot_acquire(ctx, &i);
      if (rc) return rc;
      RG_CUDA_OK(cudaMemcpyAsync(ctx->slot[i], seq + b0, len, cudaMemcpyHostToDevice, ctx->copy_stream));
      RG_CUDA_OK(cudaMemcpyAsync(ctx->slot[i] + half, qual + b0, len, cudaMemcpyHostToDevice, ctx->copy_stream));
      rc = slot_publish(ctx, i);
      if (rc == 0)
        rc = rg_launch_filter_hits(s, ctx->slot[i], ctx->slot[i] + half, offs_dev, 0, n_reads, len, b0, false, minq,
                                   skip_last, hits_dev, hit_flags_dev);
      if (rc == 0) rc = slot_release(ctx, i);
      if (rc) return rc;
    }
    r0 = r1;
  }
  return 0;
}

static int upload_offsets(rg_ctx* ctx, int pool_idx, const uint64_t* offs, uint64_t n_reads) {
  if (offs[0] != 0) {
    rg_set_error("filter: offsets must start at 0");
    return RG_ERR_ARG;
  }
  int rc = pool_ensure(ctx, pool_idx, (n_reads + 1) * 8);
  if (rc) return rc;
  RG_CUDA_OK(cudaMemcpyAsync(ctx->pool[pool_idx], offs, (n_reads + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}

int rg_filter_pairs(rg_set* s, const char* s1, const char* q1, const uint64_t* o1, const char* s2, const char* q2,
                    const uint64_t* o2, uint64_t n_pairs, int minq, int thresh, uint8_t* keep_host) {
  if (!s) return RG_ERR_ARG;
  if (n_pairs == 0) return 0;
  rg_ctx* ctx = s->ctx;
  cudaSetDevice(ctx->device);
  int rc = slots_ensure(ctx);
  if (rc == 0) rc = upload_offsets(ctx, 0, o1, n_pairs);
  if (rc == 0) rc = upload_offsets(ctx, 1, o2, n_pairs);
  if (rc == 0) rc = pool_ensure(ctx, 2, n_pairs * 4);
  if (rc == 0) rc = pool_ensure(ctx, 3, n_pairs * 4);
  if (rc == 0) rc = pool_ensure(ctx, 4, n_pairs);
  if (rc) return rc;
  int32_t* h1 = static_cast<int32_t*>(ctx->pool[2]);
  int32_t* h2 = static_cast<int32_t*>(ctx->pool[3]);
  uint8_t* d_keep = static_cast<uint8_t*>(ctx->pool[4]);
  rc = filter_host_mate(s, s1, q1, o1, static_cast<const uint64_t*>(ctx->pool[0]), n_pairs, minq, 1, h1);
  if (rc == 0) rc = filter_host_mate(s, s2, q2, o2, static_cast<const uint64_t*>(ctx->pool[1]), n_pairs, minq, 1, h2);
  if (rc == 0) rc = rg_launch_keep(ctx, h1, h2, n_pairs, thresh, d_keep);
  if (rc) return rc;
  RG_CUDA_OK(cudaMemcpyAsync(keep_host, d_keep, n_pairs, cudaMemcpyDeviceToHost, ctx->stream));
  RG_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// AnnotateOverlap (src/AnnotateOverlap.cpp:95-146): per-base depth of mutant k-mer windows. A window
// counts when it has no 'N', every quality is >= minq (the tool uses 3), it or its reverse complement
// is in the set, and it is not the last window of the read (the tool's `i < length - K` bound).
int rg_annotate(rg_set* s, const char* seq, const char* qual, const uint64_t* offs, uint64_t n_reads, int minq,
                uint8_t* depth_host) {
  if (!s) return RG_ERR_ARG;
  if (n_reads == 0) return 0;
  rg_ctx* ctx = s->ctx;
  cudaSetDevice(ctx->device);
  const uint64_t n_bytes = offs[n_reads];
  int rc = slots_ensure(ctx);
  if (rc == 0) rc = upload_offsets(ctx, 0, offs, n_reads);
  if (rc == 0) rc = pool_ensure(ctx, 2, n_reads * 4);
  if (rc == 0) rc = pool_ensure(ctx, 4, n_bytes + 16);
  if (rc == 0) rc = pool_ensure(ctx, 5, n_bytes + 16);
  if (rc) return rc;
  uint8_t* flags = static_cast<uint8_t*>(ctx->pool[4]);
  uint8_t* cov = static_cast<uint8_t*>(ctx->pool[5]);
  RG_CUDA_OK(cudaMemsetAsync(flags, 0, n_bytes + 16, ctx->stream));
  rc = filter_host_mate(s, seq, qual, offs, static_cast<const uint64_t*>(ctx->pool[0]), n_reads, minq, 1,
                        static_cast<int32_t*>(ctx->pool[2]), flags);
  if (rc == 0) rc = rg_launch_coverage(ctx, flags, static_cast<const uint64_t*>(ctx->pool[0]), n_reads, n_bytes, s->k, cov);
  if (rc) return rc;
  RG_CUDA_OK(cudaMemcpyAsync(depth_host, cov, n_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  RG_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int rg_filter_single(rg_set* s, const char* seq, const char* qual, const uint64_t* offs, uint64_t n_reads, int minq,
                     int32_t* hits_host) {
  if (!s) return RG_ERR_ARG;
  if (n_reads == 0) return 0;
  rg_ctx* ctx = s->ctx;
  cudaSetDevice(ctx->device);
  int rc = slots_ensure(ctx);
  if (rc == 0) rc = upload_offsets(ctx, 0, offs, n_reads);
  if (rc == 0) rc = pool_ensure(ctx, 2, n_reads * 4);
  if (rc) return rc;
  int32_t* h = static_cast<int32_t*>(ctx->pool[2]);
  rc = filter_host_mate(s, seq, qual, offs, static_cast<const uint64_t*>(ctx->pool[0]), n_reads, minq, 0, h);
  if (rc) return rc;
  RG_CUDA_OK(cudaMemcpyAsync(hits_host, h, n_reads * 4, cudaMemcpyDeviceToHost, ctx->stream));
  RG_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

}  // extern "C"
