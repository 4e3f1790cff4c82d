// rg_count.cu — k-mer extraction/canonicalisation and the two insert paths.
//
// Replaces the reference's hot loop `for(;mers;++mers) ary_.add(*mers,1)`
// (src/modifiedJellyfish/sub_commands/count_main.cc:154-157):
//   mer_iterator::operator++   mer_iterator.hpp:61-89   -> stage_chunk + window_key
//   array::add / claim_key     large_hash_array.hpp:298-302,513-601 -> rg_table_add / tile_insert_kernel
#include "rg_internal.cuh"

#include "rg_extract.cuh"

// ---------------------------------------------------------------------------
// Path 1 (RG_MODE_DIRECT): extraction fused with global-memory CAS/atomic insert.
// ---------------------------------------------------------------------------
constexpr int COUNT_THREADS = 256;
constexpr int COUNT_R = 16;                       // windows per thread
constexpr int CHUNK = COUNT_THREADS * COUNT_R;    // window starts per CTA chunk
using CountStage = StageT<COUNT_THREADS, COUNT_R>;

__global__ void __launch_bounds__(COUNT_THREADS)
count_direct_kernel(ExtractParams ep, TableView tv, uint64_t* status, uint64_t* spill, uint64_t spill_cap) {
  __shared__ CountStage sm;
  const uint64_t n_chunks = (ep.n_bytes + CHUNK - 1) / CHUNK;
  for (uint64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    __syncthreads();
    stage_super(sm, ep, c * CHUNK);
    __syncthreads();
    roll_windows(sm, threadIdx.x, ep.k, ep.canonical, [&](int, uint64_t key) {
      if (key == RG_EMPTY) return;
      const int rc = rg_table_add(tv, key, 1);
      if (rc == 1) {
        const uint64_t at = atomicAdd((unsigned long long*)&status[0], 1ull);
        if (at < spill_cap) spill[at] = key;
      } else if (rc == 2) {
        atomicAdd((unsigned long long*)&status[1], 1ull);
      }
    }, c == 0 ? ep.skip : 0u);
  }
}

int rg_launch_count_direct(rg_table* t, const ExtractParams& ep) {
  rg_ctx* ctx = t->ctx;
  if (ep.n_bytes == 0) return 0;
  const uint64_t n_chunks = (ep.n_bytes + CHUNK - 1) / CHUNK;
  unsigned grid = rg_grid(ctx, 8);
  if (n_chunks < grid) grid = (unsigned)n_chunks;
  {
    rg_launch_scope _ls(ctx, "count_direct_kernel");
    count_direct_kernel<<<grid, COUNT_THREADS, 0, ctx->stream>>>(ep, rg_view(t), t->d_status, t->spill, t->spill_cap);
  }
  RG_CUDA_OK(cudaGetLastError());
  return 0;
}

// insert a list of packed keys (spill resolution, received all-to-all buffers, .Jhash loads)
__global__ void __launch_bounds__(256)
insert_keys_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ counts, uint64_t n, TableView tv,
                   uint64_t* status, uint64_t* spill, uint64_t spill_cap) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t key = __ldg(keys + i);
    const uint32_t add = counts ? __ldg(counts + i) : 1u;
    if (add == 0) continue;
    const int rc = rg_table_add(tv, key, add);
    if (rc == 1) {
      const uint64_t at = atomicAdd((unsigned long long*)&status[0], 1ull);
      if (at < spill_cap) spill[at] = key;   // (count lost for counted inserts: caller sizes tables to avoid it)
    } else if (rc == 2) {
      atomicAdd((unsigned long long*)&status[1], 1ull);
    }
  }
}

int rg_launch_insert_keys(rg_table* t, const uint64_t* dev_keys, uint64_t n, const uint32_t* dev_counts) {
  if (n == 0) return 0;
  rg_ctx* ctx = t->ctx;
  unsigned grid = rg_grid(ctx, 8);
  const uint64_t need = (n + 255) / 256;
  if (need < grid) grid = (unsigned)need;
  {
    rg_launch_scope _ls(ctx, "insert_keys_kernel");
    insert_keys_kernel<<<grid, 256, 0, ctx->stream>>>(dev_keys, dev_counts, n, rg_view(t), t->d_status, t->spill,
                                                      t->spill_cap);
  }
  RG_CUDA_OK(cudaGetLastError());
  return 0;
}

__global__ void rehash_kernel(const uint64_t* __restrict__ old_keys, const uint32_t* __restrict__ old_counts,
                              uint64_t n_slots, TableView tv, uint64_t* status) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_slots;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t key = old_keys[i];
    if (key == RG_EMPTY) continue;
    const uint32_t c = old_counts[i];
    if (rg_table_add(tv, key, c) != 0) atomicAdd((unsigned long long*)&status[2], 1ull);
  }
}

int rg_launch_rehash(rg_table* dst, const uint64_t* old_keys, const uint32_t* old_counts, uint64_t n_slots) {
  rg_ctx* ctx = dst->ctx;
  {
    rg_launch_scope _ls(ctx, "rehash_kernel");
    rehash_kernel<<<rg_grid(ctx, 8), 256, 0, ctx->stream>>>(old_keys, old_counts, n_slots, rg_view(dst), dst->d_status);
  }
  RG_CUDA_OK(cudaGetLastError());
  return 0;
}

// read-start bitmap from an offsets array
__global__ void mark_breaks_kernel(const uint64_t* __restrict__ offs, uint64_t n_reads, uint32_t* brk) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_reads;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t p = offs[i];
    atomicOr(brk + (p >> 5), 1u << (p & 31));
  }
}

int rg_launch_mark_breaks(rg_ctx* ctx, const uint64_t* d_offs, uint64_t n_reads, uint32_t* brk) {
  if (n_reads == 0) return 0;
  {
    rg_launch_scope _ls(ctx, "mark_breaks_kernel");
    mark_breaks_kernel<<<rg_grid(ctx, 4), 256, 0, ctx->stream>>>(d_offs, n_reads, brk);
  }
  RG_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// Path 2, step 2: one CTA per tile. The tile (128 KiB keys + 64 KiB counts) is
// brought into shared memory with TMA bulk copies (or initialised empty), the
// bucket's keys are streamed with coalesced 16-byte loads and applied with
// shared-memory CAS/atomics, and the tile is written back with TMA bulk stores.
// HBM traffic per tile: bucket read once + tile written once (+ read once when
// the table already holds data) — two streaming passes instead of one random
// DRAM round trip per k-mer.
// ---------------------------------------------------------------------------
#ifndef TILE_THREADS_N
#define TILE_THREADS_N 1024
#endif
constexpr int TILE_THREADS = TILE_THREADS_N;
#ifndef TILE_AHEAD
#define TILE_AHEAD 4                 // steps of 32 keys a warp keeps in flight in registers (one block)
#endif
// retry queue of a warp: 64 x 16 bytes (key, slot | probes << RG_TILE_BITS, -)
constexpr uint32_t TILE_Q = 64;
constexpr size_t TILE_SMEM = RG_TILE_KEY_BYTES + RG_TILE_CNT_BYTES + 64 + (TILE_THREADS / 32) * TILE_Q * 16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_store_1d(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}

__global__ void __launch_bounds__(TILE_THREADS, 1)
tile_insert_kernel(uint64_t* __restrict__ keys, uint32_t* __restrict__ counts, const uint64_t* __restrict__ part_keys,
                   const uint32_t* __restrict__ cursor, uint32_t bucket_cap, RgGeo geo,
                   int load_existing, uint64_t* status, uint64_t* spill, uint64_t spill_cap) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* skeys = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* scounts = reinterpret_cast<uint32_t*>(smem_raw + RG_TILE_KEY_BYTES);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + RG_TILE_KEY_BYTES + RG_TILE_CNT_BYTES);

  const uint32_t tile = blockIdx.x;
  uint32_t n = cursor[tile];
  if (n > bucket_cap) n = bucket_cap;
  uint64_t* gkeys = keys + ((uint64_t)tile << RG_TILE_BITS);
  uint32_t* gcounts = counts + ((uint64_t)tile << RG_TILE_BITS);
  if (n == 0 && load_existing) return;   // nothing to apply, tile in HBM is already right

  // the first block of every warp is loaded while the tile is being initialised
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr uint32_t NWARPS = TILE_THREADS / 32, BLOCK_KEYS = 32 * TILE_AHEAD;
  const uint64_t* bucket = part_keys + (uint64_t)tile * bucket_cap;
  auto load = [&](uint32_t blk, int u) -> uint64_t {       // key of this lane in step u of block blk
    const uint32_t pos = blk * BLOCK_KEYS + u * 32 + lane;
    return pos < n ? __ldg(bucket + pos) : RG_EMPTY;
  };
  uint64_t kreg[TILE_AHEAD];
#pragma unroll
  for (int u = 0; u < TILE_AHEAD; ++u) kreg[u] = load(warp, u);

  if (load_existing) {
    if (threadIdx.x == 0) {
      mbar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_expect_tx(bar, RG_TILE_KEY_BYTES + RG_TILE_CNT_BYTES);
      constexpr uint32_t PIECE = 32768;
      for (uint32_t off = 0; off < RG_TILE_KEY_BYTES; off += PIECE)
        tma_load_1d(reinterpret_cast<unsigned char*>(skeys) + off, reinterpret_cast<const unsigned char*>(gkeys) + off,
                    PIECE, bar);
      for (uint32_t off = 0; off < RG_TILE_CNT_BYTES; off += PIECE)
        tma_load_1d(reinterpret_cast<unsigned char*>(scounts) + off,
                    reinterpret_cast<const unsigned char*>(gcounts) + off, PIECE, bar);
    }
    mbar_wait(bar, 0);
  } else {
    uint4* k4 = reinterpret_cast<uint4*>(skeys);
    for (uint32_t i = threadIdx.x; i < RG_TILE_KEY_BYTES / 16; i += blockDim.x)
      k4[i] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
    uint4* c4 = reinterpret_cast<uint4*>(scounts);
    for (uint32_t i = threadIdx.x; i < RG_TILE_CNT_BYTES / 16; i += blockDim.x) c4[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
  }

  // Stream the bucket in blocks of 32 * TILE_AHEAD keys, dealt round-robin to the warps (the first
  // block of a warp is loaded during the tile initialisation). A step = 32 consecutive keys, one
  // coalesced 256-byte load issued one block ahead. Every step makes ONE probe per lane; the lanes that hit a slot occupied by another key
  // push (key, next slot) onto the warp's retry queue (ballot + popc) and the warp moves on, so no
  // lane ever waits for another lane's probe sequence. Whenever the queue holds 32 entries the
  // warp spends a step on those instead of the stream (warp-uniform choice), and drains what is
  // left at the end.
  {
    uint4* q = reinterpret_cast<uint4*>(smem_raw + RG_TILE_KEY_BYTES + RG_TILE_CNT_BYTES + 64) + warp * TILE_Q;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t qhead = 0, qlen = 0;     // warp-uniform

    // one probe of `key` at slot (sp & mask) for the lanes with `have`; failures are queued
    auto step = [&](uint64_t key, uint32_t sp, bool have) {
      const uint32_t slot = sp & (RG_TILE_SLOTS - 1);
      unsigned long long seen = 0;                          // lanes without a key never see an empty slot
      if (have) seen = *(volatile unsigned long long*)(skeys + slot);
      if (seen == RG_EMPTY)
        seen = atomicCAS((unsigned long long*)(skeys + slot), (unsigned long long)RG_EMPTY, (unsigned long long)key);
      const bool hit = (seen == RG_EMPTY) | (seen == key);
      if (hit && have) atomicAdd(scounts + slot, 1u);
      const bool fail = have && !hit;
      const uint32_t fm = __ballot_sync(0xFFFFFFFFu, fail);
      if (fm) {                                             // warp-uniform
        if (fail) {
          // next slot, one more probe counted in the bits above the slot
          const uint32_t nsp = ((sp + 1u) & (RG_TILE_SLOTS - 1)) | ((sp + RG_TILE_SLOTS) & ~(RG_TILE_SLOTS - 1));
          q[(qhead + qlen + __popc(fm & lt_mask)) & (TILE_Q - 1)] = make_uint4((uint32_t)key, (uint32_t)(key >> 32), nsp, 0u);
        }
        qlen += __popc(fm);
      }
    };
    // a step on (up to) 32 queued keys
    auto drain = [&]() {
      __syncwarp();                                         // the queue writes of the previous step are visible
      const uint32_t take = min(32u, qlen);
      const uint4 e = q[(qhead + lane) & (TILE_Q - 1)];
      __syncwarp();                                         // every lane has its entry before the step overwrites the ring
      bool have = lane < take;
      const uint64_t key = ((uint64_t)e.y << 32) | e.x;
      qhead = (qhead + take) & (TILE_Q - 1);
      qlen -= take;
      if (have && (e.z >> RG_TILE_BITS) >= RG_TILE_SLOTS) {  // every slot of the tile probed: tile full
        const uint64_t spi = atomicAdd((unsigned long long*)&status[0], 1ull);
        if (spi < spill_cap) spill[spi] = key;
        have = false;
      }
      step(key, e.z, have);
    };
    // blocks are dealt round-robin to the warps. (A shared ticket counter balanced the warps no
    // better: the compiler emits its warp-aggregated atomic, whose shuffle waits for the shared
    // atomic's round trip once per block: 0.58 ms against 0.49.)
    for (uint32_t cur = warp; cur * BLOCK_KEYS < n; cur += NWARPS) {
#pragma unroll
      for (int u = 0; u < TILE_AHEAD; ++u) {
        const uint64_t key = kreg[u];
        kreg[u] = load(cur + NWARPS, u);
        uint32_t tile_unused, slot;
        rg_locate_in_group(key, geo, tile_unused, slot);
        step(key, slot, key != RG_EMPTY);                   // (neutralised slots and the bucket's end are skipped)
        while (qlen >= 32) drain();
      }
    }
    while (qlen) drain();
  }
  __syncthreads();
  // generic-proxy writes -> visible to the async proxy, then bulk store the tile
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    constexpr uint32_t PIECE = 32768;
    for (uint32_t off = 0; off < RG_TILE_KEY_BYTES; off += PIECE)
      tma_store_1d(reinterpret_cast<unsigned char*>(gkeys) + off, reinterpret_cast<unsigned char*>(skeys) + off, PIECE);
    for (uint32_t off = 0; off < RG_TILE_CNT_BYTES; off += PIECE)
      tma_store_1d(reinterpret_cast<unsigned char*>(gcounts) + off, reinterpret_cast<unsigned char*>(scounts) + off,
                   PIECE);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}

int rg_launch_tile_insert(rg_table* t) {
  rg_ctx* ctx = t->ctx;
  const size_t smem = TILE_SMEM;
  static bool attr_set = false;
  if (!attr_set) {
    RG_CUDA_OK(cudaFuncSetAttribute(tile_insert_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  {
    rg_launch_scope _ls(ctx, "tile_insert_kernel");
    tile_insert_kernel<<<t->n_tiles, TILE_THREADS, smem, ctx->stream>>>(t->keys, t->counts, t->part_keys, t->part_cursor,
                                                                        t->bucket_cap, rg_view(t).geo,
                                                                        t->tiles_dirty ? 1 : 0, t->d_status, t->spill,
                                                                        t->spill_cap);
  }
  RG_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// Multi-GPU building block: extract + partition into n_parts CONTIGUOUS ranges
// (exact sizes: a counting pass, an exclusive scan, a scatter pass), ready to be
// the send buffer of an NCCL all-to-all with split sizes = range lengths.
// ---------------------------------------------------------------------------
template <bool SCATTER>
__global__ void __launch_bounds__(COUNT_THREADS)
extract_parts_kernel(ExtractParams ep, uint32_t n_parts, unsigned long long* __restrict__ counters,
                     uint64_t* __restrict__ keys_out) {
  __shared__ CountStage sm;
  __shared__ unsigned int local_cnt[64];
  __shared__ unsigned long long local_base[64];
  const uint64_t n_chunks = (ep.n_bytes + CHUNK - 1) / CHUNK;
  for (uint64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    __syncthreads();
    stage_super(sm, ep, c * CHUNK);
    if (threadIdx.x < 64) local_cnt[threadIdx.x] = 0;
    __syncthreads();
    // the thread's COUNT_R keys stay in registers between the rank and the scatter phase
    uint64_t mykeys[COUNT_R];
    uint32_t myslot[COUNT_R];
    roll_windows(sm, threadIdx.x, ep.k, ep.canonical, [&](int w, uint64_t key) {
      mykeys[w] = key;
      myslot[w] = 0xFFFFFFFFu;
      if (key != RG_EMPTY) {
        const uint32_t part = rg_tile_of_key(key, n_parts);
        myslot[w] = (part << 16) | atomicAdd(&local_cnt[part], 1u);
      }
    });
    __syncthreads();
    if (threadIdx.x < n_parts) {
      const unsigned int cnt = local_cnt[threadIdx.x];
      local_base[threadIdx.x] = cnt ? atomicAdd(&counters[threadIdx.x], (unsigned long long)cnt) : 0ull;
    }
    __syncthreads();
    if (SCATTER) {
#pragma unroll
      for (int w = 0; w < COUNT_R; ++w)
        if (myslot[w] != 0xFFFFFFFFu) keys_out[local_base[myslot[w] >> 16] + (myslot[w] & 0xFFFFu)] = mykeys[w];
    }
  }
}

__global__ void parts_scan_kernel(uint32_t n_parts, const unsigned long long* counts, uint64_t* offsets,
                                  unsigned long long* cursors) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    uint64_t s = 0;
    for (uint32_t i = 0; i < n_parts; ++i) {
      offsets[i] = s;
      cursors[i] = s;
      s += counts[i];
    }
    offsets[n_parts] = s;
  }
}

int rg_launch_extract_partition(rg_ctx* ctx, const ExtractParams& ep, uint32_t n_parts, uint64_t* keys_out,
                                uint64_t* part_offsets, uint64_t* d_tmp) {
  if (n_parts == 0 || n_parts > 64) {
    rg_set_error("rg_extract_partition: n_parts must be 1..64");
    return RG_ERR_ARG;
  }
  // d_tmp: [0..63] counts, [64..127] cursors
  RG_CUDA_OK(cudaMemsetAsync(d_tmp, 0, 128 * sizeof(uint64_t), ctx->stream));
  const uint64_t n_chunks = (ep.n_bytes + CHUNK - 1) / CHUNK;
  unsigned grid = rg_grid(ctx, 6);
  if (n_chunks < grid) grid = (unsigned)(n_chunks ? n_chunks : 1);
  unsigned long long* cnt = reinterpret_cast<unsigned long long*>(d_tmp);
  unsigned long long* cur = cnt + 64;
  {
    rg_launch_scope _ls(ctx, "extract_parts_kernel<count>");
    extract_parts_kernel<false><<<grid, COUNT_THREADS, 0, ctx->stream>>>(ep, n_parts, cnt, nullptr);
  }
  {
    rg_launch_scope _ls(ctx, "parts_scan_kernel");
    parts_scan_kernel<<<1, 32, 0, ctx->stream>>>(n_parts, cnt, part_offsets, cur);
  }
  {
    rg_launch_scope _ls(ctx, "extract_parts_kernel<scatter>");
    extract_parts_kernel<true><<<grid, COUNT_THREADS, 0, ctx->stream>>>(ep, n_parts, cur, keys_out);
  }
  RG_CUDA_OK(cudaGetLastError());
  return 0;
}
