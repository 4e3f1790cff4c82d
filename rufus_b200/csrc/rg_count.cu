// rg_count.cu — k-mer extraction/canonicalisation and the two insert paths.
//
// Replaces the reference's hot loop `for(;mers;++mers) ary_.add(*mers,1)`
// (src/modifiedJellyfish/sub_commands/count_main.cc:154-157):
//   mer_iterator::operator++   mer_iterator.hpp:61-89   -> stage_chunk + window_key
//   array::add / claim_key     large_hash_array.hpp:298-302,513-601 -> rg_table_add / tile_insert_kernel
#include <type_traits>

#include "rg_internal.cuh"

#include "rg_extract.cuh"
#include "rg_sk.cuh"

// ---------------------------------------------------------------------------
// Path 1 (RG_MODE_DIRECT): extraction fused with global-memory CAS/atomic insert. The tile of every
// k-mer comes from its minimizer, recomputed per key (rg_min_hash): the slow, simple path for small
// inputs and for the callers that bring packed keys.
// ---------------------------------------------------------------------------
constexpr int COUNT_THREADS = 256;
constexpr int COUNT_R = 16;                       // windows per thread
constexpr int CHUNK = COUNT_THREADS * COUNT_R;    // window starts per CTA chunk
using CountStage = StageT<COUNT_THREADS, COUNT_R>;

__global__ void __launch_bounds__(COUNT_THREADS)
count_direct_kernel(ExtractParams ep, TableView tv) {
  __shared__ CountStage sm;
  const uint64_t n_chunks = (ep.n_bytes + CHUNK - 1) / CHUNK;
  if (blockIdx.x == 0 && threadIdx.x == 0) tv.status[RG_ST_DIRECT] = 1;
  for (uint64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    __syncthreads();
    stage_super(sm, ep, c * CHUNK);
    __syncthreads();
    roll_windows(sm, threadIdx.x, ep.k, ep.canonical, [&](int, uint64_t key) {
      if (key == RG_EMPTY) return;
      rg_table_add(tv, key, 1);
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
    count_direct_kernel<<<grid, COUNT_THREADS, 0, ctx->stream>>>(ep, rg_view(t));
  }
  RG_CUDA_OK(cudaGetLastError());
  return 0;
}

// insert a list of packed keys, each `counts[i]` (or 1) times (received all-to-all buffers, .Jhash loads)
__global__ void __launch_bounds__(256)
insert_keys_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ counts, uint64_t n, TableView tv) {
  if (blockIdx.x == 0 && threadIdx.x == 0) tv.status[RG_ST_DIRECT] = 1;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t key = __ldg(keys + i);
    const uint32_t add = counts ? __ldg(counts + i) : 1u;
    if (add == 0) continue;
    rg_table_add(tv, key, add);
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
    insert_keys_kernel<<<grid, 256, 0, ctx->stream>>>(dev_keys, dev_counts, n, rg_view(t));
  }
  RG_CUDA_OK(cudaGetLastError());
  return 0;
}

// the k-mers of spilled super-k-mer records (peer-to-peer counting: what other GPUs could not place in a slab)
__global__ void __launch_bounds__(256)
insert_records_kernel(const uint4* __restrict__ recs, uint64_t n, TableView tv, int canonical) {
  if (blockIdx.x == 0 && threadIdx.x == 0) tv.status[RG_ST_DIRECT] = 1;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint4 rec = __ldg(recs + i);
    const uint32_t tile = rg_tile_of_mh(rec.w & ~15u, tv.geo) - tv.tile_first;
    if (tile >= tv.n_tiles) {
      atomicAdd((unsigned long long*)&tv.status[RG_ST_FOREIGN], 1ull);
      continue;
    }
    const uint32_t len = (rec.w & 15u) + 1;
    for (uint32_t o = 0; o < len; ++o) rg_table_add_at(tv, tile, sk_key(rec, o, tv.geo.k, canonical), 1u);
  }
}

int rg_launch_insert_records(rg_table* t, const uint4* dev_recs, uint64_t n) {
  if (n == 0) return 0;
  rg_ctx* ctx = t->ctx;
  unsigned grid = rg_grid(ctx, 8);
  const uint64_t need = (n + 255) / 256;
  if (need < grid) grid = (unsigned)need;
  {
    rg_launch_scope _ls(ctx, "insert_records_kernel");
    insert_records_kernel<<<grid, 256, 0, ctx->stream>>>(dev_recs, n, rg_view(t), t->canonical);
  }
  RG_CUDA_OK(cudaGetLastError());
  return 0;
}

__global__ void rehash_kernel(const uint64_t* __restrict__ old_keys, const uint32_t* __restrict__ old_counts,
                              uint64_t n_slots, TableView tv) {
  if (blockIdx.x == 0 && threadIdx.x == 0) tv.status[RG_ST_DIRECT] = 1;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_slots;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t key = old_keys[i];
    if (key == RG_EMPTY) continue;
    rg_table_add(tv, key, old_counts[i]);
  }
}

int rg_launch_rehash(rg_table* dst, const uint64_t* old_keys, const uint32_t* old_counts, uint64_t n_slots) {
  rg_ctx* ctx = dst->ctx;
  {
    rg_launch_scope _ls(ctx, "rehash_kernel");
    rehash_kernel<<<rg_grid(ctx, 8), 256, 0, ctx->stream>>>(old_keys, old_counts, n_slots, rg_view(dst));
  }
  RG_CUDA_OK(cudaGetLastError());
  return 0;
}

// peer-to-peer counting: owner o's cursor[g][me] = my local fill of its group g, for every owner, one launch
struct PeerCursors {
  uint32_t* p[8];
};
__global__ void publish_cursors_kernel(const uint32_t* __restrict__ src, PeerCursors peers, uint32_t groups_local,
                                       uint32_t n_parts, uint32_t me) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= groups_local * n_parts) return;
  const uint32_t o = i / groups_local, g = i - o * groups_local;
  peers.p[o][g * n_parts + me] = src[i];
}

int rg_launch_publish_cursors(rg_table* t) {
  rg_ctx* ctx = t->ctx;
  const uint32_t gl = t->n_tiles / t->group_tiles, n = gl * t->shard_parts;
  PeerCursors pc = {};
  for (uint32_t o = 0; o < t->shard_parts; ++o) pc.p[o] = t->peer_cursor[o];
  {
    rg_launch_scope _ls(ctx, "publish_cursors_kernel");
    publish_cursors_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(t->p2p_src_cursor, pc, gl, t->shard_parts, t->shard_part);
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
// Path 2, step 2: one CTA per tile. The tile (128 KiB keys + 64 KiB counts) is brought into shared
// memory with TMA bulk copies (or initialised empty), the bucket's super-k-mer RECORDS are streamed with
// coalesced 16-byte loads, their k-mers are re-extracted, canonicalised and applied with shared-memory
// CAS/atomics, and the tile is written back with TMA bulk stores. HBM traffic per tile: bucket read once
// (~3 bytes per k-mer) + tile written once (+ read once when the table already holds data).
// ---------------------------------------------------------------------------
#ifndef TILE_WARPS_N
#define TILE_WARPS_N 22
#endif
constexpr int TILE_WARPS = TILE_WARPS_N;
constexpr int TILE_THREADS = TILE_WARPS * 32;
// per warp: retry queue of 64 x (8-byte key, 4-byte slot | probes << RG_TILE_BITS, 4-byte multiplicity), the 32
// records of the current batch, and the 512-bit map of the k-mer positions where a record starts
constexpr uint32_t TILE_Q = 64;
constexpr uint32_t TILE_WARP_SMEM = TILE_Q * 16 + 32 * 16 + 16 * 4;
constexpr size_t TILE_SMEM = RG_TILE_KEY_BYTES + RG_TILE_CNT_BYTES + 64 + (size_t)TILE_WARPS * TILE_WARP_SMEM;
static_assert(TILE_SMEM <= 232448, "tile insert: shared memory");

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
tile_insert_kernel(TableView tv, const uint4* __restrict__ part_recs, const uint32_t* __restrict__ cursor,
                   uint32_t bucket_cap, int canonical, int host_dirty) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* skeys = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* scounts = reinterpret_cast<uint32_t*>(smem_raw + RG_TILE_KEY_BYTES);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + RG_TILE_KEY_BYTES + RG_TILE_CNT_BYTES);

  const uint32_t tile = blockIdx.x;
  uint32_t n = cursor[tile];
  if (n > bucket_cap) n = bucket_cap;
  uint64_t* gkeys = tv.keys + ((uint64_t)tile << RG_TILE_BITS);
  uint32_t* gcounts = tv.counts + ((uint64_t)tile << RG_TILE_BITS);
  // the tile in HBM holds data when an earlier push filled it, or when global-memory inserts ran since the
  // table was last empty (records that found their slab full, direct pushes): load it, else start empty
  const bool load_existing = host_dirty || tv.status[RG_ST_DIRECT] != 0;
  if (n == 0 && load_existing) return;   // nothing to apply, tile in HBM is already right

  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint4* bucket = part_recs + (uint64_t)tile * bucket_cap;
  const uint4 none = make_uint4(0u, 0u, 0u, 0u);
  // the first batch of every warp is loaded while the tile is being initialised
  uint4 rec_next = warp * 32 + lane < n ? __ldg(bucket + warp * 32 + lane) : none;

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

  // The bucket is streamed in batches of 32 records, dealt round-robin to the warps (one coalesced 512-byte
  // load, issued one batch ahead). A batch holds T <= 512 k-mers; lane l takes k-mers l, l+32, ... of the
  // batch, whatever record they belong to (a 512-bit map of the record starts + popc finds the record), so
  // the lanes stay busy although records hold 1..16 k-mers. Every step makes ONE probe per lane (LDS.64,
  // CAS if empty, shared atomic on the count); the lanes that met another key push (key, next slot) onto the
  // warp's retry queue (ballot + popc) and the warp moves on: no lane waits for another lane's probe
  // sequence. Whenever the queue holds 32 entries the warp spends a step on those, and drains what is left
  // at the end. A key that has probed RG_PROBE_MAX slots goes to the table's overflow region.
  {
    unsigned char* wbase = smem_raw + RG_TILE_KEY_BYTES + RG_TILE_CNT_BYTES + 64 + (size_t)warp * TILE_WARP_SMEM;
    uint64_t* qkey = reinterpret_cast<uint64_t*>(wbase);                       // [TILE_Q]
    uint4* rq = reinterpret_cast<uint4*>(wbase + TILE_Q * 8);                  // [32]
    uint32_t* qsp = reinterpret_cast<uint32_t*>(wbase + TILE_Q * 8 + 512);     // [TILE_Q]
    uint32_t* qadd = qsp + TILE_Q;                                             // [TILE_Q]
    uint32_t* bm = qadd + TILE_Q;                                              // [16]
    const uint32_t lt_mask = (1u << lane) - 1u, le_mask = lt_mask | (1u << lane);
    uint32_t qhead = 0, qlen = 0;     // warp-uniform
    const uint32_t k = tv.geo.k;

    const uint32_t skeys_a = smem_u32(skeys), scounts_a = smem_u32(scounts);
    // NR probes per lane (NR independent dependency chains: with ~6 warps per scheduler the fixed latencies of
    // one chain leave issue slots empty): key[i] at slot (sp[i] & mask) for the lanes with have[i]; fail[i] =
    // the lane met another key. The CAS (slot seen empty) and the count increment are predicated.
    auto probe = [&](auto nr_tag, const uint64_t* key, const uint32_t* sp, const uint32_t* add, const bool* have, bool* fail) {
      constexpr int NR = decltype(nr_tag)::value;
      unsigned long long seen[NR];
      uint32_t ka[NR];
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        ka[i] = skeys_a + (sp[i] & (RG_TILE_SLOTS - 1)) * 8u;
        asm volatile("ld.volatile.shared.b64 %0, [%1];" : "=l"(seen[i]) : "r"(ka[i]) : "memory");
      }
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        const uint32_t try_claim = (have[i] && seen[i] == RG_EMPTY) ? 1u : 0u;
        asm volatile(
            "{\n"
            ".reg .pred q;\n"
            "setp.ne.u32 q, %2, 0;\n"
            "@q atom.shared.cas.b64 %0, [%1], %3, %4;\n"
            "}\n"
            : "+l"(seen[i])
            : "r"(ka[i]), "r"(try_claim), "l"((unsigned long long)RG_EMPTY), "l"((unsigned long long)key[i])
            : "memory");
      }
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        const bool hit = have[i] && ((seen[i] == RG_EMPTY) | (seen[i] == key[i]));
        asm volatile(
            "{\n"
            ".reg .pred q;\n"
            "setp.ne.u32 q, %1, 0;\n"
            "@q red.shared.add.u32 [%0], %2;\n"
            "}\n" ::"r"(scounts_a + (sp[i] & (RG_TILE_SLOTS - 1)) * 4u),
            "r"(hit ? 1u : 0u), "r"(add[i])
            : "memory");
        fail[i] = have[i] && !hit;
      }
    };
    // the failed lanes push (key, next slot) onto the warp's retry queue (room for 32 more entries required)
    auto push = [&](uint64_t key, uint32_t sp, uint32_t add, bool fail) {
      const uint32_t fm = __ballot_sync(0xFFFFFFFFu, fail);
      if (fm) {                                             // warp-uniform
        if (fail) {
          // next slot, one more probe counted in the bits above the slot (a wrap of the slot counts one more: harmless)
          const uint32_t at = (qhead + qlen + __popc(fm & lt_mask)) & (TILE_Q - 1);
          qkey[at] = key;
          qsp[at] = sp + RG_TILE_SLOTS + 1u;
          qadd[at] = add;
        }
        qlen += __popc(fm);
      }
    };
    // a step on (up to) 32 queued keys
    auto drain = [&]() {
      __syncwarp();                                         // the queue writes of the previous step are visible
      const uint32_t take = min(32u, qlen);
      const uint32_t at = (qhead + lane) & (TILE_Q - 1);
      const uint64_t key = qkey[at];
      const uint32_t sp = qsp[at];
      const uint32_t add = qadd[at];
      __syncwarp();                                         // every lane has its entry before the step overwrites the ring
      bool have = lane < take;
      qhead = (qhead + take) & (TILE_Q - 1);
      qlen -= take;
      if (have && (sp >> RG_TILE_BITS) >= RG_PROBE_MAX) {   // probe window exhausted: the key lives in the overflow region
        rg_ovf_add(tv, key, add);
        have = false;
      }
      bool fail;
      probe(std::integral_constant<int, 1>{}, &key, &sp, &add, &have, &fail);
      push(key, sp, add, fail);
    };

    for (uint32_t b0 = warp * 32; b0 < n; b0 += TILE_WARPS * 32) {
      const uint4 rec = rec_next;
      const uint32_t nb = b0 + TILE_WARPS * 32 + lane;
      rec_next = nb < n ? __ldg(bucket + nb) : none;
      const uint32_t len = b0 + lane < n ? (rec.w & 15u) + 1u : 0u;
      uint32_t incl = len;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += v;
      }
      const uint32_t start = incl - len;
      const uint32_t T = __shfl_sync(0xFFFFFFFFu, incl, 31);
      __syncwarp();                                         // the previous batch's readers are done with rq / bm
      rq[lane] = rec;
      if (lane < 16) bm[lane] = 0;
      __syncwarp();
      if (len) atomicOr(&bm[start >> 5], 1u << (start & 31));
      __syncwarp();
      uint32_t carry = 0xFFFFFFFFu;                         // (records that start before this round) - 1, warp-uniform
      // k-mer j of the batch -> (key, home slot): the record it belongs to = (number of record starts at or
      // before j) - 1; lanes past the batch's end (last round only) wrap to some valid record and are masked
      auto fetch = [&](uint32_t j0, uint64_t& key, uint32_t& slot, uint32_t& add) {
        const uint32_t word = bm[j0 >> 5];
        const uint32_t r = (carry + __popc(word & le_mask)) & 31u;
        carry += __popc(word);
        const uint32_t o = (j0 + lane - __shfl_sync(0xFFFFFFFFu, start, r)) & 15u;
        const uint4 rr = rq[r];
        key = sk_key(rr, o, k, canonical);
        slot = rg_slot_of_key(key);
        add = rr.w >> 4;                                    // how many identical records this one stands for (dedup_kernel)
      };
      uint32_t j0 = 0;
      for (; j0 + 32 < T; j0 += 64) {                       // two rounds at a time, their chains interleaved
        uint64_t key[2];
        uint32_t slot[2], add[2];
        fetch(j0, key[0], slot[0], add[0]);
        fetch(j0 + 32, key[1], slot[1], add[1]);
        const bool have[2] = {true, j0 + 32 + lane < T};
        bool fail[2];
        probe(std::integral_constant<int, 2>{}, key, slot, add, have, fail);
        push(key[0], slot[0], add[0], fail[0]);
        while (qlen >= 32) drain();
        push(key[1], slot[1], add[1], fail[1]);
        while (qlen >= 32) drain();
      }
      if (j0 < T) {
        uint64_t key;
        uint32_t slot, add;
        fetch(j0, key, slot, add);
        const bool have = j0 + lane < T;
        bool fail;
        probe(std::integral_constant<int, 1>{}, &key, &slot, &add, &have, &fail);
        push(key, slot, add, fail);
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

// ---------------------------------------------------------------------------
// Path 2, step 1.5: identical records of a tile's bucket collapse into one record + multiplicity. At 30x the
// same locus is read ~30 times and a super-k-mer record is a pure function of the locus' bases (and of where
// the 16-base block cut fell), so a bucket holds each record several times: the tile insert then probes the
// hash table once per DISTINCT record k-mer, with the multiplicity as the increment. One CTA per bucket, a
// shared-memory set of DD_SLOTS records keyed by the full 96 bits of bases + length (two 64-bit CAS words per
// slot: either word is written once, and a record counts itself only in a slot where BOTH match), emptied into
// the bucket IN PLACE whenever it could fill up (outputs never pass the inputs consumed). Record format out:
// w = (multiplicity << 4) | (len - 1): the hash prefix is no longer needed once a record sits in its bucket.
// ---------------------------------------------------------------------------
constexpr uint32_t DD_SLOTS = 13 * 1024;                    // 16-byte entries: 208 KiB of shared memory
constexpr int DD_THREADS = 1024;
constexpr uint32_t DD_RPT = 4;                              // records per thread per chunk
constexpr uint32_t DD_CHUNK = DD_THREADS * DD_RPT;
constexpr uint32_t DD_PROBES = 16;
constexpr uint32_t DD_FLUSH_AT = DD_SLOTS * 7 / 16;         // keep the load low: a warp pays for its longest probe sequence
constexpr size_t DD_SMEM = (size_t)DD_SLOTS * 16 + 64;

__global__ void __launch_bounds__(DD_THREADS, 1)
dedup_kernel(uint4* __restrict__ part_recs, uint32_t* __restrict__ cursor, uint32_t bucket_cap, uint32_t k) {
  extern __shared__ __align__(16) unsigned char dd_raw[];
  // entry = A: bases 0..31 | B: bases 32..47 << 32 | multiplicity << 4 | len - 1. A is claimed by CAS, then B by CAS
  // (multiplicity 0), then the multiplicity is raised by atomic adds of 16 on B's low word: either word's key part is
  // written once, and a record counts itself only in a slot where BOTH match.
  unsigned long long* A = reinterpret_cast<unsigned long long*>(dd_raw);
  unsigned long long* B = A + DD_SLOTS;
  uint32_t* s_out = reinterpret_cast<uint32_t*>(B + DD_SLOTS);                             // [0] records written, [1] set entries
  constexpr unsigned long long KEYPART = 0xFFFFFFFF0000000Full;
  const uint32_t tile = blockIdx.x, t = threadIdx.x;
  uint32_t n = cursor[tile];
  if (n > bucket_cap) n = bucket_cap;
  if (n == 0) return;
  uint4* bucket = part_recs + (uint64_t)tile * bucket_cap;
  for (uint32_t i = t; i < DD_SLOTS; i += DD_THREADS) {
    A[i] = ~0ull;
    B[i] = ~0ull;
  }
  if (t < 2) s_out[t] = 0;
  __syncthreads();
  const uint32_t lane = t & 31u, lt_mask = (1u << lane) - 1u;
  // a warp reserves output positions with ONE shared atomic (thousands of atomics on one address serialise)
  auto reserve = [&](bool want) -> uint32_t {
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, want);
    uint32_t base = 0;
    if (m && lane == 0) base = atomicAdd(&s_out[0], (uint32_t)__popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    return base + __popc(m & lt_mask);
  };
  auto flush = [&]() {                      // every set entry -> one record with its multiplicity; the set is emptied
    for (uint32_t i = t; i < DD_SLOTS; i += DD_THREADS) {
      const unsigned long long a = A[i], b = B[i];
      const bool used = a != ~0ull;          // (a claimed A always gets its B and a multiplicity >= 1 before the next barrier)
      const uint32_t pos = reserve(used);
      if (used) {
        bucket[pos] = make_uint4((uint32_t)(a >> 32), (uint32_t)a, (uint32_t)(b >> 32), (uint32_t)b);
        A[i] = ~0ull;
        B[i] = ~0ull;
      }
    }
  };
  // one chunk ahead: the next chunk's loads are in flight while this one is inserted
  uint4 nxt[DD_RPT];
#pragma unroll
  for (uint32_t w = 0; w < DD_RPT; ++w) {
    const uint32_t i = w * DD_THREADS + t;
    nxt[w] = i < n ? bucket[i] : make_uint4(0u, 0u, 0u, 0u);
  }
  for (uint32_t base = 0; base < n; base += DD_CHUNK) {
    // a chunk's records are in registers before anything is written back: outputs may land on their positions
    // (outputs never pass the records consumed, and the prefetched chunk counts as consumed)
    uint4 rec[DD_RPT];
#pragma unroll
    for (uint32_t w = 0; w < DD_RPT; ++w) {
      rec[w] = nxt[w];
      const uint32_t i = base + DD_CHUNK + w * DD_THREADS + t;
      nxt[w] = i < n ? bucket[i] : make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    // empty the set before it gets crowded: a chunk may still overfill it, records that find their neighbourhood
    // full are written out as they are
    if (s_out[1] > DD_FLUSH_AT) {
      __syncthreads();
      flush();
      __syncthreads();
      if (t == 0) s_out[1] = 0;
      __syncthreads();
    }
#pragma unroll
    for (uint32_t w = 0; w < DD_RPT; ++w) {
      const bool have = base + w * DD_THREADS + t < n;
      const uint32_t len1 = rec[w].w & 15u;
      // bases beyond the record's len + k - 1 belong to what follows in the read: zero them, they are not part of the
      // record (measured: comparing them as they are costs a quarter of the merges)
      const uint32_t nb = len1 + k;                          // bases held (<= 47)
      uint32_t x = rec[w].x;
      unsigned long long yz = ((unsigned long long)rec[w].y << 32) | rec[w].z;
      if (nb >= 16) {
        yz &= ~(~0ull >> (2 * (nb - 16)));                   // (nb == 16: nothing of y, z is kept)
      } else {
        yz = 0;
        x &= ~(0xFFFFFFFFu >> (2 * nb));
      }
      const uint32_t y = (uint32_t)(yz >> 32), z = (uint32_t)yz;
      const unsigned long long a = ((unsigned long long)x << 32) | y, b = ((unsigned long long)z << 32) | len1;
      bool placed = false, fresh = false;
      if (have && a != ~0ull) {
        uint32_t h = x * 0x9E3779B1u + y * 0x85EBCA77u + (z ^ len1) * 0xC2B2AE35u;
        h ^= h >> 15;
        uint32_t slot = (uint32_t)(((uint64_t)(h * 0x2C1B3C6Du) * DD_SLOTS) >> 32);
        uint32_t probes = 0;
        while (!placed && probes < DD_PROBES) {
          // (1) walk past the slots other records own: a tiny loop, a warp pays for its longest walk
          unsigned long long sa;
          for (;;) {
            sa = *(volatile unsigned long long*)&A[slot];
            if (sa == a || sa == ~0ull || ++probes >= DD_PROBES) break;
            slot = slot + 1 == DD_SLOTS ? 0u : slot + 1;
          }
          if (probes >= DD_PROBES) break;
          // (2) claim or join; most records repeat one that is already in the set (plain loads, one atomic add)
          if (sa == ~0ull) sa = atomicCAS(&A[slot], ~0ull, a);
          if (sa == ~0ull || sa == a) {
            unsigned long long sb = *(volatile unsigned long long*)&B[slot];
            if (sb == ~0ull) sb = atomicCAS(&B[slot], ~0ull, b);
            if (sb == ~0ull || ((sb ^ b) & KEYPART) == 0) {
              const uint32_t old = atomicAdd(reinterpret_cast<uint32_t*>(&B[slot]), 16u);   // low word (little-endian)
              fresh = (old >> 4) == 0;
              placed = true;
            }
          }
          if (!placed) {                                      // lost the slot to another record: go on
            ++probes;
            slot = slot + 1 == DD_SLOTS ? 0u : slot + 1;
          }
        }
      }
      const uint32_t fm = __ballot_sync(0xFFFFFFFFu, fresh);
      if (fm && lane == 0) atomicAdd(&s_out[1], (uint32_t)__popc(fm));
      // crowded neighbourhood (or the all-T sentinel): as it is, once
      const uint32_t pos = reserve(have && !placed);
      if (have && !placed) bucket[pos] = make_uint4(x, y, z, (1u << 4) | len1);
    }
    __syncthreads();
  }
  flush();
  __syncthreads();
  if (t == 0) cursor[tile] = s_out[0];
}

int rg_launch_tile_insert(rg_table* t) {
  rg_ctx* ctx = t->ctx;
  RG_CUDA_OK(cudaFuncSetAttribute(dedup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DD_SMEM));
  {
    rg_launch_scope _ls(ctx, "dedup_kernel");
    dedup_kernel<<<t->n_tiles, DD_THREADS, DD_SMEM, ctx->stream>>>(t->part_keys, t->part_cursor, t->bucket_cap, (uint32_t)t->k);
  }
  RG_CUDA_OK(cudaGetLastError());
  const size_t smem = TILE_SMEM;
  // (the attribute is per device: set on every launch, a process may drive several GPUs)
  RG_CUDA_OK(cudaFuncSetAttribute(tile_insert_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  {
    rg_launch_scope _ls(ctx, "tile_insert_kernel");
    tile_insert_kernel<<<t->n_tiles, TILE_THREADS, smem, ctx->stream>>>(rg_view(t), t->part_keys, t->part_cursor,
                                                                        t->bucket_cap, t->canonical,
                                                                        t->tiles_dirty ? 1 : 0);
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
extract_parts_kernel(ExtractParams ep, RgGeo geo, uint32_t groups_per_part, unsigned long long* __restrict__ counters,
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
        // owner = shard of the key's hash-prefix group (slow per-key minimizer: this exchange moves packed keys)
        const uint32_t part = rg_group_of_mh(rg_min_hash(key, geo.k, geo.m), geo.groups) / groups_per_part;
        myslot[w] = (part << 16) | atomicAdd(&local_cnt[part], 1u);
      }
    });
    __syncthreads();
    if (threadIdx.x < 64) {
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

int rg_launch_extract_partition(rg_ctx* ctx, const ExtractParams& ep, RgGeo geo, uint32_t n_parts, uint64_t* keys_out,
                                uint64_t* part_offsets, uint64_t* d_tmp) {
  if (n_parts == 0 || n_parts > 64 || geo.groups % n_parts != 0) {
    rg_set_error("rg_extract_partition: n_parts must be 1..64 and divide the table's groups");
    return RG_ERR_ARG;
  }
  const uint32_t gpp = geo.groups / n_parts;
  // d_tmp: [0..63] counts, [64..127] cursors
  RG_CUDA_OK(cudaMemsetAsync(d_tmp, 0, 128 * sizeof(uint64_t), ctx->stream));
  const uint64_t n_chunks = (ep.n_bytes + CHUNK - 1) / CHUNK;
  unsigned grid = rg_grid(ctx, 6);
  if (n_chunks < grid) grid = (unsigned)(n_chunks ? n_chunks : 1);
  unsigned long long* cnt = reinterpret_cast<unsigned long long*>(d_tmp);
  unsigned long long* cur = cnt + 64;
  {
    rg_launch_scope _ls(ctx, "extract_parts_kernel<count>");
    extract_parts_kernel<false><<<grid, COUNT_THREADS, 0, ctx->stream>>>(ep, geo, gpp, cnt, nullptr);
  }
  {
    rg_launch_scope _ls(ctx, "parts_scan_kernel");
    parts_scan_kernel<<<1, 32, 0, ctx->stream>>>(n_parts, cnt, part_offsets, cur);
  }
  {
    rg_launch_scope _ls(ctx, "extract_parts_kernel<scatter>");
    extract_parts_kernel<true><<<grid, COUNT_THREADS, 0, ctx->stream>>>(ep, geo, gpp, cur, keys_out);
  }
  RG_CUDA_OK(cudaGetLastError());
  return 0;
}
