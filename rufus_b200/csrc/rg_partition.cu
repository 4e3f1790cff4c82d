// rg_partition.cu — RG_MODE_TILED, step 1: radix partition of packed k-mers by hash
// prefix into one bucket per table tile (the reference has no counterpart: its
// large_hash_array.hpp:298-302 insert goes straight to one DRAM-resident table).
//
// What the B200 measurements said (profiles/microbench_r1.txt, profiles/README.md):
// shared-memory atomics are cheap (0.14 cyc/lane/SM), MATCH.ANY is not (2 cyc/lane),
// and per-key 8-byte scattered global stores are the wall (a write footprint of
// (#CTAs x #buckets) open lines overflows L2). So the CTA counting-sorts a
// SUPER-CHUNK of windows in shared memory and writes each bucket's keys as ONE
// contiguous, 32-byte-aligned run:
//   1. extract keys -> ukeys[] (smem), histogram the bucket ids (ATOMS)
//   2. block scan -> local offsets; one global atomicAdd per non-empty bucket reserves
//      a run rounded up to 4 keys (a full 32-byte sector)
//   3. place window indices into spos[] in bucket order (ATOMS rank)
//   4. half-warp per bucket copies its run out; the pad slots get RG_EMPTY (the tile
//      insert skips them), so every global write is a fully covered sector.
#include "rg_extract.cuh"

constexpr int PART_THREADS = 1024;
constexpr uint32_t PART_MAX_TILES = 8192;
constexpr size_t PART_SMEM_LIMIT = 232448;         // 227 KiB opt-in limit per CTA (static + dynamic)

template <int R>
constexpr size_t part_static_smem() { return sizeof(StageT<PART_THREADS, R>) + 33 * sizeof(uint32_t) + 128; }
static size_t part_dynamic_smem(int r, uint32_t n_tiles) { return (size_t)r * PART_THREADS * 10 + (size_t)n_tiles * 12; }

struct PartGeom {
  uint32_t n_tiles_global, tile_first, n_tiles;
  uint64_t* part_keys;
  uint32_t* cursor;
  uint32_t bucket_cap;
  uint64_t* status;
  uint64_t* spill;
  uint64_t spill_cap;
};

// exclusive scan of v over the block (blockDim.x multiple of 32); returns the prefix, total via *total
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* warp_sums, uint32_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= o) incl += n;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < nwarps ? warp_sums[lane] : 0;
    uint32_t wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, wi, o);
      if (lane >= o) wi += n;
    }
    warp_sums[lane] = wi - w;          // exclusive prefix of each warp
    if (lane == 31) *total = wi;
  }
  __syncthreads();
  return warp_sums[warp] + incl - v;
}

// FROM_KEYS = false: windows of the ASCII stream; true: an already extracted key list.
// R = windows (keys) per thread per super-chunk; K = PART_THREADS * R keys are sorted at once.
// ukeys is TRANSPOSED (key w of thread t at w*PART_THREADS + t) so that a warp's accesses are
// conflict-free.
template <bool FROM_KEYS, int R>
__global__ void __launch_bounds__(PART_THREADS, 1)
partition_sorted_kernel(ExtractParams ep, const uint64_t* __restrict__ in_keys, uint64_t n_in, PartGeom g) {
  extern __shared__ __align__(16) unsigned char dyn[];
  using Stage = StageT<PART_THREADS, R>;
  __shared__ Stage sm;
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t scan_total;
  constexpr uint32_t K = PART_THREADS * R;
  uint64_t* ukeys = reinterpret_cast<uint64_t*>(dyn);                 // [K]
  uint16_t* spos = reinterpret_cast<uint16_t*>(dyn + (size_t)K * 8);  // [K] key indices in bucket order
  uint32_t* hist = reinterpret_cast<uint32_t*>(dyn + (size_t)K * 10); // [n_tiles] count, then remaining rank
  uint32_t* loff = hist + g.n_tiles;                                  // [n_tiles] local offset of the run
  uint32_t* gbase = loff + g.n_tiles;                                 // [n_tiles] offset of the run in its bucket
  const uint32_t per_thread = (g.n_tiles + PART_THREADS - 1) / PART_THREADS;

  const uint64_t n_units = FROM_KEYS ? n_in : ep.n_bytes;
  const uint64_t n_super = (n_units + K - 1) / K;
  for (uint64_t sc = blockIdx.x; sc < n_super; sc += gridDim.x) {
    const uint64_t s0 = sc * K;
    for (uint32_t i = threadIdx.x; i < g.n_tiles; i += blockDim.x) hist[i] = 0;
    if (!FROM_KEYS) stage_super(sm, ep, s0);
    __syncthreads();
    // ---- 1. keys -> ukeys, histogram of bucket ids
    uint32_t foreign = 0;
    auto take = [&](int w, uint64_t key) {
      if (key != RG_EMPTY) {
        const uint32_t tile = rg_tile_of_key(key, g.n_tiles_global) - g.tile_first;
        if (tile >= g.n_tiles) {
          ++foreign;
          key = RG_EMPTY;
        } else {
          atomicAdd(&hist[tile], 1u);
        }
      }
      ukeys[w * PART_THREADS + threadIdx.x] = key;
    };
    if (FROM_KEYS) {
#pragma unroll
      for (int w = 0; w < R; ++w) {
        const uint64_t i = s0 + (uint64_t)w * PART_THREADS + threadIdx.x;
        take(w, i < n_in ? __ldg(in_keys + i) : RG_EMPTY);
      }
    } else {
      roll_windows(sm, threadIdx.x, ep.k, ep.canonical, take);
    }
    if (foreign) atomicAdd((unsigned long long*)&g.status[1], (unsigned long long)foreign);
    __syncthreads();
    // ---- 2. local offsets (exclusive scan of the counts) + global run reservation (padded to 4 keys)
    {
      const uint32_t t0 = threadIdx.x * per_thread;
      uint32_t sum = 0;
      for (uint32_t j = 0; j < per_thread; ++j)
        if (t0 + j < g.n_tiles) sum += hist[t0 + j];
      uint32_t run = block_exclusive_scan(sum, warp_sums, &scan_total);
      for (uint32_t j = 0; j < per_thread; ++j) {
        const uint32_t t = t0 + j;
        if (t < g.n_tiles) {
          const uint32_t cnt = hist[t];
          loff[t] = run;
          run += cnt;
          uint32_t gb = 0;
          if (cnt) {
            const uint32_t padded = (cnt + 3u) & ~3u;
            gb = atomicAdd(g.cursor + t, padded);
            uint64_t* dst = g.part_keys + (uint64_t)t * g.bucket_cap;
            if (gb + padded <= g.bucket_cap) {
              for (uint32_t q = cnt; q < padded; ++q) dst[gb + q] = RG_EMPTY;      // pad slots of the run
            } else {
              // bucket overflow (heavy hitters): this run goes to the spill list (step 4); what it
              // reserved below the capacity is neutralised with RG_EMPTY
              for (uint32_t q = gb; q < g.bucket_cap; ++q) dst[q] = RG_EMPTY;
              gb = 0xFFFFFFFFu;
            }
          }
          gbase[t] = gb;
        }
      }
    }
    __syncthreads();
    // ---- 3. bucket-order placement of the key indices (hist counts down to the rank)
#pragma unroll
    for (int w = 0; w < R; ++w) {
      const uint32_t q = w * PART_THREADS + threadIdx.x;
      const uint64_t key = ukeys[q];
      if (key == RG_EMPTY) continue;
      const uint32_t tile = rg_tile_of_key(key, g.n_tiles_global) - g.tile_first;
      const uint32_t r = atomicSub(&hist[tile], 1u) - 1u;      // cnt-1 .. 0
      spos[loff[tile] + r] = (uint16_t)q;
    }
    __syncthreads();
    // ---- 4. copy-out in bucket order: consecutive threads write consecutive keys of a run
    {
      const uint32_t total = scan_total;
      for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) {
        const uint64_t key = ukeys[spos[i]];
        const uint32_t tile = rg_tile_of_key(key, g.n_tiles_global) - g.tile_first;
        const uint32_t gb = gbase[tile];
        if (gb != 0xFFFFFFFFu) {
          g.part_keys[(uint64_t)tile * g.bucket_cap + gb + (i - loff[tile])] = key;
        } else {
          const uint64_t sp = atomicAdd((unsigned long long*)&g.status[0], 1ull);
          if (sp < g.spill_cap) g.spill[sp] = key;
        }
      }
    }
    __syncthreads();
  }
}

// Fallback for tables with more tiles than the sorted kernel can histogram in shared memory:
// one global cursor atomic and one scattered 8-byte store per key (slow; large-table path).
template <bool FROM_KEYS>
__global__ void __launch_bounds__(256)
partition_scatter_kernel(ExtractParams ep, const uint64_t* __restrict__ in_keys, uint64_t n_in, PartGeom g) {
  using Stage = StageT<256, 16>;
  __shared__ Stage sm;
  auto put = [&](uint64_t key) {
    const uint32_t tile = rg_tile_of_key(key, g.n_tiles_global) - g.tile_first;
    if (tile >= g.n_tiles) {
      atomicAdd((unsigned long long*)&g.status[1], 1ull);
      return;
    }
    const uint32_t at = atomicAdd(g.cursor + tile, 1u);
    if (at < g.bucket_cap) {
      g.part_keys[(uint64_t)tile * g.bucket_cap + at] = key;
    } else {
      const uint64_t sp = atomicAdd((unsigned long long*)&g.status[0], 1ull);
      if (sp < g.spill_cap) g.spill[sp] = key;
    }
  };
  if (FROM_KEYS) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_in; i += (uint64_t)gridDim.x * blockDim.x)
      put(__ldg(in_keys + i));
  } else {
    const uint64_t n_chunks = (ep.n_bytes + Stage::WINDOWS - 1) / Stage::WINDOWS;
    for (uint64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
      __syncthreads();
      stage_super(sm, ep, c * Stage::WINDOWS);
      __syncthreads();
      roll_windows(sm, threadIdx.x, ep.k, ep.canonical, [&](int, uint64_t key) {
        if (key != RG_EMPTY) put(key);
      });
    }
  }
}

// keys per thread per super-chunk (R) that fits shared memory for this tile count; 0 = scatter fallback
static int part_r_for(uint32_t n_tiles) {
  if (n_tiles > PART_MAX_TILES) return 0;
  if (part_dynamic_smem(20, n_tiles) + part_static_smem<20>() <= PART_SMEM_LIMIT) return 20;
  if (part_dynamic_smem(16, n_tiles) + part_static_smem<16>() <= PART_SMEM_LIMIT) return 16;
  if (part_dynamic_smem(12, n_tiles) + part_static_smem<12>() <= PART_SMEM_LIMIT) return 12;
  if (part_dynamic_smem(8, n_tiles) + part_static_smem<8>() <= PART_SMEM_LIMIT) return 8;
  return 0;
}

// upper bound on the pad keys one bucket can receive in a push of n_units windows/keys:
// up to 3 per (super-chunk, bucket) run
uint64_t rg_partition_pad_keys(uint32_t n_tiles, uint64_t n_units) {
  const int r = part_r_for(n_tiles);
  if (r == 0) return 0;
  const uint64_t K = (uint64_t)r * PART_THREADS;
  return 3 * ((n_units + K - 1) / K);
}

template <bool FROM_KEYS, int R>
static int launch_sorted(rg_ctx* ctx, const char* name, const ExtractParams& ep, const uint64_t* keys, uint64_t n_keys,
                         const PartGeom& g, uint64_t n_units) {
  const size_t smem = part_dynamic_smem(R, g.n_tiles);
  RG_CUDA_OK(cudaFuncSetAttribute(partition_sorted_kernel<FROM_KEYS, R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
  const uint64_t K = (uint64_t)R * PART_THREADS;
  const uint64_t n_super = (n_units + K - 1) / K;
  unsigned grid = rg_grid(ctx, 1);
  if (n_super < grid) grid = (unsigned)n_super;
  rg_launch_scope _ls(ctx, name);
  partition_sorted_kernel<FROM_KEYS, R><<<grid, PART_THREADS, smem, ctx->stream>>>(ep, keys, n_keys, g);
  return 0;
}

template <bool FROM_KEYS>
static int launch_partition(rg_table* t, const ExtractParams& ep, const uint64_t* keys, uint64_t n_keys) {
  rg_ctx* ctx = t->ctx;
  PartGeom g;
  g.n_tiles_global = t->n_tiles_global;
  g.tile_first = t->tile_first;
  g.n_tiles = t->n_tiles;
  g.part_keys = t->part_keys;
  g.cursor = t->part_cursor;
  g.bucket_cap = t->bucket_cap;
  g.status = t->d_status;
  g.spill = t->spill;
  g.spill_cap = t->spill_cap;
  const uint64_t n_units = FROM_KEYS ? n_keys : ep.n_bytes;
  if (n_units == 0) return 0;
  const char* name = FROM_KEYS ? "partition_keys_kernel" : "partition_kernel";
  int rc = 0;
  switch (part_r_for(t->n_tiles)) {
    case 20: rc = launch_sorted<FROM_KEYS, 20>(ctx, name, ep, keys, n_keys, g, n_units); break;
    case 16: rc = launch_sorted<FROM_KEYS, 16>(ctx, name, ep, keys, n_keys, g, n_units); break;
    case 12: rc = launch_sorted<FROM_KEYS, 12>(ctx, name, ep, keys, n_keys, g, n_units); break;
    case 8: rc = launch_sorted<FROM_KEYS, 8>(ctx, name, ep, keys, n_keys, g, n_units); break;
    default: {
      rg_launch_scope _ls(ctx, name);
      partition_scatter_kernel<FROM_KEYS><<<rg_grid(ctx, 8), 256, 0, ctx->stream>>>(ep, keys, n_keys, g);
    }
  }
  if (rc) return rc;
  RG_CUDA_OK(cudaGetLastError());
  return 0;
}

int rg_launch_partition(rg_table* t, const ExtractParams& ep) { return launch_partition<false>(t, ep, nullptr, 0); }

int rg_launch_partition_keys(rg_table* t, const uint64_t* dev_keys, uint64_t n) {
  ExtractParams ep = {};
  return launch_partition<true>(t, ep, dev_keys, n);
}
