// rg_count.cu — k-mer extraction/canonicalisation and the two insert paths.
//
// Replaces the reference's hot loop `for(;mers;++mers) ary_.add(*mers,1)`
// (src/modifiedJellyfish/sub_commands/count_main.cc:154-157):
//   mer_iterator::operator++   mer_iterator.hpp:61-89   -> stage_chunk + window_key
//   array::add / claim_key     large_hash_array.hpp:298-302,513-601 -> rg_table_add / tile_insert_kernel
#include "rg_internal.cuh"

// ---------------------------------------------------------------------------
// Staging: a CTA converts CHUNK+64 ASCII bytes of the read stream into three
// big-endian bit streams in shared memory:
//   codes : 2 bits/base, A0 C1 G2 T3 (mer_dna.hpp:46-63), first base in the MSBs
//   bad   : 1 bit/base, set for every byte outside ACGTacgt (window restart)
//   brk   : 1 bit/base, set where a new read starts (offset / fixed-stride inputs)
// A window starting at base p is then one funnel shift away.
// ---------------------------------------------------------------------------
constexpr int CHUNK = 4096;                       // window starts per chunk
constexpr int STAGE_BASES = CHUNK + 64;           // staged bases (k-1 <= 31 tail + word slack)
constexpr int CODE_WORDS = STAGE_BASES / 16;      // u32 words of 16 bases
constexpr int BIT_WORDS = STAGE_BASES / 32;       // u32 words of 32 flags
constexpr int COUNT_THREADS = 256;

struct StageSmem {
  uint32_t codes[CODE_WORDS + 2];
  uint32_t bad[BIT_WORDS + 2];
  uint32_t brk[BIT_WORDS + 2];
};

__device__ __forceinline__ void convert16(const uint4 v, uint32_t& codes, uint32_t& bad) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  codes = 0;
  bad = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t x = (w[i] >> 1) & 0x03030303u;             // A0 C1 G3 T2 per byte
    const uint32_t c4 = x ^ ((x >> 1) & 0x01010101u);         // A0 C1 G2 T3
    // gather the four 2-bit codes, first char (low byte) into the top bits of a byte
    const uint32_t byte = (c4 * 0x40100401u) >> 24;
    codes = (codes << 8) | byte;
    // validity: upper-cased byte must be one of A C G T
    const uint32_t u = w[i] & 0xDFDFDFDFu;
    uint32_t ok = __vcmpeq4(u, 0x41414141u) | __vcmpeq4(u, 0x43434343u) | __vcmpeq4(u, 0x47474747u) |
                  __vcmpeq4(u, 0x54545454u);                  // 0xFF per valid byte
    const uint32_t nb = ~ok & 0x01010101u;                    // 1 per bad byte
    const uint32_t nib = (nb * 0x08040201u) >> 24;            // bit3=char0 .. bit0=char3
    bad = (bad << 4) | (nib & 0xF);
  }
}

// Stage bases [c0, c0+STAGE_BASES) of the stream. c0 is a multiple of CHUNK.
__device__ __forceinline__ void stage_chunk(StageSmem& sm, const ExtractParams& ep, uint64_t c0) {
  uint16_t* bad16 = reinterpret_cast<uint16_t*>(sm.bad);
  for (int i = threadIdx.x; i < CODE_WORDS; i += blockDim.x) {
    const uint64_t p = c0 + (uint64_t)i * 16;
    uint4 v = make_uint4(0x0a0a0a0au, 0x0a0a0a0au, 0x0a0a0a0au, 0x0a0a0a0au);
    if (p + 16 <= ep.n_bytes) {
      v = __ldg(reinterpret_cast<const uint4*>(ep.stream + p));
    } else if (p < ep.n_bytes) {
      unsigned char tmp[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) tmp[j] = (p + j < ep.n_bytes) ? (unsigned char)ep.stream[p + j] : (unsigned char)'\n';
      v = *reinterpret_cast<uint4*>(tmp);
    }
    uint32_t c, b;
    convert16(v, c, b);
    sm.codes[i] = c;
    bad16[i ^ 1] = (uint16_t)b;  // big-endian halves inside the little-endian u32
  }
  for (int i = threadIdx.x; i < BIT_WORDS; i += blockDim.x) {
    const uint64_t p = c0 + (uint64_t)i * 32;
    uint32_t w = 0;
    if (ep.brk != nullptr) {
      if (p < ep.n_bytes) w = __brev(__ldg(ep.brk + (p >> 5)));
    } else if (ep.read_stride != 0) {
      uint32_t r = (uint32_t)(p % ep.read_stride);
      uint32_t first = r == 0 ? 0 : ep.read_stride - r;      // offset of the first read start >= p
      for (uint32_t j = first; j < 32; j += ep.read_stride) w |= 0x80000000u >> j;
    }
    sm.brk[i] = w;
  }
  if (threadIdx.x < 2) {
    sm.codes[CODE_WORDS + threadIdx.x] = 0;
    sm.bad[BIT_WORDS + threadIdx.x] = 0xFFFFFFFFu;
    sm.brk[BIT_WORDS + threadIdx.x] = 0;
  }
}

// k-mer of the window starting at local base p; returns false if the window is not counted
__device__ __forceinline__ bool window_key(const StageSmem& sm, int p, int k, int canonical, uint64_t& key) {
  const int bj = p >> 5, bs = p & 31;
  const uint64_t badw = (((uint64_t)sm.bad[bj] << 32) | sm.bad[bj + 1]) << bs;
  if (badw >> (64 - k)) return false;
  const uint64_t brkw = (((uint64_t)sm.brk[bj] << 32) | sm.brk[bj + 1]) << bs;
  if (k > 1 && ((brkw << 1) >> (65 - k))) return false;   // a read starts inside (p, p+k)
  const int j = p >> 4, s = (p & 15) * 2;
  uint64_t x = ((uint64_t)sm.codes[j] << 32) | sm.codes[j + 1];
  if (s) x = (x << s) | (sm.codes[j + 2] >> (32 - s));
  const uint64_t fwd = x >> (64 - 2 * k);
  if (canonical) {
    const uint64_t rc = rg_revcomp_jf(fwd, k);
    key = fwd < rc ? fwd : rc;
  } else {
    key = fwd;
  }
  return true;
}

// ---------------------------------------------------------------------------
// Path 1 (RG_MODE_DIRECT): extraction fused with global-memory CAS/atomic insert.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(COUNT_THREADS)
count_direct_kernel(ExtractParams ep, TableView tv, uint64_t* status, uint64_t* spill, uint64_t spill_cap) {
  __shared__ StageSmem sm;
  const uint64_t n_chunks = (ep.n_bytes + CHUNK - 1) / CHUNK;
  for (uint64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const uint64_t c0 = c * CHUNK;
    __syncthreads();
    stage_chunk(sm, ep, c0);
    __syncthreads();
    for (int p = threadIdx.x; p < CHUNK; p += blockDim.x) {
      uint64_t key;
      if (c0 + p + ep.k > ep.n_bytes) break;
      if (!window_key(sm, p, ep.k, ep.canonical, key)) continue;
      const int rc = rg_table_add(tv, key, 1);
      if (rc == 1) {
        const uint64_t at = atomicAdd((unsigned long long*)&status[0], 1ull);
        if (at < spill_cap) spill[at] = key;
      } else if (rc == 2) {
        atomicAdd((unsigned long long*)&status[1], 1ull);
      }
    }
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
// Path 2 (RG_MODE_TILED), step 1: extraction fused with the radix partition by
// hash prefix. Bucket b (one per table tile) is a fixed-capacity slab of
// part_keys; a key that finds its bucket full goes to the spill list and is
// inserted by the direct path afterwards (heavy-hitter safety valve).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(COUNT_THREADS)
partition_kernel(ExtractParams ep, uint32_t n_tiles_global, uint32_t tile_first, uint32_t n_tiles,
                 uint64_t* __restrict__ part_keys, uint32_t* __restrict__ cursor, uint32_t bucket_cap,
                 uint64_t* status, uint64_t* spill, uint64_t spill_cap) {
  __shared__ StageSmem sm;
  const uint64_t n_chunks = (ep.n_bytes + CHUNK - 1) / CHUNK;
  for (uint64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const uint64_t c0 = c * CHUNK;
    __syncthreads();
    stage_chunk(sm, ep, c0);
    __syncthreads();
    for (int p = threadIdx.x; p < CHUNK; p += blockDim.x) {
      uint64_t key;
      if (c0 + p + ep.k > ep.n_bytes) break;
      if (!window_key(sm, p, ep.k, ep.canonical, key)) continue;
      const uint32_t tile = rg_tile_of(rg_mix64(key), n_tiles_global) - tile_first;
      if (tile >= n_tiles) {
        atomicAdd((unsigned long long*)&status[1], 1ull);
        continue;
      }
      const uint32_t at = atomicAdd(cursor + tile, 1u);
      if (at < bucket_cap) {
        part_keys[(uint64_t)tile * bucket_cap + at] = key;
      } else {
        const uint64_t sp = atomicAdd((unsigned long long*)&status[0], 1ull);
        if (sp < spill_cap) spill[sp] = key;
      }
    }
  }
}

// CTA-aggregated variant (the one normally used): a CTA walks a SUPER-CHUNK of
// SUPER x CHUNK windows twice. Pass 1 histograms the bucket ids in shared memory;
// one global atomicAdd per (super-chunk, non-empty bucket) then reserves a
// contiguous run in each bucket; pass 2 recomputes the keys (ALU is cheap, the
// ASCII is L2-hot) and scatters them at run_base + shared-memory rank. Global
// atomics drop from one per k-mer to ~n_tiles per 32 Ki k-mers.
constexpr int SUPER = 8;

__global__ void __launch_bounds__(COUNT_THREADS)
partition_agg_kernel(ExtractParams ep, uint32_t n_tiles_global, uint32_t tile_first, uint32_t n_tiles,
                     uint64_t* __restrict__ part_keys, uint32_t* __restrict__ cursor, uint32_t bucket_cap,
                     uint64_t* status, uint64_t* spill, uint64_t spill_cap) {
  __shared__ StageSmem sm;
  extern __shared__ uint32_t dyn[];
  uint32_t* hist = dyn;               // [n_tiles] pass 1: counts; pass 2: running rank
  uint32_t* base = dyn + n_tiles;     // [n_tiles] start of this CTA's run inside the bucket
  const uint64_t n_chunks = (ep.n_bytes + CHUNK - 1) / CHUNK;
  const uint64_t n_super = (n_chunks + SUPER - 1) / SUPER;
  for (uint64_t sc = blockIdx.x; sc < n_super; sc += gridDim.x) {
    for (uint32_t i = threadIdx.x; i < n_tiles; i += blockDim.x) hist[i] = 0;
    const uint64_t c_begin = sc * SUPER;
    const uint64_t c_end = (c_begin + SUPER < n_chunks) ? c_begin + SUPER : n_chunks;
    unsigned long long foreign = 0;
    for (uint64_t c = c_begin; c < c_end; ++c) {
      const uint64_t c0 = c * CHUNK;
      __syncthreads();
      stage_chunk(sm, ep, c0);
      __syncthreads();
      for (int p = threadIdx.x; p < CHUNK; p += blockDim.x) {
        uint64_t key;
        if (c0 + p + ep.k > ep.n_bytes) break;
        if (!window_key(sm, p, ep.k, ep.canonical, key)) continue;
        const uint32_t tile = rg_tile_of(rg_mix64(key), n_tiles_global) - tile_first;
        if (tile >= n_tiles) { ++foreign; continue; }
        atomicAdd(&hist[tile], 1u);
      }
    }
    if (foreign) atomicAdd((unsigned long long*)&status[1], foreign);
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n_tiles; i += blockDim.x) {
      const uint32_t cnt = hist[i];
      base[i] = cnt ? atomicAdd(cursor + i, cnt) : 0u;
      hist[i] = 0;
    }
    for (uint64_t c = c_begin; c < c_end; ++c) {
      const uint64_t c0 = c * CHUNK;
      __syncthreads();
      stage_chunk(sm, ep, c0);
      __syncthreads();
      for (int p = threadIdx.x; p < CHUNK; p += blockDim.x) {
        uint64_t key;
        if (c0 + p + ep.k > ep.n_bytes) break;
        if (!window_key(sm, p, ep.k, ep.canonical, key)) continue;
        const uint32_t tile = rg_tile_of(rg_mix64(key), n_tiles_global) - tile_first;
        if (tile >= n_tiles) continue;
        const uint32_t at = base[tile] + atomicAdd(&hist[tile], 1u);
        if (at < bucket_cap) {
          part_keys[(uint64_t)tile * bucket_cap + at] = key;
        } else {
          const uint64_t sp = atomicAdd((unsigned long long*)&status[0], 1ull);
          if (sp < spill_cap) spill[sp] = key;
        }
      }
    }
    __syncthreads();
  }
}

constexpr int KEYS_SUPER = 32768;   // keys per CTA round of the key-list variant

__global__ void __launch_bounds__(256)
partition_keys_agg_kernel(const uint64_t* __restrict__ keys, uint64_t n, uint32_t n_tiles_global, uint32_t tile_first,
                          uint32_t n_tiles, uint64_t* __restrict__ part_keys, uint32_t* __restrict__ cursor,
                          uint32_t bucket_cap, uint64_t* status, uint64_t* spill, uint64_t spill_cap) {
  extern __shared__ uint32_t dyn[];
  uint32_t* hist = dyn;
  uint32_t* base = dyn + n_tiles;
  const uint64_t n_super = (n + KEYS_SUPER - 1) / KEYS_SUPER;
  for (uint64_t sc = blockIdx.x; sc < n_super; sc += gridDim.x) {
    for (uint32_t i = threadIdx.x; i < n_tiles; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const uint64_t lo = sc * KEYS_SUPER;
    const uint64_t hi = lo + KEYS_SUPER < n ? lo + KEYS_SUPER : n;
    unsigned long long foreign = 0;
    for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
      const uint32_t tile = rg_tile_of(rg_mix64(__ldg(keys + i)), n_tiles_global) - tile_first;
      if (tile >= n_tiles) { ++foreign; continue; }
      atomicAdd(&hist[tile], 1u);
    }
    if (foreign) atomicAdd((unsigned long long*)&status[1], foreign);
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n_tiles; i += blockDim.x) {
      const uint32_t cnt = hist[i];
      base[i] = cnt ? atomicAdd(cursor + i, cnt) : 0u;
      hist[i] = 0;
    }
    __syncthreads();
    for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
      const uint64_t key = __ldg(keys + i);
      const uint32_t tile = rg_tile_of(rg_mix64(key), n_tiles_global) - tile_first;
      if (tile >= n_tiles) continue;
      const uint32_t at = base[tile] + atomicAdd(&hist[tile], 1u);
      if (at < bucket_cap) {
        part_keys[(uint64_t)tile * bucket_cap + at] = key;
      } else {
        const uint64_t sp = atomicAdd((unsigned long long*)&status[0], 1ull);
        if (sp < spill_cap) spill[sp] = key;
      }
    }
    __syncthreads();
  }
}

constexpr uint32_t AGG_MAX_TILES = 16384;   // 8 B of shared memory per tile

// same bucketing for an already-extracted key list (the receive side of the all-to-all)
__global__ void __launch_bounds__(256)
partition_keys_kernel(const uint64_t* __restrict__ keys, uint64_t n, uint32_t n_tiles_global, uint32_t tile_first,
                      uint32_t n_tiles, uint64_t* __restrict__ part_keys, uint32_t* __restrict__ cursor,
                      uint32_t bucket_cap, uint64_t* status, uint64_t* spill, uint64_t spill_cap) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t key = __ldg(keys + i);
    const uint32_t tile = rg_tile_of(rg_mix64(key), n_tiles_global) - tile_first;
    if (tile >= n_tiles) {
      atomicAdd((unsigned long long*)&status[1], 1ull);
      continue;
    }
    const uint32_t at = atomicAdd(cursor + tile, 1u);
    if (at < bucket_cap) {
      part_keys[(uint64_t)tile * bucket_cap + at] = key;
    } else {
      const uint64_t sp = atomicAdd((unsigned long long*)&status[0], 1ull);
      if (sp < spill_cap) spill[sp] = key;
    }
  }
}

int rg_launch_partition_keys(rg_table* t, const uint64_t* dev_keys, uint64_t n) {
  rg_ctx* ctx = t->ctx;
  if (n == 0) return 0;
  unsigned grid = rg_grid(ctx, 8);
  const uint64_t need = (n + 255) / 256;
  if (need < grid) grid = (unsigned)need;
  if (t->n_tiles <= AGG_MAX_TILES) {
    const size_t smem = (size_t)t->n_tiles * 8;
    static bool attr_set = false;
    if (!attr_set) {
      RG_CUDA_OK(cudaFuncSetAttribute(partition_keys_agg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(AGG_MAX_TILES * 8)));
      attr_set = true;
    }
    const uint64_t n_super = (n + KEYS_SUPER - 1) / KEYS_SUPER;
    unsigned g2 = rg_grid(ctx, smem > 48 * 1024 ? 1 : 4);
    if (n_super < g2) g2 = (unsigned)n_super;
    rg_launch_scope _ls(ctx, "partition_keys_kernel");
    partition_keys_agg_kernel<<<g2, 256, smem, ctx->stream>>>(dev_keys, n, t->n_tiles_global, t->tile_first,
                                                              t->n_tiles, t->part_keys, t->part_cursor, t->bucket_cap,
                                                              t->d_status, t->spill, t->spill_cap);
  } else {
    rg_launch_scope _ls(ctx, "partition_keys_kernel");
    partition_keys_kernel<<<grid, 256, 0, ctx->stream>>>(dev_keys, n, t->n_tiles_global, t->tile_first, t->n_tiles,
                                                         t->part_keys, t->part_cursor, t->bucket_cap, t->d_status,
                                                         t->spill, t->spill_cap);
  }
  RG_CUDA_OK(cudaGetLastError());
  return 0;
}

int rg_launch_partition(rg_table* t, const ExtractParams& ep) {
  rg_ctx* ctx = t->ctx;
  if (ep.n_bytes == 0) return 0;
  const uint64_t n_chunks = (ep.n_bytes + CHUNK - 1) / CHUNK;
  unsigned grid = rg_grid(ctx, 8);
  if (n_chunks < grid) grid = (unsigned)n_chunks;
  if (t->n_tiles <= AGG_MAX_TILES) {
    const size_t smem = (size_t)t->n_tiles * 8;
    static bool attr_set = false;
    if (!attr_set) {
      RG_CUDA_OK(cudaFuncSetAttribute(partition_agg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(AGG_MAX_TILES * 8)));
      attr_set = true;
    }
    const uint64_t n_super = (n_chunks + SUPER - 1) / SUPER;
    unsigned g2 = rg_grid(ctx, smem > 96 * 1024 ? 1 : (smem > 40 * 1024 ? 2 : 4));
    if (n_super < g2) g2 = (unsigned)n_super;
    rg_launch_scope _ls(ctx, "partition_kernel");
    partition_agg_kernel<<<g2, COUNT_THREADS, smem, ctx->stream>>>(ep, t->n_tiles_global, t->tile_first, t->n_tiles,
                                                                   t->part_keys, t->part_cursor, t->bucket_cap,
                                                                   t->d_status, t->spill, t->spill_cap);
  } else {
    rg_launch_scope _ls(ctx, "partition_kernel");
    partition_kernel<<<grid, COUNT_THREADS, 0, ctx->stream>>>(ep, t->n_tiles_global, t->tile_first, t->n_tiles,
                                                              t->part_keys, t->part_cursor, t->bucket_cap, t->d_status,
                                                              t->spill, t->spill_cap);
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
constexpr int TILE_THREADS = 1024;

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
                   const uint32_t* __restrict__ cursor, uint32_t bucket_cap, int load_existing, uint64_t* status,
                   uint64_t* spill, uint64_t spill_cap) {
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

  const uint64_t* src = part_keys + (uint64_t)tile * bucket_cap;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    const uint64_t key = __ldg(src + i);
    uint32_t slot = (uint32_t)rg_mix64(key) & (RG_TILE_SLOTS - 1);
    bool done = false;
    for (uint32_t probe = 0; probe < RG_TILE_SLOTS; ++probe) {
      unsigned long long cur = *(volatile unsigned long long*)(skeys + slot);
      if (cur == RG_EMPTY)
        cur = atomicCAS((unsigned long long*)(skeys + slot), (unsigned long long)RG_EMPTY, (unsigned long long)key);
      if (cur == RG_EMPTY || cur == key) {
        atomicAdd(scounts + slot, 1u);
        done = true;
        break;
      }
      slot = (slot + 1) & (RG_TILE_SLOTS - 1);
    }
    if (!done) {
      const uint64_t sp = atomicAdd((unsigned long long*)&status[0], 1ull);
      if (sp < spill_cap) spill[sp] = key;
    }
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
  const size_t smem = RG_TILE_KEY_BYTES + RG_TILE_CNT_BYTES + 64;
  static bool attr_set = false;
  if (!attr_set) {
    RG_CUDA_OK(cudaFuncSetAttribute(tile_insert_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  {
    rg_launch_scope _ls(ctx, "tile_insert_kernel");
    tile_insert_kernel<<<t->n_tiles, TILE_THREADS, smem, ctx->stream>>>(t->keys, t->counts, t->part_keys, t->part_cursor,
                                                                        t->bucket_cap, t->tiles_dirty ? 1 : 0,
                                                                        t->d_status, t->spill, t->spill_cap);
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
  __shared__ StageSmem sm;
  __shared__ unsigned int local_cnt[64];
  __shared__ unsigned long long local_base[64];
  const uint64_t n_chunks = (ep.n_bytes + CHUNK - 1) / CHUNK;
  for (uint64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const uint64_t c0 = c * CHUNK;
    __syncthreads();
    stage_chunk(sm, ep, c0);
    if (threadIdx.x < 64) local_cnt[threadIdx.x] = 0;
    __syncthreads();
    // CHUNK / COUNT_THREADS windows per thread, kept in registers between the two phases
    constexpr int PER = CHUNK / COUNT_THREADS;
    uint64_t mykeys[PER];
    uint32_t myslot[PER];
#pragma unroll
    for (int it = 0; it < PER; ++it) {
      const int p = threadIdx.x + it * COUNT_THREADS;
      myslot[it] = 0xFFFFFFFFu;
      uint64_t key;
      if (c0 + p + ep.k <= ep.n_bytes && window_key(sm, p, ep.k, ep.canonical, key)) {
        const uint32_t part = rg_tile_of(rg_mix64(key), n_parts);
        const uint32_t r = atomicAdd(&local_cnt[part], 1u);
        mykeys[it] = key;
        myslot[it] = (part << 16) | r;
      }
    }
    __syncthreads();
    if (threadIdx.x < n_parts) {
      const unsigned int cnt = local_cnt[threadIdx.x];
      local_base[threadIdx.x] = cnt ? atomicAdd(&counters[threadIdx.x], (unsigned long long)cnt) : 0ull;
    }
    __syncthreads();
    if (SCATTER) {
#pragma unroll
      for (int it = 0; it < PER; ++it) {
        if (myslot[it] != 0xFFFFFFFFu) {
          const uint32_t part = myslot[it] >> 16, r = myslot[it] & 0xFFFFu;
          keys_out[local_base[part] + r] = mykeys[it];
        }
      }
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
