// rg_partition.cu — RG_MODE_TILED, step 1: reads -> super-k-mer records, radix-partitioned by the
// minimizer's hash prefix into one bucket per table tile (the reference has no counterpart: its
// large_hash_array.hpp:298-302 insert goes straight to one DRAM-resident table, one k-mer at a time).
//
// Round 1 partitioned the 8-byte k-mers themselves (~104 thread-instructions and 16 bytes of slab
// traffic per k-mer: histogram atomic, scatter and copy-out PER KEY). Here a CTA turns a SUPER-CHUNK of
// 16384 window starts into ~3000 records (rg_sk.cuh: the k-mers of a read that share a minimizer and
// start in the same 16-base block travel as one 16-byte record), so the counting sort, its atomics and
// the copy-out handle ~5x fewer items and the slabs (and NVLink, multi-GPU) carry ~3 bytes per k-mer:
//   1. ASCII -> 2-bit streams in shared memory (forward big-endian + complemented little-endian);
//   2. every thread hashes the 16 m-mers that start in its block: y = ((fwd * C1) ^ (rc * C1)) + (fwd & rc) * C2,
//      two funnel shifts, three multiply-adds, two logic ops each (strand-symmetric without a compare); y goes to shared
//      memory, transposed (conflict-free);
//   3. sliding minimum over the W = k-m+1 m-mers of each k-mer (log-step doubling in registers, W is a
//      template parameter) -> the minimizer value of the thread's 16 k-mers;
//   4. a record starts where the minimizer changes, at the first k-mer of the block and after a window
//      without a k-mer; its hash prefix picks the bucket; ONE shared atomic per record gives its rank;
//   5. block scan, one global atomicAdd per non-empty bucket reserves the run; records leave in bucket
//      order, consecutive threads writing consecutive 16-byte records of a run.
// The same sort serves the second level (group slabs -> tiles of the group) and the multi-GPU first
// level (runs written straight into the owner GPU's slabs over NVLink peer memory).
#include "rg_extract.cuh"
#include "rg_sk.cuh"

#ifndef SK_CTAS
#define SK_CTAS 1                                   // CTAs per SM
#endif
constexpr int SK_THREADS = 1024 / SK_CTAS;
constexpr int SK_R = 16;                            // window starts per thread: one 16-base word, every shift is static
constexpr uint32_t SK_POS = SK_THREADS * SK_R;      // window starts per super-chunk
constexpr uint32_t SK_RMAX = SK_POS / 4;            // records counting-sorted at once (a denser super-chunk takes several rounds)
constexpr uint32_t SK_RBITS = SK_CTAS == 1 ? 12 : 11;
static_assert((1u << SK_RBITS) == SK_RMAX, "record index bits");
constexpr uint32_t SK_YS = SK_THREADS + 2;          // row stride of the transposed m-mer hash array (+2 virtual threads of tail)
constexpr int SK_BUCKETS_PER_THREAD = 8;            // the scan handles up to 8 * SK_THREADS buckets (shared memory caps it earlier)
// 227 KiB opt-in limit per CTA (static + dynamic); 228 KiB per SM with 1 KiB reserved per resident CTA
constexpr size_t SK_SMEM_LIMIT = SK_CTAS == 1 ? 232448 : (233472 / SK_CTAS - 1024);

// Staging of one super-chunk (SK_POS window starts + 48 bases of tail) as four bit streams:
//   fwd : 16 bases per word, BIG-endian (base i of the word at bits 31-2i..30-2i), A0 C1 G2 T3
//   rcw : 16 bases per word, LITTLE-endian (base i at bits 2i..2i+1), complemented codes
//   bad : 1 bit per base, set for every byte outside ACGTacgt (window restart, mer_iterator.hpp:72-79)
//   brk : 1 bit per base, set where a new read starts (offset / fixed-stride inputs)
struct PartStage {
  static constexpr int CODE_WORDS = SK_POS / 16 + 3;
  static constexpr int BIT_WORDS = SK_POS / 32 + 2;
  uint32_t fwd[CODE_WORDS + 1];
  uint32_t rcw[CODE_WORDS + 1];
  uint32_t bad[BIT_WORDS + 2];
  uint32_t brk[BIT_WORDS + 2];
};

constexpr size_t sk_static_smem() { return sizeof(PartStage) + 80 * sizeof(uint32_t) + 3 * (SK_THREADS + 2) * 2 + 512; }
// y/kmin rows, records, (bucket, rank) and sorted entries; run length -> slab delta and local offset per
// bucket; overflow bitmap; level 2: chunk starts of the input slabs
static size_t sk_dynamic_smem(uint32_t n_buckets, uint32_t n_groups = 0) {
  return (size_t)SK_R * SK_YS * 4 + (size_t)SK_RMAX * (16 + 4 + 4) + (size_t)n_buckets * 8 +
         (size_t)((n_buckets + 31) / 32) * 4 + (size_t)(n_groups + 1) * 4 + 64;
}

constexpr int PART_MAX_OWNERS = 8;

// What one launch sorts into. A "bucket" is a hash-prefix GROUP (level 1: from the ASCII stream) or a
// TILE inside one group (level 2: from the group slabs the level-1 pass filled).
struct PartGeom {
  uint32_t n_buckets_global;   // level 1: groups over all shards (the multiplier of the hash prefix)
  uint32_t bucket_first;       // level 1: first bucket of this shard
  uint32_t n_buckets;          // buckets sorted per super-chunk: level 1 local groups (MULTI: all groups), level 2 group_tiles
  RgGeo geo;                   // level 2: locates the tile inside its group
  uint4* out_keys;             // bucket slabs [n_buckets (x n_groups at level 2)][out_cap] records
  uint32_t* out_cursor;        // fill of every slab
  uint32_t out_cap;
  // level 2 input: the group slabs
  const uint4* in_keys;
  const uint32_t* in_cursor;
  uint32_t in_cap, n_groups;     // n_groups input slabs, in_per_group of them per group (1, or the GPU count)
  uint32_t in_per_group;
  TableView tv;                  // records that do not fit their slab are inserted straight into the table
  int canonical;
  int sharded;                   // the table holds one hash-prefix shard: records of other shards are counted as foreign
  // MULTI: the slabs of bucket b live on GPU b / buckets_per_owner (peer memory), same layout everywhere:
  // [buckets_per_owner][n_sources][out_cap]; this GPU writes sub-slab my_source and counts in a LOCAL cursor
  int multi;
  uint32_t buckets_per_owner, n_sources, my_source;
  uint64_t spill_cap;
  uint4* owner_keys[PART_MAX_OWNERS];
  uint64_t* owner_status[PART_MAX_OWNERS];
  uint4* owner_spill[PART_MAX_OWNERS];
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

// 16 ASCII bytes -> big-endian forward codes, little-endian complemented codes, 16 bad flags.
// The per-byte fields are gathered with one multiply each (no carries: every partial product lands
// on its own bits), which keeps this off the integer ALU pipe.
__device__ __forceinline__ void convert16_both(const uint4 v, uint32_t& fwd_be, uint32_t& rc_le, uint32_t& bad) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  uint32_t le = 0;
  bad = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t x = (w[i] >> 1) & 0x03030303u;             // A0 C1 G3 T2 per byte
    const uint32_t c4 = x ^ ((x >> 1) & 0x01010101u);         // A0 C1 G2 T3
    le |= ((c4 * 0x01041040u) >> 24) << (8 * i);              // char j (byte j) -> bits 2j..2j+1
    const uint32_t u = w[i] & 0xDFDFDFDFu;                    // upper-case
    const uint32_t ok = __vcmpeq4(u, 0x41414141u) | __vcmpeq4(u, 0x43434343u) | __vcmpeq4(u, 0x47474747u) |
                        __vcmpeq4(u, 0x54545454u);            // 0xFF per valid byte
    const uint32_t nb = ~ok & 0x01010101u;                    // bit 0 of byte j set if char j is bad
    bad |= ((nb * 0x10204080u) >> 28) << (4 * i);             // bit j = char j
  }
  rc_le = ~le;
  const uint32_t r = __brev(le);                              // reverses bits: 2-bit groups end up bit-swapped
  fwd_be = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
}

__device__ __forceinline__ uint4 load16_stream(const ExtractParams& ep, uint64_t p) {
  uint4 v = make_uint4(0x0a0a0a0au, 0x0a0a0a0au, 0x0a0a0a0au, 0x0a0a0a0au);
  if (p + 16 <= ep.n_bytes) {
    v = __ldg(reinterpret_cast<const uint4*>(ep.stream + p));
  } else if (p < ep.n_bytes) {
    unsigned char tmp[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) tmp[j] = (p + j < ep.n_bytes) ? (unsigned char)ep.stream[p + j] : (unsigned char)'\n';
    v = *reinterpret_cast<uint4*>(tmp);
  }
  return v;
}

// read-start flags of the 32 bases starting at stream position p (p multiple of 32);
// r_p = p % read_stride (tracked incrementally by the caller: no 64-bit division per word)
__device__ __forceinline__ uint32_t brk_word(const ExtractParams& ep, uint64_t p, uint32_t r_p) {
  uint32_t w = 0;
  if (ep.brk != nullptr) {
    if (p < ep.n_bytes) w = __ldg(ep.brk + (p >> 5));
  } else if (ep.read_stride != 0) {
    const uint32_t first = r_p == 0 ? 0 : ep.read_stride - r_p;  // offset of the first read start >= p
    for (uint32_t j = first; j < 32; j += ep.read_stride) w |= 1u << j;
  }
  return w;
}

// What one thread prefetches for the next super-chunk: its 16 bases, (threads 0..2) one tail word,
// and (threads < BIT_WORDS) one word of read-start flags.
struct PartPrefetch {
  uint4 body, tail;
  uint32_t brk;
};

// r0 = s0 % read_stride (read_stride < 2^31, checked on the host)
__device__ __forceinline__ void part_prefetch(PartPrefetch& pf, const ExtractParams& ep, uint64_t s0, uint32_t r0) {
  if (s0 + SK_POS + 64 <= ep.n_bytes) {           // CTA-uniform: no per-load bounds checks
    const uint4* src = reinterpret_cast<const uint4*>(ep.stream + s0);
    pf.body = __ldg(src + threadIdx.x);
    if (threadIdx.x < 3) pf.tail = __ldg(src + SK_THREADS + threadIdx.x);
  } else {
    pf.body = load16_stream(ep, s0 + (uint64_t)threadIdx.x * 16);
    if (threadIdx.x < 3) pf.tail = load16_stream(ep, s0 + (uint64_t)(SK_THREADS + threadIdx.x) * 16);
  }
  if (threadIdx.x < PartStage::BIT_WORDS) {
    uint32_t r = 0;
    if (ep.read_stride != 0) r = (r0 + threadIdx.x * 32u) % ep.read_stride;
    pf.brk = brk_word(ep, s0 + (uint64_t)threadIdx.x * 32, r);
  }
}

__device__ __forceinline__ void part_stage_store(PartStage& sm, const PartPrefetch& pf) {
  uint16_t* bad16 = reinterpret_cast<uint16_t*>(sm.bad);
  uint32_t f, r, b;
  convert16_both(pf.body, f, r, b);
  sm.fwd[threadIdx.x] = f;
  sm.rcw[threadIdx.x] = r;
  bad16[threadIdx.x] = (uint16_t)b;
  if (threadIdx.x < 3) {
    convert16_both(pf.tail, f, r, b);
    sm.fwd[SK_THREADS + threadIdx.x] = f;
    sm.rcw[SK_THREADS + threadIdx.x] = r;
    bad16[SK_THREADS + threadIdx.x] = (uint16_t)b;
  }
  if (threadIdx.x < PartStage::BIT_WORDS) sm.brk[threadIdx.x] = pf.brk;
  if (threadIdx.x == 3) {
    sm.fwd[PartStage::CODE_WORDS] = 0;
    sm.rcw[PartStage::CODE_WORDS] = 0;
    bad16[SK_THREADS + 3] = 0xFFFFu;
    sm.bad[PartStage::BIT_WORDS] = 0xFFFFFFFFu;
    sm.bad[PartStage::BIT_WORDS + 1] = 0xFFFFFFFFu;
    sm.brk[PartStage::BIT_WORDS] = 0;
    sm.brk[PartStage::BIT_WORDS + 1] = 0;
  }
}

// order hash y of the 16 m-mers that start in the 16-base block `t` of the staged super-chunk, stored to
// column t of the transposed array (row w = position inside the block)
__device__ __forceinline__ void mmer_hashes(const PartStage& sm, uint32_t t, uint32_t sh, uint32_t cm, uint32_t cm2,
                                            uint32_t* Yt, uint32_t (&y)[SK_R]) {
  const uint32_t W0 = sm.fwd[t], W1 = sm.fwd[t + 1];
  const uint32_t V0 = sm.rcw[t], V1 = sm.rcw[t + 1];
  // (W0:W1) >> sh puts the m-mer of base 0 in the low 2m bits of the high word
  const uint32_t X0 = W0 >> sh, X1 = __funnelshift_r(W1, W0, sh);
#pragma unroll
  for (int w = 0; w < SK_R; ++w) {
    const uint32_t f = w ? __funnelshift_l(X1, X0, 2 * w) : X0;      // forward m-mer in the low 2m bits (garbage above)
    const uint32_t r = w ? __funnelshift_r(V0, V1, 2 * w) : V0;      // its reverse complement, same alignment
    y[w] = ((f * cm) ^ (r * cm)) + (f & r) * cm2;                    // garbage above 2m bits is multiplied away
    Yt[w * SK_YS + t] = y[w];
  }
}

// minimum over the W values x[i .. i+W-1] for i = 0..15 (x[0 .. 16+W-2] valid), log-step doubling
template <int W>
__device__ __forceinline__ void sliding_min(uint32_t (&x)[32], uint32_t (&out)[SK_R]) {
  constexpr int P = W >= 16 ? 16 : W >= 8 ? 8 : W >= 4 ? 4 : W >= 2 ? 2 : 1;
#pragma unroll
  for (int step = 1; step < P; step *= 2) {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i + 2 * step - 1 <= SK_R + W - 2) x[i] = min(x[i], x[i + step]);
  }
#pragma unroll
  for (int i = 0; i < SK_R; ++i) out[i] = min(x[i], x[i + W - P]);
}

// the k-mers of a record that found no room in its slab: straight into the table (global atomics)
__device__ __noinline__ void sk_insert_direct(const TableView& tv, const uint4 rec, int canonical) {
  const uint32_t mh = rec.w & ~15u;
  const uint32_t tile = rg_tile_of_mh(mh, tv.geo) - tv.tile_first;
  if (tile >= tv.n_tiles) {
    atomicAdd((unsigned long long*)&tv.status[RG_ST_FOREIGN], 1ull);
    return;
  }
  tv.status[RG_ST_DIRECT] = 1;
  const uint32_t len = (rec.w & 15u) + 1;
  for (uint32_t o = 0; o < len; ++o) rg_table_add_at(tv, tile, sk_key(rec, o, tv.geo.k, canonical), 1u);
}

// SRC: 0 = windows of the ASCII stream (W = k-m+1: 16, 14, 12, 10, 8 specialised; 0 = any window, slower),
//      2 = the group slabs of a level-1 pass (work items = chunks of SK_RMAX records over all slabs).
template <int SRC, int W>
__global__ void __launch_bounds__(SK_THREADS, SK_CTAS)
sk_partition_kernel(ExtractParams ep, PartGeom g) {
  extern __shared__ __align__(16) unsigned char dyn[];
  __shared__ PartStage sm;
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t scan_total;
  __shared__ uint32_t any_overflow;
  __shared__ uint32_t rec_total;
  __shared__ uint16_t blk_valid[SK_THREADS + 2], blk_bnd[SK_THREADS + 2], blk_cut[SK_THREADS + 2];
  // MULTI: the owners' base pointers, indexed per bucket / per record (a dynamically indexed kernel
  // parameter array would be copied to local memory)
  __shared__ uint4* s_owner_keys[PART_MAX_OWNERS];
  if (g.multi && threadIdx.x < PART_MAX_OWNERS) s_owner_keys[threadIdx.x] = g.owner_keys[threadIdx.x];
  uint32_t* Yt = reinterpret_cast<uint32_t*>(dyn);                                   // [SK_R][SK_YS] y, then kmin
  uint4* recs = reinterpret_cast<uint4*>(dyn + (size_t)SK_R * SK_YS * 4);            // [SK_RMAX]
  uint32_t* metaS = reinterpret_cast<uint32_t*>(recs + SK_RMAX);                     // [SK_RMAX] (bucket << RBITS) | rank
  uint32_t* sorted = metaS + SK_RMAX;                                                // [SK_RMAX] (bucket << RBITS) | record index
  uint32_t* hist = sorted + SK_RMAX;      // [n_buckets] run length, then (step 5b) index of the run in its slab array - loff
  uint32_t* loff = hist + g.n_buckets;    // [n_buckets] local offset of the run
  uint32_t* ovf_bits = loff + g.n_buckets;  // [(n_buckets+31)/32] buckets whose slab overflowed in this round
  const uint32_t per_thread = (g.n_buckets + SK_THREADS - 1) / SK_THREADS;
  const uint32_t ovf_words = (g.n_buckets + 31) >> 5;
  const uint32_t t = threadIdx.x, lane = threadIdx.x & 31;

  PartPrefetch pf;
  uint32_t r_next = 0, r_step = 0;   // (start of the next super-chunk to prefetch) % read_stride, and its increment
  if (SRC == 0 && ep.read_stride != 0) {
    r_next = (uint32_t)(((uint64_t)blockIdx.x * SK_POS) % ep.read_stride);
    r_step = (uint32_t)(((uint64_t)gridDim.x * SK_POS) % ep.read_stride);
  }
  auto advance_r = [&]() {
    r_next += r_step;
    if (r_next >= ep.read_stride) r_next -= ep.read_stride;
  };
  if (SRC == 0 && (uint64_t)blockIdx.x * SK_POS < ep.n_bytes) {
    part_prefetch(pf, ep, (uint64_t)blockIdx.x * SK_POS, r_next);
    advance_r();
    part_stage_store(sm, pf);
  }
  if (t == 0) any_overflow = 0;
  for (uint32_t i = t; i < ovf_words; i += SK_THREADS) ovf_bits[i] = 0;
  const uint32_t k = g.geo.k, m = g.geo.m;
  const uint32_t sh = 32u - 2u * m, cm = RG_MM_C << sh, cm2 = RG_MM_C2 << sh;

  // Work items = super-chunks. Level 2: chunks of SK_RMAX records over ALL input slabs, numbered slab after
  // slab (gstart[] = exclusive prefix of the slabs' chunk counts, built here from the slab cursors), so
  // that the whole grid is busy whatever the number of groups.
  uint32_t* gstart = ovf_bits + ovf_words;        // [n_groups + 1], level 2 only
  uint64_t total_items;
  if (SRC == 2) {
    __syncthreads();
    uint32_t run_g = 0;
    for (uint32_t base = 0; base < g.n_groups; base += SK_THREADS) {
      const uint32_t gi = base + t;
      const uint32_t c = gi < g.n_groups ? (min(g.in_cursor[gi], g.in_cap) + SK_RMAX - 1) / SK_RMAX : 0u;
      const uint32_t ex = block_exclusive_scan(c, warp_sums, &scan_total);
      if (gi < g.n_groups) gstart[gi] = run_g + ex;
      __syncthreads();
      run_g += scan_total;
      __syncthreads();
    }
    if (t == 0) gstart[g.n_groups] = run_g;
    __syncthreads();
    total_items = run_g;
  } else {
    total_items = (ep.n_bytes + SK_POS - 1) / SK_POS;
  }
  uint32_t og = 0;                                 // level 2: input slab of the current item
  for (uint64_t item = blockIdx.x; item < total_items; item += gridDim.x) {
    if (SRC == 2)
      while (item >= gstart[og + 1]) ++og;
    // first output slab of this group (peer-to-peer: the input slabs are (group, source GPU) pairs)
    const uint32_t slab0 = SRC == 2 ? (og / g.in_per_group) * g.n_buckets : 0u;
    const uint64_t sc = SRC == 2 ? item - gstart[og] : item;
#pragma unroll 1
    for (uint32_t i = t; i < g.n_buckets; i += SK_THREADS) hist[i] = 0;
    if (t == 0) rec_total = 0;
    const bool more = item + gridDim.x < total_items;
    if (SRC == 0 && more) {
      part_prefetch(pf, ep, (sc + gridDim.x) * SK_POS, r_next);
      advance_r();
    }
    __syncthreads();

    // ---- records of this thread: start mask, first record index, the three forward words
    uint32_t starts = 0, valid16 = 0, my_first = 0, own_cuts = 0, next_cuts = 0;
    bool stage_pending = SRC == 0 && more;            // the next super-chunk's bases wait in registers (pf)
    uint32_t n_in_chunk = 0;                          // level 2: records of this chunk
    if (SRC == 0) {
      // ---- 2. m-mer order hashes
      uint32_t x[32];
      {
        uint32_t y[SK_R];
        mmer_hashes(sm, t, sh, cm, cm2, Yt, y);
#pragma unroll
        for (int i = 0; i < SK_R; ++i) x[i] = y[i];
        if (t < 2) {                                  // the two blocks of tail after the super-chunk
          uint32_t yt[SK_R];
          mmer_hashes(sm, SK_THREADS + t, sh, cm, cm2, Yt, yt);
        }
      }
      // validity of each of the 16 windows: k good bases and no read start strictly inside
      {
        const int p0 = t * SK_R;
        const uint64_t good = ~bits_at(sm.bad, p0);
        const uint64_t nobrk = ~bits_at(sm.brk, p0 + 1);          // bit j: no read starts at base p0+j+1
        // window i holds a k-mer iff bases i..i+k-1 are good and no read starts at bases i+1..i+k-1: one run of
        // k-1 over (good & nobrk) and the last base's good bit
        uint64_t valid = good;
        if (k > 1) valid = and_run_uniform(good & nobrk, (int)k - 1) & (good >> (k - 1));
        valid16 = (uint32_t)valid & 0xFFFFu;
        // first super-chunk of a split push: the windows that start inside the halo belong to the previous piece
        const uint32_t skip_local = sc == 0 ? ep.skip : 0u;
        if ((uint32_t)p0 < skip_local) valid16 &= ~((skip_local - p0 >= 32u) ? 0xFFFFFFFFu : ((1u << (skip_local - p0)) - 1u));
      }
      __syncthreads();
      // ---- 3. minimizer value of every k-mer: sliding minimum over W m-mers
      uint32_t kmin[SK_R];
      if (W > 0) {
#pragma unroll
        for (int j = 0; j + 2 <= W; ++j) x[SK_R + j] = Yt[j * SK_YS + t + 1];
        sliding_min<W == 0 ? 1 : W>(x, kmin);
      } else {
        const uint32_t w_run = k - m + 1;
#pragma unroll
        for (int i = 0; i < SK_R; ++i) {
          uint32_t mn = 0xFFFFFFFFu;
          for (uint32_t j = 0; j < w_run; ++j) {
            const uint32_t p = t * SK_R + i + j;
            mn = min(mn, Yt[(p & 15u) * SK_YS + (p >> 4)]);
          }
          kmin[i] = mn;
        }
      }
      __syncthreads();                                 // every neighbour has read y: the rows are reused for kmin
      // ---- 4. record starts. A record = a maximal run of consecutive k-mers with the same minimizer value, cut only
      // where the read says so (a window without a k-mer), every 16 k-mers of an over-long run, and at the edges of
      // the super-chunk: NOT at the 16-base blocks of the threads, so that every read covering a locus produces the
      // same records (dedup_kernel collapses them). Three exchanges through shared memory:
      //   (a) kmin (own column of Yt) + the validity bits of the block        -> bit 0 of the boundary mask
      //   (b) the boundary masks (position without a k-mer, or minimizer != previous position's)
      //       -> forced cuts: a position 16, 32, .. k-mers after the last boundary (looked up in the two blocks before)
      //   (c) boundary | forced masks                                           -> the length of each record
#pragma unroll
      for (int i = 0; i < SK_R; ++i) Yt[i * SK_YS + t] = kmin[i];   // own column (its y values have been consumed)
      blk_valid[t] = (uint16_t)valid16;
      uint32_t bnd = 0;
#pragma unroll
      for (int i = 1; i < SK_R; ++i) bnd |= (kmin[i] != kmin[i - 1] ? 1u : 0u) << i;
      bnd |= ~valid16 | ~(valid16 << 1);                   // no k-mer here, or none just before
      bnd &= 0xFFFEu;
      __syncthreads();
      {
        const bool prev_ok = t > 0 && ((blk_valid[t - 1] >> 15) & 1u);
        const uint32_t prev_k = t > 0 ? Yt[15 * SK_YS + t - 1] : 0u;
        if (!(valid16 & 1u) || !prev_ok || prev_k != kmin[0]) bnd |= 1u;      // (t == 0: the super-chunk's edge)
      }
      blk_bnd[t] = (uint16_t)bnd;
      __syncthreads();
      uint32_t cuts = bnd;
      {
        // positions whose 16 predecessors-and-self hold no boundary: an over-long run (repeated minimizer value)
        const uint64_t m = (uint64_t)(t > 1 ? blk_bnd[t - 2] : 0xFFFFu) | ((uint64_t)(t > 0 ? blk_bnd[t - 1] : 0xFFFFu) << 16) |
                           ((uint64_t)bnd << 32);
        uint64_t o = m | (m << 1);
        o |= o << 2;
        o |= o << 4;
        o |= o << 8;
        uint32_t cand = (uint32_t)(~o >> 32) & 0xFFFFu;
        while (cand) {                                      // rare
          const int i = __ffs(cand) - 1;
          cand &= cand - 1;
          const uint64_t below = m & ((2ull << (32 + i)) - 1ull);
          // distance to the last boundary seen (none within the two blocks before: cut at the block's first position)
          const bool force = below ? ((32 + i - (63 - __clzll((long long)below))) & 15) == 0 : i == 0;
          if (force) cuts |= 1u << i;
        }
      }
      blk_cut[t] = (uint16_t)cuts;
      if (t == 0) blk_cut[SK_THREADS] = 0xFFFFu;           // nothing continues past the super-chunk
      __syncthreads();
      starts = cuts & valid16;
      next_cuts = blk_cut[t + 1];
      own_cuts = cuts;
      const uint32_t nrec = __popc(starts);
      uint32_t incl = nrec;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += n;
      }
      uint32_t wbase = 0;
      if (lane == 31) wbase = atomicAdd(&rec_total, incl);
      wbase = __shfl_sync(0xFFFFFFFFu, wbase, 31);
      my_first = wbase + incl - nrec;
    } else {
      const uint64_t n_slab = min(g.in_cursor[og], g.in_cap);
      n_in_chunk = (uint32_t)min((uint64_t)SK_RMAX, n_slab - sc * SK_RMAX);
      if (t == 0) rec_total = n_in_chunk;
    }

    // one round sorts the records [round0, round0 + SK_RMAX) of the super-chunk (almost always the only round)
    for (uint32_t round0 = 0;; round0 += SK_RMAX) {
      uint32_t foreign = 0;
      if (SRC == 0) {
        // (a) every thread lists its records: (block, first window, length) descriptors at their record index.
        // Divergent (0..16 records per thread, ~2.7 on average), so it only does the cheap part.
        uint32_t st = starts, idx = my_first;
        while (st) {
          const uint32_t s = __ffs(st) - 1;
          st &= st - 1;
          const uint32_t rel = idx - round0;
          ++idx;
          if (rel >= SK_RMAX) continue;
          // the record runs to the next cut (this block's or the next one's), at most 16 k-mers
          const uint32_t nb = (((own_cuts & 0xFFFFu) | (next_cuts << 16)) >> (s + 1)) | (1u << 15);
          sorted[rel] = (t << 8) | (s << 4) | (__ffs(nb) - 1u);
        }
        __syncthreads();
        // (b) the records themselves, one per thread-iteration whatever block they come from (balanced)
        const uint32_t n_here = min(SK_RMAX, rec_total - round0);
        for (uint32_t i = t; i < n_here; i += SK_THREADS) {
          const uint32_t d = sorted[i];
          const uint32_t tt = d >> 8, s = (d >> 4) & 15u;
          const uint32_t mh = rg_mix_min(Yt[s * SK_YS + tt]);
          uint32_t b = rg_group_of_mh(mh, g.n_buckets_global) - g.bucket_first;
          if (b >= g.n_buckets) {                      // another shard's record (sharded single-GPU push)
            ++foreign;
            metaS[i] = 0xFFFFFFFFu;
            continue;
          }
          const uint32_t rank = atomicAdd(&hist[b], 1u);
          const uint32_t W0 = sm.fwd[tt], W1 = sm.fwd[tt + 1], W2 = sm.fwd[tt + 2], W3 = sm.fwd[tt + 3];
          recs[i] = make_uint4(__funnelshift_l(W1, W0, 2 * s), __funnelshift_l(W2, W1, 2 * s), __funnelshift_l(W3, W2, 2 * s),
                               mh | (d & 15u));
          metaS[i] = (b << SK_RBITS) | rank;
        }
      } else {
#pragma unroll
        for (uint32_t w = 0; w < SK_RMAX / SK_THREADS; ++w) {
          const uint32_t i = w * SK_THREADS + t;
          if (i < n_in_chunk) {
            const uint4 rec = __ldg(g.in_keys + (uint64_t)og * g.in_cap + sc * SK_RMAX + i);
            const uint32_t b = rg_tile_in_group_of_mh(rec.w & ~15u, g.geo);
            const uint32_t rank = atomicAdd(&hist[b], 1u);
            recs[i] = rec;
            metaS[i] = (b << SK_RBITS) | rank;
          }
        }
      }
      if (foreign) atomicAdd((unsigned long long*)&g.tv.status[RG_ST_FOREIGN], (unsigned long long)foreign);
      __syncthreads();
      const uint32_t n_all = rec_total;
      // the stage buffers are free once the last round's records are built: store the next super-chunk
      if (stage_pending && round0 + SK_RMAX >= n_all) {
        part_stage_store(sm, pf);
        stage_pending = false;
      }
      const uint32_t n_round = min(SK_RMAX, n_all - round0);
      // ---- 5. local offsets (exclusive scan of the run lengths); the global run reservations (one
      // atomicAdd per non-empty bucket) are only ISSUED here: their results are consumed after the
      // scatter below, which hides the round trip to L2
      const uint32_t t0 = t * per_thread;
      uint32_t cnt[SK_BUCKETS_PER_THREAD], gb[SK_BUCKETS_PER_THREAD];
      uint32_t sum = 0;
#pragma unroll
      for (int j = 0; j < SK_BUCKETS_PER_THREAD; ++j) {
        cnt[j] = 0;
        if ((uint32_t)j < per_thread) {                  // CTA-uniform
          cnt[j] = t0 + j < g.n_buckets ? hist[t0 + j] : 0u;
          sum += cnt[j];
        }
      }
      uint32_t run = block_exclusive_scan(sum, warp_sums, &scan_total);
      const uint32_t run0 = run;
#pragma unroll
      for (int j = 0; j < SK_BUCKETS_PER_THREAD; ++j) {
        gb[j] = 0;
        if ((uint32_t)j < per_thread && cnt[j]) {
          const uint32_t b = t0 + j;
          loff[b] = run;
          // MULTI: every source GPU fills its OWN sub-slab of the owner's group slab: the cursor is local
          // (no atomics over NVLink), only the records travel
          gb[j] = atomicAdd(g.out_cursor + (g.multi ? b : slab0 + b), cnt[j]);
          run += cnt[j];
        }
      }
      __syncthreads();
      // ---- (bucket, record index) entries in bucket order
      for (uint32_t i = t; i < n_round; i += SK_THREADS) {
        const uint32_t mt = metaS[i];
        if (mt != 0xFFFFFFFFu) {
          const uint32_t b = mt >> SK_RBITS;
          sorted[loff[b] + (mt & (SK_RMAX - 1))] = (b << SK_RBITS) | i;
        }
      }
      // ---- 5b. the reservations have landed: sorted index -> slab-array index (into hist[])
      run = run0;
#pragma unroll
      for (int j = 0; j < SK_BUCKETS_PER_THREAD; ++j) {
        if ((uint32_t)j < per_thread && cnt[j]) {
          const uint32_t b = t0 + j;
          const uint32_t o = g.multi ? b / g.buckets_per_owner : 0u;
          // slab index inside its owner's array (MULTI: [group of the owner][source GPU])
          const uint32_t slab = g.multi ? (b - o * g.buckets_per_owner) * g.n_sources + g.my_source : slab0 + b;
          hist[b] = slab * g.out_cap + gb[j] - run;                            // index of the run in the slab array - loff
          if (gb[j] + cnt[j] > g.out_cap) {
            atomicOr(&ovf_bits[b >> 5], 1u << (b & 31));
            any_overflow = 1;
          }
          run += cnt[j];
        }
      }
      __syncthreads();
      const bool ovf = any_overflow != 0;
      // ---- copy-out in bucket order
      const uint32_t total = scan_total;
      if (!ovf) {
#pragma unroll 4
        for (uint32_t i = t; i < total; i += SK_THREADS) {
          const uint32_t e = sorted[i];
          const uint32_t b = e >> SK_RBITS;
          uint4* base = g.multi ? s_owner_keys[b / g.buckets_per_owner] : g.out_keys;
          base[hist[b] + i] = recs[e & (SK_RMAX - 1)];
        }
      } else {
        // some slab is full (a hot minimizer, or a push denser in records than the slabs were sized for):
        // what does not fit goes straight into the table (MULTI: to the owner's exported spill list)
        for (uint32_t i = t; i < total; i += SK_THREADS) {
          const uint32_t e = sorted[i];
          const uint32_t b = e >> SK_RBITS;
          const uint4 rec = recs[e & (SK_RMAX - 1)];
          const uint32_t o = g.multi ? b / g.buckets_per_owner : 0u;
          const uint32_t slab = g.multi ? (b - o * g.buckets_per_owner) * g.n_sources + g.my_source : slab0 + b;
          const uint32_t at = hist[b] + i;
          const bool fits = !((ovf_bits[b >> 5] >> (b & 31)) & 1u) || at < (slab + 1) * g.out_cap;
          if (fits) {
            uint4* base = g.multi ? s_owner_keys[o] : g.out_keys;
            base[at] = rec;
          } else if (g.multi) {
            const uint64_t sp = atomicAdd_system((unsigned long long*)&g.owner_status[o][0], 1ull);
            if (sp < g.spill_cap) g.owner_spill[o][sp] = rec;
          } else {
            sk_insert_direct(g.tv, rec, g.canonical);
          }
        }
        __syncthreads();
        if (t == 0) any_overflow = 0;
        for (uint32_t i = t; i < ovf_words; i += SK_THREADS) ovf_bits[i] = 0;
      }
      __syncthreads();
      if (round0 + SK_RMAX >= n_all) break;
      for (uint32_t i = t; i < g.n_buckets; i += SK_THREADS) hist[i] = 0;
      __syncthreads();
    }
  }
}

// ---------------------------------------------------------------------------
// Level 2 with few buckets (group slabs -> the <= 128 tiles of each group): a multi-split without the
// shared-memory staging of the general kernel. With so few buckets one shared histogram would serialise
// (hundreds of atomics per address per chunk: measured 0.72 ms per 25 M records); here every WARP owns a
// histogram row, so an address sees at most the 32 lanes of one instruction. A CTA (256 threads, four per
// SM) takes chunks of 1024 records (4 per thread, kept in registers); rank inside (warp, bucket) from the atomic, exclusive scan in
// (bucket, warp) order, one global atomicAdd per non-empty bucket reserves the run, then every thread
// writes its records straight to their slots (the ~8 records a warp holds for one bucket are contiguous).
// ---------------------------------------------------------------------------
constexpr int RG2_THREADS = 256;                         // small CTAs, several per SM: their load / atomic / store latencies overlap
constexpr int RG2_CTAS = 4;
constexpr int RG2_WARPS = RG2_THREADS / 32;
constexpr int RG2_RPT = 4;                               // records per thread
constexpr uint32_t RG2_CHUNK = RG2_THREADS * RG2_RPT;
constexpr uint32_t RG2_MAXB = 128;                       // buckets (tiles per group) it handles
constexpr int RG2_EPT = RG2_WARPS * RG2_MAXB / RG2_THREADS;   // (bucket, warp) cells per thread in the scan

__global__ void __launch_bounds__(RG2_THREADS, RG2_CTAS)
regroup_kernel(PartGeom g) {
  extern __shared__ __align__(16) unsigned char dyn[];
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t scan_total;
  __shared__ uint32_t boff[RG2_MAXB + 1];      // first sorted position of every bucket in the chunk
  __shared__ uint32_t gdelta[RG2_MAXB];        // slab-array index of the bucket's run - boff (0xFFFFFFFF: checked per record)
  __shared__ uint32_t glimit[RG2_MAXB];        // end of the bucket's slab in the slab array
  const uint32_t B = g.n_buckets;
  uint32_t* whist = reinterpret_cast<uint32_t*>(dyn);          // [B][RG2_WARPS]: cell (b, w) at b * RG2_WARPS + w
  uint32_t* gstart = whist + RG2_WARPS * RG2_MAXB;             // [n_groups + 1]
  const uint32_t t = threadIdx.x, lane = t & 31, warp = t >> 5;
  // chunk starts of the input slabs (exclusive prefix of their chunk counts)
  uint32_t run_g = 0;
  for (uint32_t base = 0; base < g.n_groups; base += RG2_THREADS) {
    const uint32_t gi = base + t;
    const uint32_t c = gi < g.n_groups ? (min(g.in_cursor[gi], g.in_cap) + RG2_CHUNK - 1) / RG2_CHUNK : 0u;
    const uint32_t ex = block_exclusive_scan(c, warp_sums, &scan_total);
    if (gi < g.n_groups) gstart[gi] = run_g + ex;
    __syncthreads();
    run_g += scan_total;
    __syncthreads();
  }
  if (t == 0) gstart[g.n_groups] = run_g;
  __syncthreads();
  const uint32_t total_items = run_g;
  uint32_t og = 0;
  for (uint32_t item = blockIdx.x; item < total_items; item += gridDim.x) {
    while (item >= gstart[og + 1]) ++og;
    const uint32_t slab0 = (og / g.in_per_group) * B;
    const uint32_t sc = item - gstart[og];
    const uint32_t n_slab = min(g.in_cursor[og], g.in_cap);
    const uint32_t n_chunk = min(RG2_CHUNK, n_slab - sc * RG2_CHUNK);
    for (uint32_t i = t; i < RG2_WARPS * B; i += RG2_THREADS) whist[i] = 0;
    __syncthreads();
    uint4 rec[RG2_RPT];
    uint32_t meta[RG2_RPT];                    // bucket << 16 | rank inside (warp, bucket)
    const uint4* in = g.in_keys + (uint64_t)og * g.in_cap + (uint64_t)sc * RG2_CHUNK;
#pragma unroll
    for (int w = 0; w < RG2_RPT; ++w) {
      const uint32_t i = w * RG2_THREADS + t;
      rec[w] = i < n_chunk ? __ldg(in + i) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int w = 0; w < RG2_RPT; ++w) {
      meta[w] = 0xFFFFFFFFu;
      if (w * RG2_THREADS + t < n_chunk) {
        const uint32_t b = rg_tile_in_group_of_mh(rec[w].w & ~15u, g.geo);
        meta[w] = (b << 16) | atomicAdd(&whist[b * RG2_WARPS + warp], 1u);
      }
    }
    __syncthreads();
    // exclusive scan of the cells in (bucket, warp) order; RG2_EPT consecutive cells per thread
    {
      uint32_t v[RG2_EPT], sum = 0;
#pragma unroll
      for (int e = 0; e < RG2_EPT; ++e) {
        const uint32_t c = t * RG2_EPT + e;
        v[e] = c < RG2_WARPS * B ? whist[c] : 0u;
        sum += v[e];
      }
      uint32_t ex = block_exclusive_scan(sum, warp_sums, &scan_total);
#pragma unroll
      for (int e = 0; e < RG2_EPT; ++e) {
        const uint32_t c = t * RG2_EPT + e;
        if (c < RG2_WARPS * B) {
          whist[c] = ex;
          if (c % RG2_WARPS == 0) boff[c / RG2_WARPS] = ex;
        }
        ex += v[e];
      }
      if (t == 0) boff[B] = scan_total;
    }
    __syncthreads();
    if (t < B) {
      const uint32_t cnt = boff[t + 1] - boff[t];
      const uint32_t slab = slab0 + t;
      uint32_t gb = 0;
      if (cnt) gb = atomicAdd(g.out_cursor + slab, cnt);
      gdelta[t] = slab * g.out_cap + gb - boff[t];
      glimit[t] = (slab + 1) * g.out_cap;
    }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < RG2_RPT; ++w) {
      if (meta[w] != 0xFFFFFFFFu) {
        const uint32_t b = meta[w] >> 16;
        const uint32_t at = gdelta[b] + whist[b * RG2_WARPS + warp] + (meta[w] & 0xFFFFu);
        if (at < glimit[b]) g.out_keys[at] = rec[w];
        else sk_insert_direct(g.tv, rec[w], g.canonical);      // the tile's bucket slab is full
      }
    }
    __syncthreads();
  }
}

// does the sort's per-bucket state fit shared memory for this bucket count?
static bool part_sorted_fits(uint32_t n_buckets, uint32_t n_groups = 0) {
  return n_buckets <= (uint32_t)SK_BUCKETS_PER_THREAD * SK_THREADS && n_buckets < (1u << (32 - SK_RBITS)) &&
         sk_dynamic_smem(n_buckets, n_groups) + sk_static_smem() <= SK_SMEM_LIMIT;
}
// the largest bucket count one sorted pass can take (level-1 fanout limit)
uint32_t rg_partition_max_buckets() {
  uint32_t n = SK_BUCKETS_PER_THREAD * SK_THREADS;
  while (n > 1 && !part_sorted_fits(n)) --n;
  return n;
}

// records one push of n_bytes stream bytes produces at most on average (sizes the slabs; what does not
// fit is inserted directly): one record per minimizer change (2 / (W + 1) of the windows), one per
// 16-base block, plus read starts.
double rg_partition_record_density(int k, int m) {
  const int w = k - m + 1;
  return 2.0 / (w + 1) + 1.0 / 16 + 0.03;
}

template <int SRC, int W>
static int launch_sk(rg_ctx* ctx, const char* name, const ExtractParams& ep, const PartGeom& g, uint64_t n_work,
                     uint32_t max_ctas) {
  const size_t smem = sk_dynamic_smem(g.n_buckets, g.n_groups);
  auto kern = sk_partition_kernel<SRC, W>;
  RG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  RG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  unsigned grid = rg_grid(ctx, SK_CTAS);
  if (max_ctas && max_ctas < grid) grid = max_ctas;
  if (n_work < grid) grid = (unsigned)n_work;
  rg_launch_scope _ls(ctx, name);
  kern<<<grid, SK_THREADS, smem, ctx->stream>>>(ep, g);
  return 0;
}

static int launch_sk_stream(rg_ctx* ctx, const char* name, const ExtractParams& ep, const PartGeom& g, uint32_t max_ctas) {
  const uint64_t n_super = (ep.n_bytes + SK_POS - 1) / SK_POS;
  const int w = (int)g.geo.k - (int)g.geo.m + 1;
  int rc;
  switch (w) {
    case 16: rc = launch_sk<0, 16>(ctx, name, ep, g, n_super, max_ctas); break;
    case 14: rc = launch_sk<0, 14>(ctx, name, ep, g, n_super, max_ctas); break;
    case 12: rc = launch_sk<0, 12>(ctx, name, ep, g, n_super, max_ctas); break;
    case 10: rc = launch_sk<0, 10>(ctx, name, ep, g, n_super, max_ctas); break;
    case 8: rc = launch_sk<0, 8>(ctx, name, ep, g, n_super, max_ctas); break;
    default: rc = launch_sk<0, 0>(ctx, name, ep, g, n_super, max_ctas); break;
  }
  if (rc) return rc;
  RG_CUDA_OK(cudaGetLastError());
  return 0;
}

static void fill_common(PartGeom& g, rg_table* t) {
  memset(&g, 0, sizeof(g));
  g.tv = rg_view(t);
  g.geo = g.tv.geo;
  g.canonical = t->canonical;
}

// level 1: stream -> the table's groups (group_tiles == 1: straight into the tile buckets)
int rg_launch_partition(rg_table* t, const ExtractParams& ep) {
  rg_ctx* ctx = t->ctx;
  if (ep.n_bytes == 0) return 0;
  PartGeom g;
  fill_common(g, t);
  const bool two_level = t->group_tiles > 1;
  g.n_buckets_global = g.geo.groups;
  g.n_buckets = t->n_tiles / t->group_tiles;
  g.bucket_first = t->tile_first / t->group_tiles;
  g.out_keys = two_level ? t->grp_keys : t->part_keys;
  g.out_cursor = two_level ? t->grp_cursor : t->part_cursor;
  g.out_cap = two_level ? t->grp_cap : t->bucket_cap;
  g.sharded = g.n_buckets_global != g.n_buckets;
  if (!part_sorted_fits(g.n_buckets) || (uint64_t)g.n_buckets * g.out_cap >= (1ull << 32)) {
    rg_set_error("partition: bucket geometry out of range");
    return RG_ERR_ARG;
  }
  return launch_sk_stream(ctx, "partition_kernel", ep, g, 0);
}

// level 2: the group slabs -> the tile buckets of every group
int rg_launch_partition_groups(rg_table* t) {
  rg_ctx* ctx = t->ctx;
  PartGeom g;
  fill_common(g, t);
  g.n_buckets = t->group_tiles;
  g.in_per_group = t->p2p_block ? t->shard_parts : 1;
  g.n_groups = t->n_tiles / t->group_tiles * g.in_per_group;
  g.in_keys = t->grp_keys;
  g.in_cursor = t->grp_cursor;
  g.in_cap = t->grp_cap;
  g.out_keys = t->part_keys;
  g.out_cursor = t->part_cursor;
  g.out_cap = t->bucket_cap;
  if (!part_sorted_fits(g.n_buckets, g.n_groups) || (uint64_t)t->n_tiles * t->bucket_cap >= (1ull << 32)) {
    rg_set_error("partition: group geometry out of range");
    return RG_ERR_ARG;
  }
  ExtractParams ep = {};
  // work items = record chunks over all groups (their number is only known on the device): a full grid
  if (g.n_buckets <= RG2_MAXB) {
    const size_t smem = (size_t)RG2_WARPS * RG2_MAXB * 4 + (size_t)(g.n_groups + 1) * 4;
    RG_CUDA_OK(cudaFuncSetAttribute(regroup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rg_launch_scope _ls(ctx, "partition_groups_kernel");
    regroup_kernel<<<rg_grid(ctx, RG2_CTAS), RG2_THREADS, smem, ctx->stream>>>(g);
  } else {
    int rc = launch_sk<2, 0>(ctx, "partition_groups_kernel", ep, g, 1ull << 32, 0);
    if (rc) return rc;
  }
  RG_CUDA_OK(cudaGetLastError());
  return 0;
}

// level 1 over ALL shards, straight into the owners' group slabs (peer memory)
int rg_launch_partition_multi(rg_table* t, const ExtractParams& ep) {
  rg_ctx* ctx = t->ctx;
  if (ep.n_bytes == 0) return 0;
  PartGeom g;
  fill_common(g, t);
  g.multi = 1;
  g.n_buckets_global = g.geo.groups;
  g.n_buckets = g.geo.groups;
  g.bucket_first = 0;
  g.buckets_per_owner = t->n_tiles / t->group_tiles;
  g.n_sources = t->shard_parts;
  g.my_source = t->shard_part;
  g.out_cap = t->grp_cap;
  g.out_cursor = t->p2p_src_cursor;
  g.spill_cap = t->p2p_spill_cap;
  for (uint32_t o = 0; o < t->shard_parts; ++o) {
    g.owner_keys[o] = t->peer_keys[o];
    g.owner_status[o] = t->peer_status[o];
    g.owner_spill[o] = t->peer_spill[o];
  }
  if (!part_sorted_fits(g.n_buckets) || (uint64_t)g.buckets_per_owner * g.n_sources * g.out_cap >= (1ull << 32)) {
    rg_set_error("partition: too many groups for the peer-to-peer partition");
    return RG_ERR_ARG;
  }
  return launch_sk_stream(ctx, "partition_p2p_kernel", ep, g, t->p2p_ctas);
}
