// rg_partition.cu — RG_MODE_TILED, step 1: radix partition of packed k-mers by hash
// prefix into one bucket per table tile (the reference has no counterpart: its
// large_hash_array.hpp:298-302 insert goes straight to one DRAM-resident table).
//
// What the B200 measurements said (profiles/microbench_r1.txt, profiles/README.md):
// shared-memory atomics are cheap (0.14 cyc/lane/SM), MATCH.ANY is not (2 cyc/lane),
// and per-key 8-byte scattered global stores are the wall (a write footprint of
// (#CTAs x #buckets) open lines overflows L2). So the CTA counting-sorts a
// SUPER-CHUNK of 16384 windows in shared memory and writes each bucket's keys as ONE
// contiguous run (see partition_sorted_kernel for the four steps). The same kernel body
// serves three sources (ASCII stream, key list, the group slabs of a first pass) and
// three destinations (tile buckets, group slabs, the group slabs of OTHER GPUs over
// NVLink peer memory): small tables take one pass, tables beyond ~2900 tiles two
// (stream -> hash-prefix groups -> tiles), multi-GPU counting uses the first pass to route.
#include "rg_extract.cuh"

#ifndef PART_CTAS
#define PART_CTAS 1                                 // CTAs per SM (2 half-size CTAs measured slower: 1.31 vs 0.84 ms, profiles/README.md)
#endif
constexpr int PART_THREADS = 1024 / PART_CTAS;
constexpr int PART_R = 16;                          // windows per thread: one 16-base word, so every shift below is static
constexpr uint32_t PART_K = PART_THREADS * PART_R;  // windows (keys) counting-sorted at once
#ifndef PART_RUN_ALIGN
#define PART_RUN_ALIGN 1                            // runs are padded to this many keys; 1 = packed (measured best: 4 -> 0.83+0.62, 2 -> 0.74+0.58, 1 -> 0.73+0.53 ms)
#endif
constexpr uint32_t PART_PAD = PART_RUN_ALIGN - 1;
constexpr int PART_TILES_PER_THREAD = 3;             // the sorted kernel handles up to 3 * 1024 tiles (shared memory caps it earlier)
// 227 KiB opt-in limit per CTA (static + dynamic); 228 KiB per SM with 1 KiB reserved per resident CTA
constexpr size_t PART_SMEM_LIMIT = PART_CTAS == 1 ? 232448 : (233472 / PART_CTAS - 1024);

// Staging of one super-chunk (PART_K window starts + 48 bases of tail) as four bit streams:
//   fwd : 16 bases per word, BIG-endian (base i of the word at bits 31-2i..30-2i), A0 C1 G2 T3:
//         a left funnel shift of three consecutive words by 2w puts window w's first base in the MSBs,
//         i.e. the Jellyfish forward key (mer_dna.hpp:330-354) after one right shift by 64-2k
//   rcw : 16 bases per word, LITTLE-endian (base i at bits 2i..2i+1), complemented codes: a right
//         funnel shift by 2w and a 2k-bit mask give the reverse-complement key directly
//   bad : 1 bit per base, set for every byte outside ACGTacgt (window restart, mer_iterator.hpp:72-79)
//   brk : 1 bit per base, set where a new read starts (offset / fixed-stride inputs)
struct PartStage {
  static constexpr int CODE_WORDS = PART_K / 16 + 3;
  static constexpr int BIT_WORDS = PART_K / 32 + 2;
  uint32_t fwd[CODE_WORDS + 1];
  uint32_t rcw[CODE_WORDS + 1];
  uint32_t bad[BIT_WORDS + 2];
  uint32_t brk[BIT_WORDS + 2];
};

constexpr size_t part_static_smem() { return sizeof(PartStage) + 65 * sizeof(uint32_t) + 256; }
// ukeys (8 B) + sorted (tile,index) entries (4 B) per key; count->gdelta and loff per tile; overflow bitmap
static size_t part_dynamic_smem(uint32_t n_tiles, uint32_t n_groups = 0) {
  return (size_t)PART_K * 12 + (size_t)n_tiles * 8 + (size_t)((n_tiles + 31) / 32) * 4 + (size_t)(n_groups + 1) * 4;
}

constexpr int PART_MAX_OWNERS = 8;

// What one launch sorts into. A "bucket" is a hash-prefix GROUP (level 1: from the ASCII stream or a
// key list) or a TILE inside one group (level 2: from the group slabs the level-1 pass filled).
struct PartGeom {
  uint32_t n_buckets_global;   // level 1: groups over all shards (the multiplier of the hash)
  uint32_t bucket_first;       // level 1: first bucket of this shard
  uint32_t n_buckets;          // buckets sorted per super-chunk: level 1 local groups (MULTI: all groups), level 2 group_tiles
  RgGeo geo;                   // level 2: locates the tile inside its group
  uint64_t* out_keys;          // bucket slabs [n_buckets (x n_groups at level 2)][out_cap]
  uint32_t* out_cursor;        // fill of every slab
  uint32_t out_cap;
  // level 2 input: the group slabs
  const uint64_t* in_keys;
  const uint32_t* in_cursor;
  uint32_t in_cap, n_groups;     // n_groups input slabs, in_per_group of them per group (1, or the GPU count)
  uint32_t in_per_group;
  // overflow / foreign keys
  uint64_t* status;
  uint64_t* spill;
  uint64_t spill_cap;
  // MULTI: the slabs of bucket b live on GPU b / buckets_per_owner (peer memory), same layout everywhere:
  // [buckets_per_owner][n_sources][out_cap]; this GPU writes sub-slab my_source and counts in a LOCAL cursor
  uint32_t buckets_per_owner, n_sources, my_source;
  uint64_t* owner_keys[PART_MAX_OWNERS];
  uint32_t* owner_cursor[PART_MAX_OWNERS];
  uint64_t* owner_status[PART_MAX_OWNERS];
  uint64_t* owner_spill[PART_MAX_OWNERS];
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
  if (s0 + PART_K + 64 <= ep.n_bytes) {           // CTA-uniform: no per-load bounds checks
    const uint4* src = reinterpret_cast<const uint4*>(ep.stream + s0);
    pf.body = __ldg(src + threadIdx.x);
    if (threadIdx.x < 3) pf.tail = __ldg(src + PART_THREADS + threadIdx.x);
  } else {
    pf.body = load16_stream(ep, s0 + (uint64_t)threadIdx.x * 16);
    if (threadIdx.x < 3) pf.tail = load16_stream(ep, s0 + (uint64_t)(PART_THREADS + threadIdx.x) * 16);
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
    sm.fwd[PART_THREADS + threadIdx.x] = f;
    sm.rcw[PART_THREADS + threadIdx.x] = r;
    bad16[PART_THREADS + threadIdx.x] = (uint16_t)b;
  }
  if (threadIdx.x < PartStage::BIT_WORDS) sm.brk[threadIdx.x] = pf.brk;
  if (threadIdx.x == 3) {
    sm.fwd[PartStage::CODE_WORDS] = 0;
    sm.rcw[PartStage::CODE_WORDS] = 0;
    bad16[PART_THREADS + 3] = 0xFFFFu;
    sm.bad[PartStage::BIT_WORDS] = 0xFFFFFFFFu;
    sm.bad[PartStage::BIT_WORDS + 1] = 0xFFFFFFFFu;
    sm.brk[PartStage::BIT_WORDS] = 0;
    sm.brk[PartStage::BIT_WORDS + 1] = 0;
  }
}

// The 16 windows starting at bases 16t .. 16t+15 of the staged super-chunk: emit(w, key, ok) with
// ok == false when window w holds no k-mer (key is then garbage).
// k > 16: the three forward words are shifted right by 64-2k ONCE per thread, so each window costs
// two funnel shifts + one mask per strand (the low word needs no mask), a 64-bit compare and a
// select: 10 integer-ALU instructions. k <= 16 takes the generic path.
template <typename Emit>
__device__ __forceinline__ void part_windows(const PartStage& sm, int t, int k, int canonical, Emit&& emit,
                                             uint32_t skip_local = 0) {
  const int p0 = t * PART_R;
  const uint64_t good = ~bits_at(sm.bad, p0);
  const uint64_t nobrk = ~bits_at(sm.brk, p0 + 1);          // bit j: no read starts at base p0+j+1
  uint64_t valid = and_run(good, k);
  if (k > 1) valid &= and_run(nobrk, k - 1);
  const uint32_t W0 = sm.fwd[t], W1 = sm.fwd[t + 1], W2 = sm.fwd[t + 2];
  const uint32_t V0 = sm.rcw[t], V1 = sm.rcw[t + 1], V2 = sm.rcw[t + 2];
  uint32_t valid16 = (uint32_t)valid;
  // first super-chunk of a split push: the windows that start inside the halo belong to the previous piece
  if ((uint32_t)p0 < skip_local) valid16 &= ~((skip_local - p0 >= 32u) ? 0xFFFFFFFFu : ((1u << (skip_local - p0)) - 1u));
  if (k > 16) {
    const int down = 64 - 2 * k;                            // 0..30
    const uint32_t himask = (k == 32) ? 0xFFFFFFFFu : ((1u << (2 * k - 32)) - 1u);
    // X = (W0:W1:W2:0) >> down, as four words
    const uint32_t X0 = W0 >> down, X1 = __funnelshift_r(W1, W0, down), X2 = __funnelshift_r(W2, W1, down),
                   X3 = __funnelshift_r(0u, W2, down);
#pragma unroll
    for (int w = 0; w < PART_R; ++w) {
      const uint32_t fhi = (w ? __funnelshift_l(X1, X0, 2 * w) : X0) & himask;
      const uint32_t flo = w ? __funnelshift_l(X2, X1, 2 * w) : X1;
      const uint32_t rlo = w ? __funnelshift_r(V0, V1, 2 * w) : V0;
      const uint32_t rhi = (w ? __funnelshift_r(V1, V2, 2 * w) : V1) & himask;
      (void)X3;
      const uint64_t fwd = ((uint64_t)fhi << 32) | flo;
      const uint64_t rc = ((uint64_t)rhi << 32) | rlo;
      const uint64_t key = (!canonical || fwd < rc) ? fwd : rc;
      emit(w, key, (valid16 & (1u << w)) != 0);
    }
  } else {
    const uint64_t mask = (1ull << (2 * k)) - 1;
    const int down = 64 - 2 * k;
#pragma unroll
    for (int w = 0; w < PART_R; ++w) {
      const uint32_t fhi = w ? __funnelshift_l(W1, W0, 2 * w) : W0;
      const uint32_t flo = w ? __funnelshift_l(W2, W1, 2 * w) : W1;
      const uint32_t rlo = w ? __funnelshift_r(V0, V1, 2 * w) : V0;
      const uint32_t rhi = w ? __funnelshift_r(V1, V2, 2 * w) : V1;
      const uint64_t fwd = (((uint64_t)fhi << 32) | flo) >> down;
      const uint64_t rc = (((uint64_t)rhi << 32) | rlo) & mask;
      const uint64_t key = (!canonical || fwd < rc) ? fwd : rc;
      emit(w, key, (valid16 & (1u << w)) != 0);
    }
  }
}

// SRC: 0 = windows of the ASCII stream, 1 = an already extracted key list, 2 = the group slabs of a
// level-1 pass (each CTA takes whole groups and sorts their keys by tile inside the group).
// PART_K keys are counting-sorted by bucket in shared memory per super-chunk:
//   1. keys -> ukeys[] (transposed: key w of thread t at w*1024+t, conflict-free) and ONE shared
//      atomic per key on the bucket histogram whose return value is the key's rank inside its run;
//      (bucket, rank) stays in a register
//   2. block scan -> local offsets; one global atomicAdd per non-empty bucket reserves the run; its
//      latency hides behind step 3
//   3. entry (bucket, index of the key in ukeys) is scattered to its sorted position
//   4. copy-out in sorted order: consecutive threads write consecutive keys of a run
// The ASCII of the next super-chunk is prefetched into registers during steps 1-2.
// SHARDED: the table holds one hash-prefix shard (multi-GPU), so keys of other shards are counted as foreign.
// MULTI:   level 1 over ALL shards: runs are written straight into the owner GPU's slabs over NVLink
//          peer memory, reserved with system-scope atomics on the owner's cursors (no collective).
template <int SRC, bool SHARDED, bool MULTI>
__global__ void __launch_bounds__(PART_THREADS, PART_CTAS)
partition_sorted_kernel(ExtractParams ep, const uint64_t* __restrict__ in_keys_arg, uint64_t n_in_arg, PartGeom g) {
  extern __shared__ __align__(16) unsigned char dyn[];
  __shared__ PartStage sm;
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t scan_total;
  __shared__ uint32_t any_overflow;
  __shared__ uint32_t dummy_hist[32];    // where the windows without a k-mer "count" (never read)
  // MULTI: the owners' base pointers, indexed per bucket / per key (a dynamically indexed kernel
  // parameter array would be copied to local memory)
  __shared__ uint64_t* s_owner_keys[PART_MAX_OWNERS];
  __shared__ uint32_t* s_owner_cursor[PART_MAX_OWNERS];
  if (MULTI && threadIdx.x < PART_MAX_OWNERS) {
    s_owner_keys[threadIdx.x] = g.owner_keys[threadIdx.x];
    s_owner_cursor[threadIdx.x] = g.owner_cursor[threadIdx.x];
  }
  uint64_t* ukeys = reinterpret_cast<uint64_t*>(dyn);                        // [PART_K]
  uint32_t* sorted = reinterpret_cast<uint32_t*>(dyn + (size_t)PART_K * 8);  // [PART_K] (bucket << 16) | key index
  uint32_t* hist = sorted + PART_K;       // [n_buckets] run length, then (step 3b) offset of the run in its slab array - loff
  uint32_t* loff = hist + g.n_buckets;    // [n_buckets] local offset of the run
  uint32_t* ovf_bits = loff + g.n_buckets;  // [(n_buckets+31)/32] buckets whose slab overflowed in this super-chunk
  const uint32_t per_thread = (g.n_buckets + PART_THREADS - 1) / PART_THREADS;
  const uint32_t ovf_words = (g.n_buckets + 31) >> 5;

  PartPrefetch pf;
  uint32_t r_next = 0, r_step = 0;   // (start of the next super-chunk to prefetch) % read_stride, and its increment
  if (SRC == 0 && ep.read_stride != 0) {
    r_next = (uint32_t)(((uint64_t)blockIdx.x * PART_K) % ep.read_stride);
    r_step = (uint32_t)(((uint64_t)gridDim.x * PART_K) % ep.read_stride);
  }
  auto advance_r = [&]() {
    r_next += r_step;
    if (r_next >= ep.read_stride) r_next -= ep.read_stride;
  };
  if (SRC == 0 && (uint64_t)blockIdx.x * PART_K < ep.n_bytes) {
    part_prefetch(pf, ep, (uint64_t)blockIdx.x * PART_K, r_next);
    advance_r();
    part_stage_store(sm, pf);
  }
  if (threadIdx.x == 0) any_overflow = 0;
  for (uint32_t i = threadIdx.x; i < ovf_words; i += PART_THREADS) ovf_bits[i] = 0;

  // Work items = super-chunks. Level 2: the super-chunks of ALL groups, numbered group after group
  // (gstart[] = exclusive prefix of the groups' chunk counts, built here from the slab cursors), so
  // that the whole grid is busy whatever the number of groups.
  uint32_t* gstart = ovf_bits + ovf_words;        // [n_groups + 1], level 2 only
  uint64_t total_items;
  if (SRC == 2) {
    __syncthreads();
    uint32_t run_g = 0;
    for (uint32_t base = 0; base < g.n_groups; base += PART_THREADS) {
      const uint32_t gi = base + threadIdx.x;
      const uint32_t c = gi < g.n_groups ? (min(g.in_cursor[gi], g.in_cap) + PART_K - 1) / PART_K : 0u;
      const uint32_t ex = block_exclusive_scan(c, warp_sums, &scan_total);
      if (gi < g.n_groups) gstart[gi] = run_g + ex;
      __syncthreads();
      run_g += scan_total;
      __syncthreads();
    }
    if (threadIdx.x == 0) gstart[g.n_groups] = run_g;
    __syncthreads();
    total_items = run_g;
  } else {
    total_items = ((SRC == 0 ? ep.n_bytes : n_in_arg) + PART_K - 1) / PART_K;
  }
  uint32_t og = 0;                                 // level 2: group of the current item
  for (uint64_t item = blockIdx.x; item < total_items; item += gridDim.x) {
    if (SRC == 2)
      while (item >= gstart[og + 1]) ++og;
    const uint64_t* in_keys = SRC == 2 ? g.in_keys + (uint64_t)og * g.in_cap : in_keys_arg;
    const uint64_t n_in = SRC == 2 ? (uint64_t)min(g.in_cursor[og], g.in_cap) : n_in_arg;
    // first output slab of this group (peer-to-peer: the input slabs are (group, source GPU) pairs)
    const uint32_t slab0 = SRC == 2 ? (og / g.in_per_group) * g.n_buckets : 0u;
    const uint64_t sc = SRC == 2 ? item - gstart[og] : item;
    const uint64_t n_super = total_items;
    const uint64_t s0 = sc * PART_K;
    for (uint32_t i = threadIdx.x; i < g.n_buckets; i += PART_THREADS) hist[i] = 0;
    const bool more = item + gridDim.x < n_super;
    if (SRC == 0 && more) {
      part_prefetch(pf, ep, (sc + gridDim.x) * PART_K, r_next);
      advance_r();
    }
    __syncthreads();
    // ---- 1. keys -> ukeys, rank inside the bucket's run
    uint32_t meta[PART_R];           // (bucket << 16) | rank, or ~0 for a window without a k-mer
    uint32_t foreign = 0;
    // branch-free: the hash is computed for every window, only the shared atomic is predicated
    auto take = [&](int w, uint64_t key, bool ok) {
      uint32_t b;
      if (SRC == 2) {
        uint32_t slot_unused;
        rg_locate_in_group(key, g.geo, b, slot_unused);
      } else {
        b = rg_group_of_key(key, g.n_buckets_global);
        if (SHARDED) {
          b -= g.bucket_first;
          const bool mine = b < g.n_buckets;
          foreign += (ok && !mine) ? 1u : 0u;
          ok = ok && mine;
        }
      }
      // the shared atomic is UNCONDITIONAL (no branch around it): windows without a k-mer count
      // into a per-lane dummy word instead of a bucket
      uint32_t* cell = ok ? hist + b : dummy_hist + (threadIdx.x & 31);
      const uint32_t rank = atomicAdd(cell, 1u);
      meta[w] = ok ? b * 65536u + rank : 0xFFFFFFFFu;
      ukeys[w * PART_THREADS + threadIdx.x] = key;
    };
    if (SRC != 0) {
      uint64_t kk[PART_R];
#pragma unroll
      for (int w = 0; w < PART_R; ++w) {
        const uint64_t i = s0 + (uint64_t)w * PART_THREADS + threadIdx.x;
        kk[w] = i < n_in ? __ldg(in_keys + i) : RG_EMPTY;
      }
#pragma unroll
      for (int w = 0; w < PART_R; ++w) take(w, kk[w], kk[w] != RG_EMPTY);
    } else {
      part_windows(sm, threadIdx.x, ep.k, ep.canonical, take, s0 == 0 ? ep.skip : 0u);
    }
    if (SHARDED && foreign) atomicAdd((unsigned long long*)&g.status[1], (unsigned long long)foreign);
    __syncthreads();
    // ---- 2. local offsets (exclusive scan of the run lengths); the global run reservations (one
    // atomicAdd per non-empty bucket) are only ISSUED here: their results are consumed after
    // step 3, which hides the round trip to L2 (or to the peer GPU)
    const uint32_t t0 = threadIdx.x * per_thread;
    uint32_t cnt[PART_TILES_PER_THREAD], gb[PART_TILES_PER_THREAD];
    uint32_t sum = 0;
#pragma unroll
    for (int j = 0; j < PART_TILES_PER_THREAD; ++j) {
      cnt[j] = ((uint32_t)j < per_thread && t0 + j < g.n_buckets) ? hist[t0 + j] : 0u;
      sum += cnt[j];
    }
    uint32_t run = block_exclusive_scan(sum, warp_sums, &scan_total);
    const uint32_t run0 = run;
#pragma unroll
    for (int j = 0; j < PART_TILES_PER_THREAD; ++j) {
      gb[j] = 0;
      if (cnt[j]) {
        const uint32_t b = t0 + j;
        const uint32_t want = (cnt[j] + PART_PAD) & ~PART_PAD;
        if (MULTI) {
          // every source GPU fills its OWN sub-slab of the owner's group slab: the cursor is local
          // (no atomics over NVLink), only the keys travel
          const uint32_t o = b / g.buckets_per_owner;
          loff[b] = run | (o << 16);                 // the owner rides in the high half (run < 2^14)
          gb[j] = atomicAdd(g.out_cursor + b, want);
        } else {
          loff[b] = run;
          gb[j] = atomicAdd(g.out_cursor + slab0 + b, want);
        }
        run += cnt[j];
      }
    }
    // the next super-chunk's bases are in registers by now; the stage buffers are free since step 1 ended
    if (SRC == 0 && more) part_stage_store(sm, pf);
    __syncthreads();
    // ---- 3. (bucket, key index) entries in bucket order
#pragma unroll
    for (int w = 0; w < PART_R; ++w) {
      const uint32_t m = meta[w];
      const uint32_t b = m >> 16;
      if (m != 0xFFFFFFFFu)
        sorted[(loff[b] & 0xFFFFu) + (m & 0xFFFFu)] = b * 65536u + (w * PART_THREADS + threadIdx.x);
    }
    // ---- 3b. the reservations have landed: pad slots, sorted index -> slab-array index (into hist[])
    run = run0;
#pragma unroll
    for (int j = 0; j < PART_TILES_PER_THREAD; ++j) {
      if (cnt[j]) {
        const uint32_t b = t0 + j;
        const uint32_t padded = (cnt[j] + PART_PAD) & ~PART_PAD;
        const uint32_t o = MULTI ? b / g.buckets_per_owner : 0u;
        // slab index inside its owner's array (MULTI: [group of the owner][source GPU])
        const uint32_t slab = MULTI ? (b - o * g.buckets_per_owner) * g.n_sources + g.my_source : slab0 + b;
        uint64_t* base = MULTI ? s_owner_keys[o] : g.out_keys;
        const uint32_t at = slab * g.out_cap + gb[j];                            // index of the run in the slab array
        if (gb[j] + padded <= g.out_cap) {
#pragma unroll
          for (uint32_t q = 1; q <= PART_PAD; ++q)                               // at most PART_PAD pad slots
            if (cnt[j] + q - 1 < padded) base[at + padded - q] = RG_EMPTY;
          hist[b] = at - run;
        } else {
          // slab overflow (heavy hitters): this run goes to the spill list (step 4); what it
          // reserved below the capacity is neutralised with RG_EMPTY
          for (uint32_t q = gb[j]; q < g.out_cap; ++q) base[(uint64_t)slab * g.out_cap + q] = RG_EMPTY;
          atomicOr(&ovf_bits[b >> 5], 1u << (b & 31));
          any_overflow = 1;
        }
        run += cnt[j];
      }
    }
    __syncthreads();
    const bool ovf = any_overflow != 0;
    // ---- 4. copy-out in sorted order
    const uint32_t total = scan_total;
    if (!ovf) {
#pragma unroll 4
      for (uint32_t i = threadIdx.x; i < total; i += PART_THREADS) {
        const uint32_t e = sorted[i];
        const uint32_t b = e >> 16;
        uint64_t* base = MULTI ? s_owner_keys[loff[b] >> 16] : g.out_keys;
        base[hist[b] + i] = ukeys[e & 0xFFFFu];
      }
    } else {
      for (uint32_t i = threadIdx.x; i < total; i += PART_THREADS) {
        const uint32_t e = sorted[i];
        const uint32_t b = e >> 16;
        const uint64_t key = ukeys[e & 0xFFFFu];
        const uint32_t o = MULTI ? loff[b] >> 16 : 0u;
        if (!((ovf_bits[b >> 5] >> (b & 31)) & 1u)) {
          uint64_t* base = MULTI ? s_owner_keys[o] : g.out_keys;
          base[hist[b] + i] = key;
        } else if (MULTI) {
          const uint64_t sp = atomicAdd_system((unsigned long long*)&g.owner_status[o][0], 1ull);
          if (sp < g.spill_cap) g.owner_spill[o][sp] = key;
        } else {
          const uint64_t sp = atomicAdd((unsigned long long*)&g.status[0], 1ull);
          if (sp < g.spill_cap) g.spill[sp] = key;
        }
      }
      __syncthreads();
      if (threadIdx.x == 0) any_overflow = 0;
      for (uint32_t i = threadIdx.x; i < ovf_words; i += PART_THREADS) ovf_bits[i] = 0;
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
    const uint32_t tile = rg_group_of_key(key, g.n_buckets_global) - g.bucket_first;
    if (tile >= g.n_buckets) {
      atomicAdd((unsigned long long*)&g.status[1], 1ull);
      return;
    }
    const uint32_t at = atomicAdd(g.out_cursor + tile, 1u);
    if (at < g.out_cap) {
      g.out_keys[(uint64_t)tile * g.out_cap + at] = key;
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
      }, c == 0 ? ep.skip : 0u);
    }
  }
}

// does the sorted kernel's per-bucket state fit shared memory for this bucket count?
static bool part_sorted_fits(uint32_t n_buckets) {
  return n_buckets <= PART_TILES_PER_THREAD * PART_THREADS && n_buckets < 65536 &&
         part_dynamic_smem(n_buckets) + part_static_smem() <= PART_SMEM_LIMIT;
}
// the largest bucket count one sorted pass can take (level-1 fanout limit)
uint32_t rg_partition_max_buckets() {
  uint32_t n = PART_TILES_PER_THREAD * PART_THREADS;
  while (n > 1 && !part_sorted_fits(n)) --n;
  return n;
}

// upper bound on the pad keys one bucket can receive in a push of n_units windows/keys
uint64_t rg_partition_pad_keys(uint32_t n_buckets, uint64_t n_units) {
  if (!part_sorted_fits(n_buckets)) return 0;
  return (uint64_t)PART_PAD * ((n_units + PART_K - 1) / PART_K);
}

template <int SRC, bool MULTI>
static int launch_sorted(rg_ctx* ctx, const char* name, const ExtractParams& ep, const uint64_t* keys, uint64_t n_keys,
                         const PartGeom& g, uint64_t n_work, bool sharded, uint32_t max_ctas = 0) {
  const size_t smem = part_dynamic_smem(g.n_buckets, g.n_groups);
  auto kern = sharded ? partition_sorted_kernel<SRC, true, MULTI> : partition_sorted_kernel<SRC, false, MULTI>;
  RG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  RG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  unsigned grid = rg_grid(ctx, PART_CTAS);
  if (max_ctas && max_ctas < grid) grid = max_ctas;
  if (n_work < grid) grid = (unsigned)n_work;
  rg_launch_scope _ls(ctx, name);
  kern<<<grid, PART_THREADS, smem, ctx->stream>>>(ep, keys, n_keys, g);
  return 0;
}

static void fill_common(PartGeom& g, rg_table* t) {
  memset(&g, 0, sizeof(g));
  g.geo.groups = t->n_tiles_global / t->group_tiles;
  g.geo.group_tiles = t->group_tiles;
  g.status = t->d_status;
  g.spill = t->spill;
  g.spill_cap = t->spill_cap;
}

// level 1: stream / key list -> the table's groups (group_tiles == 1: straight into the tile buckets)
template <bool FROM_KEYS>
static int launch_partition(rg_table* t, const ExtractParams& ep, const uint64_t* keys, uint64_t n_keys) {
  rg_ctx* ctx = t->ctx;
  const uint64_t n_units = FROM_KEYS ? n_keys : ep.n_bytes;
  if (n_units == 0) return 0;
  PartGeom g;
  fill_common(g, t);
  const bool two_level = t->group_tiles > 1;
  g.n_buckets_global = g.geo.groups;
  g.n_buckets = t->n_tiles / t->group_tiles;
  g.bucket_first = t->tile_first / t->group_tiles;
  g.out_keys = two_level ? t->grp_keys : t->part_keys;
  g.out_cursor = two_level ? t->grp_cursor : t->part_cursor;
  g.out_cap = two_level ? t->grp_cap : t->bucket_cap;
  const bool sharded = g.n_buckets_global != g.n_buckets;
  const char* name = FROM_KEYS ? "partition_keys_kernel" : "partition_kernel";
  int rc = 0;
  if (part_sorted_fits(g.n_buckets) && (uint64_t)g.n_buckets * g.out_cap < (1ull << 32)) {
    const uint64_t n_super = (n_units + PART_K - 1) / PART_K;
    rc = launch_sorted<FROM_KEYS ? 1 : 0, false>(ctx, name, ep, keys, n_keys, g, n_super, sharded);
  } else {
    rg_launch_scope _ls(ctx, name);
    partition_scatter_kernel<FROM_KEYS><<<rg_grid(ctx, 8), 256, 0, ctx->stream>>>(ep, keys, n_keys, g);
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
  if (!part_sorted_fits(g.n_buckets) || (uint64_t)t->n_tiles * t->bucket_cap >= (1ull << 32) ||
      part_dynamic_smem(g.n_buckets, g.n_groups) + part_static_smem() > PART_SMEM_LIMIT) {
    rg_set_error("partition: group geometry out of range");
    return RG_ERR_ARG;
  }
  ExtractParams ep = {};
  // work items = super-chunks over all groups (their number is only known on the device): a full grid
  int rc = launch_sorted<2, false>(ctx, "partition_groups_kernel", ep, nullptr, 0, g, 1ull << 32, false);
  if (rc) return rc;
  RG_CUDA_OK(cudaGetLastError());
  return 0;
}

// level 1 over ALL shards, straight into the owners' group slabs (peer memory)
int rg_launch_partition_multi(rg_table* t, const ExtractParams& ep) {
  rg_ctx* ctx = t->ctx;
  if (ep.n_bytes == 0) return 0;
  PartGeom g;
  fill_common(g, t);
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
    g.owner_cursor[o] = t->peer_cursor[o];
    g.owner_status[o] = t->peer_status[o];
    g.owner_spill[o] = t->peer_spill[o];
  }
  if (!part_sorted_fits(g.n_buckets) || (uint64_t)g.buckets_per_owner * g.n_sources * g.out_cap >= (1ull << 32)) {
    rg_set_error("partition: too many groups for the peer-to-peer partition");
    return RG_ERR_ARG;
  }
  const uint64_t n_super = (ep.n_bytes + PART_K - 1) / PART_K;
  int rc = launch_sorted<0, true>(ctx, "partition_p2p_kernel", ep, nullptr, 0, g, n_super, false, t->p2p_ctas);
  if (rc) return rc;
  RG_CUDA_OK(cudaGetLastError());
  return 0;
}
