// rg_filter.cu — the read pull. Replaces RUFUS.Filter's per-pair loop
// (src/RUFUS.Filter.cpp:196-277) and RUFUS.Filter.single (src/RUFUS.Filter.ss.cpp:164-201).
//
// Reference semantics per read (RUFUS.Filter.cpp:203-221):
//   streak resets at i when qual[i]-33 < MinQ or base == 'N', else streak++;
//   when streak >= K the K-mer ending at i is looked up (Util::HashToLong packing:
//   base j of the window at bits (2j,2j+1); C=2 G=1 T=3, every other byte = A);
//   paired tool: i < length()-1 (the window ending on the last base is never tested).
// "streak >= K at i" == "the K bases ending at i are all good and inside the read",
// so the kernel is position-parallel over the concatenated read stream: a window is
// tested iff its K flag bits are clear; set membership goes through a shared-memory
// bitmap prefilter, then the (L2-resident) open-addressed set; only on a hit is the
// owning read located (binary search over offsets, or a division for fixed-length
// reads) and the read-boundary / last-base rule applied.
#include "rg_internal.cuh"

#include "rg_extract.cuh"

constexpr int FILTER_THREADS = 512;
constexpr int FILTER_R = 16;                         // windows per thread: one 16-base word, all shifts static
constexpr int FCHUNK = FILTER_THREADS * FILTER_R;    // window starts per CTA chunk
constexpr int FCODE_WORDS = FCHUNK / 16 + 3;         // + 48 bases of tail
constexpr int FBIT_WORDS = FCHUNK / 32 + 2;

// one multiply: the high word picks the prefilter bit, the middle bits the set slot
__host__ __device__ __forceinline__ uint64_t rg_set_mix(uint64_t v) {
  return (v ^ (v >> 29)) * 0x9E3779B97F4A7C15ull;
}
__host__ __device__ __forceinline__ uint32_t rg_set_bit(uint64_t mix, uint32_t bitmap_bits) {
  return (uint32_t)(mix >> 32) & (bitmap_bits - 1);
}
__host__ __device__ __forceinline__ uint32_t rg_set_slot(uint64_t mix, uint32_t slots) {
  return (uint32_t)(mix >> 17) & (slots - 1);
}

// ---- set construction -----------------------------------------------------
__global__ void set_build_kernel(const uint64_t* __restrict__ vals, uint64_t n, uint64_t* table, uint32_t slots,
                                 uint32_t* bitmap, uint32_t bitmap_bits, unsigned long long* n_distinct) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t v = vals[i];
    if (v == RG_EMPTY) continue;   // handled by has_all_ones on the host
    const uint64_t mix = rg_set_mix(v);
    const uint32_t bit = rg_set_bit(mix, bitmap_bits);
    atomicOr(bitmap + (bit >> 5), 1u << (bit & 31));
    uint32_t slot = rg_set_slot(mix, slots);
    for (;;) {
      const unsigned long long cur = atomicCAS((unsigned long long*)(table + slot), (unsigned long long)RG_EMPTY,
                                               (unsigned long long)v);
      if (cur == RG_EMPTY) {
        atomicAdd(n_distinct, 1ull);
        break;
      }
      if (cur == v) break;
      slot = (slot + 1) & (slots - 1);
    }
  }
}

int rg_launch_set_build(rg_set* s, const uint64_t* d_vals, uint64_t n) {
  if (n == 0) return 0;
  rg_ctx* ctx = s->ctx;
  RG_CUDA_OK(cudaMemsetAsync(ctx->d_scratch, 0, sizeof(uint64_t), ctx->stream));
  unsigned grid = (unsigned)((n + 255) / 256);
  if (grid > rg_grid(ctx, 8)) grid = rg_grid(ctx, 8);
  {
    rg_launch_scope _ls(ctx, "set_build_kernel");
    set_build_kernel<<<grid, 256, 0, ctx->stream>>>(d_vals, n, s->table, s->slots, s->bitmap, s->bitmap_bits,
                                                    reinterpret_cast<unsigned long long*>(ctx->d_scratch));
  }
  RG_CUDA_OK(cudaGetLastError());
  return 0;
}

// canonical jf keys -> {rufus key of the k-mer, rufus key of its reverse complement}
__global__ void jf_to_rufus_kernel(const uint64_t* __restrict__ jf, uint64_t n, int k, uint64_t* out) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t f = jf[i];
    if (f == RG_EMPTY) {            // padding of a fixed-capacity key list (never a canonical key): skipped by the set build
      out[2 * i] = RG_EMPTY;
      out[2 * i + 1] = RG_EMPTY;
      continue;
    }
    const uint64_t r = rg_revcomp_jf(f, k);
    uint64_t a = 0, b = 0;
    for (int j = 0; j < k; ++j) {
      // base j (from the left) of the forward k-mer: jf code A0 C1 G2 T3 -> rufus code A0 C2 G1 T3
      const uint32_t cf = (uint32_t)(f >> (2 * (k - 1 - j))) & 3u;
      const uint32_t cr = (uint32_t)(r >> (2 * (k - 1 - j))) & 3u;
      a |= (uint64_t)(((cf & 1u) << 1) | (cf >> 1)) << (2 * j);
      b |= (uint64_t)(((cr & 1u) << 1) | (cr >> 1)) << (2 * j);
    }
    out[2 * i] = a;
    out[2 * i + 1] = b;
  }
}

int rg_launch_jf_to_rufus(rg_ctx* ctx, const uint64_t* d_jf, uint64_t n, int k, uint64_t* d_out2n) {
  if (n == 0) return 0;
  unsigned grid = (unsigned)((n + 255) / 256);
  if (grid > rg_grid(ctx, 8)) grid = rg_grid(ctx, 8);
  {
    rg_launch_scope _ls(ctx, "jf_to_rufus_kernel");
    jf_to_rufus_kernel<<<grid, 256, 0, ctx->stream>>>(d_jf, n, k, d_out2n);
  }
  RG_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- the filter -------------------------------------------------------------
struct FilterStage {
  uint32_t codes[FCODE_WORDS + 1];   // little-endian: base i of a word at bits (2i, 2i+1), rufus codes
  uint32_t bad[FBIT_WORDS + 2];      // little-endian flag bits: low quality, 'N', or past the end of the stream
};

// 16 sequence + 16 quality bytes -> 16 rufus codes (Util::HashToLong: C -> bit 2i+1, G -> bit 2i,
// T -> both, anything else 0) and 16 streak-reset flags (qual-33 < MinQ as signed char, or 'N').
// thr4 = (MinQ + 33) replicated in the four bytes; all_bad / none_bad cover thresholds outside a char.
__device__ __forceinline__ void fconvert16(const uint4 s, const uint4 q, uint32_t thr4, bool all_bad, uint32_t& codes,
                                           uint32_t& bad) {
  const uint32_t sw[4] = {s.x, s.y, s.z, s.w};
  const uint32_t qw[4] = {q.x, q.y, q.z, q.w};
  codes = 0;
  bad = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t c4 = (__vcmpeq4(sw[i], 0x43434343u) & 0x02020202u) | (__vcmpeq4(sw[i], 0x47474747u) & 0x01010101u) |
                        (__vcmpeq4(sw[i], 0x54545454u) & 0x03030303u);
    codes |= ((c4 | (c4 >> 6) | (c4 >> 12) | (c4 >> 18)) & 0xFFu) << (8 * i);
    const uint32_t b4 = (__vcmplts4(qw[i], thr4) | __vcmpeq4(sw[i], 0x4E4E4E4Eu)) & 0x01010101u;
    bad |= ((b4 | (b4 >> 7) | (b4 >> 14) | (b4 >> 21)) & 0xFu) << (4 * i);
  }
  if (all_bad) bad = 0xFFFFu;
}

__device__ __forceinline__ uint4 load16_guarded(const char* base, uint64_t p, uint64_t n, unsigned char fill) {
  if (p + 16 <= n) return __ldg(reinterpret_cast<const uint4*>(base + p));
  const uint32_t f4 = fill * 0x01010101u;
  if (p >= n) return make_uint4(f4, f4, f4, f4);
  unsigned char tmp[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) tmp[j] = (p + j < n) ? (unsigned char)base[p + j] : fill;
  return *reinterpret_cast<uint4*>(tmp);
}

struct FilterPrefetch {
  uint4 s, q, ts, tq;
};

__device__ __forceinline__ void filter_prefetch(FilterPrefetch& pf, const char* seq, const char* qual, uint64_t n_bytes,
                                                uint64_t c0) {
  // past-the-end bases read as 'N': the streak resets there, like at the end of the reference's read
  pf.s = load16_guarded(seq, c0 + (uint64_t)threadIdx.x * 16, n_bytes, 'N');
  pf.q = load16_guarded(qual, c0 + (uint64_t)threadIdx.x * 16, n_bytes, 0);
  if (threadIdx.x < 3) {
    pf.ts = load16_guarded(seq, c0 + (uint64_t)(FILTER_THREADS + threadIdx.x) * 16, n_bytes, 'N');
    pf.tq = load16_guarded(qual, c0 + (uint64_t)(FILTER_THREADS + threadIdx.x) * 16, n_bytes, 0);
  }
}

__device__ __forceinline__ void filter_stage_store(FilterStage& sm, const FilterPrefetch& pf, uint32_t thr4,
                                                   bool all_bad) {
  uint16_t* bad16 = reinterpret_cast<uint16_t*>(sm.bad);
  uint32_t c, b;
  fconvert16(pf.s, pf.q, thr4, all_bad, c, b);
  sm.codes[threadIdx.x] = c;
  bad16[threadIdx.x] = (uint16_t)b;
  if (threadIdx.x < 3) {
    fconvert16(pf.ts, pf.tq, thr4, all_bad, c, b);
    sm.codes[FILTER_THREADS + threadIdx.x] = c;
    bad16[FILTER_THREADS + threadIdx.x] = (uint16_t)b;
  }
  if (threadIdx.x == 3) {
    sm.codes[FCODE_WORDS] = 0;
    bad16[FILTER_THREADS + 3] = 0xFFFFu;
    sm.bad[FBIT_WORDS] = 0xFFFFFFFFu;
    sm.bad[FBIT_WORDS + 1] = 0xFFFFFFFFu;
  }
}

// Position-parallel: thread t owns the 16 windows starting at bases 16t..16t+15 of the chunk. A
// window is tested iff its K flag bits are clear; its rufus key is two static funnel shifts of the
// staged code words; membership = shared-memory bitmap prefilter (one multiply), then the
// L2-resident set. Only on a hit is the owning read located and the read-boundary rule applied.
__global__ void __launch_bounds__(FILTER_THREADS)
filter_hits_kernel(const char* __restrict__ seq, const char* __restrict__ qual, const uint64_t* __restrict__ offs,
                   uint32_t read_len, uint64_t n_reads, uint64_t n_bytes, uint64_t pos_base, int k, int minq,
                   int skip_last, const uint64_t* __restrict__ table, uint32_t slots, const uint32_t* __restrict__ g_bitmap,
                   uint32_t bitmap_bits, int has_all_ones, int32_t* __restrict__ hits,
                   uint8_t* __restrict__ hit_flags) {
  extern __shared__ __align__(16) uint32_t s_bitmap[];   // bitmap_bits-bit prefilter of the set
  __shared__ FilterStage sm;
  {
    const uint4* src = reinterpret_cast<const uint4*>(g_bitmap);
    uint4* dst = reinterpret_cast<uint4*>(s_bitmap);
    for (uint32_t i = threadIdx.x; i < bitmap_bits / 128; i += blockDim.x) dst[i] = __ldg(src + i);
  }
  // RUFUS.Filter.cpp:207: `qual[i]-33 < MinQ` on a (signed) char
  const int thr = minq + 33;
  const bool all_bad = thr > 127;
  const uint32_t thr4 = (uint32_t)(uint8_t)(int8_t)(thr < -128 ? -128 : (thr > 127 ? 127 : thr)) * 0x01010101u;
  const uint64_t kmask = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1);
  const uint64_t n_chunks = (n_bytes + FCHUNK - 1) / FCHUNK;
  FilterPrefetch pf;
  if (blockIdx.x < n_chunks) filter_prefetch(pf, seq, qual, n_bytes, (uint64_t)blockIdx.x * FCHUNK);
  for (uint64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const uint64_t c0 = c * FCHUNK;
    __syncthreads();                         // previous chunk's windows are done with the stage
    filter_stage_store(sm, pf, thr4, all_bad);
    if (c + gridDim.x < n_chunks) filter_prefetch(pf, seq, qual, n_bytes, (c + gridDim.x) * FCHUNK);
    __syncthreads();
    const int t = threadIdx.x;
    const uint32_t valid = (uint32_t)and_run(~bits_at(sm.bad, t * FILTER_R), k) & 0xFFFFu;
    if (valid == 0) continue;
    const uint32_t V0 = sm.codes[t], V1 = sm.codes[t + 1], V2 = sm.codes[t + 2];
    // pass 1 (branch-free): which windows pass the prefilter
    uint32_t maybe = 0;
#pragma unroll
    for (int w = 0; w < FILTER_R; ++w) {
      const uint32_t lo = w ? __funnelshift_r(V0, V1, 2 * w) : V0;
      const uint32_t hi = w ? __funnelshift_r(V1, V2, 2 * w) : V1;
      const uint64_t v = (((uint64_t)hi << 32) | lo) & kmask;
      const uint32_t bit = rg_set_bit(rg_set_mix(v), bitmap_bits);
      maybe |= ((s_bitmap[bit >> 5] >> (bit & 31)) & 1u) << w;
    }
    maybe &= valid;
    // pass 2 (rare): exact lookup, then the read that owns the window
    while (maybe) {
      const int w = __ffs(maybe) - 1;
      maybe &= maybe - 1;
      uint64_t x = ((uint64_t)V1 << 32) | V0;
      if (w) x = (x >> (2 * w)) | ((uint64_t)V2 << (64 - 2 * w));
      const uint64_t v = x & kmask;
      bool hit;
      if (v == RG_EMPTY) {
        hit = has_all_ones != 0;
      } else {
        hit = false;
        uint32_t slot = rg_set_slot(rg_set_mix(v), slots);
        for (;;) {
          const uint64_t cur = __ldg(table + slot);
          if (cur == v) { hit = true; break; }
          if (cur == RG_EMPTY) break;
          slot = (slot + 1) & (slots - 1);
        }
      }
      if (!hit) continue;
      const uint64_t gp = pos_base + c0 + (uint64_t)t * FILTER_R + w;   // absolute stream position
      uint64_t r, r1;
      if (offs != nullptr) {
        uint64_t lo = 0, hi = n_reads;   // largest r with offs[r] <= gp
        while (hi - lo > 1) {
          const uint64_t mid = (lo + hi) >> 1;
          if (__ldg(offs + mid) <= gp) lo = mid; else hi = mid;
        }
        r = lo;
        r1 = __ldg(offs + r + 1);
        // empty reads share their offset with the next read: skip forward to the owner
        while (r1 <= gp && r + 1 < n_reads) { ++r; r1 = __ldg(offs + r + 1); }
      } else {
        r = gp / read_len;
        r1 = (r + 1) * read_len;
      }
      const uint64_t end_excl = skip_last ? (r1 > 0 ? r1 - 1 : 0) : r1;   // window must end before this
      if (gp + k <= end_excl) {
        atomicAdd(hits + r, 1);
        if (hit_flags) hit_flags[gp] = 1;      // AnnotateOverlap: which window starts hit (per-base depth comes next)
      }
    }
  }
}

int rg_launch_filter_hits(rg_set* s, const char* seq, const char* qual, const uint64_t* offs, uint32_t read_len,
                          uint64_t n_reads, uint64_t n_bytes, uint64_t pos_base, bool zero_hits, int minq, int skip_last,
                          int32_t* hits, uint8_t* hit_flags) {
  rg_ctx* ctx = s->ctx;
  if (zero_hits) RG_CUDA_OK(cudaMemsetAsync(hits, 0, n_reads * sizeof(int32_t), ctx->stream));
  if (n_bytes == 0 || n_reads == 0) return 0;
  const size_t smem = s->bitmap_bits / 8;
  // (the attribute is per device: set on every launch, a process may drive several GPUs)
  RG_CUDA_OK(cudaFuncSetAttribute(filter_hits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(RG_SET_BITMAP_BITS / 8)));
  const uint64_t n_chunks = (n_bytes + FCHUNK - 1) / FCHUNK;
  unsigned grid = rg_grid(ctx, smem >= 65536 ? 3 : 4);
  if (n_chunks < grid) grid = (unsigned)n_chunks;
  {
    rg_launch_scope _ls(ctx, "filter_hits_kernel");
    filter_hits_kernel<<<grid, FILTER_THREADS, smem, ctx->stream>>>(seq, qual, offs, read_len, n_reads, n_bytes,
                                                                    pos_base, s->k, minq, skip_last, s->table, s->slots,
                                                                    s->bitmap,
                                                                    s->bitmap_bits, s->has_all_ones, hits, hit_flags);
  }
  RG_CUDA_OK(cudaGetLastError());
  return 0;
}

// AnnotateOverlap (src/AnnotateOverlap.cpp:117-146): depth[j] = number of hit windows covering base j,
// i.e. hit window starts in [j-k+1, j] that belong to the same read.
__global__ void coverage_kernel(const uint8_t* __restrict__ hit_flags, const uint64_t* __restrict__ offs,
                                uint64_t n_reads, uint64_t n_bytes, int k, uint8_t* __restrict__ cov) {
  for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < n_bytes;
       j += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t lo = 0, hi = n_reads;            // read that owns base j: largest r with offs[r] <= j
    while (hi - lo > 1) {
      const uint64_t mid = (lo + hi) >> 1;
      if (offs[mid] <= j) lo = mid; else hi = mid;
    }
    uint64_t r = lo;
    while (offs[r + 1] <= j && r + 1 < n_reads) ++r;
    const uint64_t r0 = offs[r];
    const uint64_t first = (j + 1 >= (uint64_t)k + r0) ? j + 1 - k : r0;
    uint32_t c = 0;
    for (uint64_t i = first; i <= j; ++i) c += hit_flags[i];
    cov[j] = (uint8_t)(c > 255 ? 255 : c);
  }
}

int rg_launch_coverage(rg_ctx* ctx, const uint8_t* hit_flags, const uint64_t* offs, uint64_t n_reads, uint64_t n_bytes,
                       int k, uint8_t* cov) {
  if (n_bytes == 0) return 0;
  unsigned grid = (unsigned)((n_bytes + 255) / 256);
  if (grid > rg_grid(ctx, 8)) grid = rg_grid(ctx, 8);
  {
    rg_launch_scope _ls(ctx, "coverage_kernel");
    coverage_kernel<<<grid, 256, 0, ctx->stream>>>(hit_flags, offs, n_reads, n_bytes, k, cov);
  }
  RG_CUDA_OK(cudaGetLastError());
  return 0;
}

__global__ void keep_kernel(const int32_t* __restrict__ h1, const int32_t* __restrict__ h2, uint64_t n, int thresh,
                            uint8_t* keep) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    keep[i] = (h1[i] >= thresh || h2[i] >= thresh) ? 1 : 0;   // RUFUS.Filter.cpp:222,262
}

int rg_launch_keep(rg_ctx* ctx, const int32_t* h1, const int32_t* h2, uint64_t n, int thresh, uint8_t* keep) {
  if (n == 0) return 0;
  unsigned grid = (unsigned)((n + 255) / 256);
  if (grid > rg_grid(ctx, 8)) grid = rg_grid(ctx, 8);
  {
    rg_launch_scope _ls(ctx, "keep_kernel");
    keep_kernel<<<grid, 256, 0, ctx->stream>>>(h1, h2, n, thresh, keep);
  }
  RG_CUDA_OK(cudaGetLastError());
  return 0;
}
