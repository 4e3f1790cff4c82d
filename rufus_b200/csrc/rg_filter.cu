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

constexpr int FCHUNK = 4096;
constexpr int FSTAGE = FCHUNK + 64;
constexpr int FCODE_WORDS = FSTAGE / 16;
constexpr int FBIT_WORDS = FSTAGE / 32;
constexpr int FILTER_THREADS = 256;

__host__ __device__ __forceinline__ uint32_t rg_set_hash(uint64_t v) { return (uint32_t)(rg_mix64(v) >> 20); }

// ---- set construction -----------------------------------------------------
__global__ void set_build_kernel(const uint64_t* __restrict__ vals, uint64_t n, uint64_t* table, uint32_t slots,
                                 uint32_t* bitmap, uint32_t bitmap_bits, unsigned long long* n_distinct) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t v = vals[i];
    if (v == RG_EMPTY) continue;   // handled by has_all_ones on the host
    const uint32_t h = rg_set_hash(v);
    atomicOr(bitmap + ((h & (bitmap_bits - 1)) >> 5), 1u << (h & 31));
    uint32_t slot = h & (slots - 1);
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
struct FilterSmem {
  uint32_t codes[FCODE_WORDS + 2];   // little-endian: base i of a word at bits (2i, 2i+1), rufus codes
  uint32_t bad[FBIT_WORDS + 2];      // little-endian flag bits
};

__device__ __forceinline__ void fconvert16(const uint4 s, const uint4 q, int minq, uint32_t& codes, uint32_t& bad) {
  const uint32_t sw[4] = {s.x, s.y, s.z, s.w};
  const uint32_t qw[4] = {q.x, q.y, q.z, q.w};
  codes = 0;
  bad = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t c = (sw[i] >> (8 * j)) & 0xFFu;
      const int qv = (int)(signed char)((qw[i] >> (8 * j)) & 0xFFu) - 33;
      // Util::HashToLong: C -> bit(2i+1), G -> bit(2i), T -> both, anything else 0
      const uint32_t code = (c == 'C') ? 2u : (c == 'G') ? 1u : (c == 'T') ? 3u : 0u;
      const int idx = i * 4 + j;
      codes |= code << (2 * idx);
      bad |= (uint32_t)((qv < minq) || (c == 'N')) << idx;
    }
  }
}

__device__ __forceinline__ uint4 load16_guarded(const char* base, uint64_t p, uint64_t n, unsigned char fill) {
  if (p + 16 <= n) return __ldg(reinterpret_cast<const uint4*>(base + p));
  unsigned char tmp[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) tmp[j] = (p + j < n) ? (unsigned char)base[p + j] : fill;
  return *reinterpret_cast<uint4*>(tmp);
}

__global__ void __launch_bounds__(FILTER_THREADS)
filter_hits_kernel(const char* __restrict__ seq, const char* __restrict__ qual, const uint64_t* __restrict__ offs,
                   uint32_t read_len, uint64_t n_reads, uint64_t n_bytes, int k, int minq, int skip_last,
                   const uint64_t* __restrict__ table, uint32_t slots, const uint32_t* __restrict__ g_bitmap,
                   uint32_t bitmap_bits, int has_all_ones, int32_t* __restrict__ hits) {
  extern __shared__ __align__(16) uint32_t s_bitmap[];   // bitmap_bits-bit prefilter of the set
  __shared__ FilterSmem sm;
  {
    const uint4* src = reinterpret_cast<const uint4*>(g_bitmap);
    uint4* dst = reinterpret_cast<uint4*>(s_bitmap);
    for (uint32_t i = threadIdx.x; i < bitmap_bits / 128; i += blockDim.x) dst[i] = __ldg(src + i);
  }
  uint16_t* bad16 = reinterpret_cast<uint16_t*>(sm.bad);
  const uint64_t kmask = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1);
  const uint64_t n_chunks = (n_bytes + FCHUNK - 1) / FCHUNK;
  for (uint64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const uint64_t c0 = c * FCHUNK;
    __syncthreads();
    for (int i = threadIdx.x; i < FCODE_WORDS; i += blockDim.x) {
      const uint64_t p = c0 + (uint64_t)i * 16;
      uint32_t cw = 0, bw = 0xFFFFu;
      if (p < n_bytes) {
        const uint4 s4 = load16_guarded(seq, p, n_bytes, 'N');
        const uint4 q4 = load16_guarded(qual, p, n_bytes, 0);
        fconvert16(s4, q4, minq, cw, bw);
      }
      sm.codes[i] = cw;
      bad16[i] = (uint16_t)bw;
    }
    if (threadIdx.x < 2) {
      sm.codes[FCODE_WORDS + threadIdx.x] = 0;
      sm.bad[FBIT_WORDS + threadIdx.x] = 0xFFFFFFFFu;
    }
    __syncthreads();
    for (int p = threadIdx.x; p < FCHUNK; p += blockDim.x) {
      const uint64_t gp = c0 + p;
      if (gp + k > n_bytes) break;
      const int bj = p >> 5, bs = p & 31;
      const uint64_t badw = (((uint64_t)sm.bad[bj + 1] << 32) | sm.bad[bj]) >> bs;
      if (badw & ((k == 64) ? ~0ull : ((1ull << k) - 1))) continue;
      const int j = p >> 4, s = (p & 15) * 2;
      uint64_t x = ((uint64_t)sm.codes[j + 1] << 32) | sm.codes[j];
      if (s) x = (x >> s) | ((uint64_t)sm.codes[j + 2] << (64 - s));
      const uint64_t v = x & kmask;
      bool hit;
      if (v == RG_EMPTY) {
        hit = has_all_ones != 0;
      } else {
        const uint32_t h = rg_set_hash(v);
        const uint32_t bit = h & (bitmap_bits - 1);
        if (!((s_bitmap[bit >> 5] >> (bit & 31)) & 1u)) continue;
        hit = false;
        uint32_t slot = h & (slots - 1);
        for (;;) {
          const uint64_t cur = __ldg(table + slot);
          if (cur == v) { hit = true; break; }
          if (cur == RG_EMPTY) break;
          slot = (slot + 1) & (slots - 1);
        }
      }
      if (!hit) continue;
      // locate the read that owns base gp and apply the boundary rules
      uint64_t r, r0, r1;
      if (offs != nullptr) {
        uint64_t lo = 0, hi = n_reads;   // largest r with offs[r] <= gp
        while (hi - lo > 1) {
          const uint64_t mid = (lo + hi) >> 1;
          if (__ldg(offs + mid) <= gp) lo = mid; else hi = mid;
        }
        r = lo;
        r0 = __ldg(offs + r);
        r1 = __ldg(offs + r + 1);
        // empty reads share their offset with the next read: skip forward to the owner
        while (r1 <= gp && r + 1 < n_reads) { ++r; r0 = r1; r1 = __ldg(offs + r + 1); }
      } else {
        r = gp / read_len;
        r0 = r * read_len;
        r1 = r0 + read_len;
      }
      (void)r0;
      const uint64_t end_excl = skip_last ? (r1 > 0 ? r1 - 1 : 0) : r1;   // window must end before this
      if (gp + k <= end_excl) atomicAdd(hits + r, 1);
    }
  }
}

int rg_launch_filter_hits(rg_set* s, const char* seq, const char* qual, const uint64_t* offs, uint32_t read_len,
                          uint64_t n_reads, uint64_t n_bytes, int minq, int skip_last, int32_t* hits) {
  rg_ctx* ctx = s->ctx;
  RG_CUDA_OK(cudaMemsetAsync(hits, 0, n_reads * sizeof(int32_t), ctx->stream));
  if (n_bytes == 0 || n_reads == 0) return 0;
  const size_t smem = s->bitmap_bits / 8;
  static bool attr_set = false;
  if (!attr_set) {
    RG_CUDA_OK(cudaFuncSetAttribute(filter_hits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(RG_SET_BITMAP_BITS / 8)));
    attr_set = true;
  }
  const uint64_t n_chunks = (n_bytes + FCHUNK - 1) / FCHUNK;
  unsigned grid = rg_grid(ctx, smem >= 65536 ? 3 : (smem >= 32768 ? 6 : 8));
  if (n_chunks < grid) grid = (unsigned)n_chunks;
  {
    rg_launch_scope _ls(ctx, "filter_hits_kernel");
    filter_hits_kernel<<<grid, FILTER_THREADS, smem, ctx->stream>>>(seq, qual, offs, read_len, n_reads, n_bytes, s->k,
                                                                    minq, skip_last, s->table, s->slots, s->bitmap,
                                                                    s->bitmap_bits, s->has_all_ones, hits);
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
