// rg_overlap.cu — the device side of the greedy overlap assembly (src/Overlap.cpp).
// The greedy loop itself is serial (step i depends on the merges of the steps before it,
// Overlap.cpp:935-1126) and stays on the host (tools/overlap.cpp); what it spends its time in runs here:
//
//   O2  seed index   RebuildHashTable (Overlap.cpp:34-76): every SearchHash-mer of every read and
//                    its reverse complement -> read index. Here: one kernel hashes all seeds
//                    (Util::HashToLong packing), one radix sort orders (seed, read) pairs.
//   O3  seed hits    the lookups of PrepairSearchList (:78-103): for each query read, the read
//                    indices > Ai sharing a seed, one entry per shared occurrence (binary search of
//                    the sorted seeds); the counting / ranking of the few survivors is host work.
//   O4  Align3       (:169-360) the ungapped placement search of A against every candidate: one CTA
//                    per candidate, one placement per thread-iteration, block reduction.
//
// Align3's per-placement early exit `(k - score) > MM` fires iff the placement ends with
// (length-1 - score) > MM (the left side never decreases along k), and its running
// `LbestScore < score` chain ends on the FIRST placement, in loop order, that reaches the maximum
// score among the accepted ones; the `break`s only skip placements that cannot beat it. So each
// candidate reduces to max over (score, earliest order) — no sequential scan is needed.
#include <cub/cub.cuh>

#include <algorithm>
#include <string>
#include <vector>

#include "rg_internal.cuh"

struct rg_overlap {
  rg_ctx* ctx = nullptr;
  int k = 0;                       // SearchHash
  uint64_t n_reads = 0;
  // current sequences: an arena of bytes + (offset, length) per read
  char* arena = nullptr;
  uint64_t arena_cap = 0, arena_used = 0;
  uint64_t* off = nullptr;         // device [n_reads]
  uint32_t* len = nullptr;         // device [n_reads]
  std::vector<uint64_t> h_off;
  std::vector<uint32_t> h_len;
  // seed index, sorted by seed
  uint64_t* seeds = nullptr;
  uint32_t* owners = nullptr;
  uint64_t n_seeds = 0;
  // scratch
  char* d_query = nullptr;
  uint64_t d_query_cap = 0;
  void* d_misc = nullptr;
  uint64_t d_misc_cap = 0;
  // pinned staging of one align call: [A | rc(A) | candidates] up, results down
  char* h_stage = nullptr;
  uint64_t h_stage_cap = 0;
};

#define OV_INVALID 0xFFFFFFFFFFFFFFFFull

// Util::HashToLong code of a base (Util.cpp:51-84); 4 = not A/C/G/T
__device__ __forceinline__ uint32_t ov_code(char c) {
  return c == 'A' ? 0u : c == 'G' ? 1u : c == 'C' ? 2u : c == 'T' ? 3u : 4u;
}

// seed starting at s[0] (k bases) and the seed of its reverse complement; OV_INVALID if a base is not ACGT
__device__ __forceinline__ void ov_seed_pair(const char* s, int k, uint64_t& fwd, uint64_t& rc) {
  uint64_t f = 0, r = 0;
  bool ok = true;
  for (int i = 0; i < k; ++i) {
    const uint32_t c = ov_code(s[i]);
    ok = ok && c < 4;
    f |= (uint64_t)(c & 3u) << (2 * i);
    // complement in this coding: A0<->T3, G1<->C2, i.e. 3 - c; the complement of base i sits at position k-1-i
    r |= (uint64_t)((3u - c) & 3u) << (2 * (k - 1 - i));
  }
  fwd = ok ? f : OV_INVALID;
  rc = ok ? r : OV_INVALID;
}

// one thread per (read, seed position): two (seed, read) pairs each
__global__ void ov_seed_kernel(const char* __restrict__ arena, const uint64_t* __restrict__ off,
                               const uint32_t* __restrict__ len, const uint64_t* __restrict__ seed_base, uint64_t first,
                               uint64_t n_reads, int k, uint64_t* __restrict__ seeds, uint32_t* __restrict__ owners) {
  for (uint64_t r = first + blockIdx.x; r < n_reads; r += gridDim.x) {
    const uint32_t L = len[r];
    const uint32_t n = L > (uint32_t)k ? L - k : 0;       // `j < size - SearchHash` (:52-53): the last seed is left out
    const char* s = arena + off[r];
    const uint64_t base = seed_base[r - first];
    for (uint32_t j = threadIdx.x; j < n; j += blockDim.x) {
      uint64_t f, rc;
      ov_seed_pair(s + j, k, f, rc);
      seeds[base + 2 * j] = f;
      owners[base + 2 * j] = (uint32_t)r;
      seeds[base + 2 * j + 1] = rc;
      owners[base + 2 * j + 1] = (uint32_t)r;
    }
  }
}

__device__ __forceinline__ uint64_t ov_lower_bound(const uint64_t* __restrict__ a, uint64_t n, uint64_t key) {
  uint64_t lo = 0, hi = n;
  while (lo < hi) {
    const uint64_t mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// one CTA per query. FILL = false: count the hits; true: write them at hit_offs[q]
template <bool FILL>
__global__ void ov_match_kernel(const char* __restrict__ qseq, const uint64_t* __restrict__ qoffs,
                                const uint64_t* __restrict__ q_ai, int k, const uint64_t* __restrict__ seeds,
                                const uint32_t* __restrict__ owners, uint64_t n_seeds,
                                unsigned long long* __restrict__ counts, const uint64_t* __restrict__ hit_offs,
                                uint32_t* __restrict__ hits) {
  const uint64_t q = blockIdx.x;
  const char* s = qseq + qoffs[q];
  const uint64_t L = qoffs[q + 1] - qoffs[q];
  const uint64_t n = L > (uint64_t)k ? L - k : 0;         // `i < Alength - SearchHash` (:84)
  const uint64_t ai = q_ai[q];
  __shared__ unsigned long long cursor;
  if (threadIdx.x == 0) cursor = 0;
  __syncthreads();
  for (uint64_t j = threadIdx.x; j < n; j += blockDim.x) {
    uint64_t f, rc;
    ov_seed_pair(s + j, k, f, rc);
    if (f == OV_INVALID) continue;
    for (uint64_t p = ov_lower_bound(seeds, n_seeds, f); p < n_seeds && seeds[p] == f; ++p) {
      const uint32_t o = owners[p];
      if ((uint64_t)o > ai) {                             // `holder > Ai` (:98)
        const unsigned long long at = atomicAdd(&cursor, 1ull);
        if (FILL) hits[hit_offs[q] + at] = o;
      }
    }
  }
  __syncthreads();
  if (!FILL && threadIdx.x == 0) counts[q] = cursor;
}

struct OvBest {
  int score, order;
};
__device__ __forceinline__ void ov_better(OvBest& b, int score, int order) {
  if (score > b.score || (score == b.score && order < b.order)) {
    b.score = score;
    b.order = order;
  }
}

// matches of x[0..n) against y[0..n): equal and not 'N' (Align3's compare, :226-233,:270-276,:309-315)
__device__ __forceinline__ int ov_matches(const char* __restrict__ x, const char* __restrict__ y, int n) {
  int m = 0;
  for (int t = 0; t < n; ++t) m += (x[t] == y[t]) & (x[t] != 'N');
  return m;
}

constexpr int OV_THREADS = 128;

__global__ void __launch_bounds__(OV_THREADS)
ov_align_kernel(const char* __restrict__ arena, const uint64_t* __restrict__ off, const uint32_t* __restrict__ len,
                const char* __restrict__ A0, const char* __restrict__ A1, uint32_t n_first, int alen,
                const uint32_t* __restrict__ cand, float minp, int min_overlap, rg_align_result* __restrict__ out) {
  // candidates [0, n_first) are aligned against A0 (the read), the rest against A1 (its reverse complement)
  const char* __restrict__ A = blockIdx.x < n_first ? A0 : A1;
  const uint32_t j = cand[blockIdx.x];
  const char* B = arena + off[j];
  const int blen = (int)len[j];
  const bool a_smaller = !(blen > alen);
  const int window = a_smaller ? blen : alen, longest = a_smaller ? alen : blen;
  // int MM = window - (window * minPercent): float arithmetic without contraction, truncated
  const int MM = (int)__fsub_rn((float)window, __fmul_rn((float)window, minp));
  const int n1 = longest - window + 1;
  const int n23 = window - 1 >= min_overlap ? window - min_overlap : 0;   // i = window-1 .. min_overlap
  OvBest b1 = {-1, 0x7FFFFFFF}, ball = {-1, 0x7FFFFFFF};
  for (int t = threadIdx.x; t < n1 + 2 * n23; t += OV_THREADS) {
    int score, length, order = t;
    bool strict = false;
    if (t < n1) {                                     // (i) the shorter inside the longer, shift t
      const int i = t;
      score = a_smaller ? ov_matches(A + i, B, window) : ov_matches(A, B + i, window);
      length = window;
    } else if (t < n1 + n23) {                        // (ii) suffix of A on prefix of B, i descending
      const int i = window - 1 - (t - n1);
      score = ov_matches(A + (alen - i - 1), B, i + 1);
      length = i + 1;
    } else {                                          // (iii) suffix of B on prefix of A, i descending
      const int i = window - 1 - (t - n1 - n23);
      score = ov_matches(B + (blen - i - 1), A, i + 1);
      length = i + 1;
      strict = true;                                  // `percent > minPercent` (:331)
    }
    if (length - 1 - score > MM) continue;            // the early exit: score = -1
    const float percent = __fdiv_rn((float)score, (float)length);
    if (strict ? !(percent > minp) : !(percent >= minp)) continue;
    if (t < n1) ov_better(b1, score, order);
    ov_better(ball, score, order);
  }
  __shared__ OvBest s1[OV_THREADS], sall[OV_THREADS];
  s1[threadIdx.x] = b1;
  sall[threadIdx.x] = ball;
  __syncthreads();
  for (int o = OV_THREADS / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      ov_better(s1[threadIdx.x], s1[threadIdx.x + o].score, s1[threadIdx.x + o].order);
      ov_better(sall[threadIdx.x], sall[threadIdx.x + o].score, sall[threadIdx.x + o].order);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    auto offset_of = [&](int order) -> int {
      if (order < n1) return a_smaller ? -order : order;                  // :247-251
      if (order < n1 + n23) return (window - 1 - (order - n1)) - alen + 1;  // :288
      return blen - (window - 1 - (order - n1 - n23)) - 1;                // :336
    };
    rg_align_result r;
    r.score1 = s1[0].score;
    r.overlap1 = s1[0].score >= 0 ? offset_of(s1[0].order) : 0;
    r.perfect1 = s1[0].score == window ? 1 : 0;                           // :253-257
    r.score_all = sall[0].score;
    r.overlap_all = sall[0].score >= 0 ? offset_of(sall[0].order) : 0;
    out[blockIdx.x] = r;
  }
}

static int ov_reserve(void** p, uint64_t* cap, uint64_t bytes, rg_ctx* ctx) {
  if (*cap >= bytes) return 0;
  if (*p) {
    RG_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    cudaFree(*p);
    *p = nullptr;
    *cap = 0;
  }
  const uint64_t want = (bytes + bytes / 2 + 4096) & ~255ull;
  RG_CUDA_OK(cudaMalloc(p, want));
  *cap = want;
  return 0;
}

extern "C" {

rg_overlap* rg_overlap_create(rg_ctx* ctx, int search_hash) {
  if (!ctx || search_hash < 1 || search_hash > 31) {
    rg_set_error("rg_overlap_create: SearchHash must be 1..31");
    return nullptr;
  }
  rg_overlap* ov = new rg_overlap();
  ov->ctx = ctx;
  ov->k = search_hash;
  return ov;
}

void rg_overlap_free(rg_overlap* ov) {
  if (!ov) return;
  cudaStreamSynchronize(ov->ctx->stream);
  cudaFree(ov->arena);
  cudaFree(ov->off);
  cudaFree(ov->len);
  cudaFree(ov->seeds);
  cudaFree(ov->owners);
  cudaFree(ov->d_query);
  cudaFree(ov->d_misc);
  cudaFreeHost(ov->h_stage);
  delete ov;
}

// O2: load the reads (concatenated, offs[n_reads+1]) and index the seeds of reads >= first
int rg_overlap_build(rg_overlap* ov, const char* seq, const uint64_t* offs, uint64_t n_reads, uint64_t first) {
  if (!ov || (n_reads && (!seq || !offs)) || first > n_reads || n_reads >= 0xFFFFFFFFull) return RG_ERR_ARG;
  rg_ctx* ctx = ov->ctx;
  RG_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  cudaFree(ov->arena); cudaFree(ov->off); cudaFree(ov->len); cudaFree(ov->seeds); cudaFree(ov->owners);
  ov->arena = nullptr; ov->off = nullptr; ov->len = nullptr; ov->seeds = nullptr; ov->owners = nullptr;
  ov->n_reads = n_reads;
  ov->n_seeds = 0;
  if (n_reads == 0) return 0;
  const uint64_t total = offs[n_reads] - offs[0];
  ov->arena_cap = total * 3 + (1u << 20);              // room for the merged contigs that replace reads
  ov->arena_used = total;
  RG_CUDA_OK(cudaMalloc(&ov->arena, ov->arena_cap));
  RG_CUDA_OK(cudaMalloc(&ov->off, n_reads * 8));
  RG_CUDA_OK(cudaMalloc(&ov->len, n_reads * 4));
  ov->h_off.resize(n_reads);
  ov->h_len.resize(n_reads);
  std::vector<uint64_t> seed_base(n_reads - first + 1, 0);
  for (uint64_t r = 0; r < n_reads; ++r) {
    ov->h_off[r] = offs[r] - offs[0];
    ov->h_len[r] = (uint32_t)(offs[r + 1] - offs[r]);
  }
  uint64_t ns = 0;
  for (uint64_t r = first; r < n_reads; ++r) {
    seed_base[r - first] = ns;
    ns += ov->h_len[r] > (uint32_t)ov->k ? 2ull * (ov->h_len[r] - ov->k) : 0;
  }
  RG_CUDA_OK(cudaMemcpyAsync(ov->arena, seq + offs[0], total, cudaMemcpyHostToDevice, ctx->stream));
  RG_CUDA_OK(cudaMemcpyAsync(ov->off, ov->h_off.data(), n_reads * 8, cudaMemcpyHostToDevice, ctx->stream));
  RG_CUDA_OK(cudaMemcpyAsync(ov->len, ov->h_len.data(), n_reads * 4, cudaMemcpyHostToDevice, ctx->stream));
  if (ns == 0) {
    RG_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    return 0;
  }
  if (ns >= (1ull << 31)) {
    rg_set_error("rg_overlap_build: too many seeds for one index");
    return RG_ERR_ARG;
  }
  uint64_t *d_base = nullptr, *k_in = nullptr;
  uint32_t* v_in = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  RG_CUDA_OK(cudaMalloc(&d_base, seed_base.size() * 8));
  RG_CUDA_OK(cudaMalloc(&k_in, ns * 8));
  RG_CUDA_OK(cudaMalloc(&v_in, ns * 4));
  RG_CUDA_OK(cudaMalloc(&ov->seeds, ns * 8));
  RG_CUDA_OK(cudaMalloc(&ov->owners, ns * 4));
  RG_CUDA_OK(cudaMemcpyAsync(d_base, seed_base.data(), seed_base.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  {
    rg_launch_scope _ls(ctx, "ov_seed_kernel");
    const unsigned grid = (unsigned)std::min<uint64_t>(n_reads - first, rg_grid(ctx, 16));
    ov_seed_kernel<<<grid, 128, 0, ctx->stream>>>(ov->arena, ov->off, ov->len, d_base, first, n_reads, ov->k, k_in, v_in);
  }
  // stable sort by seed: the owners of one seed stay in read order, like the reference's push_back order
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k_in, ov->seeds, v_in, ov->owners, (int64_t)ns, 0, 64, ctx->stream);
  RG_CUDA_OK(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16));
  ctx->launches++;
  RG_CUDA_OK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k_in, ov->seeds, v_in, ov->owners, (int64_t)ns, 0, 64,
                                             ctx->stream));
  RG_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_base); cudaFree(k_in); cudaFree(v_in); cudaFree(tmp);
  ov->n_seeds = ns;
  RG_CUDA_OK(cudaGetLastError());
  return 0;
}

// a merge replaced read idx by a longer contig (Overlap.cpp:1066-1070); the seed index is NOT updated,
// like the reference's (:1083,:1101 `if (found = false)`)
int rg_overlap_update(rg_overlap* ov, uint64_t idx, const char* seq, uint64_t len) {
  if (!ov || idx >= ov->n_reads || len >= 0xFFFFFFFFull) return RG_ERR_ARG;
  rg_ctx* ctx = ov->ctx;
  if (ov->arena_used + len > ov->arena_cap) {          // grow: compact the live sequences into a bigger arena
    RG_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    uint64_t live = len;
    for (uint64_t r = 0; r < ov->n_reads; ++r) live += ov->h_len[r];
    const uint64_t cap = live * 2 + (1u << 20);
    char* fresh = nullptr;
    RG_CUDA_OK(cudaMalloc(&fresh, cap));
    uint64_t at = 0;
    for (uint64_t r = 0; r < ov->n_reads; ++r) {
      if (r == idx) continue;
      RG_CUDA_OK(cudaMemcpyAsync(fresh + at, ov->arena + ov->h_off[r], ov->h_len[r], cudaMemcpyDeviceToDevice, ctx->stream));
      ov->h_off[r] = at;
      at += ov->h_len[r];
    }
    RG_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    cudaFree(ov->arena);
    ov->arena = fresh;
    ov->arena_cap = cap;
    ov->arena_used = at;
    RG_CUDA_OK(cudaMemcpyAsync(ov->off, ov->h_off.data(), ov->n_reads * 8, cudaMemcpyHostToDevice, ctx->stream));
  }
  ov->h_off[idx] = ov->arena_used;
  ov->h_len[idx] = (uint32_t)len;
  RG_CUDA_OK(cudaMemcpyAsync(ov->arena + ov->arena_used, seq, len, cudaMemcpyHostToDevice, ctx->stream));
  RG_CUDA_OK(cudaMemcpyAsync(ov->off + idx, &ov->h_off[idx], 8, cudaMemcpyHostToDevice, ctx->stream));
  RG_CUDA_OK(cudaMemcpyAsync(ov->len + idx, &ov->h_len[idx], 4, cudaMemcpyHostToDevice, ctx->stream));
  ov->arena_used += len;
  // (pageable sources are staged by the runtime before cudaMemcpyAsync returns: no synchronisation needed)
  return 0;
}

// O3 lookups: for every query, the owners (> q_ai[q]) of its seeds in the index, one entry per shared
// occurrence. hit_offs_host gets nq+1 offsets; returns the total, or RG_ERR_SMALL with hit_offs filled.
int64_t rg_overlap_seed_hits(rg_overlap* ov, const char* qseq, const uint64_t* qoffs, const uint64_t* q_ai, uint64_t nq,
                             uint32_t* hits_host, uint64_t cap, uint64_t* hit_offs_host) {
  if (!ov || !hit_offs_host || (nq && (!qseq || !qoffs || !q_ai))) return RG_ERR_ARG;
  rg_ctx* ctx = ov->ctx;
  hit_offs_host[0] = 0;
  if (nq == 0) return 0;
  if (ov->n_seeds == 0) {
    for (uint64_t q = 0; q <= nq; ++q) hit_offs_host[q] = 0;
    return 0;
  }
  const uint64_t qbytes = qoffs[nq] - qoffs[0];
  int rc = ov_reserve((void**)&ov->d_query, &ov->d_query_cap, qbytes + 16, ctx);
  if (rc == 0) rc = ov_reserve(&ov->d_misc, &ov->d_misc_cap, (nq + 1) * 8 * 4, ctx);
  if (rc) return rc;
  uint64_t* d_qoffs = static_cast<uint64_t*>(ov->d_misc);
  uint64_t* d_ai = d_qoffs + (nq + 1);
  unsigned long long* d_counts = reinterpret_cast<unsigned long long*>(d_ai + (nq + 1));
  uint64_t* d_hoffs = reinterpret_cast<uint64_t*>(d_counts + (nq + 1));
  std::vector<uint64_t> rel(nq + 1);
  for (uint64_t q = 0; q <= nq; ++q) rel[q] = qoffs[q] - qoffs[0];
  RG_CUDA_OK(cudaMemcpyAsync(ov->d_query, qseq + qoffs[0], qbytes, cudaMemcpyHostToDevice, ctx->stream));
  RG_CUDA_OK(cudaMemcpyAsync(d_qoffs, rel.data(), (nq + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
  RG_CUDA_OK(cudaMemcpyAsync(d_ai, q_ai, nq * 8, cudaMemcpyHostToDevice, ctx->stream));
  {
    rg_launch_scope _ls(ctx, "ov_match_kernel<count>");
    ov_match_kernel<false><<<(unsigned)nq, 128, 0, ctx->stream>>>(ov->d_query, d_qoffs, d_ai, ov->k, ov->seeds, ov->owners,
                                                                 ov->n_seeds, d_counts, nullptr, nullptr);
  }
  std::vector<uint64_t> counts(nq);
  RG_CUDA_OK(cudaMemcpyAsync(counts.data(), d_counts, nq * 8, cudaMemcpyDeviceToHost, ctx->stream));
  RG_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  uint64_t total = 0;
  for (uint64_t q = 0; q < nq; ++q) {
    hit_offs_host[q] = total;
    total += counts[q];
  }
  hit_offs_host[nq] = total;
  if (total > cap) {
    rg_set_error("rg_overlap_seed_hits: output buffer too small");
    return RG_ERR_SMALL;
  }
  if (total == 0) return 0;
  uint32_t* d_hits = nullptr;
  RG_CUDA_OK(cudaMalloc(&d_hits, total * 4));
  RG_CUDA_OK(cudaMemcpyAsync(d_hoffs, hit_offs_host, (nq + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
  {
    rg_launch_scope _ls(ctx, "ov_match_kernel<fill>");
    ov_match_kernel<true><<<(unsigned)nq, 128, 0, ctx->stream>>>(ov->d_query, d_qoffs, d_ai, ov->k, ov->seeds, ov->owners,
                                                                ov->n_seeds, d_counts, d_hoffs, d_hits);
  }
  cudaError_t e = cudaMemcpyAsync(hits_host, d_hits, total * 4, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d_hits);
  if (e != cudaSuccess) {
    rg_set_error(cudaGetErrorString(e));
    return RG_ERR_CUDA;
  }
  return (int64_t)total;
}

// O4: A (and optionally its reverse complement) against the CURRENT sequences of the candidate reads,
// one launch and one synchronisation for both strands
int rg_overlap_align_both(rg_overlap* ov, const char* a, const char* a_rev, uint64_t a_len, const uint32_t* cand,
                          uint64_t n_cand, const uint32_t* cand_rev, uint64_t n_cand_rev, float min_percent,
                          int min_overlap, rg_align_result* out_host, rg_align_result* out_rev_host) {
  const uint64_t n_all = n_cand + n_cand_rev;
  if (!ov || a_len >= 0x7FFFFFFFull || (n_cand && (!a || !cand || !out_host)) ||
      (n_cand_rev && (!a_rev || !cand_rev || !out_rev_host)))
    return RG_ERR_ARG;
  if (n_all == 0) return 0;
  rg_ctx* ctx = ov->ctx;
  for (uint64_t c = 0; c < n_all; ++c)
    if ((c < n_cand ? cand[c] : cand_rev[c - n_cand]) >= ov->n_reads) {
      rg_set_error("rg_overlap_align: candidate index out of range");
      return RG_ERR_ARG;
    }
  // one staged upload [A | rc(A) | candidates] and one staged download per call
  const uint64_t a_pad = (a_len + 63) & ~63ull;
  const uint64_t up_bytes = 2 * a_pad + ((n_all * 4 + 63) & ~63ull), down_bytes = n_all * sizeof(rg_align_result);
  int rc = ov_reserve(&ov->d_misc, &ov->d_misc_cap, up_bytes + down_bytes, ctx);
  if (rc) return rc;
  if (ov->h_stage_cap < up_bytes + down_bytes) {
    RG_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    cudaFreeHost(ov->h_stage);
    ov->h_stage = nullptr;
    ov->h_stage_cap = 0;
    const uint64_t want = 2 * (up_bytes + down_bytes) + 4096;
    RG_CUDA_OK(cudaMallocHost(&ov->h_stage, want));
    ov->h_stage_cap = want;
  }
  char* h_up = ov->h_stage;
  if (n_cand) memcpy(h_up, a, a_len);
  if (n_cand_rev) memcpy(h_up + a_pad, a_rev, a_len);
  if (n_cand) memcpy(h_up + 2 * a_pad, cand, n_cand * 4);
  if (n_cand_rev) memcpy(h_up + 2 * a_pad + n_cand * 4, cand_rev, n_cand_rev * 4);
  char* d_up = static_cast<char*>(ov->d_misc);
  rg_align_result* d_out = reinterpret_cast<rg_align_result*>(d_up + up_bytes);
  rg_align_result* h_out = reinterpret_cast<rg_align_result*>(ov->h_stage + up_bytes);
  RG_CUDA_OK(cudaMemcpyAsync(d_up, h_up, up_bytes, cudaMemcpyHostToDevice, ctx->stream));
  {
    rg_launch_scope _ls(ctx, "ov_align_kernel");
    ov_align_kernel<<<(unsigned)n_all, OV_THREADS, 0, ctx->stream>>>(
        ov->arena, ov->off, ov->len, d_up, d_up + a_pad, (uint32_t)n_cand, (int)a_len,
        reinterpret_cast<const uint32_t*>(d_up + 2 * a_pad), min_percent, min_overlap, d_out);
  }
  RG_CUDA_OK(cudaMemcpyAsync(h_out, d_out, down_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  RG_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  RG_CUDA_OK(cudaGetLastError());
  if (n_cand) memcpy(out_host, h_out, n_cand * sizeof(rg_align_result));
  if (n_cand_rev) memcpy(out_rev_host, h_out + n_cand, n_cand_rev * sizeof(rg_align_result));
  return 0;
}

int rg_overlap_align(rg_overlap* ov, const char* a, uint64_t a_len, const uint32_t* cand, uint64_t n_cand,
                     float min_percent, int min_overlap, rg_align_result* out_host) {
  return rg_overlap_align_both(ov, a, nullptr, a_len, cand, n_cand, nullptr, 0, min_percent, min_overlap, out_host,
                               nullptr);
}

}  // extern "C"
