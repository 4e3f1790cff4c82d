// rg_ops.cu — streaming scans over a counted table: histogram, dump, query,
// subtraction. All are one pass over the SoA table with coalesced 16-byte loads;
// probes into the other tables are random but hit the same hash geometry.
#include "rg_internal.cuh"

// ---------------------------------------------------------------------------
// histo (histo_main.cc:34-44): rows base..ceil, count<base -> row 0, count>ceil -> last row
// ---------------------------------------------------------------------------
constexpr int HISTO_SMEM_BINS = 10240;   // covers the reference's -l 1 -h 10000 (10002 rows)

__global__ void __launch_bounds__(512)
histo_kernel(const uint32_t* __restrict__ counts, uint64_t n_slots, uint32_t min_count, uint64_t base, uint64_t nb,
             unsigned long long* __restrict__ out) {
  __shared__ unsigned int sh[HISTO_SMEM_BINS];
  const bool use_smem = nb <= HISTO_SMEM_BINS;
  if (use_smem) {
    for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) sh[i] = 0;
    __syncthreads();
  }
  const uint64_t ceil = base + nb - 1;
  const uint4* c4 = reinterpret_cast<const uint4*>(counts);
  const uint64_t n4 = n_slots / 4;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint4 v = __ldg(c4 + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint64_t c = w[j];
      if (c == 0 || c < min_count) continue;
      uint64_t row = c < base ? 0 : (c > ceil ? nb - 1 : c - base);
      if (use_smem) atomicAdd(&sh[row], 1u);
      else atomicAdd(&out[row], 1ull);
    }
  }
  if (use_smem) {
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x)
      if (sh[i]) atomicAdd(&out[i], (unsigned long long)sh[i]);
  }
}

int rg_launch_histo(rg_table* t, uint64_t min_count, uint64_t base, uint64_t nb, uint64_t* d_out) {
  rg_ctx* ctx = t->ctx;
  const uint64_t n_slots = rg_table_slots(t);
  RG_CUDA_OK(cudaMemsetAsync(d_out, 0, nb * sizeof(uint64_t), ctx->stream));
  {
    rg_launch_scope _ls(ctx, "histo_kernel");
    histo_kernel<<<rg_grid(ctx, 2), 512, 0, ctx->stream>>>(t->counts, n_slots, (uint32_t)min_count, base, nb,
                                                           reinterpret_cast<unsigned long long*>(d_out));
  }
  RG_CUDA_OK(cudaGetLastError());
  return 0;
}

__global__ void __launch_bounds__(512)
stats_kernel(const uint32_t* __restrict__ counts, uint64_t n_slots, unsigned long long* out2) {
  unsigned long long d = 0, occ = 0;
  const uint4* c4 = reinterpret_cast<const uint4*>(counts);
  const uint64_t n4 = n_slots / 4;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint4 v = __ldg(c4 + i);
    d += (v.x != 0) + (v.y != 0) + (v.z != 0) + (v.w != 0);
    occ += (unsigned long long)v.x + v.y + v.z + v.w;
  }
  for (int o = 16; o > 0; o >>= 1) {
    d += __shfl_down_sync(0xFFFFFFFFu, d, o);
    occ += __shfl_down_sync(0xFFFFFFFFu, occ, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&out2[0], d);
    atomicAdd(&out2[1], occ);
  }
}

int rg_launch_stats(rg_table* t, uint64_t* d_out2) {
  rg_ctx* ctx = t->ctx;
  const uint64_t n_slots = rg_table_slots(t);
  RG_CUDA_OK(cudaMemsetAsync(d_out2, 0, 2 * sizeof(uint64_t), ctx->stream));
  {
    rg_launch_scope _ls(ctx, "stats_kernel");
    stats_kernel<<<rg_grid(ctx, 2), 512, 0, ctx->stream>>>(t->counts, n_slots,
                                                           reinterpret_cast<unsigned long long*>(d_out2));
  }
  RG_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// warp-aggregated compaction of selected slots
// ---------------------------------------------------------------------------
__device__ __forceinline__ void emit_compact(bool sel, uint64_t key, uint32_t cnt, uint64_t* out_keys,
                                             uint32_t* out_counts, uint64_t cap, unsigned long long* cursor) {
  const unsigned m = __ballot_sync(0xFFFFFFFFu, sel);
  if (m == 0) return;
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(m) - 1;
  unsigned long long base = 0;
  if (lane == leader) base = atomicAdd(cursor, (unsigned long long)__popc(m));
  base = __shfl_sync(0xFFFFFFFFu, base, leader);
  if (sel) {
    const unsigned long long at = base + __popc(m & ((1u << lane) - 1));
    if (at < cap) {
      out_keys[at] = key;
      out_counts[at] = cnt;
    }
  }
}

__global__ void __launch_bounds__(256)
dump_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ counts, uint64_t n_slots, uint64_t lo,
            uint64_t hi, uint32_t slice_mod, uint32_t slice_rem, uint64_t* out_keys, uint32_t* out_counts, uint64_t cap,
            unsigned long long* cursor) {
  const uint64_t n_round = (n_slots + 31) & ~31ull;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_round;
       i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t c = 0;
    uint64_t key = 0;
    if (i < n_slots) c = __ldg(counts + i);
    bool sel = c != 0 && c >= lo && c <= hi;
    if (sel) {
      key = __ldg(keys + i);
      // a slice of the key space (a fixed function of the key, whatever the table geometry): sliced parity of
      // tables too large to check whole
      if (slice_mod) sel = rg_hash32b(key) % slice_mod == slice_rem;
    }
    emit_compact(sel, key, c, out_keys, out_counts, cap, cursor);
  }
}

int rg_launch_dump(rg_table* t, uint64_t min_count, uint64_t max_count, uint64_t* d_keys, uint32_t* d_counts,
                   uint64_t cap, uint64_t* d_cursor, uint32_t slice_mod, uint32_t slice_rem) {
  rg_ctx* ctx = t->ctx;
  const uint64_t n_slots = rg_table_slots(t);
  RG_CUDA_OK(cudaMemsetAsync(d_cursor, 0, sizeof(uint64_t), ctx->stream));
  {
    rg_launch_scope _ls(ctx, "dump_kernel");
    dump_kernel<<<rg_grid(ctx, 8), 256, 0, ctx->stream>>>(t->keys, t->counts, n_slots, min_count, max_count, slice_mod,
                                                          slice_rem, d_keys, d_counts, cap,
                                                          reinterpret_cast<unsigned long long*>(d_cursor));
  }
  RG_CUDA_OK(cudaGetLastError());
  return 0;
}

__global__ void __launch_bounds__(256)
query_kernel(TableView tv, const uint64_t* __restrict__ keys, uint64_t n, uint32_t min_count, uint32_t* out) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t c = rg_table_get(tv, keys[i]);
    out[i] = c >= min_count ? c : 0;
  }
}

int rg_launch_query(rg_table* t, const uint64_t* d_keys, uint64_t n, uint64_t min_count, uint32_t* d_out) {
  if (n == 0) return 0;
  rg_ctx* ctx = t->ctx;
  unsigned grid = rg_grid(ctx, 8);
  const uint64_t need = (n + 255) / 256;
  if (need < grid) grid = (unsigned)need;
  {
    rg_launch_scope _ls(ctx, "query_kernel");
    query_kernel<<<grid, 256, 0, ctx->stream>>>(rg_view(t), d_keys, n, (uint32_t)min_count, d_out);
  }
  RG_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// subtraction (merge_files.cc:86-154 fused with CheckJellyHashList.sh:12):
// stream the subject table; a slot is a candidate if lo <= count <= hi; it
// survives if every other table holds the key with count < that table's -L
// (i.e. the key would be absent from its .Jhash). Survivors are compacted.
// ---------------------------------------------------------------------------
// count of `key` (found in slot i of the subject) in another table. A table with the subject's geometry
// holds the key in the same tile (no minimizer to recompute) unless the subject keeps it in its overflow region.
__device__ __forceinline__ uint32_t other_count(const OtherTables& others, int o, uint64_t key, uint64_t i,
                                                uint64_t subj_main_slots) {
  const TableView& tv = others.tv[o];
  if (((others.same_geo >> o) & 1u) && i < subj_main_slots) return rg_table_get_at(tv, (uint32_t)(i >> RG_TILE_BITS), key);
  return rg_table_get(tv, key);
}

__global__ void __launch_bounds__(256)
subtract_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ counts, uint64_t n_slots,
                uint64_t main_slots, uint64_t lo, uint64_t hi, OtherTables others, uint64_t* out_keys,
                uint32_t* out_counts, uint64_t cap, unsigned long long* cursor) {
  const uint64_t n_round = (n_slots + 31) & ~31ull;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_round;
       i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t c = 0;
    uint64_t key = 0;
    if (i < n_slots) c = __ldg(counts + i);
    bool sel = c != 0 && c >= lo && c <= hi;
    if (sel) {
      key = __ldg(keys + i);
      for (int o = 0; o < others.n && sel; ++o) {
        const uint32_t oc = other_count(others, o, key, i, main_slots);
        if (oc != 0 && oc >= others.min_count[o]) sel = false;
      }
    }
    emit_compact(sel, key, c, out_keys, out_counts, cap, cursor);
  }
}

// ---------------------------------------------------------------------------
// RUFUS.Build (src/RUFUS.Build.cpp:224-283, the legacy selection): subject entries with
// lo <= c_S <= hi whose SUMMED count over the control tables is <= max_ctl_sum. Emits the key, the
// subject count and the control sum (out_sums, parallel to out_counts).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
build_select_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ counts, uint64_t n_slots,
                    uint64_t main_slots, uint64_t lo, uint64_t hi, OtherTables others, uint64_t max_ctl_sum, uint64_t* out_keys,
                    uint32_t* out_counts, uint32_t* out_sums, uint64_t cap, unsigned long long* cursor) {
  const uint64_t n_round = (n_slots + 31) & ~31ull;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_round;
       i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t c = 0;
    uint64_t key = 0, sum = 0;
    if (i < n_slots) c = __ldg(counts + i);
    bool sel = c != 0 && c >= lo && c <= hi;
    if (sel) {
      key = __ldg(keys + i);
      for (int o = 0; o < others.n; ++o) sum += other_count(others, o, key, i, main_slots);
      sel = sum <= max_ctl_sum;
    }
    const unsigned m = __ballot_sync(0xFFFFFFFFu, sel);
    if (m == 0) continue;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(cursor, (unsigned long long)__popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    if (sel) {
      const unsigned long long at = base + __popc(m & ((1u << lane) - 1));
      if (at < cap) {
        out_keys[at] = key;
        out_counts[at] = c;
        out_sums[at] = sum > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)sum;
      }
    }
  }
}

int rg_launch_build_select(rg_table* subj, uint64_t lo, uint64_t hi, const OtherTables& others, uint64_t max_ctl_sum,
                           uint64_t* d_keys, uint32_t* d_counts, uint32_t* d_sums, uint64_t cap, uint64_t* d_cursor) {
  rg_ctx* ctx = subj->ctx;
  const uint64_t n_slots = rg_table_slots(subj);
  RG_CUDA_OK(cudaMemsetAsync(d_cursor, 0, sizeof(uint64_t), ctx->stream));
  {
    rg_launch_scope _ls(ctx, "build_select_kernel");
    build_select_kernel<<<rg_grid(ctx, 8), 256, 0, ctx->stream>>>(subj->keys, subj->counts, n_slots,
                                                                  (uint64_t)subj->n_tiles << RG_TILE_BITS, lo, hi, others,
                                                                  max_ctl_sum, d_keys, d_counts, d_sums, cap,
                                                                  reinterpret_cast<unsigned long long*>(d_cursor));
  }
  RG_CUDA_OK(cudaGetLastError());
  return 0;
}

int rg_launch_subtract(rg_table* subj, uint64_t lo, uint64_t hi, const OtherTables& others, uint64_t* d_keys,
                       uint32_t* d_counts, uint64_t cap, uint64_t* d_cursor) {
  rg_ctx* ctx = subj->ctx;
  const uint64_t n_slots = rg_table_slots(subj);
  RG_CUDA_OK(cudaMemsetAsync(d_cursor, 0, sizeof(uint64_t), ctx->stream));
  {
    rg_launch_scope _ls(ctx, "subtract_kernel");
    subtract_kernel<<<rg_grid(ctx, 8), 256, 0, ctx->stream>>>(subj->keys, subj->counts, n_slots,
                                                              (uint64_t)subj->n_tiles << RG_TILE_BITS, lo, hi, others, d_keys,
                                                              d_counts, cap,
                                                              reinterpret_cast<unsigned long long*>(d_cursor));
  }
  RG_CUDA_OK(cudaGetLastError());
  return 0;
}
