// rg_synth.cu — "synth-trio v2": the counter-based synthetic read generator of SURVEY §8d, on the device
// and (same bytes) on the host. Not on the hot path: it builds the inputs of the large benchmark configs
// (BASELINE configs[2..4]) in HBM, where a host generator + upload would cost minutes of GPU-box time, and
// gives the CPU restatement the identical reads for sliced parity (tests/golden/make_big_digests.py).
//
// Everything is a pure function of (seed, coordinates) through one 64-bit mixing hash (integer only, so host
// and device agree bit for bit; the survey suggests Philox, any counter-based generator does):
//   genome      base(p) iid uniform; with `repeats` 5 % of the 1 KiB regions are copies of one of 100 consensus
//               families at 2 % divergence (high-count k-mers, exercises MaxHashDepth and heavy tiles)
//   haplotypes  genome + private SNVs 1/2000 per parental haplotype (1/1000 for the tumor/normal individual)
//   trio/quad   child = maternal hap 0 + paternal hap 1 + de-novo SNVs (max(20, G/1e5), on either haplotype);
//               sibling = maternal hap 1 + paternal hap 0
//   pair        tumor = normal + max(20, G/1e5)... somatic SNVs (on either haplotype), sequenced deeper by the caller
//   reads       150 bp pairs, insert U[300,400), mate 2 reverse-complemented, uniform start, fair-coin haplotype;
//               substitution error 0.2 %/base with quality '#', 1 % ',' (Q11), else 'J' (Q41); 'N' at 2e-4
#include "rg_internal.cuh"

__host__ __device__ __forceinline__ uint64_t sy_mix(uint64_t x) {
  x ^= x >> 30;
  x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27;
  x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}
__host__ __device__ __forceinline__ uint64_t sy_h(uint64_t seed, uint64_t tag, uint64_t a) {
  return sy_mix(sy_mix(seed + tag * 0x9E3779B97F4A7C15ull) + a);
}

enum : uint64_t { SY_GENOME = 1, SY_REGION = 2, SY_FAMILY = 3, SY_DIVERGE = 4, SY_HAP = 16, SY_DENOVO = 5, SY_READ = 64, SY_ERR = 128 };

__host__ __device__ __forceinline__ uint32_t sy_genome_base(const rg_synth_spec& sp, uint64_t p) {
  if (sp.repeats) {
    const uint64_t hr = sy_h(sp.seed, SY_REGION, p >> 10);
    if ((hr & 0xFFFFu) < 3277u) {                                   // 5 % of the regions
      const uint64_t fam = (hr >> 16) % 100u;
      uint32_t b = (uint32_t)sy_h(sp.seed, SY_FAMILY, fam * 1024u + (p & 1023u)) & 3u;
      const uint64_t hd = sy_h(sp.seed, SY_DIVERGE, p);
      if ((hd & 0xFFFFu) < 1311u) b = (b + 1u + (uint32_t)((hd >> 16) % 3u)) & 3u;   // 2 % divergence
      return b;
    }
  }
  return (uint32_t)sy_h(sp.seed, SY_GENOME, p) & 3u;
}

// haplotype ids: 1, 2 = mother; 3, 4 = father; 5, 6 = the tumor/normal individual
__host__ __device__ __forceinline__ uint32_t sy_hap_base(const rg_synth_spec& sp, uint32_t hap, uint32_t snv_thr20, uint64_t p) {
  uint32_t b = sy_genome_base(sp, p);
  const uint64_t hv = sy_h(sp.seed, SY_HAP + hap, p);
  if ((hv & 0xFFFFFu) < snv_thr20) b = (b + 1u + (uint32_t)((hv >> 20) % 3u)) & 3u;
  return b;
}

// the two haplotypes of a sample and whether it carries the de-novo / somatic layer
__host__ __device__ __forceinline__ void sy_sample(const rg_synth_spec& sp, uint32_t& h0, uint32_t& h1, bool& extra,
                                                   uint32_t& snv_thr20) {
  snv_thr20 = 524u;                          // 2^20 / 2000
  extra = false;
  if (sp.kind == RG_SYNTH_PAIR) {            // sample 0 = tumor, 1 = normal
    h0 = 5;
    h1 = 6;
    snv_thr20 = 1049u;                       // 2^20 / 1000
    extra = sp.sample == 0;
  } else {                                   // trio / quad: 0 = child / proband, 1 = mother, 2 = father, 3 = sibling
    switch (sp.sample) {
      case 0: h0 = 1; h1 = 4; extra = true; break;
      case 1: h0 = 1; h1 = 2; break;
      case 2: h0 = 3; h1 = 4; break;
      default: h0 = 2; h1 = 3; break;
    }
  }
}

// where pair `pair` of the sample comes from: fragment start, insert size, haplotype
__host__ __device__ __forceinline__ void sy_fragment(const rg_synth_spec& sp, uint64_t pair, uint64_t& start, uint64_t& ins,
                                                     uint32_t& which) {
  const uint64_t r0 = sy_h(sp.seed, SY_READ + sp.sample, pair);
  const uint64_t r1 = sy_mix(r0);
  start = r0 % (sp.genome_len - 400u);
  ins = 300u + r1 % 100u;
  which = (uint32_t)(r1 >> 32) & 1u;
}

__host__ __device__ __forceinline__ void sy_base(const rg_synth_spec& sp, uint64_t dn_thr28, uint64_t pair, uint64_t start,
                                                 uint64_t ins, uint32_t which, uint32_t mate, uint32_t j, char* seq,
                                                 char* qual) {
  uint32_t h0, h1, snv;
  bool extra;
  sy_sample(sp, h0, h1, extra, snv);
  const uint64_t p = mate == 0 ? start + j : start + ins - 1u - j;
  uint32_t b = sy_hap_base(sp, which ? h1 : h0, snv, p);
  if (extra) {                               // de-novo (child) / somatic (tumor) SNVs, each on one of the two haplotypes
    const uint64_t hd = sy_h(sp.seed, SY_DENOVO, p);
    if ((hd & 0xFFFFFFFu) < dn_thr28 && ((uint32_t)(hd >> 28) & 1u) == which) b = (b + 1u + (uint32_t)((hd >> 32) % 3u)) & 3u;
  }
  if (mate) b = 3u - b;
  const uint64_t e = sy_h(sp.seed, SY_ERR + sp.sample, (pair * 2u + mate) * 1024u + j);
  const uint32_t u = (uint32_t)e & 0xFFFFFu;
  char q = 'J';
  if (u < 2097u) {                           // 0.2 % substitution errors
    b = (b + 1u + (uint32_t)((e >> 20) % 3u)) & 3u;
    q = '#';
  } else if (u < 2097u + 10486u) {           // 1 % low quality
    q = ',';
  }
  char c = "ACGT"[b];
  if (((uint32_t)(e >> 40) & 0xFFFFFu) < 210u) {   // 'N' at 2e-4
    c = 'N';
    q = '#';
  }
  *seq = c;
  *qual = q;
}

static inline uint64_t sy_denovo_thr28(const rg_synth_spec& sp) {
  const double n = sp.genome_len / 100000.0 < 20.0 ? 20.0 : sp.genome_len / 100000.0;
  return (uint64_t)(n / (double)sp.genome_len * 268435456.0 + 0.5);
}

__global__ void __launch_bounds__(256)
synth_pairs_kernel(rg_synth_spec sp, uint64_t dn_thr28, char* __restrict__ s1, char* __restrict__ q1, char* __restrict__ s2,
                   char* __restrict__ q2) {
  const uint64_t total = sp.n_pairs * 2ull * sp.read_len;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t per_mate = sp.n_pairs * (uint64_t)sp.read_len;
    const uint32_t mate = i >= per_mate ? 1u : 0u;
    const uint64_t o = i - mate * per_mate;
    const uint64_t pr = o / sp.read_len;
    const uint32_t j = (uint32_t)(o - pr * sp.read_len);
    char c, q;
    uint64_t start, ins;
    uint32_t which;
    sy_fragment(sp, sp.first_pair + pr, start, ins, which);
    sy_base(sp, dn_thr28, sp.first_pair + pr, start, ins, which, mate, j, &c, &q);
    (mate ? s2 : s1)[o] = c;
    (mate ? q2 : q1)[o] = q;
  }
}

static int sy_check(const rg_synth_spec* sp) {
  if (!sp || sp->genome_len < 1000 || sp->read_len < 1 || sp->read_len > 300 || sp->sample > 3 ||
      (sp->kind != RG_SYNTH_TRIO && sp->kind != RG_SYNTH_PAIR)) {
    rg_set_error("rg_synth: bad spec (genome >= 1000, read_len 1..300, sample 0..3, kind trio|pair)");
    return RG_ERR_ARG;
  }
  return 0;
}

extern "C" {

int rg_synth_pairs_dev(rg_ctx* ctx, const rg_synth_spec* sp, char* s1, char* q1, char* s2, char* q2) {
  if (!ctx || sy_check(sp)) return RG_ERR_ARG;
  if (sp->n_pairs == 0) return 0;
  cudaSetDevice(ctx->device);
  {
    rg_launch_scope _ls(ctx, "synth_pairs_kernel");
    synth_pairs_kernel<<<rg_grid(ctx, 8), 256, 0, ctx->stream>>>(*sp, sy_denovo_thr28(*sp), s1, q1, s2, q2);
  }
  RG_CUDA_OK(cudaGetLastError());
  return 0;
}

int rg_synth_pairs_host(const rg_synth_spec* sp, char* s1, char* q1, char* s2, char* q2) {
  if (sy_check(sp)) return RG_ERR_ARG;
  const uint64_t thr = sy_denovo_thr28(*sp);
  for (uint64_t pr = 0; pr < sp->n_pairs; ++pr) {
    uint64_t start, ins;
    uint32_t which;
    sy_fragment(*sp, sp->first_pair + pr, start, ins, which);
    for (uint32_t mate = 0; mate < 2; ++mate)
      for (uint32_t j = 0; j < sp->read_len; ++j) {
        const uint64_t o = pr * sp->read_len + j;
        sy_base(*sp, thr, sp->first_pair + pr, start, ins, which, mate, j, (mate ? s2 : s1) + o, (mate ? q2 : q1) + o);
      }
  }
  return 0;
}

}  // extern "C"
