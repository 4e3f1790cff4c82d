// rg_extract.cuh — ASCII read stream -> canonical 2-bit k-mers (shared by the count kernels).
// Replaces mer_iterator::operator++ / operator* (src/modifiedJellyfish/include/jellyfish/mer_iterator.hpp:59-89)
// and the code table of mer_dna.hpp:46-63.
//
// A CTA of NT threads stages NT*R window starts (+64 bases of tail) of the stream into
// three little-endian bit streams in shared memory (base i of the chunk at bit i resp.
// bits 2i,2i+1):
//   codes : A0 C1 G2 T3
//   bad   : 1 for every byte outside ACGTacgt (window restart, mer_iterator.hpp:72-79)
//   brk   : 1 where a new read starts (offset / fixed-stride inputs; the reference's
//           parser puts an 'N' there, mer_overlap_sequence_parser.hpp:195)
// Thread t then owns the R consecutive windows starting at bases t*R .. t*R+R-1 and
// ROLLS through them (mer_dna shift_left/shift_right, mer_dna.hpp:330-377): ~25
// instructions per window instead of re-extracting and re-reversing every window.
#pragma once
#include "rg_internal.cuh"

template <int NT, int R>
struct StageT {
  static constexpr int WINDOWS = NT * R;
  static constexpr int BASES = WINDOWS + 64;
  static constexpr int CODE_WORDS = BASES / 16;
  static constexpr int BIT_WORDS = BASES / 32;
  uint32_t codes[CODE_WORDS + 2];
  uint32_t bad[BIT_WORDS + 2];
  uint32_t brk[BIT_WORDS + 2];
};

// 16 ASCII bytes -> 16 2-bit codes (base j at bits 2j) + 16 bad flags (bit j)
__device__ __forceinline__ void convert16_le(const uint4 v, uint32_t& codes, uint32_t& bad) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  codes = 0;
  bad = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t x = (w[i] >> 1) & 0x03030303u;             // A0 C1 G3 T2 per byte
    const uint32_t c4 = x ^ ((x >> 1) & 0x01010101u);         // A0 C1 G2 T3
    // gather the four 2-bit codes: char j (byte j) -> bits 2j..2j+1 of a byte
    const uint32_t byte = (c4 | (c4 >> 6) | (c4 >> 12) | (c4 >> 18)) & 0xFFu;
    codes |= byte << (8 * i);
    const uint32_t u = w[i] & 0xDFDFDFDFu;                    // upper-case
    const uint32_t ok = __vcmpeq4(u, 0x41414141u) | __vcmpeq4(u, 0x43434343u) | __vcmpeq4(u, 0x47474747u) |
                        __vcmpeq4(u, 0x54545454u);            // 0xFF per valid byte
    const uint32_t nb = ~ok & 0x01010101u;                    // bit 0 of byte j set if char j is bad
    const uint32_t nib = (nb | (nb >> 7) | (nb >> 14) | (nb >> 21)) & 0xFu;   // bit j = char j
    bad |= nib << (4 * i);
  }
}

// Stage bases [c0, c0 + BASES) of the stream; c0 is a multiple of 16. Caller syncs around it.
template <int NT, int R>
__device__ __forceinline__ void stage_super(StageT<NT, R>& sm, const ExtractParams& ep, uint64_t c0) {
  using S = StageT<NT, R>;
  uint16_t* bad16 = reinterpret_cast<uint16_t*>(sm.bad);
  for (int i = threadIdx.x; i < S::CODE_WORDS; i += NT) {
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
    convert16_le(v, c, b);
    sm.codes[i] = c;
    bad16[i] = (uint16_t)b;
  }
  for (int i = threadIdx.x; i < S::BIT_WORDS; i += NT) {
    const uint64_t p = c0 + (uint64_t)i * 32;
    uint32_t w = 0;
    if (ep.brk != nullptr) {
      if (p < ep.n_bytes) w = __ldg(ep.brk + (p >> 5));
    } else if (ep.read_stride != 0) {
      const uint32_t r = (uint32_t)(p % ep.read_stride);
      const uint32_t first = r == 0 ? 0 : ep.read_stride - r;      // offset of the first read start >= p
      for (uint32_t j = first; j < 32; j += ep.read_stride) w |= 1u << j;
    }
    sm.brk[i] = w;
  }
  if (threadIdx.x < 2) {
    sm.codes[S::CODE_WORDS + threadIdx.x] = 0;
    sm.bad[S::BIT_WORDS + threadIdx.x] = 0xFFFFFFFFu;
    sm.brk[S::BIT_WORDS + threadIdx.x] = 0;
  }
}

// 64 flag bits starting at base p (little-endian stream)
__device__ __forceinline__ uint64_t bits_at(const uint32_t* s, int p) {
  const int j = p >> 5, sh = p & 31;
  const uint64_t lo = ((uint64_t)s[j + 1] << 32) | s[j];
  return sh ? (lo >> sh) | ((uint64_t)s[j + 2] << (64 - sh)) : lo;
}

// bit s of the result = AND of x[s .. s+n-1]
__device__ __forceinline__ uint64_t and_run(uint64_t x, int n) {
  int len = 1;
  while (2 * len <= n) {
    x &= x >> len;
    len *= 2;
  }
  if (n > len) x &= x >> (n - len);
  return x;
}

// same for a warp-uniform n in 1..64, static shifts behind uniform branches (the partition kernel's validity test)
__device__ __forceinline__ uint64_t and_run_uniform(uint64_t x, int n) {
  if (n >= 2) {
    x &= x >> 1;
    if (n >= 4) {
      x &= x >> 2;
      if (n >= 8) {
        x &= x >> 4;
        if (n >= 16) {
          x &= x >> 8;
          if (n >= 32) x &= x >> 16;
        }
      }
    }
  }
  const int len = 1 << (31 - __clz(n));
  if (n > len) x &= x >> (n - len);
  return x;
}

// reverse the order of the 2-bit groups of a 64-bit word
__device__ __forceinline__ uint64_t reverse_groups(uint64_t x) {
  x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
  x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
  const uint32_t lo = __byte_perm((uint32_t)x, 0, 0x0123), hi = __byte_perm((uint32_t)(x >> 32), 0, 0x0123);
  return ((uint64_t)lo << 32) | hi;
}

// Thread `t` rolls through its R windows (starting at local bases t*R .. t*R+R-1) and calls
// emit(w, key) for w = 0..R-1 with key == RG_EMPTY when the window holds no k-mer.
// skip_local: windows starting before this chunk-local position are dropped (first chunk of a split push)
template <int NT, int R, typename Emit>
__device__ __forceinline__ void roll_windows(const StageT<NT, R>& sm, int t, int k, int canonical, Emit&& emit,
                                             uint32_t skip_local = 0) {
  const int p0 = t * R;
  // validity of each of the R windows: k good bases and no read start strictly inside
  const uint64_t good = ~bits_at(sm.bad, p0);
  const uint64_t nobrk = ~bits_at(sm.brk, p0 + 1);          // bit j: no read starts at base p0+j+1
  uint64_t valid = and_run(good, k);
  if (k > 1) valid &= and_run(nobrk, k - 1);
  if ((uint32_t)p0 < skip_local) valid &= ~((skip_local - p0 >= 64u) ? ~0ull : ((1ull << (skip_local - p0)) - 1ull));
  // little-endian window X (base i at bits 2i): rc (Jellyfish order) = ~X, fwd = reverse of X
  const int j = p0 >> 4, sh = (p0 & 15) * 2;
  uint64_t lo = ((uint64_t)sm.codes[j + 1] << 32) | sm.codes[j];
  if (sh) lo = (lo >> sh) | ((uint64_t)sm.codes[j + 2] << (64 - sh));
  const uint64_t mask = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1);
  const int top = 2 * (k - 1);
  uint64_t rc = ~lo & mask;                                 // complement of bases p0..p0+k-1, last base on top
  uint64_t fwd = reverse_groups(lo & mask) >> (64 - 2 * k);
  // the bases that enter for windows 1..R-1 start at p0+k
  const int q0 = p0 + k;
  const int qj = q0 >> 4, qs = (q0 & 15) * 2;
  uint64_t nxt = ((uint64_t)sm.codes[qj + 1] << 32) | sm.codes[qj];
  if (qs) nxt = (nxt >> qs) | ((uint64_t)sm.codes[qj + 2] << (64 - qs));   // 32 upcoming bases, R-1 <= 31 needed
#pragma unroll
  for (int w = 0; w < R; ++w) {
    uint64_t key = RG_EMPTY;
    if ((valid >> w) & 1ull) key = (!canonical || fwd < rc) ? fwd : rc;
    emit(w, key);
    if (w + 1 < R) {
      const uint64_t c = (nxt >> (2 * w)) & 3ull;
      fwd = ((fwd << 2) | c) & mask;
      rc = (rc >> 2) | ((3ull - c) << top);
    }
  }
}
