// rg_sk.cuh — super-k-mer records: what travels between the partition pass and the tile insert.
//
// The reference inserts every k-mer occurrence on its own (count_main.cc:154-157 ->
// large_hash_array.hpp:298-302). Here the consecutive k-mers of a read that share a minimizer (same hash
// prefix mh, rg_internal.cuh) are kept together as ONE 16-byte record, so the partition moves ~2.3 bytes
// per k-mer instead of 8, and the tile insert re-extracts the k-mers from the record:
//   x, y, z : the record's bases, 2 bits each, A0 C1 G2 T3, BIG-endian (base 0 in bits 31..30 of x);
//             a record of len k-mers holds len + k - 1 <= 47 bases (what follows in x, y, z is whatever the read
//             continued with: not part of the record)
//   w       : out of the partition: mh (low 4 bits clear) | (len - 1), len = 1..16 k-mers;
//             out of dedup_kernel (what the tile insert reads): (multiplicity << 4) | (len - 1)
// A record is a PURE super-k-mer: a maximal run of k-mers with the same minimizer value, cut only where the read
// has no k-mer, every 16 k-mers of an over-long run, and at the edges of a 16384-window super-chunk — never at
// the 16-base blocks of the partition's threads. Every read that covers a locus therefore produces the same
// records, and dedup_kernel collapses the ~30 copies a 30x sample holds of each into one record + multiplicity.
// The minimizer only ROUTES (tile = f(mh)); the counted keys are the canonical k-mers themselves, so the
// table content is independent of how reads were cut into records.
#pragma once
#include "rg_internal.cuh"

// k-mer at offset o (0..15) of a record: forward key (first base in the MSBs of 2k bits, mer_dna.hpp:330-354)
__device__ __forceinline__ uint64_t sk_fwd_key(const uint4& rec, uint32_t o, uint32_t k) {
  const uint32_t s = 2u * o;                                   // 0..30
  const uint32_t a = __funnelshift_l(rec.y, rec.x, s);         // bases o .. o+15
  const uint32_t b = __funnelshift_l(rec.z, rec.y, s);         // bases o+16 .. o+31
  return (((uint64_t)a << 32) | b) >> (64u - 2u * k);
}
// the counted key: canonical = numerically smaller of forward and reverse complement (mer_iterator.hpp:59)
__device__ __forceinline__ uint64_t sk_key(const uint4& rec, uint32_t o, uint32_t k, int canonical) {
  const uint64_t f = sk_fwd_key(rec, o, k);
  if (!canonical) return f;
  const uint64_t r = rg_revcomp_jf(f, (int)k);
  return f < r ? f : r;
}

// Minimizer length for a table of n_tiles_global tiles: the window W = k - m + 1 should be one of the
// values the partition kernel is specialised for (16, 14, 12, 10, 8), with enough minimizer classes per tile
// for an even load; otherwise the generic (slower) window with m = min(k, 16). Only the ~2/(W+1) of the 4^m/2
// classes with the smallest order values ever win a window, and a tile's load is the sum over the classes it
// owns: 4^m >= 2^13 * n_tiles leaves ~500 winning classes per tile.
static inline int sk_choose_m(int k, uint64_t n_tiles_global) {
  int lg = 0;
  while ((1ull << lg) < n_tiles_global) ++lg;
  int need = (lg + 13 + 1) / 2;
  if (need > 16) need = 16;
  const int ws[5] = {16, 14, 12, 10, 8};
  for (int i = 0; i < 5; ++i) {
    const int m = k - ws[i] + 1;
    if (m >= need && m <= 16 && m >= 1) return m;
  }
  return k < 16 ? k : 16;
}
