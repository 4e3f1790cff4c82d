// rg_crc32.hpp — CRC-32 (the gzip one: polynomial 0x04C11DB7, bit-reflected) of a BGZF member's inflated bytes.
// zlib's table-driven crc32 runs at ~1.9 GB/s here: 12 % of a BAM decode thread's time once the DEFLATE decoder is
// fast (tools/rg_inflate.hpp). This one folds the message with carry-less multiplies (PCLMULQDQ; Gopal et al., "Fast
// CRC Computation for Generic Polynomials Using PCLMULQDQ Instruction", Intel 2009): four 128-bit lanes advance 64
// bytes per step, are folded into one, reduced to 64 bits and Barrett-reduced to 32. The constants are
// bitreflect(x^n mod P) << 1 for n = 4·128+32, 4·128−32 (the 64-byte step), 128+32, 128−32 (the 16-byte step), 64;
// μ = bitreflect(⌊x^64 / P⌋), P' = bitreflect(P) — re-derived by tools/sim/crc_constants.py. Runtime dispatch: without
// PCLMULQDQ + SSE4.1 (or on another architecture) zlib's crc32 does the work. tests/test_inflate.py compares the two
// on random lengths and alignments.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <zlib.h>

#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define RG_CRC32_X86 1
#endif

namespace rg {

#ifdef RG_CRC32_X86
// CRC state in, CRC state out (both in the "inverted" domain zlib keeps internally); n >= 64 and n % 16 == 0
__attribute__((target("pclmul,sse4.1"))) static inline uint32_t crc32_fold(const unsigned char* p, size_t n, uint32_t state) {
  const __m128i k1k2 = _mm_set_epi64x(0x01c6e41596ll, 0x0154442bd4ll);
  const __m128i k3k4 = _mm_set_epi64x(0x00ccaa009ell, 0x01751997d0ll);
  const __m128i k5 = _mm_set_epi64x(0, 0x0163cd6124ll);
  const __m128i poly_mu = _mm_set_epi64x(0x01f7011641ll, 0x01db710641ll);
  const __m128i low32 = _mm_setr_epi32(~0, 0, ~0, 0);
  __m128i a = _mm_xor_si128(_mm_loadu_si128((const __m128i*)p), _mm_cvtsi32_si128((int)state));
  __m128i b = _mm_loadu_si128((const __m128i*)(p + 16));
  __m128i c = _mm_loadu_si128((const __m128i*)(p + 32));
  __m128i d = _mm_loadu_si128((const __m128i*)(p + 48));
  p += 64;
  n -= 64;
  // lane = lane.lo * k1 ^ lane.hi * k2 ^ next 16 message bytes of the lane, 64 bytes further on
  auto fold = [](__m128i lane, __m128i k, __m128i next) __attribute__((target("pclmul,sse4.1"))) {
    return _mm_xor_si128(_mm_xor_si128(_mm_clmulepi64_si128(lane, k, 0x00), _mm_clmulepi64_si128(lane, k, 0x11)), next);
  };
  while (n >= 64) {
    a = fold(a, k1k2, _mm_loadu_si128((const __m128i*)p));
    b = fold(b, k1k2, _mm_loadu_si128((const __m128i*)(p + 16)));
    c = fold(c, k1k2, _mm_loadu_si128((const __m128i*)(p + 32)));
    d = fold(d, k1k2, _mm_loadu_si128((const __m128i*)(p + 48)));
    p += 64;
    n -= 64;
  }
  // four lanes -> one (each lane is 16 bytes ahead of the next), then the 16-byte pieces that are left
  a = fold(a, k3k4, b);
  a = fold(a, k3k4, c);
  a = fold(a, k3k4, d);
  while (n >= 16) {
    a = fold(a, k3k4, _mm_loadu_si128((const __m128i*)p));
    p += 16;
    n -= 16;
  }
  // 128 -> 64 bits
  __m128i t = _mm_clmulepi64_si128(a, k3k4, 0x10);
  a = _mm_xor_si128(_mm_srli_si128(a, 8), t);
  t = _mm_srli_si128(a, 4);
  a = _mm_xor_si128(_mm_clmulepi64_si128(_mm_and_si128(a, low32), k5, 0x00), t);
  // Barrett: 64 -> 32 bits
  t = _mm_clmulepi64_si128(_mm_and_si128(a, low32), poly_mu, 0x10);
  t = _mm_clmulepi64_si128(_mm_and_si128(t, low32), poly_mu, 0x00);
  a = _mm_xor_si128(a, t);
  return (uint32_t)_mm_extract_epi32(a, 1);
}
static inline bool crc32_have_clmul() {
  static const bool have = __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1");
  return have;
}
#endif

// the CRC-32 of p[0, n), as crc32(crc32(0, NULL, 0), p, n) of zlib returns it
static inline uint32_t crc32_of(const unsigned char* p, size_t n) {
#ifdef RG_CRC32_X86
  if (n >= 64 && crc32_have_clmul()) {
    const size_t body = n & ~(size_t)15;
    const uint32_t c = ~crc32_fold(p, body, ~0u);
    return body == n ? c : (uint32_t)crc32(c, p + body, (uInt)(n - body));
  }
#endif
  return (uint32_t)crc32(crc32(0L, Z_NULL, 0), p, (uInt)n);
}

}  // namespace rg
