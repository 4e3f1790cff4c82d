// rg_inflate.hpp — raw DEFLATE (RFC 1951) decoder for BGZF members: whole compressed payload in, known
// inflated size out. Host ingest of BAM input (tools/bamview.cpp) spends ~85 % of its time inflating
// (profiles/r2_host_ingest.txt): zlib's decoder refills its bit buffer a byte at a time and goes back to its
// state machine for every symbol; this one keeps a 64-bit bit buffer refilled with one unaligned load, resolves a
// symbol with one lookup in an 11-bit (litlen) / 8-bit (distance) table (+ one sub-table lookup for longer codes),
// decodes up to three literals per refill, looks the next symbol's entry up before it copies a match (the load overlaps
// the copy) and copies the first 24 bytes of a match without a loop test (a BAM's matches average 13 bytes). The caller
// checks the gzip CRC-32 of every member, so a decoder bug cannot pass silently.
#pragma once
#include <stdint.h>
#include <string.h>

namespace rg {

class Inflater {
 public:
  // Inflates exactly `out_len` bytes from the raw deflate stream in[0, in_len). false: corrupt stream, or the
  // stream does not end exactly at out_len bytes.
  bool run(const unsigned char* in, size_t in_len, unsigned char* out, size_t out_len) {
    in_ = in;
    in_end_ = in + in_len;
    bitbuf_ = 0;
    bitcnt_ = 0;
    over_ = 0;
    unsigned char* const out_begin = out;
    unsigned char* const out_end = out + out_len;
    for (;;) {
      need(3);
      const unsigned final_block = (unsigned)bitbuf_ & 1, type = ((unsigned)bitbuf_ >> 1) & 3;
      drop(3);
      if (type == 0) {
        drop(bitcnt_ & 7);  // to the next byte boundary
        need(32);
        const unsigned len = (unsigned)bitbuf_ & 0xFFFF, nlen = ((unsigned)(bitbuf_ >> 16)) & 0xFFFF;
        drop(32);
        if ((len ^ nlen) != 0xFFFF || over_) return false;
        // bytes still in the bit buffer belong to the stored data
        size_t left = len;
        if ((size_t)(out_end - out) < left) return false;
        while (left && bitcnt_ >= 8) {
          *out++ = (unsigned char)bitbuf_;
          drop(8);
          --left;
        }
        if ((size_t)(in_end_ - in_) < left) return false;
        memcpy(out, in_, left);
        out += left;
        in_ += left;
      } else if (type == 1) {
        if (!fixed_ready_) build_fixed();
        if (!decode_block(fixed_litlen_, fixed_dist_, out, out_begin, out_end)) return false;
      } else if (type == 2) {
        if (!read_dynamic_tables()) return false;
        if (!decode_block(litlen_, dist_, out, out_begin, out_end)) return false;
      } else {
        return false;
      }
      if (final_block) break;
    }
    // phantom zero bytes fed past the end of the input must not have been consumed
    if ((unsigned)over_ * 8 > bitcnt_) return false;
    return out == out_end;
  }

 private:
  enum { LL_BITS = 11, D_BITS = 8, PRE_BITS = 7 };
  // entry: [5:0] bits to drop; [12:8] extra bits (length / distance base) or sub-table bits; [30:16] value (literal
  // byte, base length, base distance, sub-table start); bit 31 = literal; bit 15 = exceptional: | SUB, | END or
  // neither (a bit pattern no code of the block maps to)
  enum : uint32_t { LITERAL = 0x80000000u, BASE = 0, EXC = 0x8000u, SUB = 0x8000u | 0x2000u, END = 0x8000u | 0x4000u, INVALID = 0x8000u };
  static inline uint32_t entry(uint32_t value, uint32_t kind, uint32_t extra, uint32_t bits) {
    return (value << 16) | kind | (extra << 8) | bits;
  }
  static inline bool is_literal(uint32_t e) { return (int32_t)e < 0; }
  static inline bool is_sub(uint32_t e) { return (e & 0x2000u) != 0; }
  static inline bool is_end(uint32_t e) { return (e & 0x4000u) != 0; }
  static inline bool is_exc(uint32_t e) { return (e & EXC) != 0; }

  const unsigned char* in_ = nullptr;
  const unsigned char* in_end_ = nullptr;
  uint64_t bitbuf_ = 0;
  unsigned bitcnt_ = 0;
  unsigned over_ = 0;
  bool fixed_ready_ = false;
  uint32_t litlen_[(1 << LL_BITS) + 288 * 16], dist_[(1 << D_BITS) + 32 * 128], pre_[1 << PRE_BITS];
  uint32_t fixed_litlen_[(1 << LL_BITS) + 288 * 16], fixed_dist_[(1 << D_BITS) + 32 * 128];

  inline void need(unsigned n) {  // careful refill: a byte at a time, zeros (counted) past the end
    while (bitcnt_ < n) {
      uint64_t b = 0;
      if (in_ < in_end_) b = *in_++;
      else ++over_;
      bitbuf_ |= b << bitcnt_;
      bitcnt_ += 8;
    }
  }
  inline void drop(unsigned n) {
    bitbuf_ >>= n;
    bitcnt_ -= n;
  }
  static inline uint64_t load64(const unsigned char* p) {
    uint64_t v;
    memcpy(&v, p, 8);
    return v;  // little-endian hosts only (x86-64, aarch64)
  }
  static inline unsigned reverse_bits(unsigned code, unsigned len) {
    unsigned r = 0;
    for (unsigned i = 0; i < len; ++i) r |= ((code >> i) & 1) << (len - 1 - i);
    return r;
  }

  // canonical Huffman code of lens[0, n) -> decode table (primary `tb` bits + sub-tables). `sym_entry(sym)` gives
  // the entry of a symbol without its bit count. false: over-subscribed code.
  template <typename SymEntry>
  static bool build_table(const unsigned char* lens, unsigned n, unsigned tb, uint32_t* tab, unsigned tab_cap, SymEntry sym_entry) {
    unsigned count[16] = {0};
    for (unsigned s = 0; s < n; ++s) count[lens[s]]++;
    count[0] = 0;
    int left = 1;
    for (unsigned l = 1; l <= 15; ++l) {
      left <<= 1;
      left -= (int)count[l];
      if (left < 0) return false;
    }
    unsigned next_code[16];
    unsigned code = 0;
    for (unsigned l = 1; l <= 15; ++l) {
      code = (code + count[l - 1]) << 1;
      next_code[l] = code;
    }
    const unsigned primary = 1u << tb;
    for (unsigned i = 0; i < primary; ++i) tab[i] = entry(0, INVALID, 0, 1);
    // longest code behind each primary prefix
    unsigned char sub_len[1 << LL_BITS];
    bool any_long = false;
    unsigned codes[288];
    for (unsigned s = 0; s < n; ++s) {
      const unsigned l = lens[s];
      if (!l) continue;
      const unsigned rev = reverse_bits(next_code[l]++, l);
      codes[s] = rev;
      if (l > tb) {
        if (!any_long) {
          memset(sub_len, 0, primary);
          any_long = true;
        }
        unsigned char& m = sub_len[rev & (primary - 1)];
        if (l > m) m = (unsigned char)l;
      }
    }
    unsigned used = primary;
    for (unsigned s = 0; s < n; ++s) {
      const unsigned l = lens[s];
      if (!l) continue;
      const unsigned rev = codes[s];
      if (l <= tb) {
        const uint32_t e = sym_entry(s) | l;
        for (unsigned i = rev; i < primary; i += 1u << l) tab[i] = e;
      } else {
        const unsigned prefix = rev & (primary - 1);
        const unsigned sub_bits = sub_len[prefix] - tb;
        if (!is_sub(tab[prefix])) {
          if (used + (1u << sub_bits) > tab_cap) return false;
          tab[prefix] = entry(used, SUB, sub_bits, tb);
          for (unsigned i = 0; i < (1u << sub_bits); ++i) tab[used + i] = entry(0, INVALID, 0, 1);
          used += 1u << sub_bits;
        }
        const unsigned start = (tab[prefix] >> 16) & 0x7FFF;
        const uint32_t e = sym_entry(s) | (l - tb);
        for (unsigned i = rev >> tb; i < (1u << sub_bits); i += 1u << (l - tb)) tab[start + i] = e;
      }
    }
    return true;
  }

  static uint32_t litlen_entry(unsigned s) {
    static const unsigned short base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const unsigned char extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    if (s < 256) return entry(s, LITERAL, 0, 0);
    if (s == 256) return entry(0, END, 0, 0);
    if (s < 286) return entry(base[s - 257], BASE, extra[s - 257], 0);
    return entry(0, INVALID, 0, 0);
  }
  static uint32_t dist_entry(unsigned s) {
    static const unsigned short base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const unsigned char extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    if (s < 30) return entry(base[s], BASE, extra[s], 0);
    return entry(0, INVALID, 0, 0);
  }
  static uint32_t pre_entry(unsigned s) { return entry(s, LITERAL, 0, 0); }

  void build_fixed() {
    unsigned char lens[288 + 32];
    for (unsigned s = 0; s < 288; ++s) lens[s] = s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8;
    for (unsigned s = 0; s < 32; ++s) lens[288 + s] = 5;
    build_table(lens, 288, LL_BITS, fixed_litlen_, sizeof(fixed_litlen_) / 4, litlen_entry);
    build_table(lens + 288, 32, D_BITS, fixed_dist_, sizeof(fixed_dist_) / 4, dist_entry);
    fixed_ready_ = true;
  }

  bool read_dynamic_tables() {
    static const unsigned char order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    need(14);
    const unsigned hlit = ((unsigned)bitbuf_ & 31) + 257, hdist = (((unsigned)bitbuf_ >> 5) & 31) + 1, hclen = (((unsigned)bitbuf_ >> 10) & 15) + 4;
    drop(14);
    if (hlit > 286 || hdist > 30) return false;
    unsigned char pre_lens[19] = {0};
    for (unsigned i = 0; i < hclen; ++i) {
      need(3);
      pre_lens[order[i]] = (unsigned char)(bitbuf_ & 7);
      drop(3);
    }
    if (!build_table(pre_lens, 19, PRE_BITS, pre_, 1 << PRE_BITS, pre_entry)) return false;
    unsigned char lens[288 + 32];
    unsigned i = 0;
    while (i < hlit + hdist) {
      need(PRE_BITS + 7);
      const uint32_t e = pre_[bitbuf_ & ((1u << PRE_BITS) - 1)];
      if (!is_literal(e)) return false;
      drop(e & 63);
      const unsigned s = (e >> 16) & 0x7FFF;
      if (s < 16) {
        lens[i++] = (unsigned char)s;
        continue;
      }
      unsigned rep, v = 0;
      if (s == 16) {
        if (i == 0) return false;
        v = lens[i - 1];
        rep = 3 + ((unsigned)bitbuf_ & 3);
        drop(2);
      } else if (s == 17) {
        rep = 3 + ((unsigned)bitbuf_ & 7);
        drop(3);
      } else {
        rep = 11 + ((unsigned)bitbuf_ & 127);
        drop(7);
      }
      if (i + rep > hlit + hdist) return false;
      memset(lens + i, (int)v, rep);
      i += rep;
    }
    if (lens[256] == 0) return false;  // no end-of-block code
    unsigned char ll[288] = {0}, dd[32] = {0};
    memcpy(ll, lens, hlit);
    memcpy(dd, lens + hlit, hdist);
    return build_table(ll, 288, LL_BITS, litlen_, sizeof(litlen_) / 4, litlen_entry) &&
           build_table(dd, 32, D_BITS, dist_, sizeof(dist_) / 4, dist_entry);
  }

  bool decode_block(const uint32_t* ll, const uint32_t* dd, unsigned char*& out_ref, unsigned char* out_begin, unsigned char* out_end) {
    unsigned char* out = out_ref;
    if (over_ == 0 && in_end_ - in_ >= 16 && out_end - out >= 274) {
      // ---- fast loop: at least 16 input bytes and 274 output bytes of room, no per-symbol bounds checks; the
      // state lives in locals (byte stores may alias the members)
      const unsigned char* in = in_;
      const unsigned char* const in_fast_end = in_end_ - 16;
      unsigned char* const out_fast_end = out_end - 274;
      uint64_t bb = bitbuf_;
      unsigned bc = bitcnt_;
      const uint64_t ll_mask = (1u << LL_BITS) - 1, d_mask = (1u << D_BITS) - 1;
#define RG_REFILL()           \
  do {                        \
    bb |= load64(in) << bc;   \
    in += (63 - bc) >> 3;     \
    bc |= 56;                 \
  } while (0)
#define RG_LOOKUP(e, tab, mask)                                              \
  do {                                                                       \
    e = tab[bb & (mask)];                                                    \
    if (__builtin_expect(is_sub(e), 0)) {                                    \
      bb >>= (e & 63);                                                       \
      bc -= (e & 63);                                                        \
      e = tab[((e >> 16) & 0x7FFF) + (bb & ((1u << ((e >> 8) & 31)) - 1))];  \
    }                                                                        \
    bb >>= (e & 63);                                                         \
    bc -= (e & 63);                                                          \
  } while (0)
      // The entry of the NEXT symbol is looked up as soon as its bits are known — before the match copy, before the
      // loop test — so that the table load's latency overlaps the copy / the refill instead of heading the next
      // symbol's dependency chain. (A looked-up entry consumes nothing until RG_CONSUME: leaving the loop with one in
      // hand is harmless.)
#define RG_CONSUME(e, tab)                                                   \
  do {                                                                       \
    if (__builtin_expect(is_sub(e), 0)) {                                    \
      bb >>= (e & 63);                                                       \
      bc -= (e & 63);                                                        \
      e = tab[((e >> 16) & 0x7FFF) + (bb & ((1u << ((e >> 8) & 31)) - 1))];  \
    }                                                                        \
    bb >>= (e & 63);                                                         \
    bc -= (e & 63);                                                          \
  } while (0)
      uint32_t e;
      RG_REFILL();
      e = ll[bb & ll_mask];
      while (in <= in_fast_end && out <= out_fast_end) {
        RG_CONSUME(e, ll);
        if (is_literal(e)) {
          *out++ = (unsigned char)(e >> 16);
          e = ll[bb & ll_mask];
          RG_CONSUME(e, ll);
          if (is_literal(e)) {
            *out++ = (unsigned char)(e >> 16);
            e = ll[bb & ll_mask];
            RG_CONSUME(e, ll);
            if (is_literal(e)) {
              *out++ = (unsigned char)(e >> 16);
              RG_REFILL();
              e = ll[bb & ll_mask];
              continue;
            }
          }
          RG_REFILL();
        }
        if (__builtin_expect(is_exc(e), 0)) {
          if (!is_end(e)) return false;
          in_ = in;
          bitbuf_ = bb & ((1ull << bc) - 1);  // (the bits above bc are bytes `in` has not been advanced over yet)
          bitcnt_ = bc;
          out_ref = out;
          return true;
        }
        unsigned xb = (e >> 8) & 31;
        const unsigned length = (e >> 16) + (unsigned)(bb & ((1u << xb) - 1));
        bb >>= xb;
        bc -= xb;
        RG_LOOKUP(e, dd, d_mask);
        if (__builtin_expect(is_exc(e), 0)) return false;
        xb = (e >> 8) & 31;
        const unsigned dist = (e >> 16) + (unsigned)(bb & ((1u << xb) - 1));
        bb >>= xb;
        bc -= xb;
        if (__builtin_expect(dist > (size_t)(out - out_begin), 0)) return false;
        RG_REFILL();
        e = ll[bb & ll_mask];                 // the next symbol's entry, while the copy runs
        const unsigned char* src = out - dist;
        unsigned char* const stop = out + length;
        if (dist >= 8) {
          // three words without a test (most matches of a BAM are shorter than 24 bytes: no loop branch to mispredict;
          // the room for the overshoot is part of the 274 bytes the loop condition keeps free), then the rest
          memcpy(out, src, 8);
          memcpy(out + 8, src + 8, 8);
          memcpy(out + 16, src + 16, 8);
          if (__builtin_expect(length > 24, 0)) {
            out += 24;
            src += 24;
            do {
              memcpy(out, src, 8);
              out += 8;
              src += 8;
            } while (out < stop);
          }
        } else if (dist == 1) {
          // a run of one byte (quality strings): whole words of it
          const uint64_t v = 0x0101010101010101ull * *src;
          memcpy(out, &v, 8);
          memcpy(out + 8, &v, 8);
          if (__builtin_expect(length > 16, 0)) {
            out += 16;
            do {
              memcpy(out, &v, 8);
              out += 8;
            } while (out < stop);
          }
        } else {
          do *out++ = *src++; while (out < stop);
        }
        out = stop;
      }
      in_ = in;
      bitbuf_ = bb & ((1ull << bc) - 1);
      bitcnt_ = bc;
#undef RG_REFILL
#undef RG_LOOKUP
#undef RG_CONSUME
    }
    // ---- careful loop: the ends of the input and of the output
    for (;;) {
      uint32_t e;
      need(48);
      if (!lookup(ll, LL_BITS, e)) return false;
      if (is_literal(e)) {
        if (out == out_end) return false;
        *out++ = (unsigned char)(e >> 16);
        continue;
      }
      if (is_end(e)) {
        out_ref = out;
        return true;
      }
      unsigned xb = (e >> 8) & 31;
      const unsigned length = (e >> 16) + (unsigned)(bitbuf_ & ((1u << xb) - 1));
      drop(xb);
      if (!lookup(dd, D_BITS, e) || is_exc(e) || is_literal(e)) return false;
      xb = (e >> 8) & 31;
      const unsigned dist = (e >> 16) + (unsigned)(bitbuf_ & ((1u << xb) - 1));
      drop(xb);
      if (dist > (size_t)(out - out_begin) || length > (size_t)(out_end - out)) return false;
      const unsigned char* src = out - dist;
      for (unsigned i = 0; i < length; ++i) out[i] = src[i];
      out += length;
    }
  }
  // one symbol through `tab` (sub-table resolved, bits dropped); false for a bit pattern without a code
  inline bool lookup(const uint32_t* tab, unsigned tb, uint32_t& e) {
    e = tab[bitbuf_ & ((1u << tb) - 1)];
    if (is_sub(e)) {
      drop(e & 63);
      e = tab[((e >> 16) & 0x7FFF) + (bitbuf_ & ((1u << ((e >> 8) & 31)) - 1))];
    }
    if (is_exc(e) && !is_end(e)) return false;
    drop(e & 63);
    return true;
  }
};

}  // namespace rg
