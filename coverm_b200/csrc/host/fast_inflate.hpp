// A DEFLATE (RFC 1951) decoder specialised for BGZF blocks: the whole compressed block and the exact output size are
// known up front, so the hot loop runs without per-symbol bounds checks until it is within one maximum match of either
// end.  64-bit bit buffer refilled with one unaligned load; 11-bit primary litlen table / 8-bit primary distance table
// with subtables for longer codes; two literals per refill.  Every block is verified by the caller against the BGZF
// footer (CRC32 + ISIZE) and falls back to zlib's inflate() on any disagreement, so a decoder bug can cost time, never
// correctness.  Written from the RFC; replaces htslib's bgzf_read/inflate for this path (bam_generator.rs:103-134).
#pragma once
#include <cstdint>
#include <cstring>

namespace cmbh {

class FastInflate {
 public:
  // Decodes in[0, in_len) into out[0, out_len).  At least 8 readable bytes must follow `in + in_len` (the BGZF footer
  // does).  Returns true iff the stream is well formed and produces exactly out_len bytes.
  bool run(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len) {
    ip_ = in;
    in_end_ = in + in_len;
    bitbuf_ = 0;
    bitcnt_ = 0;
    overrun_ = 0;
    uint8_t* op = out;
    uint8_t* const out_end = out + out_len;
    for (;;) {
      refill_safe();
      const uint32_t bfinal = (uint32_t)bitbuf_ & 1, btype = ((uint32_t)bitbuf_ >> 1) & 3;
      consume(3);
      if (btype == 0) {
        consume(bitcnt_ & 7);  // to the next byte boundary
        refill_safe();
        if (bitcnt_ < 32) return false;
        const uint32_t len = (uint32_t)bitbuf_ & 0xffff, nlen = ((uint32_t)bitbuf_ >> 16) & 0xffff;
        consume(32);
        if ((len ^ nlen) != 0xffff) return false;
        // give the whole (real) bytes still in the bit buffer back to the input
        if (!unload()) return false;
        if ((size_t)(in_end_ - ip_) < len || (size_t)(out_end - op) < len) return false;
        memcpy(op, ip_, len);
        op += len;
        ip_ += len;
      } else if (btype == 1 || btype == 2) {
        if (btype == 1) {
          if (!build_fixed()) return false;
        } else if (!read_dynamic_header()) {
          return false;
        }
        if (!decode_block(out, op, out_end)) return false;
      } else {
        return false;
      }
      if (bfinal) break;
    }
    return op == out_end && bitcnt_ >= overrun_ * 8;
  }

 private:
  static constexpr int LITLEN_BITS = 11, DIST_BITS = 8, PRE_BITS = 7;
  static constexpr uint32_t F_LITERAL = 0x8000, F_EOB = 0x4000, F_SUB = 0x2000;
  // entry: value << 16 | flags | extra_bits << 8 | bits_to_consume
  uint32_t litlen_[(1 << LITLEN_BITS) + 1024];
  uint32_t dist_[(1 << DIST_BITS) + 512];
  uint32_t pre_[1 << PRE_BITS];
  bool fixed_ready_ = false;
  uint32_t fixed_litlen_[(1 << LITLEN_BITS) + 1024];
  uint32_t fixed_dist_[(1 << DIST_BITS) + 512];
  bool using_fixed_ = false;

  const uint8_t* ip_ = nullptr;
  const uint8_t* in_end_ = nullptr;
  uint64_t bitbuf_ = 0;
  int bitcnt_ = 0;
  int overrun_ = 0;  // bytes of implicit zero padding consumed past the end of the input

  static uint64_t load64(const uint8_t* p) {
    uint64_t v;
    memcpy(&v, p, 8);
    return v;
  }
  void consume(int n) {
    bitbuf_ >>= n;
    bitcnt_ -= n;
  }
  // Return the unread whole bytes to the input; false if bits past the end of the input have been consumed.
  bool unload() {
    const int held = bitcnt_ >> 3;
    if (overrun_ > held) return false;
    ip_ -= held - overrun_;
    overrun_ = 0;
    bitcnt_ &= 7;
    bitbuf_ &= (1ull << bitcnt_) - 1;
    return true;
  }
  void refill_safe() {
    while (bitcnt_ <= 56) {
      if (ip_ < in_end_) bitbuf_ |= (uint64_t)*ip_++ << bitcnt_;
      else ++overrun_;
      bitcnt_ += 8;
    }
  }

  static uint32_t reverse_bits(uint32_t code, int len) {
    uint32_t r = 0;
    for (int i = 0; i < len; ++i) {
      r = (r << 1) | (code & 1);
      code >>= 1;
    }
    return r;
  }

  // Canonical Huffman decode table.  `make(sym)` gives the entry payload (value << 16 | flags | extra << 8) of a symbol.
  template <class Make>
  static bool build_table(const uint8_t* lens, int n_syms, int table_bits, uint32_t* table, int table_cap, Make make) {
    int count[16] = {0};
    for (int s = 0; s < n_syms; ++s) count[lens[s]]++;
    count[0] = 0;
    int max_len = 15;
    while (max_len > 0 && count[max_len] == 0) --max_len;
    if (max_len == 0) {  // no codes at all: every lookup is invalid (entry 0 consumes 0 bits -> caught as error)
      for (int i = 0; i < (1 << table_bits); ++i) table[i] = 0;
      return true;
    }
    // over-subscription / completeness check (incomplete codes are allowed only for a single code, as zlib does)
    int left = 1;
    for (int l = 1; l <= 15; ++l) {
      left <<= 1;
      left -= count[l];
      if (left < 0) return false;
    }
    uint32_t next_code[16];
    uint32_t code = 0;
    for (int l = 1; l <= 15; ++l) {
      code = (code + count[l - 1]) << 1;
      next_code[l] = code;
    }
    const int primary = 1 << table_bits;
    for (int i = 0; i < primary; ++i) table[i] = 0;
    // longest code per primary slot -> subtable sizes
    int sub_next = primary;
    if (max_len > table_bits) {
      static thread_local uint8_t sub_bits[1 << 11];
      memset(sub_bits, 0, (size_t)primary);
      uint32_t nc[16];
      memcpy(nc, next_code, sizeof nc);
      for (int s = 0; s < n_syms; ++s) {
        const int l = lens[s];
        if (l <= table_bits) {
          if (l) nc[l]++;
          continue;
        }
        const uint32_t rev = reverse_bits(nc[l]++, l);
        const uint32_t slot = rev & (primary - 1);
        if (l - table_bits > sub_bits[slot]) sub_bits[slot] = (uint8_t)(l - table_bits);
      }
      for (int slot = 0; slot < primary; ++slot) {
        if (!sub_bits[slot]) continue;
        if (sub_next + (1 << sub_bits[slot]) > table_cap) return false;
        table[slot] = ((uint32_t)sub_next << 16) | F_SUB | ((uint32_t)sub_bits[slot] << 8) | (uint32_t)table_bits;
        for (int k = 0; k < (1 << sub_bits[slot]); ++k) table[sub_next + k] = 0;
        sub_next += 1 << sub_bits[slot];
      }
    }
    for (int s = 0; s < n_syms; ++s) {
      const int l = lens[s];
      if (!l) continue;
      const uint32_t rev = reverse_bits(next_code[l]++, l);
      const uint32_t payload = make(s);
      if (l <= table_bits) {
        const uint32_t e = payload | (uint32_t)l;
        for (uint32_t i = rev; i < (uint32_t)primary; i += 1u << l) table[i] = e;
      } else {
        const uint32_t slot = rev & (primary - 1);
        const uint32_t se = table[slot];
        const int sb = (se >> 8) & 0xf;
        const uint32_t base = se >> 16;
        const uint32_t e = payload | (uint32_t)(l - table_bits);
        for (uint32_t i = rev >> table_bits; i < (1u << sb); i += 1u << (l - table_bits)) table[base + i] = e;
      }
    }
    return true;
  }

  static uint32_t litlen_payload(int sym) {
    static const uint16_t base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint8_t extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    if (sym < 256) return ((uint32_t)sym << 16) | F_LITERAL;
    if (sym == 256) return F_EOB;
    if (sym > 285) return 0xffffu << 16;  // invalid length symbol: marked by an impossible base
    return ((uint32_t)base[sym - 257] << 16) | ((uint32_t)extra[sym - 257] << 8);
  }
  static uint32_t dist_payload(int sym) {
    static const uint16_t base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint8_t extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    if (sym > 29) return 0;  // invalid: base 0 is rejected by the decoder
    return ((uint32_t)base[sym] << 16) | ((uint32_t)extra[sym] << 8);
  }

  bool build_fixed() {
    if (!fixed_ready_) {
      uint8_t lens[288 + 32];
      for (int i = 0; i < 144; ++i) lens[i] = 8;
      for (int i = 144; i < 256; ++i) lens[i] = 9;
      for (int i = 256; i < 280; ++i) lens[i] = 7;
      for (int i = 280; i < 288; ++i) lens[i] = 8;
      if (!build_table(lens, 288, LITLEN_BITS, fixed_litlen_, (int)(sizeof fixed_litlen_ / 4), litlen_payload)) return false;
      for (int i = 0; i < 32; ++i) lens[i] = 5;
      if (!build_table(lens, 32, DIST_BITS, fixed_dist_, (int)(sizeof fixed_dist_ / 4), dist_payload)) return false;
      fixed_ready_ = true;
    }
    using_fixed_ = true;
    return true;
  }

  bool read_dynamic_header() {
    using_fixed_ = false;
    refill_safe();
    const int hlit = ((int)bitbuf_ & 31) + 257, hdist = (((int)bitbuf_ >> 5) & 31) + 1, hclen = (((int)bitbuf_ >> 10) & 15) + 4;
    consume(14);
    if (hlit > 286 || hdist > 30) return false;
    static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    uint8_t pre_lens[19] = {0};
    for (int i = 0; i < hclen; ++i) {
      refill_safe();
      pre_lens[order[i]] = (uint8_t)(bitbuf_ & 7);
      consume(3);
    }
    if (!build_table(pre_lens, 19, PRE_BITS, pre_, 1 << PRE_BITS, [](int s) { return (uint32_t)s << 16; })) return false;
    uint8_t lens[286 + 30 + 16];
    int n = 0;
    const int total = hlit + hdist;
    while (n < total) {
      refill_safe();
      const uint32_t e = pre_[bitbuf_ & ((1 << PRE_BITS) - 1)];
      const int l = e & 0xff;
      if (!l) return false;
      consume(l);
      const int sym = (int)(e >> 16);
      if (sym < 16) {
        lens[n++] = (uint8_t)sym;
      } else {
        int rep, val = 0;
        if (sym == 16) {
          if (n == 0) return false;
          val = lens[n - 1];
          rep = 3 + ((int)bitbuf_ & 3);
          consume(2);
        } else if (sym == 17) {
          rep = 3 + ((int)bitbuf_ & 7);
          consume(3);
        } else {
          rep = 11 + ((int)bitbuf_ & 127);
          consume(7);
        }
        if (n + rep > total) return false;
        while (rep--) lens[n++] = (uint8_t)val;
      }
    }
    if (bitcnt_ < overrun_ * 8) return false;
    if (lens[256] == 0) return false;  // no end-of-block code
    if (!build_table(lens, hlit, LITLEN_BITS, litlen_, (int)(sizeof litlen_ / 4), litlen_payload)) return false;
    if (!build_table(lens + hlit, hdist, DIST_BITS, dist_, (int)(sizeof dist_ / 4), dist_payload)) return false;
    return true;
  }

  bool decode_block(uint8_t* out_begin, uint8_t*& op_ref, uint8_t* out_end) {
    const uint32_t* lt = using_fixed_ ? fixed_litlen_ : litlen_;
    const uint32_t* dt = using_fixed_ ? fixed_dist_ : dist_;
    uint8_t* op = op_ref;
    uint64_t bitbuf = bitbuf_;
    int bitcnt = bitcnt_;
    const uint8_t* ip = ip_;
    // two 8-byte loads per iteration, the second up to 7 bytes further on: stay 8 bytes clear of the end so that both
    // remain inside the 8 readable bytes the caller guarantees after in_end_
    const uint8_t* const in_fast_end = in_end_ - ip_ >= 8 ? in_end_ - 8 : ip_ - 1;
    constexpr uint32_t LMASK = (1u << LITLEN_BITS) - 1, DMASK = (1u << DIST_BITS) - 1;
    bool ok = true, done = false;
    // ---- fast loop: far from both ends, no per-symbol bounds checks
    while (ip <= in_fast_end && out_end - op > 320) {
      bitbuf |= load64(ip) << bitcnt;
      ip += (63 - bitcnt) >> 3;
      bitcnt |= 56;
      uint32_t e = lt[bitbuf & LMASK];
      if (e & F_SUB) {
        bitbuf >>= LITLEN_BITS;
        bitcnt -= LITLEN_BITS;
        e = lt[(e >> 16) + (bitbuf & ((1u << ((e >> 8) & 0xf)) - 1))];
      }
      int l = e & 0xff;
      bitbuf >>= l;
      bitcnt -= l;
      if (e & F_LITERAL) {
        *op++ = (uint8_t)(e >> 16);
        // a second (and third) literal from the bits already in hand
        uint32_t e2 = lt[bitbuf & LMASK];
        if ((e2 & (F_LITERAL | F_SUB)) == F_LITERAL) {
          l = e2 & 0xff;
          bitbuf >>= l;
          bitcnt -= l;
          *op++ = (uint8_t)(e2 >> 16);
          e2 = lt[bitbuf & LMASK];
          if ((e2 & (F_LITERAL | F_SUB)) == F_LITERAL) {
            l = e2 & 0xff;
            bitbuf >>= l;
            bitcnt -= l;
            *op++ = (uint8_t)(e2 >> 16);
          }
        }
        continue;
      }
      if (l == 0) { ok = false; break; }
      if (e & F_EOB) { done = true; break; }
      // length (+extra), distance (+extra): at most 15+5+15+13 = 48 bits; a literal may have used 15 before: refill
      const int xl = (e >> 8) & 0xf;
      uint32_t len = (e >> 16) + ((uint32_t)bitbuf & ((1u << xl) - 1));
      bitbuf >>= xl;
      bitcnt -= xl;
      if ((e >> 16) == 0xffff) { ok = false; break; }
      if (bitcnt < 32) {
        bitbuf |= load64(ip) << bitcnt;
        ip += (63 - bitcnt) >> 3;
        bitcnt |= 56;
      }
      uint32_t d = dt[bitbuf & DMASK];
      if (d & F_SUB) {
        bitbuf >>= DIST_BITS;
        bitcnt -= DIST_BITS;
        d = dt[(d >> 16) + (bitbuf & ((1u << ((d >> 8) & 0xf)) - 1))];
      }
      l = d & 0xff;
      if (l == 0 || (d >> 16) == 0) { ok = false; break; }
      bitbuf >>= l;
      bitcnt -= l;
      const int xd = (d >> 8) & 0xf;
      const uint32_t dist = (d >> 16) + ((uint32_t)bitbuf & ((1u << xd) - 1));
      bitbuf >>= xd;
      bitcnt -= xd;
      if (dist > (size_t)(op - out_begin)) { ok = false; break; }
      const uint8_t* src = op - dist;
      uint8_t* dst = op;
      op += len;
      if (dist >= 8) {
        do {  // may write up to 7 bytes past op: inside the 320-byte margin
          memcpy(dst, src, 8);
          dst += 8;
          src += 8;
        } while (dst < op);
      } else if (dist == 1) {
        memset(dst, *src, len);
      } else {
        do {
          *dst++ = *src++;
        } while (dst < op);
      }
    }
    // hand the bit reader back (whole unread bytes return to the input so that the safe loop can bound-check)
    ip_ = ip;
    bitbuf_ = bitbuf;
    bitcnt_ = bitcnt;
    if (!unload()) return false;
    if (ip_ > in_end_) return false;  // bits of the footer were consumed as stream bits
    if (!ok) return false;
    if (done) {
      op_ref = op;
      return true;
    }
    // ---- safe loop: byte-wise refill, every write checked
    for (;;) {
      refill_safe();
      uint32_t e = lt[bitbuf_ & LMASK];
      if (e & F_SUB) {
        consume(LITLEN_BITS);
        e = lt[(e >> 16) + (bitbuf_ & ((1u << ((e >> 8) & 0xf)) - 1))];
      }
      const int l = e & 0xff;
      if (l == 0) return false;
      consume(l);
      if (e & F_LITERAL) {
        if (op >= out_end) return false;
        *op++ = (uint8_t)(e >> 16);
        continue;
      }
      if (e & F_EOB) break;
      if ((e >> 16) == 0xffff) return false;
      const int xl = (e >> 8) & 0xf;
      const uint32_t len = (e >> 16) + ((uint32_t)bitbuf_ & ((1u << xl) - 1));
      consume(xl);
      refill_safe();
      uint32_t d = dt[bitbuf_ & DMASK];
      if (d & F_SUB) {
        consume(DIST_BITS);
        d = dt[(d >> 16) + (bitbuf_ & ((1u << ((d >> 8) & 0xf)) - 1))];
      }
      const int dl = d & 0xff;
      if (dl == 0 || (d >> 16) == 0) return false;
      consume(dl);
      const int xd = (d >> 8) & 0xf;
      const uint32_t dist = (d >> 16) + ((uint32_t)bitbuf_ & ((1u << xd) - 1));
      consume(xd);
      if (dist > (size_t)(op - out_begin) || len > (size_t)(out_end - op)) return false;
      const uint8_t* src = op - dist;
      for (uint32_t k = 0; k < len; ++k) op[k] = src[k];
      op += len;
    }
    if (bitcnt_ < overrun_ * 8) return false;  // consumed bits that were never there
    op_ref = op;
    return true;
  }
};

}  // namespace cmbh
