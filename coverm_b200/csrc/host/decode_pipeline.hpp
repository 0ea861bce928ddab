// Region-parallel BAM decode for libcoverm_b200's host side.
//
// The file is cut into work items of ~1 MB of uncompressed data (whole BGZF blocks).  Worker threads take items in
// order; each one inflates its item into a private buffer (so the data stays in that core's cache), waits for the
// record alignment handed over by its predecessor (the bytes of the record that straddles the item boundary), walks
// the block_size chain of its own records, hands the alignment on, and only then extracts its tuples straight into
// the pinned SoA staging batch of the device library.  The only sequential part is the block_size walk (one load
// per record on cache-hot data); inflate (zlib) and tuple extraction run on all cores.
// The calling thread is the coordinator: it alone talks to the (not thread-safe) device ABI — acquiring staging
// batches ahead of the workers and submitting each batch once every item assigned to it has been extracted.
//
// Replaces for this path: bam::Reader::read + htslib's BGZF thread pool (bam_generator.rs:103-134, 125-129).
#pragma once
#include <atomic>
#include <chrono>
#include <climits>
#include <exception>
#include <memory>
#include <mutex>
#include <thread>

#include "bam_source.hpp"

namespace cmbh {

struct BlockRef {
  size_t cdata, clen;  // compressed payload (BGZF) or raw slice [cdata, cdata+clen)
  uint32_t isize;      // uncompressed size
};

// All blocks of the input with their uncompressed offsets.
class BlockIndex {
 public:
  std::vector<BlockRef> blocks;
  std::vector<uint64_t> ustart;  // blocks.size()+1 cumulative uncompressed offsets
  bool bgzf = false;
  const uint8_t* p = nullptr;
  size_t n = 0;

  void build(const uint8_t* data, size_t size) {
    p = data;
    n = size;
    blocks.clear();
    bgzf = size >= 18 && data[0] == 0x1f && data[1] == 0x8b;
    if (bgzf) {
      size_t o = 0;
      while (o < size) {
        if (o + 18 > size || data[o] != 0x1f || data[o + 1] != 0x8b || data[o + 2] != 8 || !(data[o + 3] & 4))
          throw Panic("Error reading BAM record: corrupt BGZF block header");
        const size_t xlen = data[o + 10] | (data[o + 11] << 8);
        size_t x = o + 12;
        const size_t xend = x + xlen;
        if (xend > size) throw Panic("Error reading BAM record: truncated BGZF block");
        int bsize = -1;
        while (x + 4 <= xend) {
          const size_t slen = data[x + 2] | (data[x + 3] << 8);
          if (data[x] == 'B' && data[x + 1] == 'C' && slen == 2) bsize = data[x + 4] | (data[x + 5] << 8);
          x += 4 + slen;
        }
        if (bsize < 0) throw Panic("Error reading BAM record: gzip member without a BGZF block size");
        const size_t end = o + (size_t)bsize + 1;
        if (end > size || end < xend + 8) throw Panic("Error reading BAM record: truncated BGZF block");
        BlockRef b;
        b.cdata = xend;
        b.clen = end - 8 - xend;
        b.isize = data[end - 4] | (data[end - 3] << 8) | (data[end - 2] << 16) | ((uint32_t)data[end - 1] << 24);
        blocks.push_back(b);
        o = end;
      }
    } else {
      for (size_t o = 0; o < size; o += 65536) blocks.push_back({o, std::min<size_t>(65536, size - o), (uint32_t)std::min<size_t>(65536, size - o)});
    }
    ustart.assign(blocks.size() + 1, 0);
    for (size_t i = 0; i < blocks.size(); ++i) ustart[i + 1] = ustart[i] + blocks[i].isize;
  }

  // Inflate blocks [b0, b1) into dst (which has room for ustart[b1]-ustart[b0] bytes).
  void inflate(size_t b0, size_t b1, uint8_t* dst, BgzfInflater& inf) const {
    for (size_t b = b0; b < b1; ++b) {
      const BlockRef& r = blocks[b];
      uint8_t* out = dst + (ustart[b] - ustart[b0]);
      if (!bgzf) {
        memcpy(out, p + r.cdata, r.clen);
        continue;
      }
      if (!inf.block(p + r.cdata, r.clen, out, r.isize)) throw Panic("Error reading BAM record: BGZF inflate failed");
    }
  }
};

// A BAM record header that could be real: sane block_size, reference ids, name length and NUL, field sizes.  Shared by
// the speculative record alignment of the host decoder (decode_runner.hpp) and the block-range probe below.
inline bool record_plausible(const uint8_t* buf, size_t s, size_t usize, uint32_t n_ref) {
  if (s + 36 > usize) return false;
  const uint32_t bs = rd_u32(buf + s);
  if (bs < 32 || bs > (64u << 20)) return false;
  const int32_t tid = (int32_t)rd_u32(buf + s + 4), pos = (int32_t)rd_u32(buf + s + 8), mtid = (int32_t)rd_u32(buf + s + 24);
  if (tid < -1 || tid >= (int32_t)n_ref || mtid < -1 || mtid >= (int32_t)n_ref || pos < -1) return false;
  const uint32_t l_name = buf[s + 12], n_cig = rd_u16(buf + s + 16), l_seq = rd_u32(buf + s + 20);
  if (l_name == 0 || l_seq > (1u << 28)) return false;
  const uint64_t fixed = 32ull + l_name + 4ull * n_cig + (l_seq + 1) / 2 + l_seq;
  if (fixed > bs) return false;
  if (s + 36 + l_name <= usize && buf[s + 36 + l_name - 1] != 0) return false;
  return true;
}

// Decode the fixed fields, CIGAR summary and NM aux of one BAM record; M/=/X blocks (contig.rs:171-186) are written to
// ivs/ivl (room for n_cigar_op entries).  Returns the number of intervals written.
inline uint32_t decode_bam_record_into(const uint8_t* rec, Tuple& t, int32_t* ivs, int32_t* ivl) {
  const uint32_t block_size = rd_u32(rec);
  const uint8_t* o = rec + 4;
  const uint8_t* end = o + block_size;
  t.tid = (int32_t)rd_u32(o);
  t.pos = (int32_t)rd_u32(o + 4);
  const uint32_t l_read_name = o[8];
  t.mapq = o[9];
  const uint32_t n_cigar = rd_u16(o + 12);
  t.flag = rd_u16(o + 14);
  t.l_seq = rd_u32(o + 16);
  t.mtid = (int32_t)rd_u32(o + 20);
  const CigarView cv = effective_cigar(rec);  // layout check + CG:B,I long-CIGAR restoration (bam_source.hpp)
  if (!cv.valid) throw_bad_record_layout();
  const uint8_t* cig = o + 32 + l_read_name;
  uint32_t aligned = 0, del = 0, ins = 0, n_iv = 0;
  int64_t cursor = t.pos;
  for (uint32_t i = 0; i < cv.n; ++i) {
    const uint32_t v = rd_u32(cv.ops + 4 * (size_t)i);
    const uint32_t op = v & 0xf, len = v >> 4;
    switch (op) {
      case 0: case 7: case 8:
        ivs[n_iv] = cursor < 0 ? -1 : (int32_t)std::min<int64_t>(cursor, INT32_MAX);
        ivl[n_iv] = (int32_t)len;
        ++n_iv;
        cursor += len;
        aligned += len;
        break;
      case 2: cursor += len; del += len; aligned += len; break;
      case 3: cursor += len; break;
      case 1: ins += len; aligned += len; break;
      default: break;
    }
  }
  t.aligned = aligned;
  t.del = del;
  t.ins = ins;
  t.n_iv = n_iv;
  const uint8_t* a = cig + 4 * (size_t)n_cigar + (t.l_seq + 1) / 2 + t.l_seq;
  t.nm_state = 0;
  t.nm = 0;
  while (a + 3 <= end) {
    const uint8_t t0 = a[0], t1 = a[1], ty = a[2];
    a += 3;
    size_t sz;
    switch (ty) {
      case 'A': case 'c': case 'C': sz = 1; break;
      case 's': case 'S': sz = 2; break;
      case 'i': case 'I': case 'f': sz = 4; break;
      case 'Z': case 'H': {
        const uint8_t* e = (const uint8_t*)memchr(a, 0, (size_t)(end - a));
        sz = e ? (size_t)(e - a) + 1 : (size_t)(end - a);
        break;
      }
      case 'B': {
        if (a + 5 > end) { sz = (size_t)(end - a); break; }
        const uint8_t sub = a[0];
        const uint32_t cnt = rd_u32(a + 1);
        const size_t es = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
        sz = 5 + es * (size_t)cnt;
        break;
      }
      default: throw Panic("Error reading BAM record: unknown aux type");
    }
    if (t0 == 'N' && t1 == 'M' && t.nm_state == 0) {
      if (ty == 'C') { t.nm_state = 1; t.nm = a[0]; }
      else if (ty == 'S') { t.nm_state = 1; t.nm = rd_u16(a); }
      else if (ty == 'I') { t.nm_state = 1; t.nm = rd_u32(a); }
      else t.nm_state = 2;
    }
    a += sz;
  }
  return n_iv;
}

}  // namespace cmbh
