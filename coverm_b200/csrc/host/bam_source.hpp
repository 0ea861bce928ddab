// Host record source of libcoverm_b200: BGZF / BAM / SAM -> per-read tuples written straight into the pinned SoA
// staging batch of the device library (include/coverm_b200.h, cmb_read_batch).
//
// Replaces, for the coverage path, what the reference gets from rust-htslib 0.46.0 (Cargo.lock:1643-1646):
// bam::Reader::from_path + read (bam_generator.rs:103-134), record.tid/pos/flags/mapq/cigar/seq().len()
// (contig.rs:124,166-168; filter.rs:251-277) and the NM aux lookup (lib.rs:138-158).  Written from the SAM/BAM
// specification (SAMv1 §4.1 BGZF, §4.2 BAM, §1.4 SAM); BGZF blocks are inflated with zlib on a thread pool
// (the reference's set_threads, bam_generator.rs:125-129) and tuples are extracted in parallel.
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include "fast_inflate.hpp"

#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../../include/coverm_b200.h"
#include "thread_pool.hpp"

namespace cmbh {

struct Panic : std::runtime_error {  // the reference would panic!() here (exit status 101)
  using std::runtime_error::runtime_error;
};

// One BGZF block -> its uncompressed bytes: the fast decoder first, checked against the block footer's CRC32; zlib's
// inflate() when the fast decoder declines or the checksum disagrees (htslib verifies the same CRC in bgzf.c).
class BgzfInflater {
 public:
  BgzfInflater() : fast_(new FastInflate) {
    memset(&zs_, 0, sizeof zs_);
    if (inflateInit2(&zs_, -15) != Z_OK) throw Panic("zlib init failed");
  }
  ~BgzfInflater() { inflateEnd(&zs_); }
  BgzfInflater(const BgzfInflater&) = delete;
  BgzfInflater& operator=(const BgzfInflater&) = delete;
  // cdata[clen .. clen+8) is the footer (CRC32, ISIZE).  Returns false on a corrupt block.
  bool block(const uint8_t* cdata, size_t clen, uint8_t* out, uint32_t isize) {
    if (isize == 0) return true;
    uint32_t want;
    memcpy(&want, cdata + clen, 4);
    if (fast_->run(cdata, clen, out, isize) && (uint32_t)crc32(0, out, isize) == want) return true;
    ++slow_blocks;
    inflateReset(&zs_);
    zs_.next_in = const_cast<Bytef*>(cdata);
    zs_.avail_in = (uInt)clen;
    zs_.next_out = out;
    zs_.avail_out = isize;
    if (inflate(&zs_, Z_FINISH) != Z_STREAM_END || zs_.avail_out != 0) return false;
    return (uint32_t)crc32(0, out, isize) == want;
  }
  uint64_t slow_blocks = 0;

 private:
  std::unique_ptr<FastInflate> fast_;
  z_stream zs_;
};
struct ExitError : std::runtime_error {  // error!(..); process::exit(code)
  int code;
  ExitError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

struct Header {
  std::vector<std::string> names;
  std::vector<uint64_t> lens;
};

// A BAM/SAM input: a file path or a caller-owned memory buffer holding the file's bytes.
struct InputSpec {
  std::string path;              // used for the sample name (file stem) and, when data == nullptr, opened
  const uint8_t* data = nullptr;
  size_t size = 0;
};

// One decoded record, fixed part (everything the device tuple needs) + qname location for mate matching.
struct Tuple {
  int32_t tid, pos;
  uint32_t nm, l_seq, aligned, del, ins;
  uint16_t flag;
  uint8_t mapq, nm_state;
  int32_t mtid;
  uint32_t n_iv;
};

class ByteSource {  // whole input mapped (file) or borrowed (memory)
 public:
  explicit ByteSource(const InputSpec& in) {
    if (in.data) {
      p_ = in.data;
      n_ = in.size;
      return;
    }
    fd_ = open(in.path.c_str(), O_RDONLY);
    if (fd_ < 0) throw Panic("Unable to find BAM file " + in.path);
    struct stat st;
    if (fstat(fd_, &st) != 0) throw Panic("Unable to stat BAM file " + in.path);
    n_ = (size_t)st.st_size;
    if (n_) {
      void* m = mmap(nullptr, n_, PROT_READ, MAP_PRIVATE, fd_, 0);
      if (m == MAP_FAILED) throw Panic("Unable to map BAM file " + in.path);
      madvise(m, n_, MADV_SEQUENTIAL);
      p_ = (const uint8_t*)m;
      mapped_ = true;
    }
  }
  ~ByteSource() {
    if (mapped_) munmap((void*)p_, n_);
    if (fd_ >= 0) close(fd_);
  }
  ByteSource(const ByteSource&) = delete;
  const uint8_t* data() const { return p_; }
  size_t size() const { return n_; }

 private:
  const uint8_t* p_ = nullptr;
  size_t n_ = 0;
  int fd_ = -1;
  bool mapped_ = false;
};

// Streams the uncompressed bytes of a BGZF (or plain gzip, or uncompressed) file in windows.
class InflateStream {
 public:
  InflateStream(const uint8_t* p, size_t n, ThreadPool& pool, size_t window_bytes)
      : p_(p), n_(n), pool_(pool), window_(window_bytes) {
    if (n_ >= 2 && p_[0] == 0x1f && p_[1] == 0x8b) {
      kind_ = probe_bgzf(0) ? BGZF : GZIP;
      if (kind_ == GZIP) {  // not block-indexable: inflate everything once (small inputs only)
        inflate_whole();
        kind_ = RAW;
      }
    } else {
      kind_ = RAW;
      raw_p_ = p_;
      raw_n_ = n_;
    }
  }
  // Appends up to ~window bytes of uncompressed data to buf (after buf.size()). Returns false at EOF (nothing appended).
  bool fill(std::vector<uint8_t>& buf) {
    if (kind_ == RAW) {
      if (raw_off_ >= raw_n_) return false;
      size_t take = std::min(window_, raw_n_ - raw_off_);
      size_t o = buf.size();
      buf.resize(o + take);
      memcpy(buf.data() + o, raw_p_ + raw_off_, take);
      raw_off_ += take;
      return true;
    }
    blocks_.clear();
    size_t total = 0;
    while (off_ < n_ && total < window_) {
      Block b;
      if (!parse_block(off_, b)) throw Panic("Error reading BAM record: corrupt BGZF block header");
      b.out_off = total;
      total += b.isize;
      blocks_.push_back(b);
      off_ = b.next;
    }
    if (blocks_.empty()) return false;
    size_t o = buf.size();
    buf.resize(o + total);
    uint8_t* out = buf.data() + o;
    std::atomic<bool> bad{false};
    // group blocks so that each task is ~256 KB of output
    const size_t per = 4;
    const size_t n_tasks = (blocks_.size() + per - 1) / per;
    pool_.parallel_for(n_tasks, [&](size_t task, int) {
      BgzfInflater inf;
      for (size_t i = task * per; i < std::min(blocks_.size(), (task + 1) * per); ++i) {
        const Block& b = blocks_[i];
        if (!inf.block(p_ + b.cdata, b.clen, out + b.out_off, b.isize)) bad = true;
      }
    });
    if (bad) throw Panic("Error reading BAM record: BGZF inflate failed");
    return true;
  }
  void set_window(size_t window_bytes) { window_ = window_bytes; }
  bool is_bgzf() const { return kind_ == BGZF; }
  size_t compressed_consumed() const { return off_; }  // BGZF: file bytes behind everything fill() has returned so far
  uint64_t compressed_bytes() const { return n_; }
  // uncompressed / fully inflated inputs: the raw byte range (BlockIndex cuts it into slices)
  bool is_raw() const { return kind_ == RAW; }
  const uint8_t* raw_data() const { return raw_p_; }
  size_t raw_size() const { return raw_n_; }

 private:
  struct Block {
    size_t cdata, clen, next, out_off;
    uint32_t isize;
  };
  enum Kind { BGZF, GZIP, RAW } kind_;
  const uint8_t* p_;
  size_t n_;
  ThreadPool& pool_;
  size_t window_;
  size_t off_ = 0;
  std::vector<Block> blocks_;
  std::vector<uint8_t> whole_;
  const uint8_t* raw_p_ = nullptr;
  size_t raw_n_ = 0, raw_off_ = 0;

  bool probe_bgzf(size_t o) {
    Block b;
    return parse_block(o, b);
  }
  bool parse_block(size_t o, Block& b) {
    if (o + 18 > n_ || p_[o] != 0x1f || p_[o + 1] != 0x8b || p_[o + 2] != 8 || !(p_[o + 3] & 4)) return false;
    if (p_[o + 3] & ~4) return false;
    size_t xlen = p_[o + 10] | (p_[o + 11] << 8);
    size_t x = o + 12, xend = x + xlen;
    if (xend > n_) return false;
    int bsize = -1;
    while (x + 4 <= xend) {
      size_t slen = p_[x + 2] | (p_[x + 3] << 8);
      if (p_[x] == 'B' && p_[x + 1] == 'C' && slen == 2) bsize = p_[x + 4] | (p_[x + 5] << 8);
      x += 4 + slen;
    }
    if (bsize < 0) return false;
    size_t end = o + (size_t)bsize + 1;
    if (end > n_ || end < xend + 8) return false;
    b.cdata = xend;
    b.clen = end - 8 - xend;
    b.isize = p_[end - 4] | (p_[end - 3] << 8) | (p_[end - 2] << 16) | ((uint32_t)p_[end - 1] << 24);
    b.next = end;
    return true;
  }
  void inflate_whole() {
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, 15 + 32) != Z_OK) throw Panic("zlib init failed");
    zs.next_in = const_cast<Bytef*>(p_);
    zs.avail_in = (uInt)n_;
    std::vector<uint8_t> chunk(1 << 20);
    for (;;) {
      zs.next_out = chunk.data();
      zs.avail_out = (uInt)chunk.size();
      int rc = inflate(&zs, Z_NO_FLUSH);
      whole_.insert(whole_.end(), chunk.data(), chunk.data() + (chunk.size() - zs.avail_out));
      if (rc == Z_STREAM_END) {
        if (zs.avail_in == 0) break;
        inflateReset(&zs);
      } else if (rc != Z_OK) {
        inflateEnd(&zs);
        throw Panic("Error reading BAM record: gzip inflate failed");
      }
    }
    inflateEnd(&zs);
    raw_p_ = whole_.data();
    raw_n_ = whole_.size();
  }
};

inline uint32_t rd_u32(const uint8_t* p) {
  uint32_t v;
  memcpy(&v, p, 4);
  return v;
}
inline uint16_t rd_u16(const uint8_t* p) {
  uint16_t v;
  memcpy(&v, p, 2);
  return v;
}


// ---- record layout checks shared by every host decoder -------------------------------------------------------------
// htslib's bam_read1 rejects a record whose fixed-size fields do not fit its block_size (the reference then panics with
// "Error reading BAM record", contig.rs:113-115); the CIGAR / aux walks below rely on that having been checked.
// It also restores CIGARs of more than 65535 operations from the CG:B,I tag (bam_tag2cigar: the in-record CIGAR is then the
// placeholder `<l_seq>S<reflen>N`).  `rec` points at block_size and the whole record is in memory.
struct CigarView {
  const uint8_t* ops = nullptr;  // n little-endian u32 operations (len << 4 | op)
  uint32_t n = 0;
  bool valid = false;            // false: the fixed fields overrun block_size
};
inline CigarView effective_cigar(const uint8_t* rec) {
  CigarView v;
  const uint32_t block_size = rd_u32(rec);
  const uint8_t* o = rec + 4;
  const uint32_t l_read_name = o[8], n_cigar = rd_u16(o + 12), l_seq = rd_u32(o + 16);
  const uint64_t fixed = 32ull + l_read_name + 4ull * n_cigar + ((uint64_t)l_seq + 1) / 2 + l_seq;
  if (block_size < 32 || fixed > block_size) return v;
  v.valid = true;
  v.ops = o + 32 + l_read_name;
  v.n = n_cigar;
  if (n_cigar == 0) return v;
  const uint32_t op0 = rd_u32(v.ops);
  if ((op0 & 0xf) != 4 || (op0 >> 4) != l_seq) return v;  // not the placeholder: the common case ends here
  if ((int32_t)rd_u32(o) < 0 || (int32_t)rd_u32(o + 4) < 0) return v;
  const uint8_t* a = o + fixed;
  const uint8_t* end = o + block_size;
  while (a + 3 <= end) {  // bam_aux_get(b, "CG")
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
        if (a + 5 > end) return v;
        const uint8_t sub = a[0];
        const uint32_t cnt = rd_u32(a + 1);
        const size_t es = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
        sz = 5 + es * (size_t)cnt;
        if (t0 == 'C' && t1 == 'G') {
          if ((sub == 'I' || sub == 'i') && cnt >= n_cigar && cnt < (1u << 29) && a + sz <= end) {
            v.ops = a + 5;
            v.n = cnt;
          }
          return v;
        }
        break;
      }
      default: return v;  // the NM walk raises the unknown-aux-type error
    }
    a += sz;
  }
  return v;
}
[[noreturn]] inline void throw_bad_record_layout() {
  throw Panic("Error reading BAM record: the record's name, CIGAR and sequence fields do not fit its block_size");
}
// Upper bound of the intervals a record can produce (the operation count of its effective CIGAR); -1: invalid layout.
inline int64_t record_cigar_ops(const uint8_t* rec) {
  const CigarView v = effective_cigar(rec);
  return v.valid ? (int64_t)v.n : -1;
}

// Decode the fixed fields, CIGAR summary and NM aux of one BAM record (`rec` points at block_size).
// Intervals (M/=/X blocks, contig.rs:171-186) are appended to iv_start/iv_len.
inline void decode_bam_record(const uint8_t* rec, Tuple& t, std::vector<int32_t>& iv_start, std::vector<int32_t>& iv_len) {
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
  const CigarView cv = effective_cigar(rec);
  if (!cv.valid) throw_bad_record_layout();
  const uint8_t* cig = o + 32 + l_read_name;
  uint32_t aligned = 0, del = 0, ins = 0, n_iv = 0;
  int64_t cursor = t.pos;
  for (uint32_t i = 0; i < cv.n; ++i) {
    const uint32_t v = rd_u32(cv.ops + 4 * (size_t)i);
    const uint32_t op = v & 0xf, len = v >> 4;
    switch (op) {
      case 0: case 7: case 8:  // M = X
        iv_start.push_back((int32_t)std::min<int64_t>(cursor, INT32_MAX));
        iv_len.push_back((int32_t)len);
        ++n_iv;
        cursor += len;
        aligned += len;
        break;
      case 2: cursor += len; del += len; aligned += len; break;  // D
      case 3: cursor += len; break;                              // N
      case 1: ins += len; aligned += len; break;                 // I
      default: break;                                            // S H P
    }
  }
  t.aligned = aligned;
  t.del = del;
  t.ins = ins;
  t.n_iv = n_iv;
  // aux: NM (lib.rs:139-156: U8/U16/U32 accepted, anything else is a type panic, absent is a panic)
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
}

inline std::string bam_qname(const uint8_t* rec) {
  const uint32_t l = rec[4 + 8];
  return std::string((const char*)rec + 4 + 32, l ? l - 1 : 0);
}

}  // namespace cmbh
