// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under coverm_b200/ may include,
// link or execute this file; only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs use it.
//
// Minimal BGZF / BAM / SAM record source for the CPU restatement of CoverM's
// coverage path.  The reference delegates this to rust-htslib 0.46.0 /
// hts-sys 2.2.0 (Cargo.lock:1643-1646, 872-875; not vendored), used at
// bam_generator.rs:103-144 (reader), contig.rs:108-125,166-168 (fields),
// lib.rs:138-158 (NM aux).  Only the record fields CoverM consumes are decoded,
// following the public SAM/BAM specification (SAMv1 §4.2 BAM, §4.1 BGZF).
#pragma once
#include <zlib.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace oracle {

struct Panic : std::runtime_error {  // Rust panic!  (exit status 101)
  using std::runtime_error::runtime_error;
};
struct ExitError : std::runtime_error {  // error!(..); process::exit(code)
  int code;
  ExitError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

struct CigarOp {
  uint8_t op;  // BAM codes: 0 M,1 I,2 D,3 N,4 S,5 H,6 P,7 =,8 X
  uint32_t len;
};

struct Record {
  std::string qname;
  uint16_t flag = 0;
  int32_t tid = -1;
  int32_t pos = -1;
  uint8_t mapq = 0;
  int32_t mtid = -1;
  uint32_t l_seq = 0;
  std::vector<CigarOp> cigar;
  // NM aux: 0 = absent, 1 = unsigned integer type (C/S/I), 2 = other type.
  int nm_state = 0;
  uint64_t nm = 0;

  bool is_unmapped() const { return flag & 0x4; }
  bool is_secondary() const { return flag & 0x100; }
  bool is_supplementary() const { return flag & 0x800; }
  bool is_proper_pair() const { return flag & 0x2; }
};

struct Header {
  std::vector<std::string> names;
  std::vector<uint64_t> lens;
  uint32_t target_count() const { return (uint32_t)names.size(); }
};

// lib.rs:138-158  nm()
inline uint64_t nm(const Record& r) {
  if (r.nm_state == 1) return r.nm;
  if (r.nm_state == 2) throw Panic("Unexpected data type of NM aux tag");
  throw Panic(
      "Mapping record encountered that does not have an 'NM' auxiliary tag in the SAM/BAM format. "
      "This is required to work out some coverage statistics.");
}

class AlignmentFile {
 public:
  Header header;

  explicit AlignmentFile(const std::string& path, int threads = 1) {
    std::vector<uint8_t> raw = slurp(path);
    if (raw.size() >= 4 && raw[0] == 0x1f && raw[1] == 0x8b) {
      data_ = inflate_all(raw, threads);
    } else {
      data_.swap(raw);
    }
    if (data_.size() >= 4 && memcmp(data_.data(), "BAM\1", 4) == 0) {
      is_bam_ = true;
      parse_bam_header();
    } else {
      is_bam_ = false;
      parse_sam_header();
    }
  }

  // Returns false at EOF.  (NamedBamReader::read, bam_generator.rs:113)
  bool read(Record& r) { return is_bam_ ? read_bam(r) : read_sam(r); }

 private:
  std::vector<uint8_t> data_;
  size_t off_ = 0;
  bool is_bam_ = true;
  std::map<std::string, int32_t> name_to_tid_;

  static std::vector<uint8_t> slurp(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) throw Panic("Unable to find BAM file " + path);
    std::vector<uint8_t> buf;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf.resize((size_t)n);
    if (n > 0 && fread(buf.data(), 1, (size_t)n, f) != (size_t)n) {
      fclose(f);
      throw Panic("short read on " + path);
    }
    fclose(f);
    return buf;
  }

  struct Member {
    size_t cdata_off, cdata_len, out_off;
    uint32_t isize;
  };

  // Walk gzip members (BGZF blocks are gzip members with a BC extra field
  // giving the block size) and inflate them, in parallel when every member
  // advertises its size.
  static std::vector<uint8_t> inflate_all(const std::vector<uint8_t>& raw, int threads) {
    std::vector<Member> members;
    size_t p = 0, out_total = 0;
    bool all_bgzf = true;
    while (p + 18 <= raw.size()) {
      if (raw[p] != 0x1f || raw[p + 1] != 0x8b) break;
      uint8_t flg = raw[p + 3];
      size_t q = p + 10;
      int bsize = -1;
      if (flg & 4) {
        size_t xlen = raw[q] | (raw[q + 1] << 8);
        size_t x = q + 2, xend = x + xlen;
        while (x + 4 <= xend) {
          size_t slen = raw[x + 2] | (raw[x + 3] << 8);
          if (raw[x] == 'B' && raw[x + 1] == 'C' && slen == 2) bsize = raw[x + 4] | (raw[x + 5] << 8);
          x += 4 + slen;
        }
        q = xend;
      }
      if (bsize < 0 || (flg & ~4)) {
        all_bgzf = false;
        break;
      }
      size_t block_end = p + (size_t)bsize + 1;
      if (block_end > raw.size()) throw Panic("truncated BGZF block");
      Member m;
      m.cdata_off = q;
      m.cdata_len = block_end - 8 - q;
      m.isize = raw[block_end - 4] | (raw[block_end - 3] << 8) | (raw[block_end - 2] << 16) |
                ((uint32_t)raw[block_end - 1] << 24);
      m.out_off = out_total;
      out_total += m.isize;
      members.push_back(m);
      p = block_end;
    }
    if (!all_bgzf) return inflate_stream(raw);
    std::vector<uint8_t> out(out_total);
    std::atomic<size_t> next{0};
    std::atomic<bool> bad{false};
    auto work = [&]() {
      z_stream zs;
      for (;;) {
        size_t i = next.fetch_add(1);
        if (i >= members.size()) break;
        const Member& m = members[i];
        if (m.isize == 0) continue;
        memset(&zs, 0, sizeof zs);
        if (inflateInit2(&zs, -15) != Z_OK) {
          bad = true;
          return;
        }
        zs.next_in = const_cast<Bytef*>(raw.data() + m.cdata_off);
        zs.avail_in = (uInt)m.cdata_len;
        zs.next_out = out.data() + m.out_off;
        zs.avail_out = m.isize;
        int rc = inflate(&zs, Z_FINISH);
        inflateEnd(&zs);
        if (rc != Z_STREAM_END) {
          bad = true;
          return;
        }
      }
    };
    int nt = threads < 1 ? 1 : threads;
    if (nt == 1 || members.size() < 4) {
      work();
    } else {
      std::vector<std::thread> pool;
      for (int t = 0; t < nt; ++t) pool.emplace_back(work);
      for (auto& th : pool) th.join();
    }
    if (bad) throw Panic("BGZF inflate failed");
    return out;
  }

  static std::vector<uint8_t> inflate_stream(const std::vector<uint8_t>& raw) {
    std::vector<uint8_t> out;
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, 15 + 32) != Z_OK) throw Panic("zlib init");
    zs.next_in = const_cast<Bytef*>(raw.data());
    zs.avail_in = (uInt)raw.size();
    std::vector<uint8_t> chunk(1 << 20);
    for (;;) {
      zs.next_out = chunk.data();
      zs.avail_out = (uInt)chunk.size();
      int rc = inflate(&zs, Z_NO_FLUSH);
      out.insert(out.end(), chunk.data(), chunk.data() + (chunk.size() - zs.avail_out));
      if (rc == Z_STREAM_END) {
        if (zs.avail_in == 0) break;
        inflateReset(&zs);
      } else if (rc != Z_OK) {
        inflateEnd(&zs);
        throw Panic("gzip inflate failed");
      }
    }
    inflateEnd(&zs);
    return out;
  }

  uint32_t u32(size_t o) const {
    uint32_t v;
    memcpy(&v, data_.data() + o, 4);
    return v;
  }
  int32_t i32(size_t o) const { return (int32_t)u32(o); }
  uint16_t u16(size_t o) const {
    uint16_t v;
    memcpy(&v, data_.data() + o, 2);
    return v;
  }

  void parse_bam_header() {
    size_t o = 4;
    uint32_t l_text = u32(o);
    o += 4 + l_text;
    uint32_t n_ref = u32(o);
    o += 4;
    for (uint32_t i = 0; i < n_ref; ++i) {
      uint32_t l_name = u32(o);
      o += 4;
      std::string name((const char*)data_.data() + o, l_name ? l_name - 1 : 0);
      o += l_name;
      uint32_t l_ref = u32(o);
      o += 4;
      header.names.push_back(name);
      header.lens.push_back(l_ref);
    }
    off_ = o;
  }

  bool read_bam(Record& r) {
    if (off_ + 4 > data_.size()) return false;
    uint32_t block_size = u32(off_);
    size_t o = off_ + 4;
    if (o + block_size > data_.size()) throw Panic("Error reading BAM record: truncated");
    size_t end = o + block_size;
    r.tid = i32(o);
    r.pos = i32(o + 4);
    uint8_t l_read_name = data_[o + 8];
    r.mapq = data_[o + 9];
    uint16_t n_cigar = u16(o + 12);
    r.flag = u16(o + 14);
    r.l_seq = u32(o + 16);
    r.mtid = i32(o + 20);
    o += 32;
    // htslib bam_read1: the fixed-size fields must fit block_size (sam.c bam_read1 returns -4), which the reference turns
    // into a panic (contig.rs:113-115)
    if (block_size < 32 || 32ull + l_read_name + 4ull * n_cigar + ((uint64_t)r.l_seq + 1) / 2 + r.l_seq > block_size)
      throw Panic("Error reading BAM record: the record's name, CIGAR and sequence fields do not fit its block_size");
    r.qname.assign((const char*)data_.data() + o, l_read_name ? l_read_name - 1 : 0);
    o += l_read_name;
    r.cigar.resize(n_cigar);
    for (uint16_t i = 0; i < n_cigar; ++i) {
      uint32_t v = u32(o + 4 * (size_t)i);
      r.cigar[i].op = v & 0xf;
      r.cigar[i].len = v >> 4;
    }
    o += 4 * (size_t)n_cigar;
    o += ((size_t)r.l_seq + 1) / 2 + r.l_seq;
    // htslib bam_tag2cigar (called by bam_read1): a `<l_seq>S<reflen>N` placeholder CIGAR plus a CG:B,I tag is a read with
    // more than 65535 CIGAR operations; the tag's array replaces the in-record CIGAR.
    const bool placeholder = n_cigar > 0 && r.tid >= 0 && r.pos >= 0 && r.cigar[0].op == 4 && r.cigar[0].len == r.l_seq;
    // aux scan for NM (record.aux("NM"), lib.rs:139)
    r.nm_state = 0;
    r.nm = 0;
    bool cg_seen = false;
    while (o + 3 <= end) {
      char t0 = data_[o], t1 = data_[o + 1], ty = data_[o + 2];
      o += 3;
      bool is_nm = (t0 == 'N' && t1 == 'M');
      size_t sz = 0;
      switch (ty) {
        case 'A': case 'c': case 'C': sz = 1; break;
        case 's': case 'S': sz = 2; break;
        case 'i': case 'I': case 'f': sz = 4; break;
        case 'Z': case 'H': {
          size_t e = o;
          while (e < end && data_[e]) ++e;
          sz = e - o + 1;
          break;
        }
        case 'B': {
          char sub = data_[o];
          uint32_t cnt = u32(o + 1);
          size_t es = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
          sz = 5 + es * (size_t)cnt;
          if (t0 == 'C' && t1 == 'G' && !cg_seen) {
            cg_seen = true;
            if (placeholder && (sub == 'I' || sub == 'i') && cnt >= n_cigar && cnt < (1u << 29) && o + sz <= end) {
              r.cigar.resize(cnt);
              for (uint32_t i = 0; i < cnt; ++i) {
                uint32_t v = u32(o + 5 + 4 * (size_t)i);
                r.cigar[i].op = v & 0xf;
                r.cigar[i].len = v >> 4;
              }
            }
          }
          break;
        }
        default: throw Panic("Error reading BAM record: bad aux type");
      }
      if (is_nm && r.nm_state == 0) {
        if (ty == 'C') { r.nm_state = 1; r.nm = data_[o]; }
        else if (ty == 'S') { r.nm_state = 1; r.nm = u16(o); }
        else if (ty == 'I') { r.nm_state = 1; r.nm = u32(o); }
        else r.nm_state = 2;
      }
      o += sz;
    }
    off_ = end;
    return true;
  }

  // ---- SAM text (htslib reads it through the same bam::Reader::from_path) ----
  std::string next_line() {
    size_t e = off_;
    while (e < data_.size() && data_[e] != '\n') ++e;
    std::string s((const char*)data_.data() + off_, e - off_);
    off_ = e < data_.size() ? e + 1 : e;
    if (!s.empty() && s.back() == '\r') s.pop_back();
    return s;
  }
  static std::vector<std::string> split_tab(const std::string& s) {
    std::vector<std::string> v;
    size_t a = 0;
    for (;;) {
      size_t b = s.find('\t', a);
      if (b == std::string::npos) { v.push_back(s.substr(a)); break; }
      v.push_back(s.substr(a, b - a));
      a = b + 1;
    }
    return v;
  }
  void parse_sam_header() {
    off_ = 0;
    while (off_ < data_.size() && data_[off_] == '@') {
      std::string line = next_line();
      if (line.compare(0, 3, "@SQ") == 0) {
        std::string sn;
        uint64_t ln = 0;
        for (auto& f : split_tab(line)) {
          if (f.compare(0, 3, "SN:") == 0) sn = f.substr(3);
          if (f.compare(0, 3, "LN:") == 0) ln = strtoull(f.c_str() + 3, nullptr, 10);
        }
        name_to_tid_[sn] = (int32_t)header.names.size();
        header.names.push_back(sn);
        header.lens.push_back(ln);
      }
    }
  }
  bool read_sam(Record& r) {
    std::string line;
    do {
      if (off_ >= data_.size()) return false;
      line = next_line();
    } while (line.empty());
    auto f = split_tab(line);
    if (f.size() < 11) throw Panic("Error reading BAM record: malformed SAM line");
    r.qname = f[0];
    r.flag = (uint16_t)strtoul(f[1].c_str(), nullptr, 10);
    auto tid_of = [&](const std::string& n) -> int32_t {
      if (n == "*") return -1;
      auto it = name_to_tid_.find(n);
      if (it == name_to_tid_.end()) throw Panic("Error reading BAM record: unknown reference " + n);
      return it->second;
    };
    r.tid = tid_of(f[2]);
    r.pos = (int32_t)strtol(f[3].c_str(), nullptr, 10) - 1;
    r.mapq = (uint8_t)strtoul(f[4].c_str(), nullptr, 10);
    r.cigar.clear();
    if (f[5] != "*") {
      const char* c = f[5].c_str();
      while (*c) {
        char* e;
        uint32_t len = (uint32_t)strtoul(c, &e, 10);
        static const char* ops = "MIDNSHP=X";
        const char* w = strchr(ops, *e);
        if (!w || !*e) throw Panic("Error reading BAM record: bad CIGAR");
        r.cigar.push_back({(uint8_t)(w - ops), len});
        c = e + 1;
      }
    }
    r.mtid = f[6] == "=" ? r.tid : tid_of(f[6]);
    r.l_seq = f[9] == "*" ? 0 : (uint32_t)f[9].size();
    r.nm_state = 0;
    r.nm = 0;
    for (size_t i = 11; i < f.size(); ++i) {
      if (f[i].size() >= 5 && f[i][0] == 'N' && f[i][1] == 'M' && f[i][2] == ':' && r.nm_state == 0) {
        if (f[i][3] == 'i') {
          long long v = strtoll(f[i].c_str() + 5, nullptr, 10);
          // htslib's SAM parser stores non-negative integers in the smallest
          // unsigned type (C/S/I) and negatives in a signed one.
          if (v >= 0) { r.nm_state = 1; r.nm = (uint64_t)v; }
          else r.nm_state = 2;
        } else {
          r.nm_state = 2;
        }
      }
    }
    return true;
  }
};

}  // namespace oracle
