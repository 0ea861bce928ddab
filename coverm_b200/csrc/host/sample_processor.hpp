// One BAM file ("stoit") through the device library: decode -> pinned SoA batches -> cmb_submit_batch -> per-contig
// integer statistics.  This is the record loop of contig.rs:107-215 / genome.rs:109-227, 516-729 with the CIGAR walk
// and the filters moved onto the GPU; the host keeps only what needs read names (mate matching, filter.rs:149-224).
#pragma once
#include <chrono>
#include <cstdio>

#include "bam_source.hpp"
#include "decode_runner.hpp"
#include "genes.hpp"
#include "shard_range.hpp"

namespace cmbh {

// NVTX range of a host-side stage (cmb_nvtx_push / cmb_nvtx_pop of the device library; no-ops without a profiler)
struct HostRange {
  bool open = true;
  explicit HostRange(const char* name) { cmb_nvtx_push(name); }
  void end() {
    if (open) {
      cmb_nvtx_pop();
      open = false;
    }
  }
  ~HostRange() { end(); }
  HostRange(const HostRange&) = delete;
  HostRange& operator=(const HostRange&) = delete;
};

inline double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

inline std::string file_stem(const std::string& path) {  // Path::file_stem (bam_generator.rs:360-365)
  size_t s = path.find_last_of('/');
  std::string base = s == std::string::npos ? path : path.substr(s + 1);
  size_t d = base.find_last_of('.');
  if (d == std::string::npos || d == 0) return base;
  return base.substr(0, d);
}

struct SampleTiming {
  double total_s = 0, decode_s = 0, submit_wait_s = 0, end_sample_s = 0;
  double header_s = 0, index_s = 0, device_call_s = 0;  // inside decode_s: header parse, BGZF block index (+ range probes), cmb_submit_bgzf
  cmb_sample_timing device{};
  uint64_t h2d_bytes = 0;
  bool device_decode = false;
  cmb_bgzf_result bgzf{};
  uint32_t decode_launches = 0;
  // multi-GPU contig sharding
  uint32_t group_ranks = 1, shard_blocks = 0, total_blocks = 0, range_probes = 0, tid_begin = 0, tid_end = 0;
  double gather_s = 0;
};

struct SampleResult {
  std::string stoit_name;
  std::shared_ptr<Header> hdr = std::make_shared<Header>();  // shared with the session's header cache
  const Header& header() const { return *hdr; }
  const cmb_contig_stats* rows = nullptr;  // n_ref rows in the session's page-locked buffer; valid until the next process()
  std::vector<cmb_hist_pair> pairs;
  uint64_t num_detected_primary_alignments = 0;  // bam_generator.rs:113-119 / filter.rs:94-96,129-131
  uint64_t n_records = 0;                        // every record read from the file
  SampleTiming timing;
  // gene mode (--gff): rows are per gene (genes->entries order)
  std::shared_ptr<const ResolvedGenes> genes;
  std::vector<uint8_t> contig_seen;  // per tid: a record that passed the filters mapped here
  uint64_t kept_primary = 0;         // primaries among those records (genes.rs:249-252)
};

[[noreturn]] inline void throw_device_error(cmb_ctx* ctx, int rc) {
  const std::string msg = cmb_last_error(ctx);
  if (rc == CMB_E_UNSORTED || rc == CMB_E_NM || rc == CMB_E_BOUNDS) throw Panic(msg);
  throw ExitError(1, "device error " + std::to_string(rc) + ": " + msg);
}

// Converts SAM text to BAM record bytes so that one decoder serves both (htslib reads SAM through the same
// bam::Reader; tests/data/mapq_test.sam).  Integer aux tags get htslib's smallest-fitting type.
class SamToBam {
 public:
  static bool looks_like_sam(const uint8_t* p, size_t n) { return n == 0 || p[0] == '@' || (n > 4 && memcmp(p, "BAM\1", 4) != 0); }
  static void convert(const uint8_t* p, size_t n, Header& header, std::vector<uint8_t>& out) {
    std::unordered_map<std::string, int32_t> name_to_tid;
    size_t o = 0;
    auto next_line = [&](std::string& line) {
      if (o >= n) return false;
      size_t e = o;
      while (e < n && p[e] != '\n') ++e;
      line.assign((const char*)p + o, e - o);
      o = e < n ? e + 1 : e;
      if (!line.empty() && line.back() == '\r') line.pop_back();
      return true;
    };
    auto split = [](const std::string& s) {
      std::vector<std::string> v;
      size_t a = 0;
      for (;;) {
        size_t b = s.find('\t', a);
        if (b == std::string::npos) { v.push_back(s.substr(a)); break; }
        v.push_back(s.substr(a, b - a));
        a = b + 1;
      }
      return v;
    };
    out.assign({'B', 'A', 'M', 1});
    std::string line;
    std::vector<uint8_t> body;
    auto put32 = [](std::vector<uint8_t>& v, uint32_t x) { for (int k = 0; k < 4; ++k) v.push_back((x >> (8 * k)) & 0xff); };
    auto put16 = [](std::vector<uint8_t>& v, uint32_t x) { v.push_back(x & 0xff); v.push_back((x >> 8) & 0xff); };
    bool header_done = false;
    std::vector<uint8_t> recs;
    while (next_line(line)) {
      if (line.empty()) continue;
      if (line[0] == '@' && !header_done) {
        if (line.compare(0, 3, "@SQ") == 0) {
          std::string sn;
          uint64_t ln = 0;
          for (auto& f : split(line)) {
            if (f.compare(0, 3, "SN:") == 0) sn = f.substr(3);
            if (f.compare(0, 3, "LN:") == 0) ln = strtoull(f.c_str() + 3, nullptr, 10);
          }
          name_to_tid[sn] = (int32_t)header.names.size();
          header.names.push_back(sn);
          header.lens.push_back(ln);
        }
        continue;
      }
      header_done = true;
      auto f = split(line);
      if (f.size() < 11) throw Panic("Error reading BAM record: malformed SAM line");
      auto tid_of = [&](const std::string& s) -> int32_t {
        if (s == "*") return -1;
        auto it = name_to_tid.find(s);
        if (it == name_to_tid.end()) throw Panic("Error reading BAM record: unknown reference " + s);
        return it->second;
      };
      const int32_t tid = tid_of(f[2]);
      std::vector<uint32_t> cigar;
      if (f[5] != "*") {
        const char* c = f[5].c_str();
        while (*c) {
          char* e;
          const uint32_t len = (uint32_t)strtoul(c, &e, 10);
          static const char* ops = "MIDNSHP=X";
          const char* w = *e ? strchr(ops, *e) : nullptr;
          if (!w) throw Panic("Error reading BAM record: bad CIGAR");
          cigar.push_back((len << 4) | (uint32_t)(w - ops));
          c = e + 1;
        }
      }
      const uint32_t l_seq = f[9] == "*" ? 0 : (uint32_t)f[9].size();
      body.clear();
      put32(body, (uint32_t)tid);
      put32(body, (uint32_t)((int32_t)strtol(f[3].c_str(), nullptr, 10) - 1));
      body.push_back((uint8_t)std::min<size_t>(255, f[0].size() + 1));
      body.push_back((uint8_t)strtoul(f[4].c_str(), nullptr, 10));
      put16(body, 0);
      put16(body, (uint32_t)cigar.size());
      put16(body, (uint32_t)strtoul(f[1].c_str(), nullptr, 10));
      put32(body, l_seq);
      put32(body, (uint32_t)(f[6] == "=" ? tid : tid_of(f[6])));
      put32(body, (uint32_t)((int32_t)strtol(f[7].c_str(), nullptr, 10) - 1));
      put32(body, (uint32_t)strtol(f[8].c_str(), nullptr, 10));
      body.insert(body.end(), f[0].begin(), f[0].begin() + std::min<size_t>(254, f[0].size()));
      body.push_back(0);
      for (uint32_t cg : cigar) put32(body, cg);
      body.insert(body.end(), (l_seq + 1) / 2 + l_seq, 0);
      for (size_t i = 11; i < f.size(); ++i) {
        if (f[i].size() < 5 || f[i][2] != ':' || f[i][4] != ':') continue;
        if (f[i][0] != 'N' || f[i][1] != 'M') continue;  // only NM matters on this path
        body.push_back('N');
        body.push_back('M');
        if (f[i][3] == 'i') {
          const long long v = strtoll(f[i].c_str() + 5, nullptr, 10);
          if (v < 0) { body.push_back('i'); put32(body, (uint32_t)(int32_t)v); }
          else if (v <= 0xff) { body.push_back('C'); body.push_back((uint8_t)v); }
          else if (v <= 0xffff) { body.push_back('S'); put16(body, (uint32_t)v); }
          else { body.push_back('I'); put32(body, (uint32_t)v); }
        } else {
          body.push_back('A');
          body.push_back('?');
        }
      }
      put32(recs, (uint32_t)body.size());
      recs.insert(recs.end(), body.begin(), body.end());
    }
    put32(out, 0);  // l_text
    put32(out, (uint32_t)header.names.size());
    for (size_t i = 0; i < header.names.size(); ++i) {
      put32(out, (uint32_t)header.names[i].size() + 1);
      out.insert(out.end(), header.names[i].begin(), header.names[i].end());
      out.push_back(0);
      put32(out, (uint32_t)header.lens[i]);
    }
    out.insert(out.end(), recs.begin(), recs.end());
    header = Header{};  // re-parsed from the BAM bytes by the caller
  }
};

class DeviceSession {
 public:
  DeviceSession(int device, int threads, uint32_t batch_records = 1u << 20) : pool_(threads) {
    if (const char* e = getenv("CMB_BATCH_RECORDS")) batch_records = std::max(40000u, (uint32_t)strtoul(e, nullptr, 10));
    cmb_device_cfg cfg{};
    cfg.device = device;
    cfg.batch_records = batch_records;
    cfg.batch_intervals = batch_records + batch_records / 2;
    cfg.n_staging = 4;
    int rc = cmb_create(&cfg, &ctx_);
    if (rc != CMB_OK) throw ExitError(1, std::string("cannot create the CUDA coverage context: ") + cmb_last_error(nullptr));
    batch_records_ = cfg.batch_records;
    batch_intervals_ = cfg.batch_intervals;
    n_staging_ = cfg.n_staging;
  }
  ~DeviceSession() {
    cmb_host_free(rows_buf_);
    cmb_destroy(ctx_);
  }
  DeviceSession(const DeviceSession&) = delete;
  ThreadPool& pool() { return pool_; }
  cmb_ctx* ctx() { return ctx_; }

  // Restrict this session to the contig shard [begin, end) (multi-GPU); (0, UINT32_MAX) = everything.
  void set_shard(uint32_t begin, uint32_t end) { shard_begin_ = begin; shard_end_ = end; ref_lens_.clear(); }

  // Per-gene coverage: the following samples report one row per gene of `defs` (resolved against each sample's header)
  // instead of one per contig; nullptr returns to contig rows.
  void set_gene_definitions(const GeneDefinitions* defs, const GenomeNamer* namer) {
    gene_defs_ = defs;
    gene_namer_ = namer;
    ref_lens_.clear();
    gene_cache_.reset();
  }

  // Makes this session rank `rank` of `n_ranks` that process every sample TOGETHER (contigs range-partitioned by summed
  // length, each rank decoding only its BGZF block range, one gather of the per-contig table).  The ranks
  // exchange either over NCCL inside the device library (`nccl_id` from cmb_comm_unique_id, shared by the caller) or, for
  // hosts without NCCL between them (MPI, gloo, tests), through the caller's own all-gather of host buffers.
  typedef int (*AllGatherFn)(void* user, const void* send, size_t bytes_per_rank, void* recv);
  void set_group(int rank, int n_ranks, const uint8_t* nccl_id, AllGatherFn fn, void* user) {
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) throw ExitError(1, "set_group: bad rank / group size");
    if (n_ranks > 1 && !nccl_id && !fn) throw ExitError(1, "set_group: a group needs an NCCL id or an all-gather callback");
    if (group_nccl_) cmb_comm_destroy(ctx_);
    group_rank_ = rank;
    group_n_ = n_ranks;
    group_nccl_ = false;
    group_fn_ = fn;
    group_user_ = user;
    if (n_ranks > 1 && nccl_id) {
      const int rc = cmb_comm_init(ctx_, nccl_id, rank, n_ranks);
      if (rc) throw ExitError(1, std::string("cannot create the NCCL communicator: ") + cmb_last_error(ctx_));
      group_nccl_ = true;
    }
    ref_lens_.clear();
  }
  // cmb_comm_init_local was called on this session's context by the owner of all the ranks (one process, several GPUs)
  void adopt_local_group(int rank, int n_ranks) {
    group_rank_ = rank;
    group_n_ = n_ranks;
    group_nccl_ = n_ranks > 1;
    group_fn_ = nullptr;
    ref_lens_.clear();
  }
  int group_rank() const { return group_rank_; }
  int group_size() const { return group_n_; }
  // After the gather every rank holds the complete table, but only one needs to turn it into text: rank 0 replays the
  // estimators and prints, the others return an empty table -- unless every_rank_prints (tests: proves the gather is complete
  // everywhere).
  void set_every_rank_prints(bool v) { every_rank_prints_ = v; }
  bool is_output_rank() const { return group_n_ <= 1 || group_rank_ == 0 || every_rank_prints_; }

  // One sample.  In a group every rank must call this for the same input (it is collective).
  SampleResult process(const InputSpec& in, const cmb_params& params) {
    if (group_n_ <= 1) return process_local(in, params, nullptr);
    // ---- local phase: this rank's contigs, from this rank's block range.  Nothing may escape before the ranks have
    //      compared notes: a rank that failed still takes part in the exchange, and then every rank fails the same way.
    ShardState sh;
    SampleResult res;
    RankSummary mine{};
    try {
      res = process_local(in, params, &sh);
      mine.n_records = res.n_records;
      mine.n_primary = res.num_detected_primary_alignments;
      mine.counts_global = sh.counts_global ? 1 : 0;
      mine.n_pairs = sh.n_pairs;
      int32_t lo = INT32_MAX, hi = INT32_MIN;
      if (!sh.counts_global) cmb_kept_tid_range(ctx_, &lo, &hi);
      mine.min_tid = lo;
      mine.max_tid = hi;
    } catch (const Panic& e) {
      mine.kind = 1;
      mine.code = 101;
      snprintf(mine.message, sizeof mine.message, "%s", e.what());
    } catch (const ExitError& e) {
      mine.kind = 2;
      mine.code = e.code;
      snprintf(mine.message, sizeof mine.message, "%s", e.what());
    } catch (const std::exception& e) {
      mine.kind = 1;
      mine.code = 101;
      snprintf(mine.message, sizeof mine.message, "%s", e.what());
    }
    HostRange nvtx_gather("host: rank summaries + table gather");
    const double t_g0 = now_s();
    std::vector<RankSummary> all((size_t)group_n_);
    group_allgather(&mine, all.data(), sizeof(RankSummary));
    for (const RankSummary& s : all) {  // the lowest failing rank's error, on every rank
      if (s.kind == 1) throw Panic(s.message);
      if (s.kind == 2) throw ExitError(s.code, s.message);
    }
    // cross-rank half of the sortedness check (contig.rs:129-132): each rank has verified its own (overlapping) stretch of
    // the stream; the kept tids of the ranks' exclusive shares must not decrease from rank to rank either
    {
      int64_t seen_max = INT64_MIN;
      for (const RankSummary& s : all) {
        if (s.counts_global || s.min_tid > s.max_tid) continue;
        if ((int64_t)s.min_tid < seen_max)
          throw Panic("BAM file appears to be unsorted. Input BAM files must be sorted by reference (i.e. by samtools sort)");
        seen_max = std::max<int64_t>(seen_max, s.max_tid);
      }
    }
    // ---- the gather of the path: every rank ends up with the complete per-contig table (+ histogram pairs)
    const uint32_t n_ref = (uint32_t)res.hdr->names.size();
    std::vector<uint64_t> pair_base((size_t)group_n_ + 1, 0);
    for (int r = 0; r < group_n_; ++r) pair_base[(size_t)r + 1] = pair_base[(size_t)r] + all[(size_t)r].n_pairs;
    const bool csr = params.want & CMB_WANT_HIST_CSR;
    res.pairs.clear();
    if (csr) res.pairs.resize(pair_base[(size_t)group_n_]);
    ensure_rows(n_ref);
    if (group_nccl_) {
      const int rc = cmb_allgather_stats(ctx_, sh.cuts.data(), csr ? pair_base.data() : nullptr, rows_buf_, csr ? res.pairs.data() : nullptr);
      if (rc) throw_device_error(ctx_, rc);
    } else {
      gather_rows_through_host(sh, res, all, pair_base, csr);
    }
    res.rows = rows_buf_;
    // whole-file counters: summed over the ranks' owned records -- or taken from a rank that had to read the whole file
    uint64_t n_rec = 0, n_pri = 0;
    bool have_global = false;
    for (const RankSummary& s : all) {
      if (s.counts_global) {
        if (!have_global) {
          n_rec = s.n_records;
          n_pri = s.n_primary;
          have_global = true;
        }
      } else if (!have_global) {
        n_rec += s.n_records;
        n_pri += s.n_primary;
      }
    }
    res.n_records = n_rec;
    res.num_detected_primary_alignments = n_pri;
    res.timing.gather_s = now_s() - t_g0;
    res.timing.total_s += res.timing.gather_s;
    res.timing.group_ranks = (uint32_t)group_n_;
    return res;
  }

 private:
  struct ShardState {
    std::vector<uint32_t> cuts;
    bool counts_global = false;  // this rank read the whole file (host decode): its counters cover every record
    uint64_t n_pairs = 0;
  };
  struct RankSummary {  // what the ranks tell each other before the table gather (fixed size: it travels by all-gather)
    int32_t kind;       // 0 fine, 1 Panic, 2 ExitError
    int32_t code;
    uint64_t n_records, n_primary, n_pairs;
    int32_t min_tid, max_tid;
    uint32_t counts_global, reserved;
    char message[208];
  };

  void group_allgather(const void* send, void* recv, size_t bytes) {
    if (group_nccl_) {
      const int rc = cmb_comm_allgather(ctx_, send, recv, bytes);
      if (rc) throw ExitError(1, std::string("all-gather over NCCL failed: ") + cmb_last_error(ctx_));
    } else if (group_fn_(group_user_, send, bytes, recv) != 0) {
      throw ExitError(1, "the caller's all-gather failed");
    }
  }

  void ensure_rows(uint32_t n_ref) {
    if (rows_cap_ < (size_t)n_ref + 1) {
      cmb_host_free(rows_buf_);
      rows_cap_ = 0;
      rows_buf_ = (cmb_contig_stats*)cmb_host_alloc(sizeof(cmb_contig_stats) * ((size_t)n_ref + 1));
      if (!rows_buf_) throw ExitError(1, "cannot allocate the page-locked result buffer");
      rows_cap_ = (size_t)n_ref + 1;
    }
  }

  // Table gather without NCCL: own row range (and own pairs, offsets made global) through the caller's all-gather, padded
  // to the largest share.
  void gather_rows_through_host(const ShardState& sh, SampleResult& res, const std::vector<RankSummary>& all,
                                const std::vector<uint64_t>& pair_base, bool csr) {
    const int N = group_n_, me = group_rank_;
    size_t max_rows = 0, max_pairs = 0;
    for (int r = 0; r < N; ++r) {
      max_rows = std::max<size_t>(max_rows, sh.cuts[(size_t)r + 1] - sh.cuts[(size_t)r]);
      max_pairs = std::max<size_t>(max_pairs, all[(size_t)r].n_pairs);
    }
    const uint32_t b = sh.cuts[(size_t)me], e = sh.cuts[(size_t)me + 1];
    if (max_rows) {
      std::vector<cmb_contig_stats> send(max_rows), recv(max_rows * (size_t)N);
      for (uint32_t t = b; t < e; ++t) {
        send[t - b] = rows_buf_[t];
        if (csr && send[t - b].hist_count) send[t - b].hist_offset += pair_base[(size_t)me];
      }
      group_allgather(send.data(), recv.data(), max_rows * sizeof(cmb_contig_stats));
      for (int r = 0; r < N; ++r)
        for (uint32_t t = sh.cuts[(size_t)r]; t < sh.cuts[(size_t)r + 1]; ++t) rows_buf_[t] = recv[(size_t)r * max_rows + (t - sh.cuts[(size_t)r])];
    }
    if (csr && max_pairs) {
      std::vector<cmb_hist_pair> send(max_pairs), recv(max_pairs * (size_t)N);
      std::copy(local_pairs_.begin(), local_pairs_.end(), send.begin());
      group_allgather(send.data(), recv.data(), max_pairs * sizeof(cmb_hist_pair));
      for (int r = 0; r < N; ++r)
        std::copy(recv.begin() + (ptrdiff_t)((size_t)r * max_pairs), recv.begin() + (ptrdiff_t)((size_t)r * max_pairs + all[(size_t)r].n_pairs),
                  res.pairs.begin() + (ptrdiff_t)pair_base[(size_t)r]);
    }
  }

  SampleResult process_local(const InputSpec& in, const cmb_params& params, ShardState* shard) {
    SampleResult res;
    HostRange nvtx_sample("host: sample");
    HostRange nvtx_header("host: open + BAM header");
    const double t0 = now_s();
    res.stoit_name = file_stem(in.path);
    ByteSource bytes(in);
    std::vector<uint8_t> sam_as_bam;
    const uint8_t* p = bytes.data();
    size_t n = bytes.size();
    if (!(n >= 2 && p[0] == 0x1f && p[1] == 0x8b) && SamToBam::looks_like_sam(p, n)) {
      SamToBam::convert(p, n, *res.hdr, sam_as_bam);
      p = sam_as_bam.data();
      n = sam_as_bam.size();
    }
    int rc;
    cmb_filter_mode mode{};
    rc = cmb_set_params(ctx_, &params, &mode);
    if (rc) throw_device_error(ctx_, rc);
    const bool pair_mode = mode.filter_pairs;
    // only the header is read through this stream, unless the host's mate-matching fallback below needs the records too
    InflateStream stream(p, n, pool_, 1u << 20);
    std::vector<uint8_t> buf;
    size_t begin = 0;  // first unconsumed byte of buf
    auto need = [&](size_t bytes_needed) {  // make buf[begin, begin+bytes_needed) available; false at EOF
      while (buf.size() - begin < bytes_needed) {
        if (begin) {
          buf.erase(buf.begin(), buf.begin() + (ptrdiff_t)begin);
          begin = 0;
        }
        if (!stream.fill(buf)) return false;
      }
      return true;
    };
    // ---- header (SAMv1 §4.2).  Samples mapped to the same reference carry byte-identical headers up to the first record:
    //      the parsed copy of the previous sample is reused then (500 000 names are not rebuilt per sample).
    uint64_t records_at = 0;
    uint32_t n_ref = 0;
    // Fastest case: the file starts with the very same COMPRESSED bytes as the previous sample's header did (the same file
    // again, or samples written by one pipeline): nothing is inflated on the host at all.
    bool hdr_fast = false;
    if (hdr_cache_ && stream.is_bgzf() && !hdr_comp_.empty() && n >= hdr_comp_.size() && memcmp(p, hdr_comp_.data(), hdr_comp_.size()) == 0) {
      hdr_fast = true;
      res.hdr = hdr_cache_;
      n_ref = (uint32_t)res.hdr->names.size();
      records_at = hdr_records_at_;
    } else {
    if (!need(12) || memcmp(buf.data() + begin, "BAM\1", 4) != 0) throw Panic("Error reading BAM header: not a BAM/SAM file: " + in.path);
    const uint32_t l_text = rd_u32(buf.data() + begin + 4);
    if (!need(12 + (size_t)l_text)) throw Panic("Error reading BAM header: truncated");
    const size_t refs_at = 8 + (size_t)l_text;  // the n_ref field; the @-lines before it (e.g. @PG) may differ between samples
    if (hdr_cache_ && hdr_raw_.size() >= 4 && need(refs_at + hdr_raw_.size()) &&
        memcmp(buf.data() + begin + refs_at, hdr_raw_.data(), hdr_raw_.size()) == 0) {
      res.hdr = hdr_cache_;
      n_ref = (uint32_t)res.hdr->names.size();
      records_at = refs_at + hdr_raw_.size();
    } else {
      n_ref = rd_u32(buf.data() + begin + refs_at);
      size_t o = refs_at + 4;
      res.hdr->names.reserve(n_ref);
      res.hdr->lens.reserve(n_ref);
      for (uint32_t i = 0; i < n_ref; ++i) {
        if (!need(o + 4)) throw Panic("Error reading BAM header: truncated");
        const uint32_t l_name = rd_u32(buf.data() + begin + o);
        if (!need(o + 8 + l_name)) throw Panic("Error reading BAM header: truncated");
        res.hdr->names.emplace_back((const char*)buf.data() + begin + o + 4, l_name ? l_name - 1 : 0);
        res.hdr->lens.push_back(rd_u32(buf.data() + begin + o + 4 + l_name));
        o += 8 + l_name;
      }
      records_at = o;  // uncompressed offset of the first record (nothing has been discarded yet)
      hdr_raw_.assign(buf.data() + begin + refs_at, buf.data() + begin + o);
      hdr_cache_ = res.hdr;
    }
    hdr_comp_.clear();  // refreshed below, once the block index is known
    begin += records_at;
    }
    res.timing.header_s = now_s() - t0;
    nvtx_header.end();

    // ---- device reference + params
    uint32_t sb = std::min<uint32_t>(shard_begin_, n_ref), se = std::min<uint32_t>(shard_end_, n_ref);
    if (shard) {
      shard->cuts = tid_cuts_by_length(res.hdr->lens, group_n_);
      sb = shard->cuts[(size_t)group_rank_];
      se = shard->cuts[(size_t)group_rank_ + 1];
    }
    res.timing.tid_begin = sb;
    res.timing.tid_end = se;
    uint32_t n_rows = n_ref;
    if (gene_defs_) {
      if (shard) throw ExitError(1, "--gff is not available together with --gpus (per-gene coverage runs on one GPU)");
      if (!gene_cache_ || res.hdr->lens != ref_lens_ || res.hdr->names != gene_cache_names_) {
        gene_cache_ = std::make_shared<ResolvedGenes>(resolve_genes_against_header(*gene_defs_, *res.hdr, gene_namer_));
        gene_cache_names_ = res.hdr->names;
        std::vector<cmb_gene> genes(gene_cache_->entries.size());
        for (size_t g = 0; g < genes.size(); ++g) genes[g] = cmb_gene{gene_cache_->entries[g].tid, gene_cache_->entries[g].start, gene_cache_->entries[g].end};
        rc = cmb_set_genes(ctx_, n_ref, res.hdr->lens.data(), (uint32_t)genes.size(), genes.data());
        if (rc) throw_device_error(ctx_, rc);
        ref_lens_ = res.hdr->lens;
      }
      res.genes = gene_cache_;
      n_rows = std::max<uint32_t>(1, (uint32_t)gene_cache_->entries.size());
    } else if (res.hdr->lens != ref_lens_ || sb != ref_sb_ || se != ref_se_) {
      ref_sb_ = sb;
      ref_se_ = se;
      rc = cmb_set_reference(ctx_, n_ref, res.hdr->lens.data(), sb, se);
      if (rc) throw_device_error(ctx_, rc);
      ref_lens_ = res.hdr->lens;
    }
    rc = cmb_begin_sample(ctx_);
    if (rc) throw_device_error(ctx_, rc);

    // ---- records
    cmb_read_batch batch{};
    bool have_batch = false;
    uint32_t used_r = 0, used_i = 0;
    double wait_s = 0;
    auto acquire = [&]() {
      const double a = now_s();
      int r2 = cmb_acquire_batch(ctx_, &batch);
      wait_s += now_s() - a;
      if (r2) throw_device_error(ctx_, r2);
      have_batch = true;
      used_r = used_i = 0;
    };
    auto submit = [&]() {
      if (!have_batch) return;
      batch.iv_begin[used_r] = used_i;
      const double a = now_s();
      int r2 = cmb_submit_batch(ctx_, used_r, used_i);
      wait_s += now_s() - a;
      if (r2) throw_device_error(ctx_, r2);
      have_batch = false;
    };
    std::vector<size_t> rec_off;
    constexpr size_t ITEM = 4096;  // records per parallel work item
    struct ItemOut {
      std::vector<int32_t> iv_start, iv_len;
      uint64_t primaries = 0;
    };
    std::vector<ItemOut> items;
    std::vector<Tuple> tuples;  // pair mode only
    // mate matching state (filter.rs:16-18)
    struct Stored {
      Tuple t;
      std::vector<int32_t> iv_start, iv_len;
    };
    std::map<std::string, Stored> first_set;
    int32_t current_reference = -1;

    auto put_record = [&](const Tuple& t, const int32_t* ivs, const int32_t* ivl) {
      const uint32_t i = used_r++;
      batch.tid[i] = t.tid; batch.pos[i] = t.pos; batch.flag[i] = t.flag; batch.mapq[i] = t.mapq;
      batch.nm_state[i] = t.nm_state; batch.nm[i] = t.nm; batch.l_seq[i] = t.l_seq; batch.aligned[i] = t.aligned;
      batch.del[i] = t.del; batch.ins[i] = t.ins; batch.iv_begin[i] = used_i;
      for (uint32_t k = 0; k < t.n_iv; ++k) {
        batch.iv_start[used_i] = ivs[k];
        batch.iv_len[used_i] = ivl[k];
        ++used_i;
      }
    };

    bool decoded_on_device = false;
    {
      // region-parallel pipeline (decode_runner.hpp): this thread only acquires / submits staging batches
      HostRange nvtx_index("host: BGZF block index (+ range probes in a group)");
      const double t_index0 = now_s();
      BlockIndex bx;
      if (stream.is_raw()) bx.build(stream.raw_data(), stream.raw_size());
      else bx.build(p, n);
      res.timing.index_s = now_s() - t_index0;
      if (!hdr_fast && bx.bgzf && !stream.is_raw()) {
        // htslib flushes the BGZF block after the header (bam_hdr_write), so the records usually start a block: then the
        // compressed bytes in front of that block ARE the header, and the next sample that begins with the same bytes needs no
        // header inflate at all.
        const size_t b = (size_t)(std::lower_bound(bx.ustart.begin(), bx.ustart.end(), records_at) - bx.ustart.begin());
        if (b > 0 && b < bx.blocks.size() && bx.ustart[b] == records_at && bx.blocks[b].cdata >= 18) {
          const size_t start = bx.blocks[b].cdata - 18;
          if (start <= (256u << 20) && p[start] == 0x1f && p[start + 1] == 0x8b) {
            hdr_comp_.assign(p, p + start);
            hdr_records_at_ = records_at;
          }
        }
      }
      // Device-side decode first (compressed blocks cross PCIe, the GPU inflates and parses them); the host pipeline
      // below runs when the input is not BGZF, when CMB_HOST_DECODE is set, or when the device declines the stream.  Pair
      // filtering included: the device matches mates itself (cmb_pairs.cuh); the host's BTreeMap-style matching further
      // down is the fallback.
      if (bx.bgzf && !stream.is_raw() && !getenv("CMB_HOST_DECODE")) {
        const size_t nb = bx.blocks.size();
        std::vector<uint64_t> coff(nb);
        std::vector<uint32_t> clen(nb), isz(nb);
        for (size_t b = 0; b < nb; ++b) {
          coff[b] = bx.blocks[b].cdata;
          clen[b] = (uint32_t)bx.blocks[b].clen;
          isz[b] = bx.blocks[b].isize;
        }
        cmb_bgzf_input bi{};
        bi.data = p;
        bi.size = n;
        bi.n_blocks = (uint32_t)nb;
        bi.n_ref = n_ref;
        bi.block_coffset = coff.data();
        bi.block_clen = clen.data();
        bi.block_isize = isz.data();
        bi.records_at = records_at;
        bi.copy_threads = (uint32_t)std::min(pool_.size(), 8);
        bool range_ok = true;
        if (shard) {  // this rank's block range (shard_range.hpp); a stream whose alignment cannot be confirmed is read whole
          try {
            BlockRangeFinder finder(bx, n_ref, records_at);
            const bool last = group_rank_ == group_n_ - 1;
            // reads cover the reference roughly evenly, so a tid's records start near its share of the summed contig length
            double frac_lo = -1.0, frac_hi = -1.0;
            {
              long double before_lo = 0, before_hi = 0, total = 0;
              const auto& lens = res.hdr->lens;
              for (size_t t = 0; t < lens.size(); ++t) {
                if (t == sb) before_lo = total;
                if (t == se) before_hi = total;
                total += (long double)lens[t];
              }
              if (se >= lens.size()) before_hi = total;
              if (total > 0) {
                frac_lo = (double)(before_lo / total);
                frac_hi = (double)(before_hi / total);
              }
            }
            const BlockRange br = finder.find(sb, se, group_rank_ == 0, last, frac_lo, frac_hi);
            bi.ranged = 1;
            bi.walk_begin_block = br.walk_begin;
            bi.walk_end_block = br.walk_end;
            bi.records_at = br.records_at;
            bi.excl_end_block = br.excl_end;
            bi.own_tid_begin = (int32_t)sb;
            bi.own_tid_end = (int32_t)se;
            bi.own_unplaced = last ? 1u : 0u;
            res.timing.shard_blocks = br.walk_end - br.walk_begin;
            res.timing.total_blocks = (uint32_t)nb;
            res.timing.range_probes = br.probes;
          } catch (const Panic&) {
            range_ok = false;
          }
        }
        cmb_bgzf_result br{};
        nvtx_index.end();
        const double a = now_s();
        res.timing.index_s = now_s() - t_index0;  // incl. the block table and, in a group, the range probes
        const int r2 = range_ok ? cmb_submit_bgzf(ctx_, &bi, &br) : CMB_E_DECLINED;
        res.timing.device_call_s = now_s() - a;
        if (r2 == CMB_OK) {
          decoded_on_device = true;
          res.timing.device_decode = true;
          res.timing.bgzf = br;
          res.timing.h2d_bytes = br.h2d_bytes;
          res.timing.decode_launches = br.n_launches;
          res.n_records = br.n_records;
          res.num_detected_primary_alignments = br.n_primary;
          if (getenv("CMB_PIPELINE_STATS"))
            fprintf(stderr, "#device_decode\tblocks=%zu\thost_blocks=%u\trepairs=%u\tcopy_inflate_ms=%.2f\tchain_ms=%.2f\textract_ms=%.2f\ttotal_ms=%.2f\tcall_s=%.4f\n",
                    nb, br.n_blocks_host, br.chain_repairs, br.ms_copy_inflate, br.ms_chain, br.ms_extract, br.ms_total, now_s() - a);
        } else if (r2 != CMB_E_DECLINED) {
          throw_device_error(ctx_, r2);
        } else if (getenv("CMB_PIPELINE_STATS")) {
          fprintf(stderr, "#device_decode\tdeclined: %s\n", cmb_last_error(ctx_));
        }
      }
      if (!decoded_on_device && !pair_mode) {
      if (shard) shard->counts_global = true;  // every rank's host decoder reads the whole file; K1 keeps the rank's own tids
      const PipelineCounts pc = run_decode_pipeline(
          bx, records_at, n_ref, pool_.size(), batch_records_, batch_intervals_, n_staging_, scratch_,
          [&](cmb_read_batch* b) {
            const double a = now_s();
            int r2 = cmb_acquire_batch(ctx_, b);
            wait_s += now_s() - a;
            if (r2) throw_device_error(ctx_, r2);
          },
          [&](uint32_t nr, uint32_t ni) {
            const double a = now_s();
            int r2 = cmb_submit_batch(ctx_, nr, ni);
            wait_s += now_s() - a;
            if (r2) throw_device_error(ctx_, r2);
          });
      res.n_records = pc.n_records;
      res.num_detected_primary_alignments = pc.primaries;
      if (getenv("CMB_PIPELINE_STATS"))
        fprintf(stderr, "#pipeline\titems=%u\tworkers=%u\tinflate_s=%.3f\tchain_s=%.3f\textract_s=%.3f\tidle_s=%.3f (summed over workers)\n",
                pc.n_items, pc.n_workers, pc.inflate_s, pc.scan_s, pc.extract_s, pc.idle_s);
      }
    }
    if (pair_mode && !decoded_on_device && hdr_fast) {  // the header was recognised without inflating it: skip over it now
      if (!need((size_t)records_at)) throw Panic("Error reading BAM header: truncated");
      begin = (size_t)records_at;
    }
    if (pair_mode && !decoded_on_device) stream.set_window(48u << 20);  // mate matching is sequential: decode window by window
    if (pair_mode && !decoded_on_device) for (;;) {
      if (shard) shard->counts_global = true;  // mate matching reads the whole file on every rank
      // complete records currently in buf
      rec_off.clear();
      size_t q = begin;
      size_t max_iv = 0;
      while (q + 4 <= buf.size()) {
        const uint32_t bs = rd_u32(buf.data() + q);
        if (bs < 32) throw Panic("Error reading BAM record: corrupt block_size");
        if (q + 4 + (size_t)bs > buf.size()) break;
        rec_off.push_back(q);
        const int64_t ops = record_cigar_ops(buf.data() + q);
        if (ops < 0) throw_bad_record_layout();
        max_iv += (size_t)ops;
        q += 4 + (size_t)bs;
        if (rec_off.size() == batch_records_ / 2) break;  // keep one window within a batch
      }
      if (rec_off.empty()) {
        if (!need((buf.size() - begin) + 1)) break;  // EOF
        continue;
      }
      const size_t nrec = rec_off.size();
      res.n_records += nrec;
      const size_t n_items = (nrec + ITEM - 1) / ITEM;
      if (items.size() < n_items) items.resize(n_items);
      if (max_iv > batch_intervals_) throw ExitError(1, "a window of records has more aligned blocks than a device batch holds");

      {
        // filter.rs:117-233: decode in parallel, then match mates in stream order; only completed pairs reach the GPU
        tuples.resize(nrec);
        std::vector<uint32_t> iv_at(nrec);
        const uint8_t* base = buf.data();
        pool_.parallel_for(n_items, [&](size_t it, int) {
          ItemOut& io = items[it];
          io.iv_start.clear();
          io.iv_len.clear();
          const size_t r0 = it * ITEM, r1 = std::min(nrec, r0 + ITEM);
          for (size_t r = r0; r < r1; ++r) {
            iv_at[r] = (uint32_t)io.iv_start.size();
            decode_bam_record(base + rec_off[r], tuples[r], io.iv_start, io.iv_len);
          }
        });
        for (size_t r = 0; r < nrec; ++r) {
          const Tuple& t = tuples[r];
          const ItemOut& io = items[r / ITEM];
          const int32_t* ivs = io.iv_start.data() + iv_at[r];
          const int32_t* ivl = io.iv_len.data() + iv_at[r];
          if (!(t.flag & 0x900)) res.num_detected_primary_alignments += 1;
          if (t.flag & 0x900) continue;   // secondary / supplementary (filter.rs:138-140)
          if (!(t.flag & 0x2)) continue;  // not a proper pair (filter.rs:141-147, filter_out = true)
          if (t.tid != current_reference) {
            current_reference = t.tid;
            first_set.clear();
          }
          std::string qname = bam_qname(base + rec_off[r]);
          auto itf = first_set.find(qname);
          if (itf == first_set.end()) {
            if (t.mtid == current_reference) {
              Stored s;
              s.t = t;
              s.iv_start.assign(ivs, ivs + t.n_iv);
              s.iv_len.assign(ivl, ivl + t.n_iv);
              first_set.emplace(std::move(qname), std::move(s));
            }
          } else {
            const Stored& s = itf->second;
            if (have_batch && (used_r + 2 > batch_records_ || used_i + s.t.n_iv + t.n_iv > batch_intervals_)) submit();
            if (!have_batch) acquire();
            put_record(s.t, s.iv_start.data(), s.iv_len.data());  // stored first mate at the even index
            put_record(t, ivs, ivl);
            first_set.erase(itf);
          }
        }
      }
      begin = q;
    }
    const double t_dec = now_s();
    submit();
    ensure_rows(n_rows);
    res.rows = rows_buf_;
    uint64_t n_pairs = 0;
    // A histogram buffer of the device overflowed (very deep coverage over many small contigs): with the sample's tuples still in
    // device memory the buffers are enlarged and the kernels run again; otherwise the error stands.
    auto end_sample = [&](cmb_contig_stats* rows_out) {
      int r2 = cmb_end_sample(ctx_, rows_out, nullptr, 0, &n_pairs);
      for (int attempt = 0; r2 == CMB_E_CAPACITY && decoded_on_device && attempt < 8; ++attempt) {
        cmb_read_batch again{};
        uint32_t nr = 0, ni = 0;
        if (cmb_last_bgzf_batch(ctx_, &again, &nr, &ni) != CMB_OK) break;
        if (getenv("CMB_PIPELINE_STATS")) fprintf(stderr, "#capacity_retry\tattempt=%d\n", attempt + 1);
        if ((r2 = cmb_grow_buffers(ctx_)) != CMB_OK) break;
        if ((r2 = cmb_begin_sample(ctx_)) != CMB_OK) break;
        if ((r2 = cmb_submit_device_batch(ctx_, &again, nr, ni)) != CMB_OK) break;
        r2 = cmb_end_sample(ctx_, rows_out, nullptr, 0, &n_pairs);
      }
      return r2;
    };
    if (shard && group_nccl_) {
      // rows (and pairs) stay on the device: cmb_allgather_stats completes the table there and copies it back once
      rc = end_sample(nullptr);
      if (rc) throw_device_error(ctx_, rc);
      shard->n_pairs = n_pairs;
    } else {
      rc = end_sample(rows_buf_);
      if (rc) throw_device_error(ctx_, rc);
      if ((params.want & CMB_WANT_HIST_CSR) && n_pairs) {
        res.pairs.resize(n_pairs);
        rc = cmb_fetch_pairs(ctx_, res.pairs.data(), n_pairs);
        if (rc) throw_device_error(ctx_, rc);
      }
      if (shard) {
        shard->n_pairs = n_pairs;
        local_pairs_ = res.pairs;
      }
    }
    if (gene_defs_) {
      res.contig_seen.assign((size_t)n_ref + 1, 0);
      rc = cmb_fetch_gene_extras(ctx_, res.contig_seen.data(), &res.kept_primary);
      if (rc) throw_device_error(ctx_, rc);
    }
    cmb_get_timing(ctx_, &res.timing.device);
    if (!res.timing.device_decode) res.timing.h2d_bytes = 40ull * res.timing.device.n_records + 4 + 8ull * res.timing.device.n_intervals;
    const double t1 = now_s();
    res.timing.total_s = t1 - t0;
    res.timing.decode_s = t_dec - t0 - wait_s;
    res.timing.submit_wait_s = wait_s;
    res.timing.end_sample_s = t1 - t_dec;
    return res;
  }

  std::shared_ptr<Header> hdr_cache_;  // parsed reference list of the previous sample and its raw bytes (n_ref .. first record)
  std::vector<uint8_t> hdr_raw_;
  std::vector<uint8_t> hdr_comp_;  // the compressed file prefix the last full header parse consumed, and where its records start
  uint64_t hdr_records_at_ = 0;
  cmb_contig_stats* rows_buf_ = nullptr;
  size_t rows_cap_ = 0;
  ThreadPool pool_;
  DecodeScratch scratch_;
  cmb_ctx* ctx_ = nullptr;
  uint32_t batch_records_ = 0, batch_intervals_ = 0, n_staging_ = 0;
  uint32_t shard_begin_ = 0, shard_end_ = 0xffffffffu;
  std::vector<uint64_t> ref_lens_;
  uint32_t ref_sb_ = 0, ref_se_ = 0;
  const GeneDefinitions* gene_defs_ = nullptr;
  const GenomeNamer* gene_namer_ = nullptr;
  std::shared_ptr<ResolvedGenes> gene_cache_;
  std::vector<std::string> gene_cache_names_;
  // group (multi-GPU contig sharding)
  int group_rank_ = 0, group_n_ = 1;
  bool group_nccl_ = false, every_rank_prints_ = false;
  AllGatherFn group_fn_ = nullptr;
  void* group_user_ = nullptr;
  std::vector<cmb_hist_pair> local_pairs_;
};

}  // namespace cmbh
