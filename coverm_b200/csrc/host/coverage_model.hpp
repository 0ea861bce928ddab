// Host finalisation of libcoverm_b200: estimators that consume the GPU's per-contig integer statistics instead of a
// dense `&[i32]`, plus the reference's sink (CoverageTaker) and printer (CoveragePrinter) protocols.
//
// Mirrors, for drop-in behaviour (same names / argument meaning / error behaviour):
//   enum CoverageEstimator + trait MosdepthGenomeCoverageEstimator   mosdepth_genome_coverage_estimators.rs:3-1061 (EST)
//   trait CoverageTaker / enum CoverageTakerType + cached iterator    coverage_takers.rs:8-377
//   enum CoveragePrinter                                              coverage_printer.rs:9-553
// Every O(contig length) loop of EST::add_contig (EST:393-404, 447-465, 494-501) has already run on the GPU; what is
// left here is the O(1) / O(max depth) float arithmetic of calculate_coverage, replayed with the reference's exact
// f32/f64 operation order (build with -ffp-contract=off).
#pragma once
#include <charconv>
#include <cmath>
#include <optional>
#include <ostream>

#include "bam_source.hpp"

namespace cmbh {

// An output buffer that can hand out room for `n` more bytes at once, so that formatted blocks are copied into place in
// parallel (implemented by the in-memory sink of the C ABI; a plain ostream takes the sequential path).
struct BulkSink {
  virtual char* append_uninitialized(size_t n) = 0;
  virtual ~BulkSink() = default;
};

// Stage timings on stderr when CMB_HOST_STATS is set (profiling aid).
inline double host_now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
inline void host_stat(const char* what, double seconds) {
  static const bool on = getenv("CMB_HOST_STATS") != nullptr;
  if (on) fprintf(stderr, "#host_stat\t%s\t%.4f\n", what, seconds);
}

template <class F>
inline std::string rust_display(F v) {  // Rust `{}` for f32/f64: shortest round-trip digits, fixed notation
  if (std::isnan(v)) return "NaN";
  if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
  char buf[512];
  auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::fixed);
  return std::string(buf, r.ptr);
}

template <class F>
inline void append_display(std::string& out, F v) {  // rust_display appended in place (no temporary string)
  if (std::isnan(v)) { out += "NaN"; return; }
  if (std::isinf(v)) { out += v < 0 ? "-inf" : "inf"; return; }
  char buf[512];
  auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::fixed);
  out.append(buf, r.ptr);
}

struct ReadsMapped {  // lib.rs:53-57
  uint64_t num_mapped_reads = 0, num_reads = 0;
};

// ------------------------------------------------------------------------------------------------ takers
struct EntryAndCoverages {
  size_t entry_index, stoit_index;
  std::vector<float> coverages;
};

class CoverageTaker {  // coverage_takers.rs:29-38 (trait) over the three CoverageTakerType variants (:8-27)
 public:
  enum class Kind { SingleFloatStreaming, PileupStreaming, CachedSingleFloat };
  Kind kind;
  std::ostream* out = nullptr;
  size_t num_coverages = 0;
  std::vector<std::string> stoit_names;
  std::vector<std::optional<std::string>> entry_names;
  // Contig mode with many entries: the names are the BAM header's, shared instead of copied entry by entry.  While set,
  // entry_names stays empty; entry_name() serves both representations.
  std::shared_ptr<const std::vector<std::string>> shared_names;
  const std::string* entry_name(size_t idx) const {
    if (shared_names) return idx < shared_names->size() ? &(*shared_names)[idx] : nullptr;
    return idx < entry_names.size() && entry_names[idx] ? &*entry_names[idx] : nullptr;
  }
  // Adopt `names` for every entry index (true), or report that they conflict with what is already recorded (false: the
  // caller then goes through start_entry(), which raises the reference's error at the first differing entry).
  bool try_share_names(const std::shared_ptr<const std::vector<std::string>>& names) {
    if (kind != Kind::CachedSingleFloat) return false;
    if (!shared_names && entry_names.empty()) {
      shared_names = names;
      return true;
    }
    return shared_names && (shared_names == names || *shared_names == *names);
  }
  void start_entry_shared(size_t order_id) { cur_entry_index_ = order_id; }
  void unshare_names() {  // back to per-entry names (a later input has entries the shared list lacks)
    if (!shared_names) return;
    entry_names.assign(shared_names->begin(), shared_names->end());
    shared_names.reset();
  }
  struct Entry { size_t entry_index; float coverage; };
  std::vector<std::vector<Entry>> coverages;

  static CoverageTaker streaming(std::ostream* o) { CoverageTaker t; t.kind = Kind::SingleFloatStreaming; t.out = o; return t; }
  static CoverageTaker pileup(std::ostream* o) { CoverageTaker t; t.kind = Kind::PileupStreaming; t.out = o; return t; }
  static CoverageTaker cached(size_t n) { CoverageTaker t; t.kind = Kind::CachedSingleFloat; t.num_coverages = n; return t; }

  // Capacity hint (cached taker): n_entries entries of k coverages are about to be added to the current stoit.
  void reserve_entries(size_t n_entries, size_t k) {
    if (kind != Kind::CachedSingleFloat) return;
    if (!shared_names && entry_names.size() < n_entries) entry_names.resize(n_entries);
    coverages[cur_stoit_index_].reserve(coverages[cur_stoit_index_].size() + n_entries * k);
  }
  void start_stoit(const std::string& name) {
    if (kind == Kind::CachedSingleFloat) {
      stoit_names.push_back(name);
      coverages.emplace_back();
      cur_stoit_index_ = stoit_names.size() - 1;
    } else {
      cur_stoit_ = name;
    }
  }
  void start_entry(size_t order_id, const std::string& name) {
    if (kind == Kind::SingleFloatStreaming) {
      *out << cur_stoit_ << '\t' << name;
    } else if (kind == Kind::PileupStreaming) {
      cur_entry_ = name;
    } else {
      if (shared_names) {
        if (order_id < shared_names->size() && (*shared_names)[order_id] == name) {
          cur_entry_index_ = order_id;
          return;
        }
        if (order_id >= shared_names->size()) unshare_names();
        else
          throw ExitError(1, "Found a difference amongst the reference sets used for mapping. For this (non-streaming) usage of "
                             "CoverM, all BAM files must have the same set of reference sequences. Previous entry was " +
                                 (*shared_names)[order_id] + ", new is " + name);
      }
      if (order_id >= entry_names.size()) entry_names.resize(order_id + 1);
      if (!entry_names[order_id]) entry_names[order_id] = name;
      if (*entry_names[order_id] != name)
        throw ExitError(1, "Found a difference amongst the reference sets used for mapping. For this (non-streaming) usage of "
                           "CoverM, all BAM files must have the same set of reference sequences. Previous entry was " +
                               *entry_names[order_id] + ", new is " + name);
      cur_entry_index_ = order_id;
    }
  }
  void add_single_coverage(float c) {
    if (kind == Kind::SingleFloatStreaming) {
      if (c == 0.0f) *out << "\t0";  // coverage_takers.rs:162-166
      else *out << '\t' << rust_display(c);
    } else if (kind == Kind::CachedSingleFloat) {
      coverages[cur_stoit_index_].push_back({cur_entry_index_, c});
    } else {
      throw Panic("internal error: entered unreachable code");
    }
  }
  void add_coverage_entry(size_t num_reads, uint64_t num_bases) {
    if (kind != Kind::PileupStreaming) throw Panic("internal error: entered unreachable code");
    *out << cur_stoit_ << '\t' << cur_entry_ << '\t' << num_reads << '\t' << num_bases << '\n';
  }
  void finish_entry() {
    if (kind == Kind::SingleFloatStreaming) *out << '\n';
  }

  // What CoverageTakerTypeIterator (coverage_takers.rs:265-377) yields: stoit by stoit, every entry index that any
  // stoit recorded (ascending), zero-filled where this stoit has none.  A Cell points at the first of the
  // num_coverages values of (stoit, entry), or is null for "all zeros".
  struct Cell {
    size_t entry_index;
    const Entry* first;
    float at(size_t c) const { return first ? first[c].coverage : 0.0f; }
  };
  std::vector<std::vector<Cell>> cells_by_stoit() const {
    const size_t ns = stoit_names.size();
    std::vector<std::vector<Cell>> out(ns);
    if (ns == 1) {  // one sample: its own entries, in the order recorded (ascending entry index)
      const auto& mine = coverages[0];
      out[0].reserve(num_coverages ? mine.size() / num_coverages : 0);
      std::optional<size_t> last;
      for (size_t i = 0; i + num_coverages <= mine.size() && num_coverages; i += num_coverages) {
        if (last && mine[i].entry_index <= *last) break;  // the reference's iterator stops at a non-ascending entry
        out[0].push_back({mine[i].entry_index, &mine[i]});
        last = mine[i].entry_index;
      }
      return out;
    }
    for (size_t s = 0; s < ns; ++s) {  // k-way walk over the stoits' lists (each ascending in entry index)
      std::vector<size_t> head(ns, 0);
      std::optional<size_t> last;
      for (;;) {
        std::optional<size_t> lowest;
        for (size_t k = 0; k < ns; ++k) {
          if (head[k] >= coverages[k].size()) continue;
          const size_t e = coverages[k][head[k]].entry_index;
          if (last && e <= *last) continue;  // the reference only considers entries beyond the last one returned
          if (!lowest || e < *lowest) lowest = e;
        }
        if (!lowest) break;
        const auto& mine = coverages[s];
        const bool have = head[s] < mine.size() && mine[head[s]].entry_index == *lowest;
        out[s].push_back({*lowest, have ? &mine[head[s]] : nullptr});
        for (size_t k = 0; k < ns; ++k)
          if (head[k] < coverages[k].size() && coverages[k][head[k]].entry_index == *lowest) head[k] += num_coverages;
        last = *lowest;
      }
    }
    return out;
  }
  std::vector<EntryAndCoverages> entries_by_stoit() const {
    std::vector<EntryAndCoverages> v;
    const auto cells = cells_by_stoit();
    for (size_t s = 0; s < cells.size(); ++s)
      for (const Cell& c : cells[s]) {
        EntryAndCoverages ec;
        ec.entry_index = c.entry_index;
        ec.stoit_index = s;
        for (size_t j = 0; j < num_coverages; ++j) ec.coverages.push_back(c.at(j));
        v.push_back(std::move(ec));
      }
    return v;
  }

 private:
  std::string cur_stoit_, cur_entry_;
  size_t cur_stoit_index_ = 0, cur_entry_index_ = 0;
};

// ------------------------------------------------------------------------------------------------ estimators
// What add_contig receives in place of `ups_and_downs: &[i32]`.
struct ContigObservation {
  uint64_t len = 0;                    // ups_and_downs.len()
  uint64_t num_mapped_reads = 0;       // the driver's read count for this contig
  uint64_t total_mismatches = 0;       // total_edit_distance - total_indels (unchecked u64 subtraction, contig.rs:59)
  double sum_identity = 0.0;
  const cmb_contig_stats* stats = nullptr;  // nullptr == an all-zero array of length `len`
  const cmb_hist_pair* hist = nullptr;      // merged window histogram of this contig (CSR), when requested
  uint32_t n_hist = 0;
};

class CoverageEstimator {
 public:
  enum class Kind { Mean, TrimmedMean, PileupCounts, CoveredFraction, CoveredBases, RPKM, TPM, Variance, Length,
                    ReadCount, ReadsPerBase, ANIr };
  Kind kind;
  float min_fraction_covered_bases = 0.0f;
  uint64_t contig_end_exclusion = 0;
  bool exclude_mismatches = false;
  float min = 0.0f, max = 0.0f;

  static CoverageEstimator make(Kind k, float minfrac = 0.0f, uint64_t excl = 0) {
    CoverageEstimator e;
    e.kind = k;
    e.min_fraction_covered_bases = minfrac;
    e.contig_end_exclusion = excl;
    return e;
  }
  bool needs_histogram() const { return kind == Kind::TrimmedMean || kind == Kind::PileupCounts || kind == Kind::Variance; }

  std::vector<std::string> column_headers() const {  // EST:84-104
    switch (kind) {
      case Kind::Mean: return {"Mean"};
      case Kind::TrimmedMean: return {"Trimmed Mean"};
      case Kind::PileupCounts: return {"Coverage", "Bases"};
      case Kind::CoveredFraction: return {"Covered Fraction"};
      case Kind::CoveredBases: return {"Covered Bases"};
      case Kind::RPKM: return {"RPKM"};
      case Kind::TPM: return {"TPM"};
      case Kind::Variance: return {"Variance"};
      case Kind::Length: return {"Length"};
      case Kind::ReadCount: return {"Read Count"};
      case Kind::ReadsPerBase: return {"Reads per base"};
      case Kind::ANIr: return {"ANIr"};
    }
    return {};
  }

  void setup() {  // EST:268-364
    total_count_ = total_bases_ = covered_ = reads_ = mismatches_ = observed_len_ = 0;
    counts_.clear();
    have_counts_ = false;
    dev_valid_ = false;
    n_window_contigs_ = 0;
    sum_identity_ = 0.0;
  }

  void add_contig(const ContigObservation& ob) {  // EST:366-528
    const cmb_contig_stats* s = ob.stats;
    const uint64_t E = contig_end_exclusion;
    switch (kind) {
      case Kind::Mean:
        reads_ += ob.num_mapped_reads;
        mismatches_ += ob.total_mismatches;
        if (!(E * 2 < ob.len)) return;
        total_bases_ += ob.len - 2 * E;
        if (s) {
          covered_ += s->covered_window;
          total_count_ += s->sum_depth_window;
        }
        break;
      case Kind::TrimmedMean:
      case Kind::PileupCounts:
      case Kind::Variance: {
        reads_ = ob.num_mapped_reads;  // assignment, EST:434
        if (!(E * 2 < ob.len)) return;
        const uint64_t T = ob.len - 2 * E;
        observed_len_ += T;
        if (s) covered_ += s->covered_window;
        if (s && ob.hist) {  // merge the GPU's (depth,count) pairs: `counts[depth] += 1` per window base
          have_counts_ = true;
          for (uint32_t i = 0; i < ob.n_hist; ++i) {
            const cmb_hist_pair& p = ob.hist[i];
            if (counts_.size() <= p.depth) counts_.resize((size_t)p.depth + 1, 0);
            counts_[p.depth] += p.count;
          }
        } else if (s) {  // single-contig use (contig mode): keep the GPU's finished integers
          dev_valid_ = n_window_contigs_ == 0;
          dev_ = *s;
        } else {  // an all-zero contig: every window base has depth 0
          have_counts_ = true;
          if (counts_.empty()) counts_.resize(1, 0);
          counts_[0] += T;
        }
        ++n_window_contigs_;
        break;
      }
      case Kind::CoveredFraction:
      case Kind::CoveredBases:
      case Kind::RPKM:
      case Kind::TPM:
        reads_ += ob.num_mapped_reads;
        total_bases_ += ob.len;
        if (s) covered_ += s->covered_full;
        break;
      case Kind::Length:
      case Kind::ReadsPerBase:
        observed_len_ += ob.len;
        reads_ += ob.num_mapped_reads;
        break;
      case Kind::ReadCount: reads_ += ob.num_mapped_reads; break;
      case Kind::ANIr:
        reads_ += ob.num_mapped_reads;
        sum_identity_ += ob.sum_identity;
        break;
    }
  }

  float calculate_coverage(const std::vector<uint64_t>& unobserved_contig_lengths) {  // EST:530-839
    uint64_t unobs_plain = 0, unobs_trimmed = 0;
    for (uint64_t l : unobserved_contig_lengths) {
      unobs_plain += l;
      unobs_trimmed += (l < 2 * contig_end_exclusion) ? l : l - 2 * contig_end_exclusion;  // EST:226-242
    }
    auto below_min = [&](uint64_t total) { return ((float)covered_ / (float)total) < min_fraction_covered_bases; };
    switch (kind) {
      case Kind::Mean: {
        const uint64_t T = total_bases_ + unobs_trimmed;
        if (T == 0 || below_min(T)) return 0.0f;
        const float num = exclude_mismatches ? (float)(total_count_ - mismatches_) : (float)total_count_;
        return num / (float)T;
      }
      case Kind::TrimmedMean: {
        const uint64_t T = observed_len_ + unobs_trimmed;
        if (T == 0) return 0.0f;
        if (below_min(T)) return 0.0f;
        const size_t min_index = (size_t)std::floor(min * (float)T);
        const size_t max_index = (size_t)std::ceil(max * (float)T);
        if (covered_ == 0) return 0.0f;
        uint64_t total;
        if (use_device_walk()) {
          if (unobs_trimmed != 0 || dev_.trim_min_index != min_index || dev_.trim_max_index != max_index)
            throw Panic("internal error: device trimmed-mean walk used outside single-contig mode");
          total = dev_.trimmed_total;
        } else {
          if (counts_.empty()) throw Panic("index out of bounds: the len is 0 but the index is 0");
          counts_[0] += unobs_trimmed;
          total = trimmed_walk(counts_, min_index, max_index);
        }
        return (float)total / (float)(max_index - min_index);
      }
      case Kind::PileupCounts: {
        if (observed_len_ == 0) return 0.0f;
        const uint64_t T = observed_len_ + unobs_trimmed;
        if (below_min(T)) return 0.0f;
        return (float)(T - covered_ + 1);
      }
      case Kind::CoveredFraction: {
        const uint64_t T = total_bases_ + unobs_plain;
        if (T == 0 || below_min(T)) return 0.0f;
        return (float)covered_ / (float)T;
      }
      case Kind::CoveredBases: {
        const uint64_t T = total_bases_ + unobs_plain;
        if (T == 0 || below_min(T)) return 0.0f;
        return (float)covered_;
      }
      case Kind::RPKM: {
        const uint64_t T = total_bases_ + unobs_plain;
        if (T == 0 || below_min(T)) return 0.0f;
        return (float)(reads_ * 1000000000ULL) / (float)T;
      }
      case Kind::TPM: {
        const uint64_t T = total_bases_ + unobs_plain;
        if (T == 0 || below_min(T)) return 0.0f;
        return (float)std::exp(std::log((double)reads_) - std::log((double)T));
      }
      case Kind::Variance: {
        const uint64_t T = observed_len_ + unobs_trimmed;
        if (T == 0) return 0.0f;
        const bool counts_empty = use_device_walk() ? false : counts_.empty();
        if (below_min(T) || T < 3 || counts_empty) return 0.0f;
        uint64_t ex, ex2;
        if (use_device_walk()) {
          if (unobs_trimmed != 0) throw Panic("internal error: device variance sums used outside single-contig mode");
          ex = dev_.var_ex;
          ex2 = dev_.var_ex2;
        } else {
          counts_[0] += unobs_trimmed;
          size_t k = 0;
          while (counts_[k] == 0) k += 1;
          ex = 0;
          ex2 = 0;
          for (size_t x = 0; x < counts_.size(); ++x) {
            if (counts_[x] == 0) continue;
            const uint64_t nc = counts_[x];
            ex += (uint64_t)(x - k) * nc;
            ex2 += (uint64_t)(x - k) * (uint64_t)(x - k) * nc;
          }
        }
        return ((float)ex2 - (float)(ex * ex) / (float)T) / (float)(T - 1);
      }
      case Kind::Length: return (float)(observed_len_ + unobs_plain);
      case Kind::ReadCount: return (float)reads_;
      case Kind::ReadsPerBase: return (float)reads_ / (float)(observed_len_ + unobs_plain);
      case Kind::ANIr: return reads_ == 0 ? 0.0f : (float)(sum_identity_ / (double)reads_);
    }
    return 0.0f;
  }

  void print_coverage(float coverage, CoverageTaker& t) const {  // EST:936-969
    if (kind != Kind::PileupCounts) {
      t.add_single_coverage(coverage);
      return;
    }
    for (size_t i = 0; i < counts_.size(); ++i) {
      uint64_t cov;
      if (i == 0) {
        const uint64_t c = (uint64_t)std::floor(coverage);
        cov = c == 0 ? 0 : c - 1;
      } else {
        cov = counts_[i];
      }
      t.add_coverage_entry(i, cov);
    }
  }
  void print_zero_coverage(CoverageTaker& t, uint64_t entry_length) const {  // EST:971-991
    if (kind == Kind::PileupCounts) return;
    t.add_single_coverage(kind == Kind::Length ? (float)entry_length : 0.0f);
  }

  // EST:598-642, over a dense ascending histogram.
  static uint64_t trimmed_walk(const std::vector<uint64_t>& counts, size_t min_index, size_t max_index) {
    size_t accounted = 0, total = 0;
    bool started = false;
    for (size_t depth = 0; depth < counts.size(); ++depth) {
      const size_t here = (size_t)counts[depth];
      accounted += here;
      if (accounted < min_index) continue;
      if (started) {
        if (accounted > max_index) {
          const size_t excess = accounted - here;
          total += (max_index >= excess ? max_index - excess + 1 : 0) * depth;
          break;
        }
        total += here * depth;
      } else if (accounted > max_index) {
        total = (max_index - min_index + 1) * depth;
        started = true;
      } else {
        total = (accounted - min_index + 1) * depth;
        started = true;
      }
    }
    return total;
  }

 private:
  bool use_device_walk() const { return dev_valid_ && !have_counts_ && n_window_contigs_ == 1; }
  uint64_t total_count_ = 0, total_bases_ = 0, covered_ = 0, reads_ = 0, mismatches_ = 0, observed_len_ = 0;
  std::vector<uint64_t> counts_;
  bool have_counts_ = false, dev_valid_ = false;
  uint32_t n_window_contigs_ = 0;
  cmb_contig_stats dev_{};
  double sum_identity_ = 0.0;
};

// ------------------------------------------------------------------------------------------------ printers
class CoveragePrinter {  // coverage_printer.rs:9-17
 public:
  enum class Kind { Streamed, SparseCached, DenseCached, MetabatAdjusted };
  Kind kind = Kind::Streamed;
  ThreadPool* pool = nullptr;  // optional: rows are formatted in parallel blocks

  void print_headers(const std::string& entry_type, const std::vector<std::string>& headers, std::ostream& os) {  // :123-152
    if (kind == Kind::Streamed || kind == Kind::SparseCached) {
      os << "Sample\t" << entry_type;
      for (auto& h : headers) os << '\t' << h;
      os << '\n';
    } else if (kind == Kind::DenseCached) {
      entry_type_ = entry_type;
      headers_ = headers;
    }
  }

  void finalise_printing(const CoverageTaker& taker, std::ostream& os, const std::vector<ReadsMapped>& reads_mapped,
                         const std::vector<size_t>& columns_to_normalise, std::optional<size_t> rpkm_column,
                         std::optional<size_t> tpm_column) {  // :20-121
    switch (kind) {
      case Kind::Streamed: break;
      case Kind::SparseCached: sparse(taker, os, reads_mapped, columns_to_normalise, rpkm_column, tpm_column); break;
      case Kind::DenseCached: dense(taker, os, reads_mapped, columns_to_normalise, rpkm_column, tpm_column); break;
      case Kind::MetabatAdjusted: metabat(taker, os); break;
    }
  }

 private:
  std::string entry_type_;
  std::vector<std::string> headers_;

  static bool in(const std::vector<size_t>& v, size_t x) { return std::find(v.begin(), v.end(), x) != v.end(); }
  static std::string strip_cr(std::string s) {
    while (!s.empty() && s.back() == '\r') s.pop_back();
    return s;
  }
  static size_t tabs(const std::string& s) { return (size_t)std::count(s.begin(), s.end(), '\t'); }
  // the NA padding around the normalised columns of the `unmapped` row (:238-263, :409-437)
  template <class PrintValue>
  static void unmapped_columns(std::ostream& os, const std::vector<size_t>& cols, size_t num_coverages, PrintValue value) {
    for (size_t i = 0; i < cols.size(); ++i) {
      const size_t from = i == 0 ? 0 : cols[i - 1] + 1;
      for (size_t k = from; k < cols[i]; ++k) os << "\tNA";
      os << '\t' << value(cols[i]);
    }
    for (size_t k = cols.back() + 1; k < num_coverages; ++k) os << "\tNA";
  }

  void metabat(const CoverageTaker& taker, std::ostream& os) {  // :57-119
    os << "contigName\tcontigLen\ttotalAvgDepth";
    for (auto& s : taker.stoit_names) os << '\t' << s << ".bam\t" << s << ".bam-var";
    os << '\n';
    std::vector<std::vector<EntryAndCoverages>> by_stoit(taker.stoit_names.size());
    for (auto& ec : taker.entries_by_stoit()) by_stoit[ec.stoit_index].push_back(ec);
    if (by_stoit.empty() || by_stoit[0].empty()) {
      if (by_stoit.empty()) throw Panic("index out of bounds: the len is 0 but the index is 0");
      return;
    }
    for (size_t e = 0; e < by_stoit[0].size(); ++e) {
      float total_depth = 0.0f;
      for (auto& st : by_stoit) total_depth += st[e].coverages[1];
      os << *taker.entry_name(e) << '\t' << rust_display(by_stoit[0][e].coverages[0]) << '\t'
         << rust_display(std::round((double)total_depth * 10000.0 / (double)taker.coverages.size()) / 10000.0);
      for (auto& st : by_stoit)
        os << '\t' << rust_display(std::round((double)st[e].coverages[1] * 10000.0) / 10000.0) << '\t'
           << rust_display(std::round((double)st[e].coverages[2] * 10000.0) / 10000.0);
      os << '\n';
    }
  }

  // Formats rows [0, n) with row(i, std::string&) in parallel blocks and writes them in order.
  template <class RowFn>
  void write_rows(std::ostream& os, size_t n, RowFn row) {
    constexpr size_t BLOCK = 4096;
    const size_t n_blocks = (n + BLOCK - 1) / BLOCK;
    std::vector<std::string> text(n_blocks);
    auto do_block = [&](size_t b, int) {
      std::string& t = text[b];
      t.reserve(BLOCK * 48);
      for (size_t i = b * BLOCK; i < std::min(n, (b + 1) * BLOCK); ++i) row(i, t);
    };
    const double t0 = host_now();
    if (pool && n_blocks > 1) pool->parallel_for(n_blocks, do_block);
    else for (size_t b = 0; b < n_blocks; ++b) do_block(b, 0);
    const double t1 = host_now();
    BulkSink* sink = dynamic_cast<BulkSink*>(os.rdbuf());
    if (sink && pool && n_blocks > 1) {
      os.flush();
      std::vector<size_t> at(n_blocks + 1, 0);
      for (size_t b = 0; b < n_blocks; ++b) at[b + 1] = at[b] + text[b].size();
      char* dst = sink->append_uninitialized(at[n_blocks]);
      pool->parallel_for(n_blocks, [&](size_t b, int) { memcpy(dst + at[b], text[b].data(), text[b].size()); });
    } else {
      for (auto& t : text) os.write(t.data(), (std::streamsize)t.size());
    }
    host_stat("write_rows.format", t1 - t0);
    host_stat("write_rows.concat", host_now() - t1);
  }

  void sparse(const CoverageTaker& taker, std::ostream& os, const std::vector<ReadsMapped>& rm,
              const std::vector<size_t>& norm, std::optional<size_t> rpkm, std::optional<size_t> tpm) {  // :155-356
    const size_t nc = taker.num_coverages;
    size_t extra_cols = 0;
    if (taker.shared_names) {  // the lowest entry index any stoit recorded
      std::optional<size_t> first;
      for (auto& c : taker.coverages)
        if (!c.empty() && (!first || c[0].entry_index < *first)) first = c[0].entry_index;
      if (first && taker.entry_name(*first)) extra_cols = tabs(*taker.entry_name(*first));
    }
    for (auto& n : taker.entry_names)
      if (n) { extra_cols = tabs(*n); break; }
    const auto cells = taker.cells_by_stoit();
    // The reference always emits the block of the first stoit (even when it is empty); later stoits only when they
    // have entries — which, by construction of the iterator, is whenever any stoit has entries.
    for (size_t stoit = 0; stoit < std::max<size_t>(1, cells.size()); ++stoit) {
      if (stoit >= taker.stoit_names.size()) throw Panic("index out of bounds");
      if (stoit > 0 && cells[stoit].empty()) continue;
      static const std::vector<CoverageTaker::Cell> none;
      const auto& rows = cells.empty() ? none : cells[stoit];
      const std::string& name = taker.stoit_names[stoit];
      std::vector<float> totals(nc, 0.0f), mult(nc, 0.0f);
      for (size_t c : norm) {
        float t = 0.0f;
        for (auto& r : rows) t += r.at(c);  // sequential f32 sum in entry order (:220-224)
        totals[c] = t;
        mult[c] = (float)rm[stoit].num_mapped_reads / (float)rm[stoit].num_reads;
      }
      if (tpm) {
        float t = 0.0f;
        for (auto& r : rows) t += r.at(*tpm);
        totals[*tpm] = t;
      }
      if (!norm.empty()) {
        os << name << "\tunmapped";
        for (size_t k = 0; k < extra_cols; ++k) os << '\t';
        unmapped_columns(os, norm, nc, [&](size_t c) { return rust_display(100.0f * (1.0f - mult[c])); });
        os << '\n';
      }
      for (auto& r : rows)
        if (!taker.entry_name(r.entry_index)) throw ExitError(1, "Didn't find entry name string as expected");
      const uint64_t n_mapped = rm.empty() ? 0 : rm[stoit].num_mapped_reads;
      write_rows(os, rows.size(), [&](size_t i, std::string& t) {
        const auto& r = rows[i];
        t += name;
        t += '\t';
        t += strip_cr(*taker.entry_name(r.entry_index));
        for (size_t c = 0; c < nc; ++c) {
          t += '\t';
          const float cov = r.at(c);
          if (in(norm, c)) append_display(t, cov * 100.0f * mult[c] / totals[c]);  // :285-286
          else if (rpkm && *rpkm == c) append_display(t, n_mapped == 0 ? 0.0f : cov / (float)n_mapped);  // :300
          else if (tpm && *tpm == c)
            append_display(t, n_mapped == 0 ? 0.0 : (double)std::exp(std::log(cov) - std::log(totals[c])) * 1000000.0);  // :320-323
          else append_display(t, cov);
        }
        t += '\n';
      });
    }
  }

  void dense(const CoverageTaker& taker, std::ostream& os, const std::vector<ReadsMapped>& rm,
             const std::vector<size_t>& norm, std::optional<size_t> rpkm, std::optional<size_t> tpm) {  // :359-553
    const size_t nc = taker.num_coverages, ns = taker.stoit_names.size();
    os << entry_type_;
    for (auto& s : taker.stoit_names)
      for (auto& h : headers_) os << '\t' << s << ' ' << h;
    os << '\n';
    std::vector<float> mult;
    for (auto& r : rm) mult.push_back((float)r.num_mapped_reads / (float)r.num_reads);
    if (!norm.empty()) {
      os << "unmapped";
      for (size_t k = 0; k < tabs(entry_type_); ++k) os << '\t';
      for (size_t s = 0; s < ns; ++s) unmapped_columns(os, norm, nc, [&](size_t) { return rust_display(100.0f * (1.0f - mult[s])); });
      os << '\n';
    }
    const double t_cells0 = host_now();
    const auto cells = taker.cells_by_stoit();
    host_stat("dense.cells_by_stoit", host_now() - t_cells0);
    std::vector<std::vector<float>> totals(ns, std::vector<float>(nc, 0.0f));
    for (size_t s = 0; s < cells.size(); ++s) {
      auto accumulate = [&](size_t c) {  // first value seeds the total, later ones are added (:457-477)
        bool seeded = false;
        float t = 0.0f;
        for (auto& r : cells[s]) {
          if (seeded) t += r.at(c);
          else { t = r.at(c); seeded = true; }
        }
        totals[s][c] = t;
      };
      for (size_t c : norm) accumulate(c);
      if (tpm) accumulate(*tpm);
    }
    size_t n_filled = 0;  // the reference only materialises stoits that yielded entries
    while (n_filled < cells.size() && !cells[n_filled].empty()) ++n_filled;
    if (n_filled == 0) throw Panic("index out of bounds: the len is 0 but the index is 0");
    write_rows(os, cells[0].size(), [&](size_t e, std::string& t) {
      t += strip_cr(*taker.entry_name(cells[0][e].entry_index));
      for (size_t s = 0; s < n_filled; ++s) {
        const auto& r = cells[s][e];
        const uint64_t n_mapped = rm.empty() ? 0 : rm[s].num_mapped_reads;
        for (size_t c = 0; c < nc; ++c) {
          t += '\t';
          const float cov = r.at(c);
          if (in(norm, c)) append_display(t, cov / totals[s][c] * 100.0f * mult[s]);  // :496-502
          else if (rpkm && *rpkm == c) append_display(t, n_mapped == 0 ? 0.0f : cov / (float)n_mapped);  // :517
          else if (tpm && *tpm == c)
            append_display(t, n_mapped == 0 ? 0.0f : std::exp(std::log(cov) - std::log(totals[s][c])) * 1000000.0f);  // :536-539
          else append_display(t, cov);
        }
      }
      t += '\n';
    });
  }
};

}  // namespace cmbh
