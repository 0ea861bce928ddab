// The three coverage drivers of the reference, re-expressed over the GPU's per-contig statistics:
//   contig_coverage                              src/contig.rs:13-253
//   mosdepth_genome_coverage_with_contig_names   src/genome.rs:17-322
//   mosdepth_genome_coverage                     src/genome.rs:419-797  (+ helpers :331-416, :799-929)
// Same signatures in spirit (readers -> taker, estimators, print-zeros flag, FlagFilter -> Vec<ReadsMapped>).  The
// record loop itself (flag filter, read filter, CIGAR walk, per-contig counters) ran on the device inside
// DeviceSession::process; what remains is the per-contig / per-genome flush logic, which only depends on WHICH
// contigs were seen (n_records > 0), in tid order (the device has verified the sort order, contig.rs:129-132).
#pragma once
#include "coverage_model.hpp"
#include "sample_processor.hpp"

namespace cmbh {

struct GenomesAndContigs {  // genomes_and_contigs.rs:7-58
  std::vector<std::string> genomes;
  std::unordered_map<std::string, size_t> contig_to_genome;
};

// genome_parsing.rs:75-142
inline GenomesAndContigs read_genome_definition_file(const std::string& path) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) throw Panic("Unable to find/read genome definition file " + path);
  std::string text;
  char chunk[1 << 16];
  size_t got;
  while ((got = fread(chunk, 1, sizeof chunk, f)) > 0) text.append(chunk, got);
  fclose(f);
  GenomesAndContigs gc;
  std::unordered_map<std::string, std::string> contig_genome;
  std::unordered_map<std::string, size_t> genome_index;
  auto ws = [](char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\x0b' || c == '\x0c'; };
  size_t a = 0;
  while (a < text.size()) {
    size_t b = text.find('\n', a);
    if (b == std::string::npos) b = text.size();
    std::string line = text.substr(a, b - a);
    a = b + 1;
    if (!line.empty() && line.back() == '\r') line.pop_back();
    const size_t tab = line.find('\t');
    if (tab == std::string::npos || line.find('\t', tab + 1) != std::string::npos)
      throw ExitError(1, "The line \"" + line + "\" in the genome definition file is not a genome name and contig name separated by a tab");
    std::string genome = line.substr(0, tab);
    while (!genome.empty() && ws(genome.back())) genome.pop_back();
    size_t g0 = 0;
    while (g0 < genome.size() && ws(genome[g0])) ++g0;
    genome = genome.substr(g0);
    size_t c0 = tab + 1;
    while (c0 < line.size() && ws(line[c0])) ++c0;
    size_t c1 = c0;
    while (c1 < line.size() && !ws(line[c1])) ++c1;
    if (c1 == c0) throw Panic("Failed to split contig name by whitespace in genome definition file");
    const std::string contig = line.substr(c0, c1 - c0);
    auto known = contig_genome.find(contig);
    if (known != contig_genome.end()) {
      if (known->second != genome) throw ExitError(1, "The contig name '" + contig + "' was assigned to multiple genomes");
    } else {
      contig_genome.emplace(contig, genome);
    }
    auto gi = genome_index.find(genome);
    if (gi == genome_index.end()) {
      gi = genome_index.emplace(genome, gc.genomes.size()).first;
      gc.genomes.push_back(genome);
    }
    // GenomesAndContigs::insert exits when a contig is inserted twice (genomes_and_contigs.rs:26-41): the reference
    // pushes every line's contig into genome_to_contig, so a repeated (genome, contig) line also trips it.
    if (!gc.contig_to_genome.emplace(contig, gi->second).second)
      throw ExitError(1, "The contig '" + contig + "' has been assigned to multiple genomes");
  }
  return gc;
}

enum class CountMode { Contig, GenomeNames, GenomeSeparator };

inline const cmb_hist_pair* no_pairs() {
  static const cmb_hist_pair none{};
  return &none;
}

inline ContigObservation observe(const SampleResult& r, uint32_t tid, CountMode mode, bool csr) {
  const cmb_contig_stats& s = r.rows[tid];
  ContigObservation ob;
  ob.len = r.header().lens[tid];
  ob.stats = &s;
  ob.total_mismatches = s.sum_edit - s.sum_indel;  // unchecked u64 subtraction (contig.rs:59)
  switch (mode) {
    case CountMode::Contig:  // contig.rs:157-159, 208-211
      ob.num_mapped_reads = s.n_primary;
      ob.sum_identity = s.sum_identity_primary;
      break;
    case CountMode::GenomeNames:  // genome.rs:173-174, 220-223
      ob.num_mapped_reads = s.n_records;
      ob.sum_identity = s.sum_identity_nonsupp;
      break;
    case CountMode::GenomeSeparator:  // genome.rs:677-682, 724-727
      ob.num_mapped_reads = s.n_nonsupp;
      ob.sum_identity = s.sum_identity_primary;
      break;
  }
  if (csr) {
    ob.hist = s.hist_count ? r.pairs.data() + s.hist_offset : no_pairs();
    ob.n_hist = s.hist_count;
  }
  return ob;
}

struct DriverIO {
  DeviceSession* session;
  cmb_params params;  // FlagFilter + filter thresholds + E + trim + want
  std::vector<SampleTiming>* timings = nullptr;
  std::vector<uint64_t>* record_counts = nullptr;
  std::ostream* log = nullptr;  // the reference's info!/warn! lines (env_logger on stderr); null = --quiet
};

// contig.rs:233-240, genome.rs:309-316, 784-791: `info!("In sample '{}', found {} reads mapped out of {} total ({:.*}%)", ..)`
// and the contig driver's warning about a sample without primary alignments (contig.rs:243-248).
inline void log_reads_mapped(const DriverIO& io, const std::string& stoit_name, const ReadsMapped& rm, bool warn_if_none) {
  if (!io.log) return;
  char pct[64];
  const double num = (double)(rm.num_mapped_reads * 100ull), den = (double)rm.num_reads;
  if (rm.num_reads == 0) snprintf(pct, sizeof pct, "%s", rm.num_mapped_reads == 0 ? "NaN" : "inf");  // Rust Display of f64
  else snprintf(pct, sizeof pct, "%.2f", num / den);
  *io.log << "[INFO] In sample '" << stoit_name << "', found " << rm.num_mapped_reads << " reads mapped out of " << rm.num_reads
          << " total (" << pct << "%)\n";
  if (warn_if_none && rm.num_reads == 0)
    *io.log << "[WARN] No primary alignments were observed for sample " << stoit_name
            << " - perhaps something went wrong in the mapping?\n";
}

inline SampleResult run_sample(const DriverIO& io, const InputSpec& in) {
  SampleResult r = io.session->process(in, io.params);
  if (io.timings) io.timings->push_back(r.timing);
  if (io.record_counts) io.record_counts->push_back(r.n_records);
  return r;
}

// ---------------------------------------------------------------------------------------------- contig.rs:13-253
inline std::vector<ReadsMapped> contig_coverage(const std::vector<InputSpec>& bam_readers, CoverageTaker& coverage_taker,
                                                std::vector<CoverageEstimator>& coverage_estimators,
                                                bool print_zero_coverage_contigs, const DriverIO& io) {
  std::vector<ReadsMapped> reads_mapped_vector;
  const bool csr = io.params.want & CMB_WANT_HIST_CSR;
  for (const InputSpec& in : bam_readers) {
    const SampleResult r = run_sample(io, in);
    if (!io.session->is_output_rank()) {  // a non-printing rank of a multi-GPU group: its part ended with the gather
      reads_mapped_vector.push_back({0, r.num_detected_primary_alignments});
      continue;
    }
    coverage_taker.start_stoit(r.stoit_name);
    uint64_t num_mapped_reads_total = 0;
    const uint32_t n = (uint32_t)r.header().names.size();
    std::vector<float> coverages(coverage_estimators.size());
    const std::vector<uint64_t> contig_mode_unobserved{0};  // calculate_coverage(&[0]), contig.rs:65
    bool any_pileup = false;
    for (auto& e : coverage_estimators) any_pileup = any_pileup || e.kind == CoverageEstimator::Kind::PileupCounts;
    if (!any_pileup && n >= 1024) {
      // Same per-contig sequence as the loop below (add_contig, calculate_coverage(&[0]), setup), but contigs are
      // independent, so the estimator maths runs on all host threads into a table; the taker is then fed in tid order.
      const size_t K = coverage_estimators.size();
      std::vector<float> vals((size_t)n * K);
      std::vector<uint8_t> state(n, 0);  // 0: never seen, 1: seen, all coverages zero, 2: seen, some coverage > 0
      ThreadPool& pool = io.session->pool();
      const size_t n_tasks = std::min<size_t>((size_t)pool.size() * 4, (n + 4095) / 4096);
      std::vector<uint64_t> task_mapped(n_tasks, 0);
      const double t_table0 = host_now();
      pool.parallel_for(n_tasks, [&](size_t task, int) {
        std::vector<CoverageEstimator> est = coverage_estimators;
        const uint32_t t0 = (uint32_t)((uint64_t)n * task / n_tasks), t1 = (uint32_t)((uint64_t)n * (task + 1) / n_tasks);
        uint64_t mapped = 0;
        for (uint32_t tid = t0; tid < t1; ++tid) {
          if (r.rows[tid].n_records == 0) continue;
          const ContigObservation ob = observe(r, tid, CountMode::Contig, csr);
          bool has_nonzero = false;
          for (size_t k = 0; k < K; ++k) {
            est[k].add_contig(ob);
            const float c = est[k].calculate_coverage(contig_mode_unobserved);
            vals[(size_t)tid * K + k] = c;
            has_nonzero = has_nonzero || c > 0.0f;
            est[k].setup();
          }
          state[tid] = has_nonzero ? 2 : 1;
          if (has_nonzero) mapped += ob.num_mapped_reads;
        }
        task_mapped[task] = mapped;
      });
      for (uint64_t m : task_mapped) num_mapped_reads_total += m;
      const double t_feed0 = host_now();
      host_stat("contig.estimator_table", t_feed0 - t_table0);
      const bool shared = coverage_taker.try_share_names(std::shared_ptr<const std::vector<std::string>>(r.hdr, &r.hdr->names));
      coverage_taker.reserve_entries(n, K);
      auto start_entry = [&](uint32_t tid) {
        if (shared) coverage_taker.start_entry_shared(tid);
        else coverage_taker.start_entry(tid, r.header().names[tid]);
      };
      for (uint32_t tid = 0; tid < n; ++tid) {
        if (state[tid] == 0) {
          if (print_zero_coverage_contigs) {
            start_entry(tid);
            for (auto& e : coverage_estimators) e.print_zero_coverage(coverage_taker, r.header().lens[tid]);
            coverage_taker.finish_entry();
          }
        } else if (print_zero_coverage_contigs || state[tid] == 2) {
          start_entry(tid);
          for (size_t k = 0; k < K; ++k) coverage_estimators[k].print_coverage(vals[(size_t)tid * K + k], coverage_taker);
          coverage_taker.finish_entry();
        }
      }
      host_stat("contig.taker_feed", host_now() - t_feed0);
      reads_mapped_vector.push_back({num_mapped_reads_total, r.num_detected_primary_alignments});
      log_reads_mapped(io, r.stoit_name, reads_mapped_vector.back(), true);
      continue;
    }
    for (uint32_t tid = 0; tid < n; ++tid) {
      if (r.rows[tid].n_records == 0) {  // never seen: print_previous_zero_coverage_contigs (:255-277)
        if (print_zero_coverage_contigs) {
          coverage_taker.start_entry(tid, r.header().names[tid]);
          for (auto& e : coverage_estimators) e.print_zero_coverage(coverage_taker, r.header().lens[tid]);
          coverage_taker.finish_entry();
        }
        continue;
      }
      const ContigObservation ob = observe(r, tid, CountMode::Contig, csr);  // process_previous_contigs (:40-104)
      bool has_nonzero = false;
      for (size_t k = 0; k < coverage_estimators.size(); ++k) {
        coverage_estimators[k].add_contig(ob);
        coverages[k] = coverage_estimators[k].calculate_coverage(contig_mode_unobserved);
        has_nonzero = has_nonzero || coverages[k] > 0.0f;
      }
      if (has_nonzero) num_mapped_reads_total += ob.num_mapped_reads;
      if (print_zero_coverage_contigs || has_nonzero) {
        coverage_taker.start_entry(tid, r.header().names[tid]);
        for (size_t k = 0; k < coverage_estimators.size(); ++k) coverage_estimators[k].print_coverage(coverages[k], coverage_taker);
        coverage_taker.finish_entry();
      }
      for (auto& e : coverage_estimators) e.setup();
    }
    reads_mapped_vector.push_back({num_mapped_reads_total, r.num_detected_primary_alignments});
    log_reads_mapped(io, r.stoit_name, reads_mapped_vector.back(), true);
  }
  return reads_mapped_vector;
}

// ---------------------------------------------------------------------------------------------- genes.rs:182-344
// Per-gene coverage.  The device laid its arena out over the genes (cmb_set_genes) and clipped every aligned block to the genes
// it overlaps, so each row already holds what emit_genes_for_contig (genes.rs:467-552) computes from the contig's arrays: the
// gene's own window statistics / histogram, the primaries whose leftmost position lies in the gene, their substitutions
// (NM.saturating_sub(indels)) and identities.  What remains is the reference's emission order and zero handling.
inline std::vector<ReadsMapped> gene_coverage(const std::vector<InputSpec>& bam_readers, CoverageTaker& coverage_taker,
                                              std::vector<CoverageEstimator>& coverage_estimators, const GeneDefinitions& gene_definitions,
                                              const GenomeNamer* genome_namer, bool print_zero_coverage_genes, const DriverIO& io) {
  std::vector<ReadsMapped> reads_mapped_vector;
  const bool csr = io.params.want & CMB_WANT_HIST_CSR;
  io.session->set_gene_definitions(&gene_definitions, genome_namer);
  struct Restore {
    DeviceSession* s;
    ~Restore() { s->set_gene_definitions(nullptr, nullptr); }
  } restore{io.session};
  const std::vector<uint64_t> contig_mode_unobserved{0};  // calculate_coverage(&[0]), genes.rs:536-539
  for (const InputSpec& in : bam_readers) {
    const SampleResult r = run_sample(io, in);
    if (!io.session->is_output_rank()) {  // a non-printing rank of a multi-GPU group: its part ended with the gather
      reads_mapped_vector.push_back({0, r.num_detected_primary_alignments});
      continue;
    }
    coverage_taker.start_stoit(r.stoit_name);
    const ResolvedGenes& genes = *r.genes;
    const uint32_t n = (uint32_t)r.header().names.size();
    std::vector<float> coverages(coverage_estimators.size());
    for (uint32_t tid = 0; tid < n; ++tid) {
      const uint32_t g0 = genes.first_of_tid[tid], g1 = genes.first_of_tid[tid + 1];
      if (g0 == g1) continue;
      if (!r.contig_seen[tid]) {  // emit_zero_coverage_genes (genes.rs:554-568)
        if (!print_zero_coverage_genes) continue;
        for (uint32_t g = g0; g < g1; ++g) {
          coverage_taker.start_entry(g, genes.entries[g].name);
          for (auto& e : coverage_estimators) e.print_zero_coverage(coverage_taker, (uint64_t)(genes.entries[g].end - genes.entries[g].start));
          coverage_taker.finish_entry();
        }
        continue;
      }
      for (uint32_t g = g0; g < g1; ++g) {  // emit_genes_for_contig (genes.rs:502-551)
        const cmb_contig_stats& st = r.rows[g];
        ContigObservation ob;
        ob.len = genes.entries[g].end - genes.entries[g].start;
        ob.stats = &st;
        ob.num_mapped_reads = st.n_primary;
        ob.total_mismatches = st.sum_edit - st.sum_indel;  // sum_indel is 0: the device already summed the saturated differences
        ob.sum_identity = st.sum_identity_primary;
        if (csr) {
          ob.hist = st.hist_count ? r.pairs.data() + st.hist_offset : no_pairs();
          ob.n_hist = st.hist_count;
        }
        bool has_nonzero = false;
        for (size_t k = 0; k < coverage_estimators.size(); ++k) {
          coverage_estimators[k].add_contig(ob);
          coverages[k] = coverage_estimators[k].calculate_coverage(contig_mode_unobserved);
          has_nonzero = has_nonzero || coverages[k] > 0.0f;
        }
        if (print_zero_coverage_genes || has_nonzero) {
          coverage_taker.start_entry(g, genes.entries[g].name);
          for (size_t k = 0; k < coverage_estimators.size(); ++k) coverage_estimators[k].print_coverage(coverages[k], coverage_taker);
          coverage_taker.finish_entry();
        }
        for (auto& e : coverage_estimators) e.setup();
      }
    }
    reads_mapped_vector.push_back({r.kept_primary, r.num_detected_primary_alignments});
    log_reads_mapped(io, r.stoit_name, reads_mapped_vector.back(), true);
  }
  return reads_mapped_vector;
}

// ---------------------------------------------------------------------------------------------- genome.rs:17-322
inline std::vector<ReadsMapped> mosdepth_genome_coverage_with_contig_names(
    const std::vector<InputSpec>& bam_readers, const GenomesAndContigs& contigs_and_genomes, CoverageTaker& coverage_taker,
    bool print_zero_coverage_genomes, std::vector<CoverageEstimator>& coverage_estimators, const DriverIO& io) {
  std::vector<ReadsMapped> reads_mapped_vector;
  const bool csr = io.params.want & CMB_WANT_HIST_CSR;
  const size_t n_genomes = contigs_and_genomes.genomes.size();
  for (const InputSpec& in : bam_readers) {
    const SampleResult r = run_sample(io, in);
    if (!io.session->is_output_rank()) {  // a non-printing rank of a multi-GPU group: its part ended with the gather
      reads_mapped_vector.push_back({0, r.num_detected_primary_alignments});
      continue;
    }
    coverage_taker.start_stoit(r.stoit_name);
    const uint32_t n = (uint32_t)r.header().names.size();
    std::vector<int64_t> genome_of(n, -1);
    std::vector<std::vector<uint32_t>> refs_of(n_genomes);
    uint32_t in_genomes = 0;
    for (uint32_t tid = 0; tid < n; ++tid) {
      auto it = contigs_and_genomes.contig_to_genome.find(r.header().names[tid]);
      if (it == contigs_and_genomes.contig_to_genome.end()) continue;
      genome_of[tid] = (int64_t)it->second;
      refs_of[it->second].push_back(tid);
      ++in_genomes;
    }
    if (in_genomes == 0) throw ExitError(1, "Error: There are no found reference sequences that are a part of a genome");
    std::vector<std::vector<CoverageEstimator>> per_genome(n_genomes, coverage_estimators);  // cloned per genome (:92-97)
    std::vector<uint64_t> reads_mapped_in_each_genome(n_genomes, 0);
    bool any_seen = false;
    for (uint32_t tid = 0; tid < n; ++tid) {
      if (r.rows[tid].n_records == 0) continue;
      any_seen = true;
      if (genome_of[tid] < 0) continue;  // reads on contigs outside every genome are ignored (:170-171)
      const ContigObservation ob = observe(r, tid, CountMode::GenomeNames, csr);
      reads_mapped_in_each_genome[(size_t)genome_of[tid]] += ob.num_mapped_reads;
      for (auto& e : per_genome[(size_t)genome_of[tid]]) e.add_contig(ob);
    }
    uint64_t num_mapped_reads_total = 0;
    if (!any_seen && r.num_detected_primary_alignments == 0) {
      // "No primary alignments were observed": nothing is printed for this sample (:230-234)
    } else {
      if (!any_seen && n > 0 && genome_of[0] >= 0) {  // the reference still records tid 0 with an empty array (:237-248)
        ContigObservation empty;
        for (auto& e : per_genome[(size_t)genome_of[0]]) e.add_contig(empty);
      }
      for (size_t g = 0; g < n_genomes; ++g) {
        std::vector<uint64_t> unobserved;
        uint64_t genome_len = 0;
        for (uint32_t tid : refs_of[g]) {
          genome_len += r.header().lens[tid];
          if (r.rows[tid].n_records == 0) unobserved.push_back(r.header().lens[tid]);
        }
        std::vector<float> coverages;
        bool any_nonzero = false;
        for (auto& e : per_genome[g]) {
          coverages.push_back(e.calculate_coverage(unobserved));
          any_nonzero = any_nonzero || coverages.back() > 0.0f;
        }
        if (any_nonzero) num_mapped_reads_total += reads_mapped_in_each_genome[g];
        if (print_zero_coverage_genomes || any_nonzero) {
          coverage_taker.start_entry(g, contigs_and_genomes.genomes[g]);
          for (size_t k = 0; k < per_genome[g].size(); ++k) {
            if (coverages[k] > 0.0f) per_genome[g][k].print_coverage(coverages[k], coverage_taker);
            else per_genome[g][k].print_zero_coverage(coverage_taker, genome_len);
          }
          coverage_taker.finish_entry();
        }
      }
    }
    reads_mapped_vector.push_back({num_mapped_reads_total, r.num_detected_primary_alignments});
    log_reads_mapped(io, r.stoit_name, reads_mapped_vector.back(), false);
  }
  return reads_mapped_vector;
}

// ---------------------------------------------------------------------------------------------- genome.rs:419-929
namespace separator_mode {

inline std::string extract_genome(uint32_t tid, const Header& h, uint8_t split_char) {  // :799-805
  const std::string& name = h.names.at(tid);
  const size_t at = name.find((char)split_char);
  if (at == std::string::npos)
    throw Panic("Contig name " + name + " does not contain split symbol, so cannot determine which genome it belongs to");
  return name.substr(0, at);
}

struct Unobserved {  // UnobservedLengthAndFirstTid (:324-328)
  std::vector<uint64_t> lengths;
  size_t first_tid = 0;
};

struct Walk {
  const Header& h;
  uint8_t split_char;
  bool single_genome;
  bool same_genome(uint32_t tid, const std::string& g) const { return single_genome || extract_genome(tid, h, split_char) == g; }

  Unobserved backwards(uint32_t current_tid, const std::string& genome) const {  // fill_genome_length_backwards (:807-853)
    Unobserved u;
    if (current_tid == 0) return u;
    uint32_t t = current_tid - 1;
    while (same_genome(t, genome)) {
      u.lengths.push_back(h.lens[t]);
      if (t == 0) return u;  // first_tid stays 0
      --t;
    }
    u.first_tid = (size_t)t + 1;
    return u;
  }
  void between(uint32_t current_tid, uint32_t last_tid, const std::string& genome, std::vector<uint64_t>& out) const {  // :477-499
    if (current_tid == 0) return;
    for (uint32_t t = last_tid + 1; t < current_tid && same_genome(t, genome); ++t) out.push_back(h.lens[t]);
  }
  void forwards(uint32_t current_tid, const std::optional<std::string>& genome, std::vector<uint64_t>& out) const {  // :448-475
    if (!genome) return;
    for (uint32_t t = current_tid + 1; t < h.names.size() && same_genome(t, *genome); ++t) out.push_back(h.lens[t]);
  }
};

// print_previous_zero_coverage_genomes2 (:859-929): zero rows for whole genomes skipped between last_genome and
// current_genome, discovered by scanning the header downwards from current_tid.
inline void print_skipped_genomes(const std::optional<std::string>& last_genome, const std::string& current_genome,
                                  uint32_t current_tid, const std::vector<CoverageEstimator>& ests, const Header& h,
                                  uint8_t split_char, CoverageTaker& taker) {
  struct Pending { std::string genome; size_t first_tid; uint64_t length; };
  std::vector<Pending> found;
  std::string open_genome = current_genome;
  std::optional<uint32_t> open_first;
  uint64_t open_len = 0;
  for (uint32_t t = current_tid;; --t) {
    const std::string g = extract_genome(t, h, split_char);
    if (last_genome && g == *last_genome) break;
    if (g != open_genome) {
      if (open_first && (!last_genome || g != *last_genome)) found.push_back({open_genome, *open_first, open_len});
      open_genome = g;
      open_first = t;
      open_len = h.lens[t];
    } else if (g != current_genome) {
      open_first = t;
      open_len += h.lens[t];
    }
    if (t == 0) break;
  }
  if (open_first) found.push_back({open_genome, *open_first, open_len});
  for (size_t k = found.size(); k-- > 0;) {
    taker.start_entry(found[k].first_tid, found[k].genome);
    for (auto& e : ests) e.print_zero_coverage(taker, found[k].length);
    taker.finish_entry();
  }
}

}  // namespace separator_mode

inline std::vector<ReadsMapped> mosdepth_genome_coverage(const std::vector<InputSpec>& bam_readers, uint8_t split_char,
                                                         CoverageTaker& coverage_taker, bool print_zero_coverage_genomes,
                                                         std::vector<CoverageEstimator>& coverage_estimators,
                                                         bool single_genome, const DriverIO& io) {
  using namespace separator_mode;
  std::vector<ReadsMapped> reads_mapped_vector;
  const bool csr = io.params.want & CMB_WANT_HIST_CSR;
  for (const InputSpec& in : bam_readers) {
    const SampleResult r = run_sample(io, in);
    if (!io.session->is_output_rank()) {  // a non-printing rank of a multi-GPU group: its part ended with the gather
      reads_mapped_vector.push_back({0, r.num_detected_primary_alignments});
      continue;
    }
    coverage_taker.start_stoit(r.stoit_name);
    const Header& h = r.header();
    const Walk walk{h, split_char, single_genome};
    const uint32_t n = (uint32_t)h.names.size();

    // print_last_genomes (:331-416): finish the open genome with its last contig `last_ob`
    auto finish_genome = [&](const ContigObservation& last_ob, const std::optional<std::string>& genome, Unobserved& unobs,
                             const std::string& next_genome, uint32_t tid_to_print_zeros_to) {
      std::vector<float> coverages;
      bool positive = false;
      for (auto& e : coverage_estimators) {
        e.add_contig(last_ob);
        coverages.push_back(e.calculate_coverage(unobs.lengths));
        positive = positive || coverages.back() > 0.0f;
      }
      if ((print_zero_coverage_genomes || positive) && genome) {
        coverage_taker.start_entry(unobs.first_tid, *genome);
        for (size_t k = 0; k < coverage_estimators.size(); ++k) {
          if (coverages[k] > 0.0f) coverage_estimators[k].print_coverage(coverages[k], coverage_taker);
          else coverage_estimators[k].print_zero_coverage(coverage_taker, 9);
        }
        coverage_taker.finish_entry();
      }
      for (auto& e : coverage_estimators) e.setup();
      if (print_zero_coverage_genomes && !single_genome)
        print_skipped_genomes(genome, next_genome, tid_to_print_zeros_to, coverage_estimators, h, split_char, coverage_taker);
      return positive;
    };

    bool doing_first = true;
    uint32_t last_tid = 0;
    std::optional<std::string> last_genome;
    Unobserved unobs;
    uint64_t num_mapped_reads_total = 0, reads_in_genome = 0;
    ContigObservation pending;  // the contig at last_tid, not yet added
    for (uint32_t tid = 0; tid < n; ++tid) {
      if (r.rows[tid].n_records == 0) continue;
      const std::string current_genome = single_genome ? std::string() : extract_genome(tid, h, split_char);
      if (doing_first) {
        for (auto& e : coverage_estimators) e.setup();
        unobs = walk.backwards(tid, current_genome);
        last_genome = current_genome;
        doing_first = false;
        if (print_zero_coverage_genomes && !single_genome)
          print_skipped_genomes(std::nullopt, current_genome, tid, coverage_estimators, h, split_char, coverage_taker);
      } else if (current_genome == *last_genome) {
        for (auto& e : coverage_estimators) e.add_contig(pending);
        walk.between(tid, last_tid, current_genome, unobs.lengths);
      } else {
        walk.between(tid, last_tid, *last_genome, unobs.lengths);
        if (finish_genome(pending, last_genome, unobs, current_genome, tid)) num_mapped_reads_total += reads_in_genome;
        reads_in_genome = 0;
        last_genome = current_genome;
        unobs = walk.backwards(tid, current_genome);
      }
      pending = observe(r, tid, CountMode::GenomeSeparator, csr);
      reads_in_genome += pending.num_mapped_reads;
      last_tid = tid;
    }
    if (doing_first && r.num_detected_primary_alignments == 0) {
      // warn only (:731-735)
    } else {
      if (doing_first) pending = ContigObservation{};  // ups_and_downs is still the empty Vec
      if (single_genome) last_genome = std::string("genome1");
      walk.forwards(last_tid, last_genome, unobs.lengths);
      if (n == 0) throw Panic("attempt to subtract with overflow");
      if (finish_genome(pending, last_genome, unobs, std::string(), n - 1)) num_mapped_reads_total += reads_in_genome;
    }
    reads_mapped_vector.push_back({num_mapped_reads_total, r.num_detected_primary_alignments});
    log_reads_mapped(io, r.stoit_name, reads_mapped_vector.back(), false);
  }
  return reads_mapped_vector;
}

}  // namespace cmbh
