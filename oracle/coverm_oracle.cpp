// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_bam.hpp header).
//
// Command-line front end of the CPU restatement: the `contig` / `genome`
// `--bam-files` branches of src/bin/coverm.rs (contig arm :473-663, genome arm
// :56-407, EstimatorsAndTaker::generate_from_clap :1315-1504, FilterParameters
// :1648-1704, parse_percentage :1296-1312, run_contig :2088-2131, run_genome
// :1539-1628), flag names/defaults from src/cli.rs:1670-2582.
//
// Library-level switches (no CLI equivalent in the reference; they reproduce
// the in-module unit-test harnesses contig.rs:290-322, genome.rs:940-1086,
// filter.rs:342-844):
//   --lib-estimators "mean:MINFRAC:EXCL:EXCLUDE_MISMATCHES;variance:MINFRAC:EXCL;trimmed_mean:MIN:MAX:MINFRAC:EXCL;..."
//   --lib-streaming            streaming taker, no header line
//   --lib-flags I,S,SEC        FlagFilter {include_improper_pairs, include_supplementary, include_secondary}
//   --print-reads-mapped       "#reads_mapped\t<sample>\t<mapped>\t<total>" lines on stderr
//   filter-names subcommand    prints qnames surviving ReferenceSortedBamFilter
//   --timing                   "#timing\t<records>\t<seconds>" on stderr (bench cpu_baseline leg)
#include <chrono>
#include <fstream>
#include <iostream>

#include "oracle_core.hpp"

using namespace oracle;

namespace {

struct Args {
  std::string sub;
  std::vector<std::string> bam_files, methods;
  bool methods_given = false;
  std::optional<uint32_t> min_read_aligned_length, min_read_aligned_length_pair;
  std::optional<float> min_read_percent_identity, min_read_aligned_percent, min_read_percent_identity_pair,
      min_read_aligned_percent_pair;
  std::optional<uint8_t> min_mapq;
  bool proper_pairs_only = false, exclude_supplementary = false, include_secondary = false, no_zeros = false;
  float min_covered_fraction = 0.0f, trim_min = 5.0f, trim_max = 95.0f;
  uint64_t contig_end_exclusion = 75;
  std::string output_format = "dense";
  std::optional<std::string> output_file, separator, genome_definition, gff, gff_feature_type;
  bool single_genome = false;
  int threads = 1;
  // library-level
  std::optional<std::string> lib_estimators, lib_flags;
  bool lib_streaming = false, print_reads_mapped = false, inverse = false, timing = false;
};

[[noreturn]] void usage_error(const std::string& m) { throw ExitError(2, "error: " + m); }

// clap value_parser!(f32): Rust's f32::from_str (correctly rounded).
float parse_f32(const std::string& s) {
  char* e;
  float v = strtof(s.c_str(), &e);
  if (*e || s.empty()) usage_error("invalid float '" + s + "'");
  return v;
}

Args parse_args(int argc, char** argv) {
  Args a;
  if (argc < 2) usage_error("missing subcommand");
  a.sub = argv[1];
  if (a.sub == "genome") { a.min_covered_fraction = 10.0f; a.methods = {"relative_abundance"}; }
  else a.methods = {"mean"};
  std::vector<std::string>* multi = nullptr;
  for (int i = 2; i < argc; ++i) {
    std::string s = argv[i];
    bool is_flag = s.size() >= 2 && s[0] == '-' && !(s[1] >= '0' && s[1] <= '9') && s[1] != '.';
    if (!is_flag) {
      if (!multi) usage_error("unexpected argument '" + s + "'");
      multi->push_back(s);
      continue;
    }
    multi = nullptr;
    auto val = [&]() -> std::string {
      if (i + 1 >= argc) usage_error("missing value for " + s);
      return argv[++i];
    };
    if (s == "-b" || s == "--bam-files") multi = &a.bam_files;
    else if (s == "-m" || s == "--methods" || s == "--method") {
      if (!a.methods_given) a.methods.clear();
      a.methods_given = true;
      multi = &a.methods;
    }
    else if (s == "--min-read-aligned-length") a.min_read_aligned_length = (uint32_t)std::stoul(val());
    else if (s == "--min-read-percent-identity") a.min_read_percent_identity = parse_f32(val());
    else if (s == "--min-read-aligned-percent") a.min_read_aligned_percent = parse_f32(val());
    else if (s == "--min-read-aligned-length-pair") a.min_read_aligned_length_pair = (uint32_t)std::stoul(val());
    else if (s == "--min-read-percent-identity-pair") a.min_read_percent_identity_pair = parse_f32(val());
    else if (s == "--min-read-aligned-percent-pair") a.min_read_aligned_percent_pair = parse_f32(val());
    else if (s == "--min-mapq") a.min_mapq = (uint8_t)std::stoul(val());
    else if (s == "--proper-pairs-only") a.proper_pairs_only = true;
    else if (s == "--exclude-supplementary") a.exclude_supplementary = true;
    else if (s == "--include-secondary") a.include_secondary = true;
    else if (s == "--no-zeros") a.no_zeros = true;
    else if (s == "--min-covered-fraction") a.min_covered_fraction = parse_f32(val());
    else if (s == "--contig-end-exclusion") a.contig_end_exclusion = std::stoull(val());
    else if (s == "--trim-min") a.trim_min = parse_f32(val());
    else if (s == "--trim-max") a.trim_max = parse_f32(val());
    else if (s == "--output-format") a.output_format = val();
    else if (s == "-o" || s == "--output-file") a.output_file = val();
    else if (s == "-s" || s == "--separator") a.separator = val();
    else if (s == "--single-genome") a.single_genome = true;
    else if (s == "--genome-definition") a.genome_definition = val();
    else if (s == "--gff") a.gff = val();
    else if (s == "--gff-feature-type") a.gff_feature_type = val();
    else if (s == "-t" || s == "--threads") a.threads = std::stoi(val());
    else if (s == "--lib-estimators") a.lib_estimators = val();
    else if (s == "--lib-flags") a.lib_flags = val();
    else if (s == "--lib-streaming") a.lib_streaming = true;
    else if (s == "--print-reads-mapped") a.print_reads_mapped = true;
    else if (s == "--inverse") a.inverse = true;
    else if (s == "--timing") a.timing = true;
    else if (s == "-q" || s == "--quiet" || s == "-v" || s == "--verbose") {}
    else usage_error("unexpected argument '" + s + "'");
  }
  if ((a.min_read_aligned_length_pair || a.min_read_percent_identity_pair || a.min_read_aligned_percent_pair) &&
      !a.proper_pairs_only && !a.lib_flags)
    usage_error("the pair filters require --proper-pairs-only");  // cli.rs:2486-2499
  return a;
}

// coverm.rs:1296-1312
float parse_percentage(std::optional<float> v) {
  if (!v.has_value()) return 0.0f;
  float percentage = *v;
  if (percentage >= 1.0f && percentage <= 100.0f) percentage /= 100.0f;
  else if (!(percentage >= 0.0f && percentage <= 100.0f))
    throw ExitError(1, "Invalid alignment percentage: '" + fmt_float(percentage) + "'");
  return percentage;
}

struct FilterParameters {  // coverm.rs:1648-1704
  FlagFilter flag_filters;
  uint32_t min_aligned_length_single = 0;
  float min_percent_identity_single = 0, min_aligned_percent_single = 0;
  uint8_t min_mapq = 255;
  uint32_t min_aligned_length_pair = 0;
  float min_percent_identity_pair = 0, min_aligned_percent_pair = 0;
  bool doing_filtering() const {
    return min_percent_identity_single > 0.0f || min_percent_identity_pair > 0.0f || min_aligned_percent_single > 0.0f ||
           min_mapq < 255 || min_aligned_percent_pair > 0.0f || min_aligned_length_single > 0 || min_aligned_length_pair > 0;
  }
};

bool doing_metabat(const Args& a) {  // coverm.rs:1630-1646
  bool has = false;
  for (auto& m : a.methods) if (m == "metabat") has = true;
  if (has && a.methods.size() > 1) throw ExitError(1, "Cannot specify the metabat method with any other coverage methods");
  return has;
}

FilterParameters filter_params_from(const Args& a, bool contig_mode) {
  FilterParameters f;
  f.flag_filters.include_improper_pairs = !a.proper_pairs_only;
  f.flag_filters.include_secondary = a.include_secondary;
  f.flag_filters.include_supplementary = !a.exclude_supplementary;
  f.min_aligned_length_single = a.min_read_aligned_length.value_or(0);
  f.min_percent_identity_single = parse_percentage(a.min_read_percent_identity);
  f.min_aligned_percent_single = parse_percentage(a.min_read_aligned_percent);
  f.min_mapq = a.min_mapq.value_or(255);
  f.min_aligned_length_pair = a.min_read_aligned_length_pair.value_or(0);
  f.min_percent_identity_pair = parse_percentage(a.min_read_percent_identity_pair);
  f.min_aligned_percent_pair = parse_percentage(a.min_read_aligned_percent_pair);
  if (contig_mode && doing_metabat(a)) {  // add_metabat_filtering_if_required, coverm.rs:1680-1693
    f.min_percent_identity_single = 0.97001f;
    f.flag_filters.include_improper_pairs = true;
    f.flag_filters.include_supplementary = true;
    f.flag_filters.include_secondary = true;
  }
  if (a.lib_flags) {
    int i, s, sec;
    if (sscanf(a.lib_flags->c_str(), "%d,%d,%d", &i, &s, &sec) != 3) usage_error("--lib-flags I,S,SEC");
    f.flag_filters.include_improper_pairs = i;
    f.flag_filters.include_supplementary = s;
    f.flag_filters.include_secondary = sec;
  }
  return f;
}

struct EstimatorsAndTaker {  // coverm.rs:1315-1504
  std::vector<CoverageEstimator> estimators;
  CoverageTaker taker;
  std::vector<size_t> columns_to_normalise;
  std::optional<size_t> rpkm_column, tpm_column;
  CoveragePrinter printer;
};

std::vector<std::string> split(const std::string& s, char d) {
  std::vector<std::string> v;
  size_t a = 0;
  for (;;) {
    size_t b = s.find(d, a);
    if (b == std::string::npos) { v.push_back(s.substr(a)); break; }
    v.push_back(s.substr(a, b - a));
    a = b + 1;
  }
  return v;
}

EstimatorsAndTaker generate_estimators(const Args& a, std::ostream* stream) {
  EstimatorsAndTaker r;
  using CE = CoverageEstimator;
  if (a.lib_estimators) {  // constructor calls as written in the reference's unit tests
    for (auto& spec : split(*a.lib_estimators, ';')) {
      auto p = split(spec, ':');
      auto f = [&](size_t i) { return parse_f32(p.at(i)); };
      auto u = [&](size_t i) { return (uint64_t)std::stoull(p.at(i)); };
      CE e;
      if (p[0] == "mean") { e = CE::make(CE::Mean, f(1), u(2)); e.exclude_mismatches = p.size() > 3 && p[3] == "1"; }
      else if (p[0] == "trimmed_mean") { e = CE::make(CE::TrimmedMean, f(3), u(4)); e.min = f(1); e.max = f(2); }
      else if (p[0] == "coverage_histogram") e = CE::make(CE::PileupCounts, f(1), u(2));
      else if (p[0] == "variance") e = CE::make(CE::Variance, f(1), u(2));
      else if (p[0] == "covered_fraction") e = CE::make(CE::CoveredFraction, f(1));
      else if (p[0] == "covered_bases") e = CE::make(CE::CoveredBases, f(1));
      else if (p[0] == "rpkm") e = CE::make(CE::RPKM, f(1));
      else if (p[0] == "tpm") e = CE::make(CE::TPM, f(1));
      else if (p[0] == "length") e = CE::make(CE::Length);
      else if (p[0] == "count") e = CE::make(CE::ReadCount);
      else if (p[0] == "reads_per_base") e = CE::make(CE::ReadsPerBase);
      else if (p[0] == "anir") e = CE::make(CE::ANIr);
      else usage_error("bad --lib-estimators spec '" + spec + "'");
      r.estimators.push_back(e);
    }
    bool pileup = r.estimators.size() == 1 && r.estimators[0].kind == CE::PileupCounts;
    r.taker = pileup ? CoverageTaker::pileup(stream) : CoverageTaker::streaming(stream);
    r.printer.kind = CoveragePrinter::Streamed;
    return r;
  }
  float min_fraction_covered = parse_percentage(a.min_covered_fraction);
  uint64_t excl = a.contig_end_exclusion;
  if (doing_metabat(a)) {
    if (a.sub != "contig") usage_error("invalid value 'metabat' for '--methods'");
    r.estimators.push_back(CE::make(CE::Length));
    r.estimators.push_back(CE::make(CE::Mean, min_fraction_covered, excl));
    r.estimators.push_back(CE::make(CE::Variance, min_fraction_covered, excl));
    r.taker = CoverageTaker::cached(r.estimators.size());
    r.printer.kind = CoveragePrinter::MetabatAdjusted;
  } else {
    bool has_hist = false;
    for (size_t i = 0; i < a.methods.size(); ++i) {
      const std::string& m = a.methods[i];
      if (m == "mean") r.estimators.push_back(CE::make(CE::Mean, min_fraction_covered, excl));
      else if (m == "coverage_histogram") { r.estimators.push_back(CE::make(CE::PileupCounts, min_fraction_covered, excl)); has_hist = true; }
      else if (m == "trimmed_mean") {
        CE e = CE::make(CE::TrimmedMean, min_fraction_covered, excl);
        e.min = parse_percentage(a.trim_min);
        e.max = parse_percentage(a.trim_max);
        r.estimators.push_back(e);
      }
      else if (m == "covered_fraction") r.estimators.push_back(CE::make(CE::CoveredFraction, min_fraction_covered));
      else if (m == "covered_bases") r.estimators.push_back(CE::make(CE::CoveredBases, min_fraction_covered));
      else if (m == "rpkm") {
        if (r.rpkm_column) throw ExitError(1, "The RPKM column cannot be specified more than once");
        r.rpkm_column = i;
        r.estimators.push_back(CE::make(CE::RPKM, min_fraction_covered));
      }
      else if (m == "tpm") {
        if (r.tpm_column) throw ExitError(1, "The TPM column cannot be specified more than once");
        r.tpm_column = i;
        r.estimators.push_back(CE::make(CE::TPM, min_fraction_covered));
      }
      else if (m == "variance") r.estimators.push_back(CE::make(CE::Variance, min_fraction_covered, excl));
      else if (m == "length") r.estimators.push_back(CE::make(CE::Length));
      else if (m == "relative_abundance" && a.sub == "genome") {
        r.columns_to_normalise.push_back(i);
        r.estimators.push_back(CE::make(CE::Mean, min_fraction_covered, excl));
      }
      else if (m == "count") r.estimators.push_back(CE::make(CE::ReadCount));
      else if (m == "reads_per_base") r.estimators.push_back(CE::make(CE::ReadsPerBase));
      else if (m == "anir") r.estimators.push_back(CE::make(CE::ANIr));
      else usage_error("invalid value '" + m + "' for '--methods'");
    }
    if (has_hist) {
      if (a.methods.size() > 1) throw ExitError(1, "Cannot specify the coverage_histogram method with any other coverage methods");
      r.taker = CoverageTaker::pileup(stream);
      r.printer.kind = CoveragePrinter::Streamed;
    } else if (r.columns_to_normalise.empty() && !r.rpkm_column && !r.tpm_column && a.output_format == "sparse") {
      r.taker = CoverageTaker::streaming(stream);
      r.printer.kind = CoveragePrinter::Streamed;
    } else {
      r.taker = CoverageTaker::cached(r.estimators.size());
      if (a.output_format == "sparse") r.printer.kind = CoveragePrinter::SparseCached;
      else if (a.output_format == "dense") r.printer.kind = CoveragePrinter::DenseCached;
      else usage_error("invalid value '" + a.output_format + "' for '--output-format'");
    }
  }
  if (min_fraction_covered != 0.0f) {  // coverm.rs:1473-1494
    for (auto& e : r.estimators) {
      const char* n = nullptr;
      if (e.kind == CE::ReadCount) n = "counts";
      if (e.kind == CE::Length) n = "length";
      if (e.kind == CE::ReadsPerBase) n = "reads_per_base";
      if (e.kind == CE::ANIr) n = "anir";
      if (n)
        throw ExitError(1, std::string("The '") + n + "' coverage estimator cannot be used when --min-covered-fraction is > 0 as it does not calculate the covered fraction. You may wish to set the --min-covered-fraction to 0 and/or run this estimator separately.");
    }
  }
  return r;
}

void print_headers(EstimatorsAndTaker& et, const std::string& entry_type, std::ostream& os) {  // coverm.rs:1506-1519
  std::vector<std::string> headers;
  for (auto& e : et.estimators) for (auto& h : e.column_headers()) headers.push_back(h);
  for (size_t i : et.columns_to_normalise) headers[i] = "Relative Abundance (%)";
  et.printer.print_headers(entry_type, headers, os);
}

std::vector<ReaderFactory> make_readers(const Args& a, const FilterParameters& fp) {
  std::vector<ReaderFactory> v;
  bool filtering = fp.doing_filtering();
  for (auto& path : a.bam_files) {
    if (filtering) {  // generate_filtered_bam_readers_from_bam_files, bam_generator.rs:563-607
      v.push_back([=]() -> std::unique_ptr<NamedBamReader> {
        return std::make_unique<FilteredBamReader>(path, a.threads, fp.flag_filters, fp.min_aligned_length_single,
                                                   fp.min_percent_identity_single, fp.min_aligned_percent_single,
                                                   fp.min_mapq, fp.min_aligned_length_pair, fp.min_percent_identity_pair,
                                                   fp.min_aligned_percent_pair, true);
      });
    } else {  // generate_named_bam_readers_from_bam_files, bam_generator.rs:356-371
      v.push_back([=]() -> std::unique_ptr<NamedBamReader> { return std::make_unique<BamFileNamedReader>(path, a.threads); });
    }
  }
  return v;
}

int run(int argc, char** argv, std::ostream& out_default) {
  Args a = parse_args(argc, argv);
  std::ofstream ofs;
  std::ostream* os = &out_default;
  if (a.output_file && *a.output_file != "-") {
    ofs.open(*a.output_file);
    if (!ofs) throw Panic("Failed to create output file: " + *a.output_file);
    os = &ofs;
  }
  if (a.bam_files.empty()) usage_error("--bam-files is required (read mapping is out of scope for the oracle)");
  auto t0 = std::chrono::steady_clock::now();

  if (a.sub == "filter-names") {
    FilterParameters fp = filter_params_from(a, false);
    for (auto& path : a.bam_files) {
      FilteredBamReader r(path, a.threads, fp.flag_filters, fp.min_aligned_length_single, fp.min_percent_identity_single,
                          fp.min_aligned_percent_single, fp.min_mapq, fp.min_aligned_length_pair,
                          fp.min_percent_identity_pair, fp.min_aligned_percent_pair, !a.inverse);
      Record rec;
      while (r.read(rec)) *os << rec.qname << "\n";
    }
    return 0;
  }

  std::vector<ReadsMapped> reads_mapped;
  std::vector<std::string> stoits;
  for (auto& p : a.bam_files) stoits.push_back(file_stem(p));
  if (a.sub == "contig") {
    FilterParameters fp = filter_params_from(a, true);
    std::optional<GeneDefinitions> gene_definitions;  // coverm.rs:488-509
    if (a.gff) {
      if (doing_metabat(a)) throw ExitError(1, "The metabat method cannot be used with --gff");
      gene_definitions = read_gff(*a.gff, a.gff_feature_type);
    }
    EstimatorsAndTaker et = generate_estimators(a, os);
    if (!a.lib_streaming) print_headers(et, gene_definitions ? "Gene\tContig" : "Contig", *os);
    auto readers = make_readers(a, fp);
    if (gene_definitions)  // run_contig, coverm.rs:2100-2111
      reads_mapped = gene_coverage(readers, et.taker, et.estimators, *gene_definitions, nullptr, !a.no_zeros, fp.flag_filters);
    else
      reads_mapped = contig_coverage(readers, et.taker, et.estimators, !a.no_zeros, fp.flag_filters);
    et.printer.finalise_printing(et.taker, *os, &reads_mapped, et.columns_to_normalise, et.rpkm_column, et.tpm_column);
  } else if (a.sub == "genome") {
    FilterParameters fp = filter_params_from(a, false);
    EstimatorsAndTaker et = generate_estimators(a, os);
    if (!a.lib_streaming) print_headers(et, a.gff ? "Gene\tContig\tGenome" : "Genome", *os);  // coverm.rs:67-74
    // parse_separator, coverm.rs:1522-1537
    std::optional<uint8_t> separator;
    if (a.single_genome) separator = (uint8_t)'0';
    else if (a.separator) {
      if (a.separator->size() != 1) usage_error("separator must be a single character");
      separator = (uint8_t)(*a.separator)[0];
    }
    auto readers = make_readers(a, fp);
    if (a.gff) {  // coverm.rs:1554-1590: per-gene coverage with a genome column; separator / single-genome win over the definition file
      const GeneDefinitions gene_definitions = read_gff(*a.gff, a.gff_feature_type);
      GenomesAndContigs gc;
      GenomeNamer namer;
      if (a.single_genome) namer = [](const std::string&) { return std::optional<std::string>("genome1"); };
      else if (separator) {
        const char sep = (char)*separator;
        namer = [sep](const std::string& contig) -> std::optional<std::string> {
          const size_t at = contig.find(sep);
          if (at == std::string::npos) return std::nullopt;
          return contig.substr(0, at);
        };
      } else {
        if (!a.genome_definition) usage_error("a genome definition is required when using --gff in genome mode");
        gc = read_genome_definition_file(*a.genome_definition);
        namer = [&gc](const std::string& contig) -> std::optional<std::string> {
          auto it = gc.contig_to_genome.find(contig);
          if (it == gc.contig_to_genome.end()) return std::nullopt;
          return gc.genomes[it->second];
        };
      }
      reads_mapped = gene_coverage(readers, et.taker, et.estimators, gene_definitions, &namer, !a.no_zeros, fp.flag_filters);
    } else if (separator.has_value() || a.single_genome) {
      reads_mapped = mosdepth_genome_coverage(readers, *separator, et.taker, !a.no_zeros, et.estimators,
                                              fp.flag_filters, a.single_genome);
    } else {
      if (!a.genome_definition)
        usage_error("one of --separator, --single-genome, --genome-definition is required (FASTA genome input is out of scope)");
      GenomesAndContigs gc = read_genome_definition_file(*a.genome_definition);
      reads_mapped = mosdepth_genome_coverage_with_contig_names(readers, gc, et.taker, !a.no_zeros, fp.flag_filters,
                                                                et.estimators);
    }
    et.printer.finalise_printing(et.taker, *os, &reads_mapped, et.columns_to_normalise, et.rpkm_column, et.tpm_column);
  } else {
    usage_error("unknown subcommand '" + a.sub + "'");
  }
  os->flush();
  if (a.print_reads_mapped)
    for (size_t i = 0; i < reads_mapped.size(); ++i)
      std::cerr << "#reads_mapped\t" << stoits[i] << "\t" << reads_mapped[i].num_mapped_reads << "\t"
                << reads_mapped[i].num_reads << "\n";
  if (a.timing) {
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::cerr << "#timing\t" << s << "\n";
  }
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  std::ios::sync_with_stdio(false);
  try {
    return run(argc, argv, std::cout);
  } catch (const Panic& p) {
    std::cout.flush();
    std::cerr << "thread 'main' panicked: " << p.what() << "\n";
    return 101;
  } catch (const ExitError& e) {
    std::cout.flush();
    std::cerr << "[ERROR] " << e.what() << "\n";
    return e.code;
  } catch (const std::exception& e) {
    std::cout.flush();
    std::cerr << "thread 'main' panicked: " << e.what() << "\n";
    return 101;
  }
}
