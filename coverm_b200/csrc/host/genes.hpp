// Per-gene coverage (`--gff`): the host half of src/genes.rs.
//   GeneDefinitions::read_gff            genes.rs:44-126   (GFF3 / GTF lines -> genes, ids from the attributes column)
//   resolve_genes_against_header         genes.rs:351-432  (genes -> tids of this BAM's header, clamped, sorted per contig)
//   gene_coverage / emit_genes_for_contig genes.rs:182-344, 434-568 -- see drivers.hpp: the per-gene arrays are built on
//   the device (cmb_set_genes: every aligned block clipped to the genes it overlaps), the estimator maths is replayed here.
#pragma once
#include <algorithm>
#include <functional>
#include <optional>
#include <string>
#include <unordered_map>
#include <vector>

#include "bam_source.hpp"

namespace cmbh {

struct Gene {  // 0-based half-open range on a named contig
  std::string id, contig;
  uint64_t start = 0, end = 0;
};
struct GeneDefinitions {
  std::vector<Gene> genes;
};
using GenomeNamer = std::function<std::optional<std::string>(const std::string&)>;

namespace gff {
inline bool is_space(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13); }
inline std::string trimmed(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && is_space((unsigned char)s[a])) ++a;
  while (b > a && is_space((unsigned char)s[b - 1])) --b;
  return s.substr(a, b - a);
}
// value of `key` in a `;`-separated attribute list: GFF3 `key=value` or GTF `key "value"` (genes.rs:146-164); the first
// entry that carries the key decides, even with an empty value
inline bool attribute(const std::string& attrs, const char* key, std::string& out) {
  const std::string k(key);
  for (size_t a = 0; a <= attrs.size();) {
    size_t b = attrs.find(';', a);
    if (b == std::string::npos) b = attrs.size();
    const std::string entry = trimmed(attrs.substr(a, b - a));
    if (entry.size() > k.size() && entry.compare(0, k.size(), k) == 0) {
      const std::string rest = entry.substr(k.size() + 1);
      if (entry[k.size()] == '=') {
        out = trimmed(rest);
        return true;
      }
      if (entry[k.size()] == ' ') {
        const std::string v = trimmed(rest);
        size_t x = 0, y = v.size();
        while (x < y && v[x] == '"') ++x;
        while (y > x && v[y - 1] == '"') --y;
        out = v.substr(x, y - x);
        return true;
      }
    }
    a = b + 1;
  }
  return false;
}
inline bool parse_coordinate(const std::string& t, uint64_t& v) {  // Rust u64::from_str: optional '+', decimal digits, no overflow
  size_t i = !t.empty() && t[0] == '+' ? 1 : 0;
  if (i >= t.size()) return false;
  unsigned __int128 acc = 0;
  for (; i < t.size(); ++i) {
    if (t[i] < '0' || t[i] > '9') return false;
    acc = acc * 10 + (unsigned)(t[i] - '0');
    if (acc > (unsigned __int128)UINT64_MAX) return false;
  }
  v = (uint64_t)acc;
  return true;
}
}  // namespace gff

inline GeneDefinitions read_gff(const std::string& path, const std::optional<std::string>& feature_type) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) throw Panic("Failed to open GFF file " + path);
  std::string text;
  char chunk[1 << 16];
  size_t got;
  while ((got = fread(chunk, 1, sizeof chunk, f)) > 0) text.append(chunk, got);
  fclose(f);
  GeneDefinitions defs;
  uint64_t auto_id = 0;
  for (size_t a = 0; a < text.size();) {
    size_t b = text.find('\n', a);
    if (b == std::string::npos) b = text.size();
    std::string line = text.substr(a, b - a);
    a = b + 1;
    while (!line.empty() && gff::is_space((unsigned char)line.back())) line.pop_back();  // lines() drops "\r\n", then trim_end
    if (line.empty() || line[0] == '#') continue;
    std::vector<std::string> col;
    for (size_t x = 0;;) {
      const size_t y = line.find('\t', x);
      col.push_back(line.substr(x, y == std::string::npos ? std::string::npos : y - x));
      if (y == std::string::npos) break;
      x = y + 1;
    }
    if (col.size() < 8) continue;                                   // malformed line: skipped with a warning
    if (feature_type && col[2] != *feature_type) continue;
    uint64_t first, last;
    if (!gff::parse_coordinate(col[3], first) || !gff::parse_coordinate(col[4], last)) continue;
    if (first == 0 || last < first) continue;
    Gene g;
    g.contig = col[0];
    g.start = first - 1;  // 1-based inclusive -> 0-based half-open
    g.end = last;
    bool named = false;
    if (col.size() > 8) {
      for (const char* key : {"ID", "locus_tag", "gene_id", "Name", "gene", "Parent"}) {
        std::string v;
        if (gff::attribute(col[8], key, v) && !v.empty()) {
          g.id = v;
          named = true;
          break;
        }
      }
    }
    if (!named) g.id = g.contig + "_gene_" + std::to_string(++auto_id);
    defs.genes.push_back(std::move(g));
  }
  return defs;
}

// The genes of one BAM header, in entry order: by tid, then by start (stable), clamped to their contigs.
struct ResolvedGenes {
  struct Entry {
    std::string name;  // "gene\tcontig" or "gene\tcontig\tgenome": the entry's output columns
    uint32_t tid, start, end;
  };
  std::vector<Entry> entries;           // index == entry id == device row
  std::vector<uint32_t> first_of_tid;   // n_ref + 1
};

inline ResolvedGenes resolve_genes_against_header(const GeneDefinitions& defs, const Header& header, const GenomeNamer* genome_namer) {
  std::unordered_map<std::string, uint32_t> tid_of;
  for (uint32_t t = 0; t < header.names.size(); ++t) tid_of[header.names[t]] = t;  // a repeated name keeps the last tid (HashMap::insert)
  std::vector<std::vector<ResolvedGenes::Entry>> per_tid(header.names.size());
  for (const Gene& g : defs.genes) {
    auto it = tid_of.find(g.contig);
    if (it == tid_of.end()) continue;
    const uint64_t L = header.lens[it->second];
    const uint64_t s = std::min(g.start, L), e = std::min(g.end, L);
    if (s >= e) continue;
    ResolvedGenes::Entry en;
    if (genome_namer) {
      const std::optional<std::string> genome = (*genome_namer)(g.contig);
      if (!genome) continue;  // genes on contigs outside every genome are not reported
      en.name = g.id + "\t" + g.contig + "\t" + *genome;
    } else {
      en.name = g.id + "\t" + g.contig;
    }
    en.tid = it->second;
    en.start = (uint32_t)s;
    en.end = (uint32_t)e;
    per_tid[it->second].push_back(std::move(en));
  }
  ResolvedGenes r;
  r.first_of_tid.assign(header.names.size() + 1, 0);
  for (uint32_t t = 0; t < per_tid.size(); ++t) {
    std::stable_sort(per_tid[t].begin(), per_tid[t].end(), [](const auto& x, const auto& y) { return x.start < y.start; });
    r.first_of_tid[t] = (uint32_t)r.entries.size();
    for (auto& en : per_tid[t]) r.entries.push_back(std::move(en));
  }
  r.first_of_tid[per_tid.size()] = (uint32_t)r.entries.size();
  return r;
}

}  // namespace cmbh
