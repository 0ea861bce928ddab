// bamgen — synthetic reference-sorted BAM generator for the BASELINE.json configurations:
// log-normal contig lengths, reads placed per contig (abundance x length), 150 bp reads with a CIGAR / NM / MAPQ /
// FLAG mixture (proper pairs sharing a qname on one contig, improper pairs, secondary, supplementary, unmapped mates,
// unmapped pairs at the end), real SEQ/QUAL bytes, BGZF level-1 64 KiB blocks, written in parallel.
// Output is deterministic in (seed, parameters) and independent of the thread count.
//
//   bamgen --out x.bam --contigs N --reads R [--median-len 4000 --sigma 0.8 --min-len 1000 --max-len 2000000]
//          [--genomes G (names magNNNN~ctgNNNNN, definition TSV via --definition-out)] [--seed S] [--threads T]
//          [--read-len 150] [--contig-table names_and_lengths.tsv]
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Rng {  // splitmix64
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed) {}
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
  uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
  double normal() {
    double u1 = uni(), u2 = uni();
    if (u1 < 1e-300) u1 = 1e-300;
    return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
  }
  uint32_t poisson(double lambda) {
    if (lambda > 30) return (uint32_t)std::max(0.0, std::floor(lambda + std::sqrt(lambda) * normal() + 0.5));
    const double l = std::exp(-lambda);
    uint32_t k = 0;
    double p = 1.0;
    do {
      ++k;
      p *= uni();
    } while (p > l);
    return k - 1;
  }
};

struct Args {
  std::string out, definition_out, contig_table;
  uint64_t contigs = 1000, reads = 100000, seed = 1;
  double median_len = 4000, sigma = 0.8;
  uint64_t min_len = 1000, max_len = 2000000;
  uint32_t genomes = 0, read_len = 150;
  int threads = 0;
};

void put32(std::vector<uint8_t>& v, uint32_t x) {
  const size_t o = v.size();
  v.resize(o + 4);
  memcpy(v.data() + o, &x, 4);
}
void put16(std::vector<uint8_t>& v, uint16_t x) {
  const size_t o = v.size();
  v.resize(o + 2);
  memcpy(v.data() + o, &x, 2);
}

// Append `raw` as BGZF blocks (<= 0xff00 payload each) to out.
void bgzf_append(const uint8_t* raw, size_t n, std::vector<uint8_t>& out, z_stream& zs) {
  const size_t BLOCK = 0xff00;
  std::vector<uint8_t> buf(BLOCK + 1024);
  for (size_t o = 0; o < n || (n == 0 && o == 0); o += BLOCK) {
    const size_t len = std::min(BLOCK, n - o);
    deflateReset(&zs);
    zs.next_in = const_cast<Bytef*>(raw + o);
    zs.avail_in = (uInt)len;
    zs.next_out = buf.data();
    zs.avail_out = (uInt)buf.size();
    if (deflate(&zs, Z_FINISH) != Z_STREAM_END) {
      fprintf(stderr, "bamgen: deflate failed\n");
      exit(1);
    }
    const size_t clen = buf.size() - zs.avail_out;
    const uint8_t hdr[12] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0};
    out.insert(out.end(), hdr, hdr + 12);
    out.push_back('B');
    out.push_back('C');
    put16(out, 2);
    put16(out, (uint16_t)(clen + 25));
    out.insert(out.end(), buf.data(), buf.data() + clen);
    put32(out, (uint32_t)crc32(crc32(0, nullptr, 0), raw + o, (uInt)len));
    put32(out, (uint32_t)len);
    if (n == 0) break;
  }
}

struct Rec {
  int32_t pos;
  uint32_t order;
  std::vector<uint8_t> bytes;
};

void emit_record(std::vector<uint8_t>& b, int32_t tid, int32_t pos, uint8_t mapq, uint16_t flag, int32_t mtid, int32_t mpos,
                 const std::string& qname, const std::vector<uint32_t>& cigar, uint32_t l_seq, uint32_t nm, Rng& rng) {
  std::vector<uint8_t> body;
  put32(body, (uint32_t)tid);
  put32(body, (uint32_t)pos);
  body.push_back((uint8_t)(qname.size() + 1));
  body.push_back(mapq);
  put16(body, 4680);
  put16(body, (uint16_t)cigar.size());
  put16(body, flag);
  put32(body, l_seq);
  put32(body, (uint32_t)mtid);
  put32(body, (uint32_t)mpos);
  put32(body, 0);
  body.insert(body.end(), qname.begin(), qname.end());
  body.push_back(0);
  for (uint32_t c : cigar) put32(body, c);
  for (uint32_t i = 0; i < (l_seq + 1) / 2; ++i) {
    const uint64_t r = rng.next();
    body.push_back((uint8_t)((1u << (r & 3)) << 4 | (1u << ((r >> 2) & 3))));
  }
  for (uint32_t i = 0; i < l_seq; ++i) body.push_back((uint8_t)(20 + (rng.next() & 15)));
  body.push_back('N');
  body.push_back('M');
  if (nm < 256) {
    body.push_back('C');
    body.push_back((uint8_t)nm);
  } else {
    body.push_back('S');
    put16(body, (uint16_t)nm);
  }
  body.push_back('A');
  body.push_back('S');
  body.push_back('C');
  body.push_back((uint8_t)std::min<uint32_t>(255, l_seq - std::min(l_seq, nm * 5)));
  put32(b, (uint32_t)body.size());
  b.insert(b.end(), body.begin(), body.end());
}

// One read's alignment: CIGAR mixture of SURVEY.md §8d.
void make_alignment(Rng& rng, uint32_t L, std::vector<uint32_t>& cigar, uint32_t& nm) {
  cigar.clear();
  const double u = rng.uni();
  uint32_t indel = 0;
  auto op = [](uint32_t len, uint32_t code) { return (len << 4) | code; };
  if (u < 0.88) {
    cigar.push_back(op(L, 0));
  } else if (u < 0.92) {
    const uint32_t a = 20 + rng.below(L - 40), b = 1 + rng.below(5);
    cigar = {op(a, 0), op(b, 2), op(L - a, 0)};
    indel = b;
  } else if (u < 0.96) {
    const uint32_t a = 20 + rng.below(L - 40), b = 1 + rng.below(5);
    cigar = {op(a, 0), op(b, 1), op(L - a - b, 0)};
    indel = b;
  } else if (u < 0.99) {
    const uint32_t s = 5 + rng.below(56);
    cigar = {op(s, 4), op(L - s, 0)};
  } else {
    const uint32_t a = 20 + rng.below(L - 40), b = 50 + rng.below(500);
    cigar = {op(a, 0), op(b, 3), op(L - a, 0)};
  }
  nm = std::min<uint32_t>(L, rng.poisson(1.5) + indel);
}

uint8_t make_mapq(Rng& rng) {
  const double u = rng.uni();
  if (u < 0.05) return 0;
  if (u < 0.20) return (uint8_t)(1 + rng.below(59));
  return 60;
}

}  // namespace

int main(int argc, char** argv) {
  Args a;
  for (int i = 1; i < argc; ++i) {
    const std::string s = argv[i];
    auto val = [&]() -> const char* {
      if (i + 1 >= argc) {
        fprintf(stderr, "bamgen: missing value for %s\n", s.c_str());
        exit(2);
      }
      return argv[++i];
    };
    if (s == "--out") a.out = val();
    else if (s == "--definition-out") a.definition_out = val();
    else if (s == "--contigs") a.contigs = strtoull(val(), nullptr, 10);
    else if (s == "--contig-table") a.contig_table = val();  // "name<TAB>length" lines: use these reference sequences
    else if (s == "--reads") a.reads = strtoull(val(), nullptr, 10);
    else if (s == "--seed") a.seed = strtoull(val(), nullptr, 10);
    else if (s == "--median-len") a.median_len = atof(val());
    else if (s == "--sigma") a.sigma = atof(val());
    else if (s == "--min-len") a.min_len = strtoull(val(), nullptr, 10);
    else if (s == "--max-len") a.max_len = strtoull(val(), nullptr, 10);
    else if (s == "--genomes") a.genomes = (uint32_t)atoi(val());
    else if (s == "--read-len") a.read_len = (uint32_t)atoi(val());
    else if (s == "--threads") a.threads = atoi(val());
    else {
      fprintf(stderr, "bamgen: unknown argument %s\n", s.c_str());
      return 2;
    }
  }
  if (a.out.empty()) {
    fprintf(stderr, "bamgen: --out is required\n");
    return 2;
  }
  if (a.threads <= 0) a.threads = (int)std::max(1u, std::thread::hardware_concurrency());
  const uint32_t RL = std::max<uint32_t>(80, a.read_len);

  // ---- reference
  std::vector<std::pair<std::string, uint32_t>> table;
  if (!a.contig_table.empty()) {
    FILE* tf = fopen(a.contig_table.c_str(), "r");
    if (!tf) {
      fprintf(stderr, "bamgen: cannot read %s\n", a.contig_table.c_str());
      return 2;
    }
    char line[4096];
    while (fgets(line, sizeof line, tf)) {
      char* tab = strchr(line, '\t');
      if (!tab) continue;
      *tab = 0;
      table.emplace_back(line, (uint32_t)strtoul(tab + 1, nullptr, 10));
    }
    fclose(tf);
    a.contigs = table.size();
    a.genomes = 0;
  }
  Rng rr(a.seed * 0x9E3779B97F4A7C15ull + 17);
  std::vector<uint32_t> len(a.contigs);
  std::vector<std::string> names(a.contigs);
  std::vector<double> weight(a.contigs);
  double wsum = 0;
  std::vector<uint32_t> genome_of(a.contigs, 0);
  {
    uint32_t g = 0, in_g = 0, g_size = 0;
    for (uint64_t c = 0; c < a.contigs; ++c) {
      double l = std::exp(std::log(a.median_len) + a.sigma * rr.normal());
      l = std::min<double>((double)a.max_len, std::max<double>((double)a.min_len, std::round(l)));
      len[c] = (uint32_t)l;
      char buf[64];
      if (a.genomes) {
        if (in_g == g_size) {
          if (c) ++g;
          g = std::min(g, a.genomes - 1);
          in_g = 0;
          const uint64_t remaining_c = a.contigs - c, remaining_g = a.genomes - g;
          g_size = (uint32_t)std::max<uint64_t>(1, remaining_g <= 1 ? remaining_c : std::min<uint64_t>(remaining_c - (remaining_g - 1), (uint64_t)(0.25 * remaining_c / remaining_g + rr.uni() * 1.5 * remaining_c / remaining_g)));
        }
        snprintf(buf, sizeof buf, "mag%04u~ctg%05u", g, in_g);
        genome_of[c] = g;
        ++in_g;
      } else {
        snprintf(buf, sizeof buf, "c%07llu", (unsigned long long)c);
      }
      names[c] = buf;
      if (!table.empty()) {
        names[c] = table[c].first;
        len[c] = std::max<uint32_t>(1, table[c].second);
      }
    }
    // abundance is per genome (or per contig when there are no genomes): log-normal sigma 1.5
    std::vector<double> gab(a.genomes ? a.genomes : 0);
    for (auto& x : gab) x = std::exp(1.5 * rr.normal());
    for (uint64_t c = 0; c < a.contigs; ++c) {
      const double ab = a.genomes ? gab[genome_of[c]] : std::exp(1.5 * rr.normal());
      weight[c] = ab * len[c];
      wsum += weight[c];
    }
  }
  const uint64_t mapped_target = (uint64_t)((double)a.reads * 0.99);  // ~1 % unmapped pairs at the end

  // ---- header block(s)
  std::vector<uint8_t> hdr;
  hdr.insert(hdr.end(), {'B', 'A', 'M', 1});
  const std::string text = "@HD\tVN:1.6\tSO:coordinate\n";
  put32(hdr, (uint32_t)text.size());
  hdr.insert(hdr.end(), text.begin(), text.end());
  put32(hdr, (uint32_t)a.contigs);
  for (uint64_t c = 0; c < a.contigs; ++c) {
    put32(hdr, (uint32_t)names[c].size() + 1);
    hdr.insert(hdr.end(), names[c].begin(), names[c].end());
    hdr.push_back(0);
    put32(hdr, len[c]);
  }
  if (!a.definition_out.empty()) {
    FILE* df = fopen(a.definition_out.c_str(), "w");
    if (!df) {
      perror("bamgen: definition-out");
      return 1;
    }
    for (uint64_t c = 0; c < a.contigs; ++c) {
      const std::string g = names[c].substr(0, names[c].find('~'));
      fprintf(df, "%s\t%s\n", g.c_str(), names[c].c_str());
    }
    fclose(df);
  }

  // ---- records: contiguous contig ranges per task, each producing complete BGZF blocks
  const size_t n_tasks = (size_t)std::max<uint64_t>(1, std::min<uint64_t>(a.contigs, (uint64_t)a.threads * 8));
  std::vector<uint64_t> task_begin(n_tasks + 1, a.contigs);
  {
    double acc = 0;
    size_t t = 0;
    task_begin[0] = 0;
    for (uint64_t c = 0; c < a.contigs && t + 1 < n_tasks; ++c) {
      acc += weight[c];
      if (acc >= wsum * (double)(t + 1) / (double)n_tasks) task_begin[++t] = c + 1;
    }
    for (size_t k = t + 1; k <= n_tasks; ++k) task_begin[k] = a.contigs;
  }
  std::vector<std::vector<uint8_t>> task_out(n_tasks);
  std::vector<uint64_t> task_records(n_tasks, 0);
  std::atomic<size_t> next{0};
  auto worker = [&]() {
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    deflateInit2(&zs, 1, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
    std::vector<uint8_t> raw;
    std::vector<Rec> recs;
    std::vector<uint32_t> cigar;
    for (;;) {
      const size_t t = next.fetch_add(1);
      if (t >= n_tasks) break;
      raw.clear();
      for (uint64_t c = task_begin[t]; c < task_begin[t + 1]; ++c) {
        Rng rng(a.seed ^ (c * 0xD1B54A32D192ED03ull + 0x2545F4914F6CDD1Dull));
        const double expect = (double)mapped_target * weight[c] / wsum;
        const uint32_t n_templates = rng.poisson(expect / 1.9);  // most templates are pairs
        recs.clear();
        uint32_t order = 0;
        const uint32_t L = len[c];
        for (uint32_t k = 0; k < n_templates; ++k) {
          char qn[48];
          snprintf(qn, sizeof qn, "r%llu.%u", (unsigned long long)c, k);
          const std::string qname = qn;
          const double u = rng.uni();
          const int32_t p1 = (int32_t)rng.below(L);
          uint32_t nm;
          auto add = [&](int32_t pos, uint16_t flag, int32_t mpos, uint8_t mapq) {
            make_alignment(rng, RL, cigar, nm);
            // every M block must START inside the contig (the reference indexes ups_and_downs[cursor], contig.rs:178)
            uint32_t ref_off = 0, last_m_start = 0;
            for (uint32_t cg : cigar) {
              const uint32_t op = cg & 0xf, l = cg >> 4;
              if (op == 0) last_m_start = ref_off;
              if (op == 0 || op == 2 || op == 3) ref_off += l;
            }
            if ((uint32_t)pos + last_m_start >= L) {
              if (L > last_m_start) pos = (int32_t)(L - 1 - last_m_start);
              else { cigar = {(RL << 4) | 0u}; nm = std::min(nm, RL); }
            }
            Rec r;
            r.pos = pos;
            r.order = order++;
            emit_record(r.bytes, (int32_t)c, pos, mapq, flag, (int32_t)c, mpos, qname, cigar, RL, nm, rng);
            recs.push_back(std::move(r));
          };
          if (u < 0.88) {  // proper pair on this contig
            const int32_t p2 = (int32_t)std::min<uint32_t>(L - 1, (uint32_t)p1 + 100 + rng.below(400));
            const uint8_t q = make_mapq(rng);
            add(p1, 99, p2, q);
            add(p2, 147, p1, make_mapq(rng));
          } else if (u < 0.94) {  // improper pair
            const int32_t p2 = (int32_t)rng.below(L);
            add(p1, 65, p2, make_mapq(rng));
            add(p2, 129, p1, make_mapq(rng));
          } else if (u < 0.96) {  // primary + secondary
            add(p1, 0, -1, make_mapq(rng));
            add((int32_t)rng.below(L), 256, -1, 0);
          } else if (u < 0.98) {  // primary + supplementary
            add(p1, 0, -1, make_mapq(rng));
            add((int32_t)rng.below(L), 2048, -1, make_mapq(rng));
          } else {  // one mate mapped, the other unmapped but placed with its mate (flag 0x4 set, tid/pos of the mate)
            add(p1, 73, p1, make_mapq(rng));
            Rec r;
            r.pos = p1;
            r.order = order++;
            emit_record(r.bytes, (int32_t)c, p1, 0, 133, (int32_t)c, p1, qname, {}, RL, 0, rng);
            recs.push_back(std::move(r));
          }
        }
        std::sort(recs.begin(), recs.end(), [](const Rec& x, const Rec& y) { return x.pos != y.pos ? x.pos < y.pos : x.order < y.order; });
        for (auto& r : recs) raw.insert(raw.end(), r.bytes.begin(), r.bytes.end());
        task_records[t] += recs.size();
      }
      bgzf_append(raw.data(), raw.size(), task_out[t], zs);
    }
    deflateEnd(&zs);
  };
  std::vector<std::thread> pool;
  for (int i = 0; i < a.threads; ++i) pool.emplace_back(worker);
  for (auto& th : pool) th.join();

  // ---- unmapped pairs at the end of the file
  std::vector<uint8_t> tail_raw, tail;
  uint64_t n_tail = 0;
  {
    Rng rng(a.seed + 999);
    const uint64_t n_unmapped_pairs = (a.reads - mapped_target) / 2;
    for (uint64_t k = 0; k < n_unmapped_pairs; ++k) {
      char qn[48];
      snprintf(qn, sizeof qn, "u%llu", (unsigned long long)k);
      emit_record(tail_raw, -1, -1, 0, 77, -1, -1, qn, {}, RL, 0, rng);
      emit_record(tail_raw, -1, -1, 0, 141, -1, -1, qn, {}, RL, 0, rng);
      n_tail += 2;
    }
  }
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  deflateInit2(&zs, 1, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
  std::vector<uint8_t> hdr_out;
  bgzf_append(hdr.data(), hdr.size(), hdr_out, zs);
  if (!tail_raw.empty()) bgzf_append(tail_raw.data(), tail_raw.size(), tail, zs);
  std::vector<uint8_t> eof;
  bgzf_append(nullptr, 0, eof, zs);
  deflateEnd(&zs);

  FILE* f = fopen(a.out.c_str(), "wb");
  if (!f) {
    perror("bamgen: out");
    return 1;
  }
  fwrite(hdr_out.data(), 1, hdr_out.size(), f);
  uint64_t total = n_tail, bases = 0;
  for (size_t t = 0; t < n_tasks; ++t) {
    if (task_records[t]) fwrite(task_out[t].data(), 1, task_out[t].size(), f);
    total += task_records[t];
  }
  fwrite(tail.data(), 1, tail.size(), f);
  fwrite(eof.data(), 1, eof.size(), f);
  fclose(f);
  for (auto l : len) bases += l;
  printf("{\"contigs\": %llu, \"bases\": %llu, \"records\": %llu}\n", (unsigned long long)a.contigs, (unsigned long long)bases,
         (unsigned long long)total);
  return 0;
}
