// TEST INFRASTRUCTURE ONLY — never linked into libcoverm_b200.so or the `coverm` product binary.
//
// A plain-C++ stand-in for the device half of include/coverm_b200.h, so that the product's HOST code (BGZF/BAM
// decode, batching, mate matching, driver replay, estimator finalisation, printers, CLI) can be checked against the
// reference's golden vectors in the GPU-less build container (`oracle/coverm_hostcheck`, tests/test_host_golden.py).
// It follows the same reference rules as the oracle (contig.rs:166-211, filter.rs:243-336, lib.rs:59-79,
// EST:366-502, 591-642, 790-805) with dense per-contig arrays; the CUDA kernels are validated separately, on the
// GPU, against the oracle through the real library.
#include <cmath>
#include <cstdlib>
#include <algorithm>
#include <climits>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include <zlib.h>

#include "../include/coverm_b200.h"

struct cmb_ctx {
  std::string err;
  cmb_device_cfg cfg{};
  std::vector<std::vector<uint8_t>> slabs;
  std::vector<cmb_read_batch> batches;
  uint32_t n_acquired = 0;
  uint32_t next = 0;
  std::vector<uint64_t> lens;
  uint32_t tid_begin = 0, tid_end = 0;
  cmb_params p{};
  cmb_filter_mode mode{};
  bool in_sample = false, ended = false;
  std::map<uint32_t, std::vector<int32_t>> arena;
  std::vector<cmb_contig_stats> rows;
  std::vector<cmb_hist_pair> pairs;
  int64_t last_kept_tid = INT64_MIN;
  int64_t excl_min = INT64_MAX, excl_max = INT64_MIN;  // kept tid range of the exclusive records (cmb_kept_tid_range)
  // gene mode (cmb_set_genes): lens / rows / arena are per gene; records carry contig tids
  bool gene_mode = false;
  std::vector<uint64_t> contig_lens;
  std::vector<cmb_gene> genes;
  std::vector<uint32_t> gene_first;
  std::vector<uint8_t> contig_seen;
  uint64_t kept_primary = 0;
  int error = 0;
  uint64_t n_records = 0, n_intervals = 0;
};

static std::string g_err;
static int fail(cmb_ctx* c, int code, const std::string& m) {
  if (c) c->err = m;
  else g_err = m;
  return code;
}

extern "C" {

int cmb_abi_version(void) { return CMB_ABI_VERSION; }
const char* cmb_last_error(const cmb_ctx* c) { return c ? c->err.c_str() : g_err.c_str(); }

int cmb_create(const cmb_device_cfg* cfg, cmb_ctx** out) {
  cmb_ctx* c = new cmb_ctx();
  c->cfg = *cfg;
  if (!c->cfg.batch_records) c->cfg.batch_records = 1 << 16;
  if (!c->cfg.batch_intervals) c->cfg.batch_intervals = c->cfg.batch_records * 2;
  if (c->cfg.n_staging < 2) c->cfg.n_staging = 2;
  const size_t nr = c->cfg.batch_records, ni = c->cfg.batch_intervals;
  for (uint32_t i = 0; i < c->cfg.n_staging; ++i) {
    c->slabs.emplace_back(4 * nr * 7 + 4 * (nr + 1) + 8 * ni + 4 * nr + 64);
    uint8_t* p = c->slabs.back().data();
    cmb_read_batch b{};
    b.capacity_records = (uint32_t)nr;
    b.capacity_intervals = (uint32_t)ni;
    auto take = [&](size_t bytes) { uint8_t* r = p; p += bytes; return r; };
    b.tid = (int32_t*)take(4 * nr); b.pos = (int32_t*)take(4 * nr); b.nm = (uint32_t*)take(4 * nr);
    b.l_seq = (uint32_t*)take(4 * nr); b.aligned = (uint32_t*)take(4 * nr); b.del = (uint32_t*)take(4 * nr);
    b.ins = (uint32_t*)take(4 * nr); b.iv_begin = (uint32_t*)take(4 * (nr + 1)); b.iv_start = (int32_t*)take(4 * ni);
    b.iv_len = (int32_t*)take(4 * ni); b.flag = (uint16_t*)take(2 * nr); b.mapq = (uint8_t*)take(nr); b.nm_state = (uint8_t*)take(nr);
    c->batches.push_back(b);
  }
  *out = c;
  return CMB_OK;
}
void cmb_destroy(cmb_ctx* c) { delete c; }

int cmb_set_reference(cmb_ctx* c, uint32_t n, const uint64_t* len, uint32_t b, uint32_t e) {
  c->lens.assign(len, len + n);
  c->gene_mode = false;
  c->tid_begin = b;
  c->tid_end = e;
  return CMB_OK;
}

int cmb_set_genes(cmb_ctx* c, uint32_t n_contigs, const uint64_t* contig_len, uint32_t n_genes, const cmb_gene* genes) {
  c->gene_mode = true;
  c->contig_lens.assign(contig_len, contig_len + n_contigs);
  c->genes.assign(genes, genes + n_genes);
  c->gene_first.assign((size_t)n_contigs + 1, 0);
  for (uint32_t g = 0; g < n_genes; ++g) c->gene_first[genes[g].tid + 1] += 1;
  for (uint32_t t = 0; t < n_contigs; ++t) c->gene_first[t + 1] += c->gene_first[t];
  c->lens.clear();
  for (uint32_t g = 0; g < n_genes; ++g) c->lens.push_back(genes[g].end - genes[g].start);
  if (c->lens.empty()) c->lens.push_back(1);
  c->tid_begin = 0;
  c->tid_end = (uint32_t)c->lens.size();
  return CMB_OK;
}
int cmb_fetch_gene_extras(cmb_ctx* c, uint8_t* contig_seen, uint64_t* n_kept_primary) {
  if (!c->contig_seen.empty()) memcpy(contig_seen, c->contig_seen.data(), c->contig_seen.size());
  *n_kept_primary = c->kept_primary;
  return CMB_OK;
}

int cmb_set_params(cmb_ctx* c, const cmb_params* p, cmb_filter_mode* mode_out) {
  c->p = *p;
  const bool si = p->min_aligned_length_single > 0 || p->min_percent_identity_single > 0.0f || p->min_aligned_percent_single > 0.0f;
  const bool pi = p->min_aligned_length_pair > 0 || p->min_percent_identity_pair > 0.0f || p->min_aligned_percent_pair > 0.0f;
  const bool fs = si || (!pi && p->min_mapq != 255);
  const bool fp = pi || ((!fs || !p->include_improper_pairs) && p->min_mapq != 255);
  c->mode.filter_single_reads = p->filtering ? fs : 0;
  c->mode.filter_pairs = p->filtering ? fp : 0;
  if (mode_out) *mode_out = c->mode;
  return CMB_OK;
}

int cmb_begin_sample(cmb_ctx* c) {
  c->arena.clear();
  c->rows.assign(c->lens.size(), cmb_contig_stats{});
  c->pairs.clear();
  c->last_kept_tid = INT64_MIN;
  c->excl_min = INT64_MAX;
  c->excl_max = INT64_MIN;
  c->contig_seen.assign(c->gene_mode ? c->contig_lens.size() : 0, 0);
  c->kept_primary = 0;
  c->error = 0;
  c->in_sample = true;
  c->ended = false;
  c->n_records = c->n_intervals = 0;
  c->n_acquired = 0;
  return CMB_OK;
}
int cmb_acquire_batch(cmb_ctx* c, cmb_read_batch* b) {
  if (c->n_acquired >= c->cfg.n_staging) return fail(c, CMB_E_ARG, "cmb_acquire_batch: every staging batch is already acquired");
  *b = c->batches[c->next];
  c->next = (c->next + 1) % c->cfg.n_staging;
  c->n_acquired += 1;
  return CMB_OK;
}

struct Rec { uint32_t flag, mapq, nm_state, nm, l_seq, aligned, del; };
static bool single_ok(const Rec& r, const cmb_params& p, bool* nm_err) {
  if (p.min_mapq != 255 && (r.mapq < p.min_mapq || r.mapq == 255)) return false;
  if (r.nm_state != 1) *nm_err = true;
  return r.aligned >= p.min_aligned_length_single && (float)r.aligned / (float)r.l_seq >= p.min_aligned_percent_single &&
         1.0f - (float)r.nm / (float)r.aligned >= p.min_percent_identity_single;
}
static bool pair_ok(const Rec& a, const Rec& b, const cmb_params& p, bool* nm_err) {
  if (p.min_mapq != 255 && (a.mapq < p.min_mapq || b.mapq < p.min_mapq || a.mapq == 255 || b.mapq == 255)) return false;
  if (a.nm_state != 1 || b.nm_state != 1) *nm_err = true;
  const uint32_t al = (a.aligned - a.del) + (b.aligned - b.del);
  return al >= p.min_aligned_length_pair && (float)al / (float)((uint64_t)a.l_seq + b.l_seq) >= p.min_aligned_percent_pair &&
         1.0f - ((float)((uint64_t)a.nm + b.nm) / (float)al) >= p.min_percent_identity_pair;
}

static int submit(cmb_ctx* c, const cmb_read_batch& b, uint32_t n, uint32_t ni, uint32_t excl_n = 0xffffffffu) {
  const cmb_params& p = c->p;
  auto rec = [&](uint32_t i) { return Rec{b.flag[i], b.mapq[i], b.nm_state[i], b.nm[i], b.l_seq[i], b.aligned[i], b.del[i]}; };
  for (uint32_t i = 0; i < n; ++i) {
    const Rec r = rec(i);
    const bool unmapped = r.flag & 4, sec = r.flag & 0x100, sup = r.flag & 0x800, proper = r.flag & 2;
    const bool flag_pass = !(sec && !p.include_secondary) && !(sup && !p.include_supplementary) && !(!proper && !p.include_improper_pairs);
    bool keep = flag_pass && !unmapped, nm_err = false;
    if (p.filtering) {
      bool passes;
      if (c->mode.filter_single_reads && !c->mode.filter_pairs) {
        passes = !unmapped && (p.include_supplementary || !sup) && (p.include_secondary || !sec) && single_ok(r, p, &nm_err);
      } else {
        const uint32_t m = i ^ 1u;
        bool ok = m < n;
        if (ok) {
          const Rec o = rec(m);
          const Rec& first = (i & 1) ? o : r;
          const Rec& second = (i & 1) ? r : o;
          if (c->mode.filter_single_reads) ok = single_ok(first, p, &nm_err) && single_ok(second, p, &nm_err);
          if (ok) ok = pair_ok(second, first, p, &nm_err);
        }
        passes = ok;
      }
      keep = keep && passes;
    }
    if (keep && r.nm_state != 1) nm_err = true;
    if (nm_err) c->error |= 2;
    if (!keep) continue;
    const int32_t tid = b.tid[i];
    if (tid < 0 || (size_t)tid >= (c->gene_mode ? c->contig_lens.size() : c->lens.size())) { c->error |= 4; continue; }
    if (tid < c->last_kept_tid) c->error |= 1;
    c->last_kept_tid = std::max<int64_t>(c->last_kept_tid, tid);
    if (i < excl_n) {
      c->excl_min = std::min<int64_t>(c->excl_min, tid);
      c->excl_max = std::max<int64_t>(c->excl_max, tid);
    }
    if (c->gene_mode) {
      // genes.rs:254-303 + 467-552 restated per gene: the gene's array is the contig's cut to [start, end) with the running
      // depth at `start` in front, i.e. every aligned block clipped to the gene; reads count for the genes holding their pos
      const bool primary = !sec && !sup;
      c->contig_seen[tid] = 1;
      c->kept_primary += primary;
      const uint64_t CL = c->contig_lens[tid];
      bool bad = false;
      for (uint32_t k = b.iv_begin[i]; k < b.iv_begin[i + 1]; ++k) {
        const int32_t s = b.iv_start[k];
        if (s == CMB_IV_PAD) continue;
        if (s < 0 || (uint64_t)s >= CL) { c->error |= 4; bad = true; }
      }
      if (bad) continue;
      const uint64_t indels = (uint64_t)b.ins[i] + r.del;
      for (uint32_t g = c->gene_first[tid]; g < c->gene_first[tid + 1]; ++g) {
        const cmb_gene& ge = c->genes[g];
        cmb_contig_stats& row = c->rows[g];
        if ((uint32_t)b.pos[i] >= ge.start && (uint32_t)b.pos[i] < ge.end) {
          row.n_records += 1;
          row.n_primary += primary;
          row.sum_edit += r.nm >= indels ? r.nm - indels : 0;
          if (primary && r.aligned > 0) row.sum_identity_primary += ((double)r.aligned - (double)r.nm) / (double)r.aligned;
        }
        for (uint32_t k = b.iv_begin[i]; k < b.iv_begin[i + 1]; ++k) {
          const int32_t s = b.iv_start[k];
          if (s == CMB_IV_PAD) continue;
          const uint64_t e = (uint64_t)s + (uint32_t)b.iv_len[k];
          if (e <= ge.start || (uint32_t)s >= ge.end) continue;
          auto& ud = c->arena[g];
          const uint64_t L = ge.end - ge.start;
          if (ud.empty()) ud.assign(L + 1, 0);
          ud[std::max<uint32_t>((uint32_t)s, ge.start) - ge.start] += 1;
          if (e - ge.start < L) ud[e - ge.start] -= 1;
        }
      }
      continue;
    }
    if ((uint32_t)tid < c->tid_begin || (uint32_t)tid >= c->tid_end) continue;
    cmb_contig_stats& row = c->rows[tid];
    const bool primary = !sec && !sup;
    row.n_records += 1;
    row.n_primary += primary;
    row.n_nonsupp += !sup;
    row.sum_edit += r.nm;
    row.sum_indel += (uint64_t)b.ins[i] + r.del;
    if (r.aligned > 0) {
      const double id = ((double)r.aligned - (double)r.nm) / (double)r.aligned;
      if (primary) row.sum_identity_primary += id;
      if (!sup) row.sum_identity_nonsupp += id;
    }
    auto& ud = c->arena[(uint32_t)tid];
    if (ud.empty()) ud.assign(c->lens[tid] + 1, 0);  // +1 sentinel so a zero-length contig is still "allocated"
    const uint64_t L = c->lens[tid];
    for (uint32_t k = b.iv_begin[i]; k < b.iv_begin[i + 1]; ++k) {
      const int32_t s = b.iv_start[k];
      if (s == CMB_IV_PAD) continue;
      if (s < 0 || (uint64_t)s >= L) { c->error |= 4; continue; }
      ud[s] += 1;
      const uint64_t e = (uint64_t)s + (uint32_t)b.iv_len[k];
      if (e < L) ud[e] -= 1;
    }
  }
  c->n_records += n;
  c->n_intervals += ni;
  return CMB_OK;
}

int cmb_submit_batch(cmb_ctx* c, uint32_t n, uint32_t ni) {
  if (c->n_acquired == 0) return fail(c, CMB_E_ARG, "cmb_submit_batch: no acquired batch");
  const uint32_t i = (c->next + c->cfg.n_staging - c->n_acquired) % c->cfg.n_staging;
  c->n_acquired -= 1;
  return submit(c, c->batches[i], n, ni);
}
int cmb_submit_device_batch(cmb_ctx* c, const cmb_read_batch* b, uint32_t n, uint32_t ni) { return submit(c, *b, n, ni); }
void* cmb_host_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void cmb_host_free(void* p) { free(p); }
// cmb_submit_bgzf.  By default the emulator declines, so that the CPU tests exercise the host decode pipeline; with
// CMB_EMU_BGZF=1 it plays the device-side decoder with zlib and a plain record walk (BAM spec, SAMv1 section 4.2), which
// lets the host's device-decode branch (block table, records_at, counters, fallback on CMB_E_DECLINED) run without a GPU.
int cmb_last_bgzf_batch(cmb_ctx* c, cmb_read_batch*, uint32_t*, uint32_t*) { return fail(c, CMB_E_ARG, "emulator: no device-resident tuples"); }

int cmb_submit_bgzf(cmb_ctx* c, const cmb_bgzf_input* in, cmb_bgzf_result* out) {
  if (!c || !in || !out) return CMB_E_ARG;
  if (!getenv("CMB_EMU_BGZF")) {
    c->err = "device emulator: no device-side decode";
    return CMB_E_DECLINED;
  }
  *out = cmb_bgzf_result{};
  // like the device: only the blocks of the range (plus a tail for its last straddling record) are inflated; `stream`
  // is indexed with absolute uncompressed offsets through `base`
  std::vector<uint64_t> ustart((size_t)in->n_blocks + 1, 0);
  for (uint32_t b = 0; b < in->n_blocks; ++b) ustart[b + 1] = ustart[b] + in->block_isize[b];
  uint32_t b_lo = 0, walk_end = in->n_blocks, data_end = in->n_blocks;
  if (in->ranged) {
    b_lo = in->walk_begin_block;
    walk_end = in->walk_end_block;
    if (walk_end <= b_lo) return CMB_OK;
    data_end = walk_end;
    uint64_t tail = 0;
    while (data_end < in->n_blocks && tail < (4u << 20)) tail += in->block_isize[data_end++];
  }
  const uint64_t base = ustart[b_lo];
  std::vector<uint8_t> stream;
  for (uint32_t b = b_lo; b < data_end; ++b) {
    const uint32_t isz = in->block_isize[b];
    const size_t at = stream.size();
    stream.resize(at + isz);
    if (!isz) continue;
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, -15) != Z_OK) return fail(c, CMB_E_NOMEM, "zlib");
    zs.next_in = const_cast<Bytef*>(in->data + in->block_coffset[b]);
    zs.avail_in = in->block_clen[b];
    zs.next_out = stream.data() + at;
    zs.avail_out = isz;
    const int zr = inflate(&zs, Z_FINISH);
    inflateEnd(&zs);
    uint32_t want;
    memcpy(&want, in->data + in->block_coffset[b] + in->block_clen[b], 4);
    if (zr != Z_STREAM_END || zs.avail_out != 0 || (uint32_t)crc32(0, stream.data() + at, isz) != want)
      return fail(c, CMB_E_DECLINED, "emulator: BGZF block does not inflate");
  }
  auto u32 = [&](size_t o) { uint32_t v; memcpy(&v, stream.data() + o, 4); return v; };
  auto u16 = [&](size_t o) { uint16_t v; memcpy(&v, stream.data() + o, 2); return (uint32_t)v; };
  std::vector<int32_t> tid, pos, ivs, ivl;
  std::vector<uint16_t> flag;
  std::vector<uint8_t> mapq, nm_state;
  std::vector<uint32_t> nm, l_seq, aligned, del, ins, iv_begin;
  std::vector<int32_t> mtid;
  std::vector<std::string> qname;
  size_t o = in->records_at - base;
  const size_t walk_stop = (size_t)(ustart[walk_end] - base);  // records starting at or after this belong to the next range
  size_t excl_n = (size_t)-1;
  uint64_t n_owned = 0;
  while (o < stream.size() && o < walk_stop) {
    if (in->ranged && excl_n == (size_t)-1 && in->excl_end_block < walk_end && o >= (size_t)(ustart[in->excl_end_block] - base)) excl_n = tid.size();
    if (o + 36 > stream.size()) return fail(c, CMB_E_DECLINED, "emulator: record cut short");
    const uint32_t bs = u32(o);
    if (bs < 32 || o + 4 + (size_t)bs > stream.size()) return fail(c, CMB_E_DECLINED, "emulator: record cut short");
    const size_t r = o + 4, end = r + bs;
    const uint32_t l_name = stream[r + 8], n_cig = u16(r + 12), ls = u32(r + 16);
    tid.push_back((int32_t)u32(r));
    pos.push_back((int32_t)u32(r + 4));
    mapq.push_back(stream[r + 9]);
    flag.push_back((uint16_t)u16(r + 14));
    l_seq.push_back(ls);
    mtid.push_back((int32_t)u32(r + 20));
    qname.emplace_back((const char*)stream.data() + r + 32, l_name ? l_name - 1 : 0);
    const bool owned = !in->ranged || (tid.back() < 0 ? in->own_unplaced != 0 : (tid.back() >= in->own_tid_begin && tid.back() < in->own_tid_end));
    n_owned += owned;
    if (owned && !(flag.back() & 0x900)) out->n_primary += 1;
    size_t cg = r + 32 + l_name;
    size_t aux = cg + 4ull * n_cig + (ls + 1) / 2 + ls;
    if (aux > end) return fail(c, CMB_E_DECLINED, "emulator: malformed record");
    iv_begin.push_back((uint32_t)ivs.size());
    uint32_t al = 0, dl = 0, in_ = 0;
    int64_t cur = pos.back();
    for (uint32_t k = 0; k < n_cig; ++k) {
      const uint32_t v = u32(cg + 4 * k), op = v & 15, len = v >> 4;
      if (op == 0 || op == 7 || op == 8) {
        ivs.push_back(cur < 0 ? -1 : (int32_t)std::min<int64_t>(cur, INT32_MAX));
        ivl.push_back((int32_t)len);
        cur += len;
        al += len;
      } else if (op == 2) { cur += len; dl += len; al += len; }
      else if (op == 3) cur += len;
      else if (op == 1) { in_ += len; al += len; }
    }
    aligned.push_back(al);
    del.push_back(dl);
    ins.push_back(in_);
    uint8_t st = 0;
    uint32_t nmv = 0;
    while (aux + 3 <= end) {
      const uint8_t t0 = stream[aux], t1 = stream[aux + 1], ty = stream[aux + 2];
      aux += 3;
      size_t sz;
      if (ty == 'A' || ty == 'c' || ty == 'C') sz = 1;
      else if (ty == 's' || ty == 'S') sz = 2;
      else if (ty == 'i' || ty == 'I' || ty == 'f') sz = 4;
      else if (ty == 'Z' || ty == 'H') { size_t e = aux; while (e < end && stream[e]) ++e; sz = e < end ? e - aux + 1 : end - aux; }
      else if (ty == 'B') {
        if (aux + 5 > end) sz = end - aux;
        else { const uint8_t sub = stream[aux]; const uint32_t cnt = u32(aux + 1); sz = 5 + (size_t)cnt * ((sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4); }
      } else return fail(c, CMB_E_DECLINED, "emulator: unknown aux type");
      if (t0 == 'N' && t1 == 'M' && st == 0) {
        if (ty == 'C') { st = 1; nmv = stream[aux]; }
        else if (ty == 'S') { st = 1; nmv = u16(aux); }
        else if (ty == 'I') { st = 1; nmv = u32(aux); }
        else st = 2;
      }
      // like kd_extract: a CG:B,I tag behind a `<l_seq>S...` placeholder (long CIGAR) goes to the host decoder
      if (t0 == 'C' && t1 == 'G' && ty == 'B' && n_cig && (u32(cg) & 15) == 4 && (u32(cg) >> 4) == ls && tid.back() >= 0 && pos.back() >= 0)
        return fail(c, CMB_E_DECLINED, "emulator: long CIGAR in a CG tag");
      aux += sz;
    }
    nm_state.push_back(st);
    nm.push_back(nmv);
    o = end;
  }
  iv_begin.push_back((uint32_t)ivs.size());
  out->n_records = n_owned;
  out->n_intervals = ivs.size();
  out->h2d_bytes = in->size;
  if (tid.empty()) return CMB_OK;
  if (ivs.empty()) { ivs.push_back(0); ivl.push_back(0); }
  if (c->mode.filter_pairs) {
    // the pair path of ReferenceSortedBamFilter::read (filter.rs:117-233) as the reference runs it -- a map of stored first
    // mates, cleared when the reference id changes -- then only completed pairs, stored mate first, in the host layout
    // K1's emulation expects (first at the even index).  The CUDA library matches mates with a hash table instead
    // (cmb_pairs.cuh) and leaves the records in file order; both must give the oracle's table.
    std::map<std::string, size_t> first_set;
    int32_t current = -1;
    std::vector<size_t> order;
    for (size_t i = 0; i < tid.size(); ++i) {
      if ((flag[i] & 0x900) || !(flag[i] & 0x2)) continue;
      if (tid[i] != current) {
        current = tid[i];
        first_set.clear();
      }
      auto it = first_set.find(qname[i]);
      if (it == first_set.end()) {
        if (mtid[i] == current) first_set.emplace(qname[i], i);
      } else {
        order.push_back(it->second);
        order.push_back(i);
        first_set.erase(it);
      }
    }
    auto pick = [&](auto& v) {
      auto w = v;
      w.clear();
      for (size_t i : order) w.push_back(v[i]);
      v.swap(w);
    };
    std::vector<int32_t> ivs2, ivl2;
    std::vector<uint32_t> ivb2;
    for (size_t i : order) {
      ivb2.push_back((uint32_t)ivs2.size());
      for (uint32_t k = iv_begin[i]; k < iv_begin[i + 1]; ++k) {
        ivs2.push_back(ivs[k]);
        ivl2.push_back(ivl[k]);
      }
    }
    ivb2.push_back((uint32_t)ivs2.size());
    pick(tid); pick(pos); pick(flag); pick(mapq); pick(nm_state); pick(nm); pick(l_seq); pick(aligned); pick(del); pick(ins);
    ivs.swap(ivs2); ivl.swap(ivl2); iv_begin.swap(ivb2);
    if (excl_n != (size_t)-1) {  // pairs are ordered by their second mate: those completed inside the exclusive share form a prefix
      size_t k = 0;
      while (k + 1 < order.size() && order[k + 1] < excl_n) k += 2;
      excl_n = k;
    }
    if (tid.empty()) return CMB_OK;
    if (ivs.empty()) { ivs.push_back(0); ivl.push_back(0); }
  }
  cmb_read_batch b{};
  b.tid = tid.data(); b.pos = pos.data(); b.flag = flag.data(); b.mapq = mapq.data(); b.nm_state = nm_state.data(); b.nm = nm.data();
  b.l_seq = l_seq.data(); b.aligned = aligned.data(); b.del = del.data(); b.ins = ins.data(); b.iv_begin = iv_begin.data();
  b.iv_start = ivs.data(); b.iv_len = ivl.data();
  return submit(c, b, (uint32_t)tid.size(), (uint32_t)out->n_intervals, excl_n == (size_t)-1 ? 0xffffffffu : (uint32_t)excl_n);
}

int cmb_grow_buffers(cmb_ctx*) { return CMB_OK; }
// `coverm filter` on the device is not emulated: the host's own filter loop (filter_command.hpp) is what the CPU tests check.
int cmb_decode_bgzf(cmb_ctx* c, const cmb_bgzf_input*, cmb_bgzf_result*) { return fail(c, CMB_E_DECLINED, "emulator: no device-side filter"); }
int cmb_filter_plan(cmb_ctx* c, int, uint64_t*, uint64_t*) { return fail(c, CMB_E_ARG, "emulator: no device-side filter"); }
int cmb_filter_fetch(cmb_ctx* c, uint8_t*, uint64_t) { return fail(c, CMB_E_ARG, "emulator: no device-side filter"); }

// No NCCL in the emulator: groups of emulated ranks exchange through the host all-gather callback of the session.
int cmb_comm_unique_id(uint8_t*) { return fail(nullptr, CMB_E_ARG, "emulator: no NCCL"); }
int cmb_comm_init(cmb_ctx* c, const uint8_t*, int, int) { return fail(c, CMB_E_ARG, "emulator: no NCCL"); }
int cmb_comm_init_local(cmb_ctx* const*, int) { return fail(nullptr, CMB_E_ARG, "emulator: no NCCL"); }
void cmb_comm_destroy(cmb_ctx*) {}
int cmb_comm_allgather(cmb_ctx* c, const void*, void*, size_t) { return fail(c, CMB_E_ARG, "emulator: no NCCL"); }
int cmb_allgather_stats(cmb_ctx* c, const uint32_t*, const uint64_t*, cmb_contig_stats*, cmb_hist_pair*) { return fail(c, CMB_E_ARG, "emulator: no NCCL"); }
int cmb_kept_tid_range(cmb_ctx* c, int32_t* lo, int32_t* hi) {
  *lo = c->excl_min == INT64_MAX ? INT32_MAX : (int32_t)c->excl_min;
  *hi = c->excl_max == INT64_MIN ? INT32_MIN : (int32_t)c->excl_max;
  return CMB_OK;
}

int cmb_end_sample_device(cmb_ctx* c, const cmb_contig_stats** out) {
  c->in_sample = false;
  const uint64_t E = c->p.contig_end_exclusion;
  const bool hist = c->p.want & (CMB_WANT_HIST | CMB_WANT_HIST_CSR), csr = c->p.want & CMB_WANT_HIST_CSR;
  if (c->gene_mode)  // every gene of a seen contig is scanned, covered or not (the device scans the whole arena)
    for (size_t t = 0; t < c->contig_seen.size(); ++t)
      if (c->contig_seen[t])
        for (uint32_t g = c->gene_first[t]; g < c->gene_first[t + 1]; ++g) {
          auto& ud = c->arena[g];
          if (ud.empty()) ud.assign(c->lens[g] + 1, 0);
        }
  for (auto& kv : c->arena) {
    const uint32_t tid = kv.first;
    const uint64_t L = c->lens[tid];
    cmb_contig_stats& row = c->rows[tid];
    std::map<uint32_t, uint64_t> h;
    int64_t d = 0;
    const bool win = 2 * E < L;
    for (uint64_t i = 0; i < L; ++i) {
      d += kv.second[i];
      if (d > 0) row.covered_full += 1;
      if (win && i >= E && i < L - E) {
        if (d > 0) row.covered_window += 1;
        row.sum_depth_window += (uint64_t)d;
        if (hist) h[(uint32_t)d] += 1;
      }
    }
    if (!hist || !win || (row.n_records == 0 && !c->gene_mode)) continue;  // a gene may be covered by reads that start before it
    const uint64_t T = L - 2 * E;
    const uint64_t min_index = (uint64_t)std::floor(c->p.trim_min * (float)T), max_index = (uint64_t)std::ceil(c->p.trim_max * (float)T);
    uint64_t cprev = 0, total = 0, s0 = 0, s1 = 0, s2 = 0, k = h.begin()->first;
    for (auto& dc : h) {
      const uint64_t depth = dc.first, cnt = dc.second, ccur = cprev + cnt;
      uint64_t w;
      if (ccur < min_index) w = 0;
      else if (cprev < min_index) w = ccur > max_index ? max_index - min_index + 1 : ccur - min_index + 1;
      else w = cprev > max_index ? 0 : (ccur > max_index ? max_index - cprev + 1 : cnt);
      total += w * depth;
      s0 += cnt; s1 += depth * cnt; s2 += depth * depth * cnt;
      cprev = ccur;
    }
    row.trimmed_total = total;
    row.trim_min_index = min_index;
    row.trim_max_index = max_index;
    row.var_k = k;
    row.var_ex = s1 - k * s0;
    row.var_ex2 = s2 - 2 * k * s1 + k * k * s0;
    row.hist_count = (uint32_t)h.size();
    if (csr) {
      row.hist_offset = c->pairs.size();
      for (auto& dc : h) c->pairs.push_back({dc.first, (uint32_t)dc.second});
    }
  }
  c->ended = true;
  if (c->error & 1) return fail(c, CMB_E_UNSORTED, "BAM file appears to be unsorted. Input BAM files must be sorted by reference (i.e. by samtools sort)");
  if (c->error & 2) return fail(c, CMB_E_NM, "Mapping record encountered that does not have an 'NM' auxiliary tag in the SAM/BAM format. This is required to work out some coverage statistics");
  if (c->error & 4) return fail(c, CMB_E_BOUNDS, "index out of bounds: an aligned block starts beyond the end of its reference sequence");
  if (out) *out = c->rows.data();
  return CMB_OK;
}

int cmb_end_sample(cmb_ctx* c, cmb_contig_stats* stats, cmb_hist_pair* pairs, uint64_t cap, uint64_t* n_pairs) {
  int rc = cmb_end_sample_device(c, nullptr);
  if (rc) return rc;
  if (stats) memcpy(stats, c->rows.data(), sizeof(cmb_contig_stats) * c->rows.size());
  if (pairs && cap >= c->pairs.size()) memcpy(pairs, c->pairs.data(), sizeof(cmb_hist_pair) * c->pairs.size());
  if (n_pairs) *n_pairs = c->pairs.size();
  return CMB_OK;
}
int cmb_fetch_pairs(cmb_ctx* c, cmb_hist_pair* pairs, uint64_t n) {
  memcpy(pairs, c->pairs.data(), sizeof(cmb_hist_pair) * std::min<uint64_t>(n, c->pairs.size()));
  return CMB_OK;
}
int cmb_get_timing(const cmb_ctx* c, cmb_sample_timing* t) {
  *t = cmb_sample_timing{};
  t->n_records = c->n_records;
  t->n_intervals = c->n_intervals;
  return CMB_OK;
}
void* cmb_stream(cmb_ctx*) { return nullptr; }
void cmb_nvtx_push(const char*) {}
void cmb_nvtx_pop(void) {}

}  // extern "C"
