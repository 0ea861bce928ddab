// Host-side check of the multi-GPU block-range search (coverm_b200/csrc/host/shard_range.hpp) on one BAM file:
//   * brute force: the whole file is inflated and walked record by record, which gives, for every BGZF block, the tid of the
//     first record that starts in it or after it; the finder's walk_end / excl_end / walk_begin must agree with that;
//   * every block's speculative "first record that starts here" must be the true one (a wrong guess would misdirect the search);
//   * a hint (where a tid's records are expected to start) must never change a result: every rank's range is computed without a
//     hint, with the caller's cumulative-length hint, and with hints that are plain wrong (0, 1, mirrored, pseudo-random).
// usage: shard_range_check file.bam [hint_min_blocks]  -> prints "ok <ranges checked> <probes without hints> <probes with length hints>"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <stdexcept>
#include <string>
#include <vector>

#define private public  // the check also calls BlockRangeFinder::first_record_in directly
#include "host/shard_range.hpp"
#undef private

using namespace cmbh;

static uint32_t u32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  std::ifstream f(argv[1], std::ios::binary);
  std::vector<uint8_t> file((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  BlockIndex bx;
  bx.build(file.data(), file.size());
  if (!bx.bgzf) { fprintf(stderr, "not BGZF\n"); return 2; }
  const uint32_t nb = (uint32_t)bx.blocks.size();
  std::vector<uint8_t> all((size_t)bx.ustart[nb] + 8);
  BgzfInflater inf;
  bx.inflate(0, nb, all.data(), inf);
  const size_t total = (size_t)bx.ustart[nb];
  // BAM header
  if (total < 12 || memcmp(all.data(), "BAM\1", 4)) { fprintf(stderr, "not BAM\n"); return 2; }
  size_t o = 8 + u32(all.data() + 4);
  const uint32_t n_ref = u32(all.data() + o);
  o += 4;
  std::vector<uint64_t> lens(n_ref);
  for (uint32_t r = 0; r < n_ref; ++r) {
    const uint32_t l_name = u32(all.data() + o);
    lens[r] = u32(all.data() + o + 4 + l_name);
    o += 8 + l_name;
  }
  const uint64_t records_at = o;
  // brute force: first_tid_from[b] = key of the first record starting at or after block b (+inf past the end)
  auto key_of = [](int32_t tid) -> int64_t { return tid < 0 ? (int64_t)INT32_MAX + 1 : (int64_t)tid; };
  std::vector<int64_t> first_key(nb + 1, INT64_MAX);
  std::vector<uint64_t> first_off(nb + 1, UINT64_MAX);
  {
    std::vector<int64_t> own(nb, INT64_MAX);
    std::vector<uint64_t> own_off(nb, UINT64_MAX);
    size_t pos = (size_t)records_at;
    uint32_t b = 0;
    while (pos + 36 <= total) {
      while (b + 1 < nb && bx.ustart[b + 1] <= pos) ++b;
      if (own[b] == INT64_MAX) {
        own[b] = key_of((int32_t)u32(all.data() + pos + 4));
        own_off[b] = pos;
      }
      pos += 4 + (size_t)u32(all.data() + pos);
    }
    for (uint32_t k = nb; k-- > 0;) {
      first_key[k] = own[k] != INT64_MAX ? own[k] : first_key[k + 1];
      first_off[k] = own_off[k];
    }
  }
  const uint32_t first_block = (uint32_t)(std::upper_bound(bx.ustart.begin(), bx.ustart.end(), records_at) - bx.ustart.begin()) - 1;
  {
    BlockRangeFinder fd(bx, n_ref, records_at);
    for (uint32_t k = first_block; k < nb; ++k) {
      int32_t tid = 0;
      uint64_t uoff = 0;
      const bool got = fd.first_record_in(k, &tid, &uoff);
      const bool want = first_off[k] != UINT64_MAX;
      if (got != want || (got && (uoff != first_off[k] || key_of(tid) != first_key[k]))) {
        fprintf(stderr, "block %u: guessed %d tid %d @%llu, true %d @%llu\n", k, (int)got, tid, (unsigned long long)uoff, (int)want,
                (unsigned long long)first_off[k]);
        return 1;
      }
    }
  }
  auto lower_bound_block = [&](uint32_t t) {
    uint32_t b = first_block;
    while (b < nb && first_key[b] < (int64_t)t) ++b;
    return b;
  };
  const uint32_t hint_min_blocks = argc > 2 ? (uint32_t)atoi(argv[2]) : 1024u;
  unsigned long checked = 0, probes_plain = 0, probes_hint = 0;
  uint64_t rng = 88172645463325252ull;
  auto rnd = [&]() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return (double)(rng >> 11) / 9007199254740992.0; };
  long double total_len = 0;
  for (uint64_t l : lens) total_len += (long double)l;
  for (int n_ranks : {2, 3, 4, 7, 8}) {
    const std::vector<uint32_t> cut = tid_cuts_by_length(lens, n_ranks);
    for (int r = 0; r < n_ranks; ++r) {
      const uint32_t lo = cut[r], hi = cut[r + 1];
      const bool first = r == 0, last = r == n_ranks - 1;
      BlockRange base;
      {
        BlockRangeFinder fd(bx, n_ref, records_at);
        base = fd.find(lo, hi, first, last);
        probes_plain += base.probes;
      }
      // against the brute force
      if (!last) {
        const uint32_t want_end = lower_bound_block(hi);
        if (base.walk_end != std::max(want_end, base.walk_begin)) {
          fprintf(stderr, "n=%d rank=%d: walk_end %u, brute force %u\n", n_ranks, r, base.walk_end, want_end);
          return 1;
        }
      } else if (base.walk_end != nb) {
        fprintf(stderr, "last rank must run to the end\n");
        return 1;
      }
      if (!(first || lo == 0)) {
        const uint32_t beta = lower_bound_block(lo);
        uint32_t b = beta > first_block ? beta - 1 : first_block;
        while (b > first_block && first_off[b] == UINT64_MAX) --b;
        if (first_off[b] != UINT64_MAX && (base.walk_begin != b || base.records_at != first_off[b])) {
          fprintf(stderr, "n=%d rank=%d: walk_begin %u @%llu, brute force %u @%llu\n", n_ranks, r, base.walk_begin,
                  (unsigned long long)base.records_at, b, (unsigned long long)first_off[b]);
          return 1;
        }
      }
      // hints must not change anything
      long double before_lo = 0, before_hi = 0, acc = 0;
      for (uint32_t t = 0; t < n_ref; ++t) {
        if (t == lo) before_lo = acc;
        if (t == hi) before_hi = acc;
        acc += (long double)lens[t];
      }
      if (hi >= n_ref) before_hi = acc;
      const double f_lo = total_len > 0 ? (double)(before_lo / total_len) : 0.0, f_hi = total_len > 0 ? (double)(before_hi / total_len) : 0.0;
      const double hints[][2] = {{f_lo, f_hi}, {0.0, 0.0}, {1.0, 1.0}, {1.0 - f_lo, 1.0 - f_hi}, {rnd(), rnd()}, {rnd(), rnd()}, {f_hi, f_lo}};
      for (size_t h = 0; h < sizeof hints / sizeof hints[0]; ++h) {
        BlockRangeFinder fd(bx, n_ref, records_at, hint_min_blocks);
        const BlockRange got = fd.find(lo, hi, first, last, hints[h][0], hints[h][1]);
        if (h == 0) probes_hint += got.probes;
        if (got.walk_begin != base.walk_begin || got.walk_end != base.walk_end || got.records_at != base.records_at || got.excl_end != base.excl_end) {
          fprintf(stderr, "n=%d rank=%d hint %zu (%.3f, %.3f): [%u,%u) @%llu excl %u, without hint [%u,%u) @%llu excl %u\n", n_ranks, r, h, hints[h][0],
                  hints[h][1], got.walk_begin, got.walk_end, (unsigned long long)got.records_at, got.excl_end, base.walk_begin, base.walk_end,
                  (unsigned long long)base.records_at, base.excl_end);
          return 1;
        }
        ++checked;
      }
    }
  }
  printf("ok %lu %lu %lu\n", checked, probes_plain, probes_hint);
  return 0;
}
