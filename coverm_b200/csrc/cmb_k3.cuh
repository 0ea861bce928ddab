// K3: per-contig histogram finalisation, one WARP per contig.
//
// K2 left, per (chunk, contig slot), a short list of (depth,count) records (plus, rarely, records on the chunk's
// overflow list).  A contig gathers the lists of the chunks it overlaps into a per-warp shared-memory window of depth
// bins [dmin, dmin+512) (repeated for deeper windows if ever needed), then walks the bins in order with warp scans:
//   * trimmed-mean `total` exactly as the reference's ascending walk (EST:598-642),
//   * S0 = sum n, S1 = sum x n, S2 = sum x^2 n (wrapping u64) and k = lowest depth -> variance sums (EST:790-805),
//   * optionally the merged (depth,count) pairs (CSR) for the host-side per-genome merge / coverage_histogram.
#pragma once

constexpr uint32_t K3_WARPS = 8;
constexpr uint32_t K3_THREADS = K3_WARPS * 32;
constexpr uint32_t K3_WINDOW = 512;  // depth bins per warp window

struct K3Args {
  const uint32_t* off_span;
  const uint32_t* len;
  const uint32_t* chunk_first;
  cmb_contig_stats* rows;
  uint32_t tid_begin, n_local, excl;
  float trim_min, trim_max;
  const uint2* rec;
  const uint2* warp_table;
  const uint4* ovf;
  const uint32_t* ovf_head;
  uint32_t ovf_capacity;
  cmb_hist_pair* pairs;
  unsigned long long* pair_count;
  uint64_t pair_capacity;
  uint32_t want_csr;
  uint32_t all_rows;  // gene mode: a gene is covered by reads that start before it, so "no record counted here" does not mean "empty"
  uint32_t* error_flags;
};

__global__ void __launch_bounds__(K3_THREADS) k3_finalize(const K3Args a) {
  __shared__ uint32_t whist_all[K3_WARPS][K3_WINDOW];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t* whist = whist_all[warp];
  const uint32_t lc = blockIdx.x * K3_WARPS + warp;
  if (lc >= a.n_local) return;
  cmb_contig_stats* row = a.rows + a.tid_begin + lc;
  if (row->n_records == 0 && !a.all_rows) return;  // unseen contig: the host never consults its histogram
  const uint32_t L = a.len[lc];
  const uint64_t E = a.excl;
  if (!(2 * E < L)) return;  // no window (EST:436-445)
  const uint64_t T = (uint64_t)L - 2 * E;
  const float Tf = __ull2float_rn(T);  // EST:591-592: f32 products, `as usize` saturating casts
  const uint64_t min_index = (uint64_t)floorf(__fmul_rn(a.trim_min, Tf));
  const uint64_t max_index = (uint64_t)ceilf(__fmul_rn(a.trim_max, Tf));
  const uint32_t k0 = a.off_span[lc] / CHUNK_SPANS, k1 = (a.off_span[lc + 1] - 1) / CHUNK_SPANS;
  const uint32_t n_chunk = k1 - k0 + 1;

  // Visits every record of this contig: the (chunk, slot) lists one after the other with the records lane-parallel,
  // then the chunks' overflow lists (one lane per chunk).
  auto for_each_record = [&](auto&& fn) {
    for (uint32_t i = 0; i < n_chunk; ++i) {
      const uint32_t k = k0 + i;
      const uint32_t slot = lc - a.chunk_first[k];
      if (slot >= HIST_SLOTS) continue;
      const uint2 ent = a.warp_table[(uint64_t)k * HIST_SLOTS + slot];
      for (uint32_t r = lane; r < ent.y; r += 32) {
        const uint2 rc = a.rec[ent.x + r];
        fn(rc.x, rc.y);
      }
    }
    for (uint32_t i = lane; i < n_chunk; i += 32) {
      for (uint32_t o = a.ovf_head[k0 + i]; o != OVF_NIL && o < a.ovf_capacity;) {
        const uint4 rc = a.ovf[o];
        if (rc.x == lc) fn(rc.y, rc.z);
        o = rc.w;
      }
    }
  };

  // ---- depth range of this contig's records
  uint32_t dmin = 0xffffffffu, dmax = 0;
  for_each_record([&](uint32_t depth, uint32_t) {
    dmin = min(dmin, depth);
    dmax = max(dmax, depth);
  });
  dmin = __reduce_min_sync(FULL, dmin);
  dmax = __reduce_max_sync(FULL, dmax);
  if (dmin > dmax) return;  // no records (cannot happen for a contig with a window)

  unsigned long long ltot = 0, l0 = 0, l1 = 0, l2 = 0;  // per-lane partial sums
  uint32_t n_pairs = 0;
  unsigned long long pair_base = 0;
  const int n_rounds = a.want_csr ? 2 : 1;  // round 0: statistics (+ count the pairs); round 1: write the pairs
  for (int round = 0; round < n_rounds; ++round) {
    unsigned long long cum = 0;  // counts below the current window
    uint32_t written = 0;
    for (uint32_t wbase = dmin;; wbase += K3_WINDOW) {
      const uint32_t nb = min(K3_WINDOW, dmax - wbase + 1);
      for (uint32_t b = lane; b < nb; b += 32) whist[b] = 0;
      __syncwarp();
      for_each_record([&](uint32_t depth, uint32_t cnt) {
        const uint32_t b = depth - wbase;
        if (depth >= wbase && b < nb) atomicAdd(&whist[b], cnt);
      });
      __syncwarp();
      for (uint32_t b0 = 0; b0 < nb; b0 += 32) {
        const uint32_t b = b0 + lane;
        const uint32_t n = b < nb ? whist[b] : 0u;
        uint32_t incl = n;  // inclusive scan of the 32 bins (a contig holds < 2^31 bases: fits u32)
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t o = __shfl_up_sync(FULL, incl, d);
          if ((int)lane >= d) incl += o;
        }
        const uint32_t nzmask = __ballot_sync(FULL, n != 0);
        if (round == 0) {
          if (n) {
            const unsigned long long depth = (unsigned long long)wbase + b;
            const unsigned long long cprev = cum + (incl - n), ccur = cum + incl;
            unsigned long long w;
            if (ccur < min_index) w = 0;
            else if (cprev < min_index) w = ccur > max_index ? max_index - min_index + 1 : ccur - min_index + 1;
            else w = cprev > max_index ? 0 : (ccur > max_index ? max_index - cprev + 1 : (unsigned long long)n);
            ltot += w * depth;
            l0 += n;
            l1 += depth * n;
            l2 += depth * depth * n;
          }
          n_pairs += __popc(nzmask);
        } else if (n) {
          const unsigned long long idx = pair_base + written + __popc(nzmask & ((1u << lane) - 1));
          if (idx < a.pair_capacity) {
            cmb_hist_pair pr;
            pr.depth = wbase + b;
            pr.count = n;
            a.pairs[idx] = pr;
          }
        }
        written += __popc(nzmask);
        cum += __shfl_sync(FULL, incl, 31);
      }
      if (dmax - wbase < K3_WINDOW) break;
      __syncwarp();
    }
    if (round == 0) {
      const unsigned long long total = warp_sum_u64(ltot), S0 = warp_sum_u64(l0), S1 = warp_sum_u64(l1), S2 = warp_sum_u64(l2);
      if (lane == 0) {
        const unsigned long long k = dmin;  // lowest depth with a non-zero count
        row->trimmed_total = total;
        row->trim_min_index = min_index;
        row->trim_max_index = max_index;
        row->var_k = k;
        row->var_ex = S1 - k * S0;                    // sum (x-k) n    (mod 2^64, as the reference's usize)
        row->var_ex2 = S2 - 2 * k * S1 + k * k * S0;  // sum (x-k)^2 n
        row->hist_count = n_pairs;
        if (a.want_csr) {
          pair_base = atomicAdd(a.pair_count, (unsigned long long)n_pairs);
          row->hist_offset = pair_base;
          if (pair_base + n_pairs > a.pair_capacity) atomicOr(a.error_flags, ERR_CAPACITY);
        }
      }
      pair_base = __shfl_sync(FULL, pair_base, 0);
    }
  }
}
