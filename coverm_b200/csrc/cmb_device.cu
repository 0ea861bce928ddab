// libcoverm_b200 — device side: hand-written sm_100a kernels + the C ABI of include/coverm_b200.h.
//
// Data layout in HBM (one cmb_ctx = one GPU = one contig shard):
//   arena        i32[arena_elems]   all contigs' `ups_and_downs` (contig.rs:144-145) back to back; every contig
//                                   starts on a 16-element span boundary (SPAN), the arena is a whole number of
//                                   8192-element chunks (CHUNK).  4 B per reference base.
//   off_span     u32[n_local+1]     padded contig offsets in span units; len u32[n_local]
//   chunk_first  u32[n_chunks+1]    contig containing the first span of each chunk
//   tail_sum     i32[n_chunks]      K1: sum of the deltas of the contig that continues past the chunk end
//   carry_in     i32[n_chunks]      K1b: running depth at the first element of each chunk
//   rows         cmb_contig_stats[n_contigs]
//   rec / warp_table / ovf          K2 -> K3 histogram records
//
// Kernels (all HBM-bound integer work, no tensor cores):
//   K1  k1_filter_accumulate  one thread per record: FlagFilter + ReferenceSortedBamFilter predicates
//                             (lib.rs:59-79, filter.rs:243-336), per-contig read counters (contig.rs:157-211),
//                             +1/-1 delta REDs into the arena (contig.rs:166-202), chunk tail sums.
//   K1b k1b_chunk_carry       segmented scan of the per-chunk tail sums -> carry_in (so K2 needs no look-back).
//   K2  k2_scan_reduce        persistent CTAs, TMA (cp.async.bulk.tensor, 128B swizzle) + mbarrier ring of 32 KB
//                             chunks, blocked 16-element spans per thread, warp-shuffle segmented scan, then every
//                             O(L) reduction of EST:366-502 in one pass: sum/covered over the end-trimmed window,
//                             covered over the full contig, window depth histogram into a shared-memory table that
//                             is flushed as (depth,count) records; optionally re-zeroes the arena as it goes.
//   K3  k3_finalize           per contig: merge the records, trimmed-mean walk (EST:598-642) and the variance sums
//                             (EST:790-805) in integers; optional CSR histogram output.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/coverm_b200.h"

namespace {

constexpr uint32_t SPAN = 16;                   // elements per thread span; contig alignment
constexpr uint32_t K2_THREADS = 512;
constexpr uint32_t CHUNK = SPAN * K2_THREADS;   // 8192 elements = 32 KB
constexpr uint32_t CHUNK_BYTES = CHUNK * 4;
constexpr uint32_t CHUNK_SPANS = K2_THREADS;    // spans per chunk
constexpr uint32_t ROW_ELEMS = 32;              // TMA row: 32 x i32 = 128 B
constexpr uint32_t CHUNK_ROWS = CHUNK / ROW_ELEMS;  // 256
constexpr uint32_t K2_STAGES = 3;
constexpr uint32_t K2_WARPS = K2_THREADS / 32;  // 16
constexpr uint32_t HIST_SLOTS = 4;              // contigs per chunk with a shared-memory histogram
constexpr uint32_t HIST_BINS = 512;             // bins per slot
constexpr uint32_t HIST_TOTAL = HIST_SLOTS * HIST_BINS;  // 2048 = 16 warps x 128
constexpr uint32_t K3_THREADS = 128;
constexpr uint32_t K3_WINDOW = 1024;            // depth bins per K3 pass
constexpr uint32_t K1_THREADS = 256;
constexpr uint32_t ROWFLAG_OVF = 1u;            // cmb_contig_stats.reserved: some records are in the overflow list

// error_flags bits (device)
constexpr uint32_t ERR_UNSORTED = 1u, ERR_NM = 2u, ERR_BOUNDS = 4u, ERR_CAPACITY = 8u, ERR_TID = 16u, ERR_INTERNAL = 32u;

#define FULL 0xffffffffu

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// TMA: 2-D tiled bulk tensor load global -> shared, completion on an mbarrier.
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tmap, int32_t x, int32_t y, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(x), "r"(y), "r"(bar)
      : "memory");
}

__device__ __forceinline__ uint64_t warp_sum_u64(uint64_t v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
  return v;
}

// ------------------------------------------------------------------------------------------------ K1
struct K1Args {
  // batch (device pointers)
  const int32_t* tid;
  const int32_t* pos;
  const uint16_t* flag;
  const uint8_t* mapq;
  const uint8_t* nm_state;
  const uint32_t* nm;
  const uint32_t* l_seq;
  const uint32_t* aligned;
  const uint32_t* del;
  const uint32_t* ins;
  const uint32_t* iv_begin;
  const int32_t* iv_start;
  const int32_t* iv_len;
  uint32_t n;
  // reference
  const uint32_t* off_span;  // [n_local+1]
  const uint32_t* len;       // [n_local]
  uint32_t n_contigs, tid_begin, tid_end;
  // outputs
  int32_t* arena;
  int32_t* tail_sum;
  cmb_contig_stats* rows;
  int2* block_minmax;  // per block {min kept tid, max kept tid} for the cross-block sortedness check
  uint32_t* error_flags;
  // params
  cmb_params p;
  uint8_t filter_single, filter_pairs;
};

struct RecView {
  uint32_t flag, mapq, nm_state, nm, l_seq, aligned, del;
};

// filter.rs:243-279.  Sets *nm_err when the reference would reach nm() on a record without a usable NM tag.
__device__ __forceinline__ bool single_read_passes(const RecView& r, const cmb_params& p, bool* nm_err) {
  if (p.min_mapq != 255 && (r.mapq < p.min_mapq || r.mapq == 255)) return false;
  if (r.nm_state != 1) *nm_err = true;
  const float aligned_f = __uint2float_rn(r.aligned);
  return r.aligned >= p.min_aligned_length_single &&
         __fdiv_rn(aligned_f, __uint2float_rn(r.l_seq)) >= p.min_aligned_percent_single &&
         __fsub_rn(1.0f, __fdiv_rn(__uint2float_rn(r.nm), aligned_f)) >= p.min_percent_identity_single;
}
// filter.rs:281-336 (D is not part of the pair aligned length).
__device__ __forceinline__ bool read_pair_passes(const RecView& a, const RecView& b, const cmb_params& p, bool* nm_err) {
  if (p.min_mapq != 255 && (a.mapq < p.min_mapq || b.mapq < p.min_mapq || a.mapq == 255 || b.mapq == 255)) return false;
  if (a.nm_state != 1 || b.nm_state != 1) *nm_err = true;
  const uint32_t aligned = (a.aligned - a.del) + (b.aligned - b.del);
  const float aligned_f = __uint2float_rn(aligned);
  const float seq_f = __ull2float_rn((unsigned long long)a.l_seq + (unsigned long long)b.l_seq);
  const float edit_f = __ull2float_rn((unsigned long long)a.nm + (unsigned long long)b.nm);
  return aligned >= p.min_aligned_length_pair && __fdiv_rn(aligned_f, seq_f) >= p.min_aligned_percent_pair &&
         __fsub_rn(1.0f, __fdiv_rn(edit_f, aligned_f)) >= p.min_percent_identity_pair;
}

__global__ void __launch_bounds__(K1_THREADS) k1_filter_accumulate(const K1Args a) {
  const uint32_t i = blockIdx.x * K1_THREADS + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool valid = i < a.n;
  const cmb_params& p = a.p;

  RecView r = {};
  int32_t tid = -1, pos = 0;
  uint32_t ins = 0, ivb = 0, ive = 0;
  if (valid) {
    tid = a.tid[i];
    pos = a.pos[i];
    r.flag = a.flag[i];
    r.mapq = a.mapq[i];
    r.nm_state = a.nm_state[i];
    r.nm = a.nm[i];
    r.l_seq = a.l_seq[i];
    r.aligned = a.aligned[i];
    r.del = a.del[i];
    ins = a.ins[i];
    ivb = a.iv_begin[i];
    ive = a.iv_begin[i + 1];
  }
  const bool unmapped = r.flag & 0x4, secondary = r.flag & 0x100, supplementary = r.flag & 0x800, proper = r.flag & 0x2;
  // FlagFilter::passes, lib.rs:67-78
  const bool flag_pass = !(secondary && !p.include_secondary) && !(supplementary && !p.include_supplementary) &&
                         !(!proper && !p.include_improper_pairs);
  bool keep = valid && flag_pass && !unmapped;  // contig.rs:119-125
  bool nm_err = false;
  if (valid && p.filtering) {
    bool passes;
    if (a.filter_single && !a.filter_pairs) {  // filter.rs:88-116
      const bool passes_filter1 = !unmapped && (p.include_supplementary || !supplementary) && (p.include_secondary || !secondary);
      passes = passes_filter1 && single_read_passes(r, p, &nm_err);
    } else {  // filter.rs:117-233: the host submits completed pairs only; stored first mate at the even index
      const uint32_t m = i ^ 1u;
      RecView o = {};
      const bool have_mate = m < a.n;
      if (have_mate) {
        o.flag = a.flag[m];
        o.mapq = a.mapq[m];
        o.nm_state = a.nm_state[m];
        o.nm = a.nm[m];
        o.l_seq = a.l_seq[m];
        o.aligned = a.aligned[m];
        o.del = a.del[m];
      }
      const RecView& first = (i & 1u) ? o : r;   // record1 (stored)
      const RecView& second = (i & 1u) ? r : o;  // record (just read)
      bool ok = have_mate;
      if (ok && a.filter_single) ok = single_read_passes(first, p, &nm_err) && single_read_passes(second, p, &nm_err);
      if (ok) ok = read_pair_passes(second, first, p, &nm_err);
      passes = ok;
    }
    keep = keep && passes;
  }
  uint32_t err = 0;
  if (keep && r.nm_state != 1) nm_err = true;  // nm(&record), contig.rs:206
  if (nm_err) err |= ERR_NM;
  if (keep && (tid < 0 || (uint32_t)tid >= a.n_contigs)) {
    err |= ERR_TID;
    keep = false;
  }

  // ---- sortedness of the kept stream (contig.rs:128-132): prefix max over the block
  __shared__ int s_wmax[K1_THREADS / 32];
  __shared__ int s_wmin[K1_THREADS / 32];
  {
    const int key = keep ? tid : INT_MIN;
    int pm = key;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(FULL, pm, d);
      if ((int)lane >= d) pm = max(pm, o);
    }
    int excl = __shfl_up_sync(FULL, pm, 1);
    if (lane == 0) excl = INT_MIN;
    int kmin = keep ? tid : INT_MAX;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) kmin = min(kmin, __shfl_xor_sync(FULL, kmin, d));
    if (lane == 31) s_wmax[warp] = pm;
    if (lane == 0) s_wmin[warp] = kmin;
    __syncthreads();
    int before = INT_MIN;
    for (uint32_t w = 0; w < warp; ++w) before = max(before, s_wmax[w]);
    if (keep && tid < max(before, excl)) err |= ERR_UNSORTED;
    if (threadIdx.x == 0) {
      int bmax = INT_MIN, bmin = INT_MAX;
      for (uint32_t w = 0; w < K1_THREADS / 32; ++w) {
        bmax = max(bmax, s_wmax[w]);
        bmin = min(bmin, s_wmin[w]);
      }
      a.block_minmax[blockIdx.x] = make_int2(bmin, bmax);
    }
  }

  const bool mine = keep && (uint32_t)tid >= a.tid_begin && (uint32_t)tid < a.tid_end;
  // ---- per-contig read counters (contig.rs:157-159, 204-211; genome.rs:173-174, 220-223, 677-682, 724-727)
  {
    const bool primary = !secondary && !supplementary;
    const uint64_t c_rec = mine ? 1 : 0, c_pri = (mine && primary) ? 1 : 0, c_ns = (mine && !supplementary) ? 1 : 0;
    const uint64_t c_edit = mine ? r.nm : 0, c_indel = mine ? (uint64_t)ins + r.del : 0;
    double idn = 0.0;
    if (mine && r.aligned > 0) idn = ((double)r.aligned - (double)r.nm) / (double)r.aligned;
    const double id_pri = primary ? idn : 0.0, id_ns = !supplementary ? idn : 0.0;
    const uint32_t mine_mask = __ballot_sync(FULL, mine);
    if (mine_mask) {
      const int leader = __ffs(mine_mask) - 1;
      const int ltid = __shfl_sync(FULL, tid, leader);
      const bool uniform = __all_sync(FULL, !mine || tid == ltid);
      if (uniform) {
        const uint64_t s_rec = warp_sum_u64(c_rec), s_pri = warp_sum_u64(c_pri), s_ns = warp_sum_u64(c_ns),
                       s_edit = warp_sum_u64(c_edit), s_indel = warp_sum_u64(c_indel);
        double s_idp = id_pri, s_idn = id_ns;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
          s_idp += __shfl_xor_sync(FULL, s_idp, d);
          s_idn += __shfl_xor_sync(FULL, s_idn, d);
        }
        if ((int)lane == leader) {
          cmb_contig_stats* row = a.rows + ltid;
          atomicAdd((unsigned long long*)&row->n_records, (unsigned long long)s_rec);
          if (s_pri) atomicAdd((unsigned long long*)&row->n_primary, (unsigned long long)s_pri);
          if (s_ns) atomicAdd((unsigned long long*)&row->n_nonsupp, (unsigned long long)s_ns);
          if (s_edit) atomicAdd((unsigned long long*)&row->sum_edit, (unsigned long long)s_edit);
          if (s_indel) atomicAdd((unsigned long long*)&row->sum_indel, (unsigned long long)s_indel);
          if (s_idp != 0.0) atomicAdd(&row->sum_identity_primary, s_idp);
          if (s_idn != 0.0) atomicAdd(&row->sum_identity_nonsupp, s_idn);
        }
      } else if (mine) {
        cmb_contig_stats* row = a.rows + tid;
        atomicAdd((unsigned long long*)&row->n_records, 1ull);
        if (c_pri) atomicAdd((unsigned long long*)&row->n_primary, 1ull);
        if (c_ns) atomicAdd((unsigned long long*)&row->n_nonsupp, 1ull);
        if (c_edit) atomicAdd((unsigned long long*)&row->sum_edit, (unsigned long long)c_edit);
        if (c_indel) atomicAdd((unsigned long long*)&row->sum_indel, (unsigned long long)c_indel);
        if (id_pri != 0.0) atomicAdd(&row->sum_identity_primary, id_pri);
        if (id_ns != 0.0) atomicAdd(&row->sum_identity_nonsupp, id_ns);
      }
    }
  }

  // ---- delta events (contig.rs:171-186)
  if (mine) {
    const uint32_t lc = (uint32_t)tid - a.tid_begin;
    const uint32_t L = a.len[lc];
    const uint64_t base = (uint64_t)a.off_span[lc] * SPAN;
    const uint64_t end_padded = (uint64_t)a.off_span[lc + 1] * SPAN;  // first element of the next contig
    (void)pos;
    for (uint32_t k = ivb; k < ive; ++k) {
      const int32_t s = a.iv_start[k];
      const uint32_t n = (uint32_t)a.iv_len[k];
      if (s < 0 || (uint32_t)s >= L) {  // `ups_and_downs[cursor] += 1` would panic
        err |= ERR_BOUNDS;
        continue;
      }
      const uint64_t gs = base + (uint32_t)s;
      const uint64_t e = (uint64_t)(uint32_t)s + n;
      const bool has_end = e < L;  // "True unless the read hits the contig end"
      atomicAdd(a.arena + gs, 1);
      const uint64_t ks = gs / CHUNK;
      const bool cont_s = end_padded > (ks + 1) * (uint64_t)CHUNK;  // this contig continues past chunk ks
      if (has_end) {
        const uint64_t ge = base + e;
        atomicAdd(a.arena + ge, -1);
        const uint64_t ke = ge / CHUNK;
        if (ke != ks) {
          if (cont_s) atomicAdd(a.tail_sum + ks, 1);
          if (end_padded > (ke + 1) * (uint64_t)CHUNK) atomicAdd(a.tail_sum + ke, -1);
        }
      } else if (cont_s) {
        atomicAdd(a.tail_sum + ks, 1);
      }
    }
  }
  err = __reduce_or_sync(FULL, err);
  if (err && lane == 0) atomicOr(a.error_flags, err);
}

// Cross-block sortedness: block b's smallest kept tid must be >= every earlier block's largest.
__global__ void __launch_bounds__(1024) k1c_check_sorted(const int2* block_minmax, uint32_t n_blocks, uint32_t* error_flags) {
  __shared__ int s_max[1024];
  const uint32_t t = threadIdx.x;
  const uint32_t per = (n_blocks + 1023) / 1024;
  const uint32_t b0 = t * per, b1 = min(n_blocks, b0 + per);
  int lmax = INT_MIN;
  bool bad = false;
  for (uint32_t b = b0; b < b1; ++b) {
    const int2 mm = block_minmax[b];
    if (mm.x != INT_MAX && mm.x < lmax) bad = true;
    lmax = max(lmax, mm.y);
  }
  s_max[t] = lmax;
  __syncthreads();
  int before = INT_MIN;
  for (uint32_t k = 0; k < t; ++k) before = max(before, s_max[k]);
  for (uint32_t b = b0; b < b1 && !bad; ++b) {
    const int2 mm = block_minmax[b];
    if (mm.x != INT_MAX && mm.x < before) bad = true;
  }
  if (bad) atomicOr(error_flags, ERR_UNSORTED);
}

// ------------------------------------------------------------------------------------------------ K1b
// carry_in[0] = 0; carry_in[k+1] = (chunk k+1 starts mid-contig ? (same contig as chunk k's first span ? carry_in[k] : 0)
//                                   + tail_sum[k] : 0).   One CTA; n_chunks is ~L/8192.
__global__ void __launch_bounds__(1024) k1b_chunk_carry(const int32_t* tail_sum, const uint32_t* chunk_first,
                                                        const uint32_t* off_span, uint32_t n_chunks, int32_t* carry_in) {
  __shared__ int s_val[1024];
  __shared__ int s_flg[1024];
  const uint32_t t = threadIdx.x;
  const uint32_t per = (n_chunks + 1023) / 1024;
  const uint32_t k0 = t * per, k1 = min(n_chunks, k0 + per);
  // element k of the recurrence: x_{k+1} = (reset_k ? 0 : x_k) + add_k with
  //   mid   = chunk k+1 starts mid-contig, same = chunk_first[k]==chunk_first[k+1]
  //   reset_k = !(mid && same), add_k = mid ? tail_sum[k] : 0
  int val = 0;
  int flg = 0;
  for (uint32_t k = k0; k < k1; ++k) {
    const uint32_t cn = chunk_first[k + 1];
    const bool mid = (k + 1 < n_chunks) && off_span[cn] < (k + 1) * CHUNK_SPANS;
    const bool same = chunk_first[k] == cn;
    const bool reset = !(mid && same);
    const int add = mid ? tail_sum[k] : 0;
    if (reset) {
      val = add;
      flg = 1;
    } else {
      val += add;
    }
  }
  s_val[t] = val;
  s_flg[t] = flg;
  __syncthreads();
  // exclusive segmented combine of the preceding threads (1024 serial steps at most; negligible)
  int x = 0;
  for (uint32_t j = 0; j < t; ++j) x = s_flg[j] ? s_val[j] : x + s_val[j];
  for (uint32_t k = k0; k < k1; ++k) {
    carry_in[k] = x;  // x_k
    const uint32_t cn = chunk_first[k + 1];
    const bool mid = (k + 1 < n_chunks) && off_span[cn] < (k + 1) * CHUNK_SPANS;
    const bool same = chunk_first[k] == cn;
    const int add = mid ? tail_sum[k] : 0;
    x = (mid && same) ? x + add : add;
  }
}

// ------------------------------------------------------------------------------------------------ K2
struct K2Args {
  const uint32_t* off_span;
  const uint32_t* len;
  const uint32_t* chunk_first;
  const int32_t* carry_in;
  cmb_contig_stats* rows;
  uint32_t tid_begin, n_local, n_chunks, excl;
  uint32_t* ticket;
  int32_t* arena;
  uint2* rec;
  uint32_t rec_capacity;
  uint32_t* rec_count;
  uint2* warp_table;  // [n_chunks * 16] {offset, count}
  uint4* ovf;         // {contig_local, depth, count, 0}
  uint32_t ovf_capacity;
  uint32_t* ovf_count;
  uint32_t* error_flags;
};

constexpr uint32_t K2_SMEM_STAGE_BYTES = K2_STAGES * CHUNK_BYTES;
constexpr uint32_t K2_SMEM_BYTES = K2_SMEM_STAGE_BYTES + 64 /*barriers+tickets*/ + 2 * K2_WARPS * 8 + HIST_TOTAL * 4 + 1024 /*align slack*/;

template <bool HIST, bool CLEAN>
__global__ void __launch_bounds__(K2_THREADS, 2) k2_scan_reduce(const __grid_constant__ CUtensorMap tmap, const K2Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);  // 128B swizzle atom = 1024 B
  uint64_t* full = (uint64_t*)(smem + K2_SMEM_STAGE_BYTES);
  uint32_t* s_chunk = (uint32_t*)(full + K2_STAGES);
  int2* wagg2 = (int2*)(smem + K2_SMEM_STAGE_BYTES + 64);  // double-buffered by iteration parity
  uint32_t* hist = (uint32_t*)(wagg2 + 2 * K2_WARPS);

  const uint32_t t = threadIdx.x, lane = t & 31, warp = t >> 5;

  auto issue = [&](uint32_t s) {  // thread 0: claim the next chunk and start its TMA load into stage s
    const uint32_t tk = atomicAdd(a.ticket, 1u);
    s_chunk[s] = tk;
    if (tk < a.n_chunks) {
      const uint32_t bar = smem_u32(full + s);
      mbar_arrive_expect_tx(bar, CHUNK_BYTES);
      tma_load_2d(smem_u32(smem + s * CHUNK_BYTES), &tmap, 0, (int32_t)(tk * CHUNK_ROWS), bar);
    }
  };

  if (t == 0) {
    for (uint32_t s = 0; s < K2_STAGES; ++s) mbar_init(smem_u32(full + s), 1);
    fence_barrier_init();
  }
  if (HIST)
    for (uint32_t b = t; b < HIST_TOTAL; b += K2_THREADS) hist[b] = 0;
  __syncthreads();
  if (t == 0)
    for (uint32_t s = 0; s < K2_STAGES; ++s) issue(s);
  __syncthreads();

  const uint32_t row = t >> 1, half = t & 1;
  const uint32_t E = a.excl;

  for (uint32_t it = 0;; ++it) {
    const uint32_t s = it % K2_STAGES;
    const uint32_t chunk = s_chunk[s];
    if (chunk >= a.n_chunks) break;
    int2* wagg = wagg2 + (it & 1) * K2_WARPS;
    mbar_wait(smem_u32(full + s), (it / K2_STAGES) & 1);

    // ---- 16 consecutive elements per thread: 4 x LDS.128 through the 128B swizzle (conflict-free)
    int v[SPAN];
    {
      const uint8_t* rowp = smem + s * CHUNK_BYTES + row * 128;
#pragma unroll
      for (uint32_t j = 0; j < 4; ++j) {
        const uint32_t unit = (half * 4 + j) ^ (row & 7);
        const int4 q = *reinterpret_cast<const int4*>(rowp + unit * 16);
        v[4 * j + 0] = q.x;
        v[4 * j + 1] = q.y;
        v[4 * j + 2] = q.z;
        v[4 * j + 3] = q.w;
      }
    }
    const uint32_t span = chunk * CHUNK_SPANS + t;
    if (CLEAN) {  // re-zero only the 16 B units that hold an event (the arena is zero everywhere else)
      int4* g = reinterpret_cast<int4*>(a.arena + (uint64_t)span * SPAN);
#pragma unroll
      for (uint32_t j = 0; j < 4; ++j)
        if (v[4 * j] | v[4 * j + 1] | v[4 * j + 2] | v[4 * j + 3]) g[j] = make_int4(0, 0, 0, 0);
    }

    // ---- which contig owns this span
    const uint32_t cf = __ldg(a.chunk_first + chunk);
    uint32_t lo = cf, hi = __ldg(a.chunk_first + chunk + 1);
    while (lo < hi) {
      const uint32_t mid = (lo + hi + 1) >> 1;
      if (__ldg(a.off_span + mid) <= span) lo = mid;
      else hi = mid - 1;
    }
    const uint32_t c = lo;
    const uint32_t cstart = __ldg(a.off_span + c);
    const uint32_t L = __ldg(a.len + c);
    const bool is_head = span == cstart;
    const uint32_t rel = (span - cstart) * SPAN;  // position in the contig of v[0]
    const uint32_t n_in = rel >= L ? 0u : min(SPAN, L - rel);
    uint32_t w0 = 0, w1 = 0;
    if (2ull * E < L) {
      const uint32_t ws = E, we = L - E;
      w0 = rel >= ws ? 0u : min(SPAN, ws - rel);
      w1 = rel >= we ? 0u : min(SPAN, we - rel);
      if (w1 < w0) w1 = w0;
    }

    // ---- thread-local inclusive prefix, event mask
    int nz = v[0];
#pragma unroll
    for (uint32_t j = 1; j < SPAN; ++j) {
      nz |= v[j];
      v[j] += v[j - 1];
    }

    // ---- segmented (by contig head) inclusive scan of span totals across the warp
    int val = v[SPAN - 1];
    int flg = is_head;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int ov = __shfl_up_sync(FULL, val, d);
      const int of = __shfl_up_sync(FULL, flg, d);
      if ((int)lane >= d) {
        if (!flg) val += ov;
        flg |= of;
      }
    }
    int pval = __shfl_up_sync(FULL, val, 1), pflg = __shfl_up_sync(FULL, flg, 1);
    if (lane == 0) {
      pval = 0;
      pflg = 0;
    }
    if (lane == 31) wagg[warp] = make_int2(val, flg);
    __syncthreads();  // (A) stage fully read, warp aggregates visible, previous chunk's histogram flush done
    if (t == 0) issue(s);

    int wv, wf;
    {
      const int2 wa = lane < K2_WARPS ? wagg[lane] : make_int2(0, 0);
      wv = wa.x;
      wf = wa.y;
#pragma unroll
      for (int d = 1; d < (int)K2_WARPS; d <<= 1) {
        const int ov = __shfl_up_sync(FULL, wv, d);
        const int of = __shfl_up_sync(FULL, wf, d);
        if ((int)lane >= d) {
          if (!wf) wv += ov;
          wf |= of;
        }
      }
      const int src = warp ? (int)warp - 1 : 0;
      wv = __shfl_sync(FULL, wv, src);
      wf = __shfl_sync(FULL, wf, src);
      if (warp == 0) {
        wv = 0;
        wf = 0;
      }
    }
    const int cin = __ldg(a.carry_in + chunk);
    int carry;
    if (is_head) carry = 0;
    else if (pflg) carry = pval;
    else if (wf) carry = wv + pval;
    else carry = cin + wv + pval;

    // ---- reductions over this span (EST:393-404, 447-465, 494-501)
    uint32_t cov_full = 0, cov_win = 0;
    uint64_t sum_win = 0;
    const uint32_t slot = c - cf;
    const int base0 = max(0, cin - (int)(HIST_BINS / 2));
    const int hbase = slot == 0 ? base0 : 0;
    auto hist_add = [&](int depth, uint32_t cnt) {
      if (!HIST || cnt == 0) return;
      const int b = depth - hbase;
      if (depth < 0) {  // impossible for a consistent arena (every -1 follows its +1 within the contig)
        atomicOr(a.error_flags, ERR_INTERNAL);
      } else if (slot < HIST_SLOTS && (uint32_t)b < HIST_BINS) {
        atomicAdd(hist + slot * HIST_BINS + b, cnt);
      } else {  // rare: more than HIST_SLOTS contigs in the chunk, or depth outside the window
        const uint32_t o = atomicAdd(a.ovf_count, 1u);
        if (o < a.ovf_capacity) a.ovf[o] = make_uint4(c, (uint32_t)depth, cnt, 0);
        else atomicOr(a.error_flags, ERR_CAPACITY);
        atomicOr(&a.rows[a.tid_begin + c].reserved, ROWFLAG_OVF);
      }
    };
    const bool uni = nz == 0;  // no event in the span: constant depth
    if (uni) {
      const uint32_t pos = carry > 0;
      cov_full = pos * n_in;
      cov_win = pos * (w1 - w0);
      sum_win = (uint64_t)(int64_t)carry * (w1 - w0);
    } else {
      int run_depth = 0;
      uint32_t run_cnt = 0;
#pragma unroll
      for (uint32_t j = 0; j < SPAN; ++j) {
        const int d = carry + v[j];
        const bool in_c = j < n_in, in_w = j >= w0 && j < w1;
        cov_full += (in_c && d > 0);
        cov_win += (in_w && d > 0);
        if (in_w) sum_win += (uint64_t)(int64_t)d;
        if (HIST && in_w) {
          if (run_cnt && d == run_depth) {
            ++run_cnt;
          } else {
            hist_add(run_depth, run_cnt);
            run_depth = d;
            run_cnt = 1;
          }
        }
      }
      if (HIST) hist_add(run_depth, run_cnt);
    }
    if (HIST) {
      // spans at one constant depth: aggregate the lanes that agree with the first such lane into one shared atomic
      const uint32_t cnt = w1 - w0;
      const bool cand = uni && cnt > 0;
      const uint32_t cm = __ballot_sync(FULL, cand);
      if (cm) {
        const int leader = __ffs(cm) - 1;
        const int d0 = __shfl_sync(FULL, carry, leader);
        const uint32_t c0 = __shfl_sync(FULL, c, leader);
        const bool same = cand && carry == d0 && c == c0;
        const uint32_t m = __ballot_sync(FULL, same);
        if (same) {
          const uint32_t tot = __reduce_add_sync(m, cnt);
          if ((int)lane == leader) hist_add(carry, tot);
        } else if (cand) {
          hist_add(carry, cnt);
        }
      }
    }

    // ---- per-contig accumulation: one RED triple per (warp, contig)
    {
      const uint32_t c0 = __shfl_sync(FULL, c, 0);
      if (__all_sync(FULL, c == c0)) {
        const uint32_t sf = __reduce_add_sync(FULL, cov_full), sw = __reduce_add_sync(FULL, cov_win);
        const uint64_t sd = warp_sum_u64(sum_win);
        if (lane == 0) {
          cmb_contig_stats* rowp = a.rows + a.tid_begin + c0;
          if (sf) atomicAdd((unsigned long long*)&rowp->covered_full, (unsigned long long)sf);
          if (sw) atomicAdd((unsigned long long*)&rowp->covered_window, (unsigned long long)sw);
          if (sd) atomicAdd((unsigned long long*)&rowp->sum_depth_window, (unsigned long long)sd);
        }
      } else {
        cmb_contig_stats* rowp = a.rows + a.tid_begin + c;
        if (cov_full) atomicAdd((unsigned long long*)&rowp->covered_full, (unsigned long long)cov_full);
        if (cov_win) atomicAdd((unsigned long long*)&rowp->covered_window, (unsigned long long)cov_win);
        if (sum_win) atomicAdd((unsigned long long*)&rowp->sum_depth_window, (unsigned long long)sum_win);
      }
    }

    // ---- flush the chunk histogram: warp w owns slot w/4, bins (w%4)*128 .. +128
    if (HIST) {
      __syncthreads();  // (C) all histogram adds of this chunk done
      const uint32_t fslot = warp >> 2;
      const uint32_t bin0 = (warp & 3) * 128 + lane;
      uint32_t cnt[4], msk[4], total = 0;
#pragma unroll
      for (uint32_t k = 0; k < 4; ++k) {
        cnt[k] = hist[fslot * HIST_BINS + bin0 + 32 * k];
        msk[k] = __ballot_sync(FULL, cnt[k] != 0);
        total += __popc(msk[k]);
      }
      uint32_t base = 0;
      if (total) {
        if (lane == 0) base = atomicAdd(a.rec_count, total);
        base = __shfl_sync(FULL, base, 0);
        const bool fits = (uint64_t)base + total <= a.rec_capacity;
        if (!fits && lane == 0) atomicOr(a.error_flags, ERR_CAPACITY);
        uint32_t before = 0;
        const int fb = fslot == 0 ? base0 : 0;
#pragma unroll
        for (uint32_t k = 0; k < 4; ++k) {
          if (cnt[k]) {
            if (fits) a.rec[base + before + __popc(msk[k] & ((1u << lane) - 1))] = make_uint2((uint32_t)(fb + (int)(bin0 + 32 * k)), cnt[k]);
            hist[fslot * HIST_BINS + bin0 + 32 * k] = 0;
          }
          before += __popc(msk[k]);
        }
        if (!fits) total = 0;
      }
      if (lane == 0) a.warp_table[(uint64_t)chunk * K2_WARPS + warp] = make_uint2(base, total);
    }
  }
}

// ------------------------------------------------------------------------------------------------ K3
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
  const uint32_t* ovf_count;
  uint32_t ovf_capacity;
  // CSR output
  cmb_hist_pair* pairs;
  unsigned long long* pair_count;
  uint64_t pair_capacity;
  uint32_t want_csr;
  uint32_t* error_flags;
};

__global__ void __launch_bounds__(K3_THREADS) k3_finalize(const K3Args a) {
  __shared__ uint32_t whist[K3_WINDOW];
  __shared__ unsigned long long s_scan[K3_THREADS];
  __shared__ unsigned long long s_red[4][K3_THREADS / 32];
  __shared__ uint32_t s_max;
  __shared__ unsigned long long s_base;

  const uint32_t t = threadIdx.x, lane = t & 31, warp = t >> 5;
  for (uint32_t lc = blockIdx.x; lc < a.n_local; lc += gridDim.x) {
    cmb_contig_stats* row = a.rows + a.tid_begin + lc;
    if (row->n_records == 0) continue;  // unseen contig: the host never consults its histogram
    const uint32_t L = a.len[lc];
    const uint64_t E = a.excl;
    if (!(2 * E < L)) continue;  // no window (EST:436-445)
    const uint64_t T = (uint64_t)L - 2 * E;
    // EST:591-592, f32 products; `as usize` saturates
    const float Tf = __ull2float_rn(T);
    const uint64_t min_index = (uint64_t)floorf(__fmul_rn(a.trim_min, Tf));
    const uint64_t max_index = (uint64_t)ceilf(__fmul_rn(a.trim_max, Tf));
    const uint32_t k0 = a.off_span[lc] / CHUNK_SPANS, k1 = (a.off_span[lc + 1] - 1) / CHUNK_SPANS;
    const bool has_ovf = row->reserved & ROWFLAG_OVF;
    const uint32_t n_ovf = has_ovf ? min(*a.ovf_count, a.ovf_capacity) : 0;

    unsigned long long cum = 0;        // counts below the current window
    unsigned long long total = 0;      // trimmed-mean `total`
    unsigned long long S0 = 0, S1 = 0, S2 = 0;
    unsigned long long kmin = ~0ull;
    uint32_t n_pairs = 0;
    unsigned long long pair_base = 0;
    const int n_rounds = a.want_csr ? 2 : 1;  // round 0: statistics (+ count pairs); round 1: write pairs
    for (int round = 0; round < n_rounds; ++round) {
      uint32_t written = 0;
      for (uint32_t wbase = 0;; wbase += K3_WINDOW) {
        for (uint32_t b = t; b < K3_WINDOW; b += K3_THREADS) whist[b] = 0;
        if (t == 0) s_max = 0;
        __syncthreads();
        uint32_t lmax = 0;
        // gather: table entries (chunk, 4 warps of my slot)
        const uint32_t n_ent = (k1 - k0 + 1) * 4;
        for (uint32_t e = t; e < n_ent; e += K3_THREADS) {
          const uint32_t k = k0 + (e >> 2);
          const uint32_t slot = lc - a.chunk_first[k];
          if (slot >= HIST_SLOTS) continue;
          const uint2 ent = a.warp_table[(uint64_t)k * K2_WARPS + slot * 4 + (e & 3)];
          for (uint32_t r = 0; r < ent.y; ++r) {
            const uint2 rc = a.rec[ent.x + r];
            lmax = max(lmax, rc.x);
            if (rc.x >= wbase && rc.x - wbase < K3_WINDOW) atomicAdd(&whist[rc.x - wbase], rc.y);
          }
        }
        for (uint32_t o = t; o < n_ovf; o += K3_THREADS) {
          const uint4 rc = a.ovf[o];
          if (rc.x != lc) continue;
          lmax = max(lmax, rc.y);
          if (rc.y >= wbase && rc.y - wbase < K3_WINDOW) atomicAdd(&whist[rc.y - wbase], rc.z);
        }
        atomicMax(&s_max, lmax);
        __syncthreads();
        const uint32_t maxd = s_max;
        // walk this window: thread t owns bins [8t, 8t+8)
        constexpr uint32_t PER = K3_WINDOW / K3_THREADS;
        uint32_t n[PER];
        unsigned long long tsum = 0;
        uint32_t nnz = 0;
#pragma unroll
        for (uint32_t j = 0; j < PER; ++j) {
          n[j] = whist[t * PER + j];
          tsum += n[j];
          nnz += n[j] != 0;
        }
        // block exclusive scan of tsum and nnz (packed: nnz <= 8*128 fits in the low 16 bits... keep separate)
        s_scan[t] = tsum;
        __syncthreads();
        unsigned long long before = 0;
        for (uint32_t k = 0; k < t; ++k) before += s_scan[k];
        unsigned long long wtotal = 0;
        if (t == K3_THREADS - 1) wtotal = before + tsum;
        __syncthreads();
        s_scan[t] = nnz;
        __syncthreads();
        uint32_t nz_before = 0, nz_total = 0;
        for (uint32_t k = 0; k < K3_THREADS; ++k) {
          const uint32_t x = (uint32_t)s_scan[k];
          if (k < t) nz_before += x;
          nz_total += x;
        }
        if (round == 0) {
          unsigned long long cprev = cum + before;
          unsigned long long ltot = 0, l0 = 0, l1 = 0, l2 = 0, lk = ~0ull;
#pragma unroll
          for (uint32_t j = 0; j < PER; ++j) {
            if (n[j] == 0) continue;
            const unsigned long long depth = (unsigned long long)wbase + t * PER + j;
            const unsigned long long cnt = n[j];
            const unsigned long long ccur = cprev + cnt;
            unsigned long long w;
            if (ccur < min_index) w = 0;
            else if (cprev < min_index) w = ccur > max_index ? max_index - min_index + 1 : ccur - min_index + 1;
            else w = cprev > max_index ? 0 : (ccur > max_index ? max_index - cprev + 1 : cnt);
            ltot += w * depth;
            l0 += cnt;
            l1 += depth * cnt;
            l2 += depth * depth * cnt;
            lk = min(lk, depth);
            cprev = ccur;
          }
          // block reduce
          ltot = warp_sum_u64(ltot);
          l0 = warp_sum_u64(l0);
          l1 = warp_sum_u64(l1);
          l2 = warp_sum_u64(l2);
#pragma unroll
          for (int d = 16; d > 0; d >>= 1) lk = min(lk, __shfl_xor_sync(FULL, lk, d));
          __syncthreads();
          if (lane == 0) {
            s_red[0][warp] = ltot;
            s_red[1][warp] = l0;
            s_red[2][warp] = l1;
            s_red[3][warp] = l2;
          }
          __syncthreads();
          for (uint32_t w = 0; w < K3_THREADS / 32; ++w) {
            total += s_red[0][w];
            S0 += s_red[1][w];
            S1 += s_red[2][w];
            S2 += s_red[3][w];
          }
          __syncthreads();
          if (lane == 0) s_red[0][warp] = lk;
          __syncthreads();
          for (uint32_t w = 0; w < K3_THREADS / 32; ++w) kmin = min(kmin, s_red[0][w]);
          n_pairs += nz_total;
        } else {
          uint32_t idx = written + nz_before;
#pragma unroll
          for (uint32_t j = 0; j < PER; ++j) {
            if (n[j] == 0) continue;
            if (pair_base + idx < a.pair_capacity) {
              cmb_hist_pair pr;
              pr.depth = wbase + t * PER + j;
              pr.count = n[j];
              a.pairs[pair_base + idx] = pr;
            }
            ++idx;
          }
          written += nz_total;
        }
        // total count in this window -> cum
        if (t == K3_THREADS - 1) s_base = wtotal;
        __syncthreads();
        cum += s_base;
        __syncthreads();
        if (maxd < wbase + K3_WINDOW || wbase + K3_WINDOW < wbase) break;
      }
      if (round == 0) {
        if (t == 0) {
          row->trimmed_total = total;
          row->trim_min_index = min_index;
          row->trim_max_index = max_index;
          const unsigned long long k = kmin == ~0ull ? 0 : kmin;
          row->var_k = k;
          row->var_ex = S1 - k * S0;                       // sum (x-k) n   (mod 2^64)
          row->var_ex2 = S2 - 2 * k * S1 + k * k * S0;     // sum (x-k)^2 n (mod 2^64)
          row->hist_count = n_pairs;
          if (a.want_csr) {
            const unsigned long long b = atomicAdd(a.pair_count, (unsigned long long)n_pairs);
            row->hist_offset = b;
            s_base = b;
            if (b + n_pairs > a.pair_capacity) atomicOr(a.error_flags, ERR_CAPACITY);
          }
        }
        __syncthreads();
        pair_base = s_base;
        cum = 0;
        __syncthreads();
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ host context
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

std::string g_create_error;

struct DevBatch {  // device mirror of one staging batch
  void* slab = nullptr;
  cmb_read_batch ptr{};
};

}  // namespace

struct cmb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  cmb_device_cfg cfg{};
  int sm_count = 0;
  // staging
  std::vector<void*> host_slab;
  std::vector<cmb_read_batch> host_batch;
  std::vector<DevBatch> dev_batch;
  std::vector<cudaEvent_t> batch_done;
  std::vector<bool> batch_busy;
  int acquired = -1;
  uint32_t next_batch = 0;
  // reference
  uint32_t n_contigs = 0, tid_begin = 0, tid_end = 0, n_local = 0;
  uint64_t arena_elems = 0;
  uint32_t n_chunks = 0;
  int32_t* d_arena = nullptr;
  uint32_t *d_off_span = nullptr, *d_len = nullptr, *d_chunk_first = nullptr;
  int32_t *d_tail_sum = nullptr, *d_carry_in = nullptr;
  cmb_contig_stats* d_rows = nullptr;
  uint32_t* d_counters = nullptr;  // [0] error flags, [1] ticket, [2] rec_count, [3] ovf_count, [4..5] pair_count (u64)
  uint2* d_rec = nullptr;
  uint32_t rec_capacity = 0;
  uint2* d_warp_table = nullptr;
  uint4* d_ovf = nullptr;
  uint32_t ovf_capacity = 0;
  cmb_hist_pair* d_pairs = nullptr;
  uint64_t pair_capacity = 0;
  int2* d_block_minmax = nullptr;
  uint32_t block_minmax_capacity = 0, block_minmax_used = 0;
  CUtensorMap tmap{};
  bool arena_dirty = true;
  bool clean_as_you_go = true;
  // params
  cmb_params params{};
  cmb_filter_mode mode{};
  bool have_params = false, in_sample = false, ended = false;
  // timing
  cudaEvent_t ev[8]{};
  cmb_sample_timing timing{};
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> k1_events;
  uint32_t k1_events_used = 0;
  uint64_t n_records = 0, n_intervals = 0;
};

namespace {

int fail(cmb_ctx* ctx, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  else g_create_error = buf;
  return code;
}

#define CU_TRY(ctx, expr)                                                                                   \
  do {                                                                                                      \
    cudaError_t e_ = (expr);                                                                                \
    if (e_ != cudaSuccess) return fail(ctx, e_ == cudaErrorMemoryAllocation ? CMB_E_NOMEM : CMB_E_CUDA,      \
                                       "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

size_t batch_slab_bytes(uint32_t nr, uint32_t ni, size_t* offs) {
  // column order: tid,pos,nm,l_seq,aligned,del,ins,iv_begin(nr+1),iv_start(ni),iv_len(ni),flag(u16),mapq(u8),nm_state(u8)
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t r = o;
    o += (bytes + 255) & ~(size_t)255;
    return r;
  };
  offs[0] = take(4ull * nr);        // tid
  offs[1] = take(4ull * nr);        // pos
  offs[2] = take(4ull * nr);        // nm
  offs[3] = take(4ull * nr);        // l_seq
  offs[4] = take(4ull * nr);        // aligned
  offs[5] = take(4ull * nr);        // del
  offs[6] = take(4ull * nr);        // ins
  offs[7] = take(4ull * (nr + 1));  // iv_begin
  offs[8] = take(4ull * ni);        // iv_start
  offs[9] = take(4ull * ni);        // iv_len
  offs[10] = take(2ull * nr);       // flag
  offs[11] = take(1ull * nr);       // mapq
  offs[12] = take(1ull * nr);       // nm_state
  return o;
}

void carve_batch(void* slab, uint32_t nr, uint32_t ni, cmb_read_batch* b) {
  size_t offs[13];
  batch_slab_bytes(nr, ni, offs);
  uint8_t* p = (uint8_t*)slab;
  b->capacity_records = nr;
  b->capacity_intervals = ni;
  b->tid = (int32_t*)(p + offs[0]);
  b->pos = (int32_t*)(p + offs[1]);
  b->nm = (uint32_t*)(p + offs[2]);
  b->l_seq = (uint32_t*)(p + offs[3]);
  b->aligned = (uint32_t*)(p + offs[4]);
  b->del = (uint32_t*)(p + offs[5]);
  b->ins = (uint32_t*)(p + offs[6]);
  b->iv_begin = (uint32_t*)(p + offs[7]);
  b->iv_start = (int32_t*)(p + offs[8]);
  b->iv_len = (int32_t*)(p + offs[9]);
  b->flag = (uint16_t*)(p + offs[10]);
  b->mapq = (uint8_t*)(p + offs[11]);
  b->nm_state = (uint8_t*)(p + offs[12]);
}

void free_reference(cmb_ctx* c) {
  cudaFree(c->d_arena);
  cudaFree(c->d_off_span);
  cudaFree(c->d_len);
  cudaFree(c->d_chunk_first);
  cudaFree(c->d_tail_sum);
  cudaFree(c->d_carry_in);
  cudaFree(c->d_rows);
  cudaFree(c->d_rec);
  cudaFree(c->d_warp_table);
  cudaFree(c->d_ovf);
  cudaFree(c->d_pairs);
  c->d_arena = nullptr;
  c->d_off_span = c->d_len = c->d_chunk_first = nullptr;
  c->d_tail_sum = c->d_carry_in = nullptr;
  c->d_rows = nullptr;
  c->d_rec = nullptr;
  c->d_warp_table = nullptr;
  c->d_ovf = nullptr;
  c->d_pairs = nullptr;
  c->pair_capacity = 0;
}

int launch_k1(cmb_ctx* c, const cmb_read_batch& b, uint32_t n_records, uint32_t n_intervals) {
  if (n_records == 0) return CMB_OK;
  const uint32_t blocks = (n_records + K1_THREADS - 1) / K1_THREADS;
  if (c->block_minmax_used + blocks > c->block_minmax_capacity) {
    // grow (rare): allocate a larger array and copy what is there
    uint32_t ncap = std::max(c->block_minmax_capacity * 2, c->block_minmax_used + blocks + 4096);
    int2* nd = nullptr;
    CU_TRY(c, cudaMalloc(&nd, sizeof(int2) * (size_t)ncap));
    if (c->block_minmax_used)
      CU_TRY(c, cudaMemcpyAsync(nd, c->d_block_minmax, sizeof(int2) * (size_t)c->block_minmax_used, cudaMemcpyDeviceToDevice, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->d_block_minmax);
    c->d_block_minmax = nd;
    c->block_minmax_capacity = ncap;
  }
  K1Args a{};
  a.tid = b.tid; a.pos = b.pos; a.flag = b.flag; a.mapq = b.mapq; a.nm_state = b.nm_state; a.nm = b.nm;
  a.l_seq = b.l_seq; a.aligned = b.aligned; a.del = b.del; a.ins = b.ins; a.iv_begin = b.iv_begin;
  a.iv_start = b.iv_start; a.iv_len = b.iv_len;
  a.n = n_records;
  a.off_span = c->d_off_span; a.len = c->d_len;
  a.n_contigs = c->n_contigs; a.tid_begin = c->tid_begin; a.tid_end = c->tid_end;
  a.arena = c->d_arena; a.tail_sum = c->d_tail_sum; a.rows = c->d_rows;
  a.block_minmax = c->d_block_minmax + c->block_minmax_used;
  a.error_flags = c->d_counters + 0;
  a.p = c->params;
  a.filter_single = c->mode.filter_single_reads;
  a.filter_pairs = c->mode.filter_pairs;
  if (c->k1_events_used == c->k1_events.size()) {
    cudaEvent_t e0, e1;
    CU_TRY(c, cudaEventCreate(&e0));
    CU_TRY(c, cudaEventCreate(&e1));
    c->k1_events.emplace_back(e0, e1);
  }
  auto& ev = c->k1_events[c->k1_events_used++];
  CU_TRY(c, cudaEventRecord(ev.first, c->stream));
  k1_filter_accumulate<<<blocks, K1_THREADS, 0, c->stream>>>(a);
  CU_TRY(c, cudaGetLastError());
  CU_TRY(c, cudaEventRecord(ev.second, c->stream));
  c->block_minmax_used += blocks;
  c->n_records += n_records;
  c->n_intervals += n_intervals;
  c->timing.k1_launches += 1;
  return CMB_OK;
}

template <bool HIST, bool CLEAN>
int launch_k2_variant(cmb_ctx* c, const K2Args& a) {
  auto kern = k2_scan_reduce<HIST, CLEAN>;
  CU_TRY(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K2_SMEM_BYTES));
  int occ = 0;
  CU_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, (int)K2_THREADS, K2_SMEM_BYTES));
  if (occ < 1) return fail(c, CMB_E_CUDA, "k2_scan_reduce does not fit on an SM");
  const uint32_t grid = std::min<uint32_t>(c->n_chunks, (uint32_t)(occ * c->sm_count));
  kern<<<grid, K2_THREADS, K2_SMEM_BYTES, c->stream>>>(c->tmap, a);
  CU_TRY(c, cudaGetLastError());
  return CMB_OK;
}

int run_end_of_sample(cmb_ctx* c) {
  const bool hist = c->params.want & (CMB_WANT_HIST | CMB_WANT_HIST_CSR);
  const bool csr = c->params.want & CMB_WANT_HIST_CSR;
  const uint32_t excl = (uint32_t)std::min<uint64_t>(c->params.contig_end_exclusion, 0x7fffffffu);
  CU_TRY(c, cudaEventRecord(c->ev[2], c->stream));
  if (c->block_minmax_used) {
    k1c_check_sorted<<<1, 1024, 0, c->stream>>>(c->d_block_minmax, c->block_minmax_used, c->d_counters + 0);
    CU_TRY(c, cudaGetLastError());
  }
  k1b_chunk_carry<<<1, 1024, 0, c->stream>>>(c->d_tail_sum, c->d_chunk_first, c->d_off_span, c->n_chunks, c->d_carry_in);
  CU_TRY(c, cudaGetLastError());
  K2Args a{};
  a.off_span = c->d_off_span; a.len = c->d_len; a.chunk_first = c->d_chunk_first; a.carry_in = c->d_carry_in;
  a.rows = c->d_rows; a.tid_begin = c->tid_begin; a.n_local = c->n_local; a.n_chunks = c->n_chunks; a.excl = excl;
  a.ticket = c->d_counters + 1; a.arena = c->d_arena;
  a.rec = c->d_rec; a.rec_capacity = c->rec_capacity; a.rec_count = c->d_counters + 2;
  a.warp_table = c->d_warp_table; a.ovf = c->d_ovf; a.ovf_capacity = c->ovf_capacity; a.ovf_count = c->d_counters + 3;
  a.error_flags = c->d_counters + 0;
  CU_TRY(c, cudaEventRecord(c->ev[3], c->stream));
  int rc;
  if (hist) rc = c->clean_as_you_go ? launch_k2_variant<true, true>(c, a) : launch_k2_variant<true, false>(c, a);
  else rc = c->clean_as_you_go ? launch_k2_variant<false, true>(c, a) : launch_k2_variant<false, false>(c, a);
  if (rc) return rc;
  c->timing.k2_launches = 1;
  c->arena_dirty = !c->clean_as_you_go;
  CU_TRY(c, cudaEventRecord(c->ev[4], c->stream));
  if (hist) {
    K3Args k{};
    k.off_span = c->d_off_span; k.len = c->d_len; k.chunk_first = c->d_chunk_first; k.rows = c->d_rows;
    k.tid_begin = c->tid_begin; k.n_local = c->n_local; k.excl = excl;
    k.trim_min = c->params.trim_min; k.trim_max = c->params.trim_max;
    k.rec = c->d_rec; k.warp_table = c->d_warp_table; k.ovf = c->d_ovf; k.ovf_count = c->d_counters + 3;
    k.ovf_capacity = c->ovf_capacity;
    k.pairs = c->d_pairs; k.pair_count = (unsigned long long*)(c->d_counters + 4); k.pair_capacity = c->pair_capacity;
    k.want_csr = csr; k.error_flags = c->d_counters + 0;
    const uint32_t grid = std::min<uint32_t>(c->n_local, (uint32_t)c->sm_count * 64u);
    k3_finalize<<<grid, K3_THREADS, 0, c->stream>>>(k);
    CU_TRY(c, cudaGetLastError());
    c->timing.k3_launches = 1;
  }
  CU_TRY(c, cudaEventRecord(c->ev[5], c->stream));
  return CMB_OK;
}

int collect_errors_and_timing(cmb_ctx* c, uint32_t* counters_out) {
  uint32_t h[6];
  CU_TRY(c, cudaMemcpyAsync(h, c->d_counters, sizeof h, cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(c, cudaEventRecord(c->ev[6], c->stream));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  memcpy(counters_out, h, sizeof h);
  float ms = 0;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]); c->timing.ms_zero = ms;
  cudaEventElapsedTime(&ms, c->ev[3], c->ev[4]); c->timing.ms_scan = ms;
  cudaEventElapsedTime(&ms, c->ev[4], c->ev[5]); c->timing.ms_finalize = ms;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[6]); c->timing.ms_total = ms;
  float acc = 0;
  for (uint32_t i = 0; i < c->k1_events_used; ++i) {
    cudaEventElapsedTime(&ms, c->k1_events[i].first, c->k1_events[i].second);
    acc += ms;
  }
  cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]);  // k1c + k1b
  c->timing.ms_accumulate = acc + ms;
  c->timing.arena_elems = c->arena_elems;
  c->timing.n_records = c->n_records;
  c->timing.n_intervals = c->n_intervals;
  const uint32_t e = h[0];
  if (e) c->arena_dirty = true;
  if (e & ERR_UNSORTED)
    return fail(c, CMB_E_UNSORTED, "BAM file appears to be unsorted. Input BAM files must be sorted by reference (i.e. by samtools sort)");
  if (e & ERR_NM)
    return fail(c, CMB_E_NM, "Mapping record encountered that does not have an 'NM' auxiliary tag in the SAM/BAM format. This is required to work out some coverage statistics");
  if (e & (ERR_BOUNDS | ERR_TID)) return fail(c, CMB_E_BOUNDS, "index out of bounds: an aligned block starts beyond the end of its reference sequence");
  if (e & ERR_CAPACITY) return fail(c, CMB_E_CAPACITY, "device histogram record buffer overflowed");
  if (e & ERR_INTERNAL) return fail(c, CMB_E_CUDA, "internal error: negative running depth (inconsistent delta arena)");
  return CMB_OK;
}

}  // namespace

// ================================================================================================ C ABI
extern "C" {

int cmb_abi_version(void) { return CMB_ABI_VERSION; }

const char* cmb_last_error(const cmb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int cmb_create(const cmb_device_cfg* cfg, cmb_ctx** out) {
  if (!cfg || !out) return fail(nullptr, CMB_E_ARG, "cmb_create: null argument");
  *out = nullptr;
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0)
    return fail(nullptr, CMB_E_CUDA, "cmb_create: no usable CUDA device (%s); libcoverm_b200 has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  if (cfg->device < 0 || cfg->device >= n_dev) return fail(nullptr, CMB_E_ARG, "cmb_create: device %d out of range", cfg->device);
  cmb_ctx* c = new cmb_ctx();
  c->device = cfg->device;
  c->cfg = *cfg;
  if (c->cfg.batch_records == 0) c->cfg.batch_records = 1u << 20;
  if (c->cfg.batch_intervals == 0) c->cfg.batch_intervals = c->cfg.batch_records + c->cfg.batch_records / 2;
  if (c->cfg.n_staging < 2) c->cfg.n_staging = 2;
  auto bail = [&](int code) {
    g_create_error = c->err;
    cmb_destroy(c);
    return code;
  };
#define CREATE_TRY(expr)                                                                           \
  do {                                                                                             \
    cudaError_t e_ = (expr);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      fail(c, CMB_E_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_));                         \
      return bail(e_ == cudaErrorMemoryAllocation ? CMB_E_NOMEM : CMB_E_CUDA);                     \
    }                                                                                              \
  } while (0)
  CREATE_TRY(cudaSetDevice(c->device));
  cudaDeviceProp prop;
  CREATE_TRY(cudaGetDeviceProperties(&prop, c->device));
  if (prop.major < 10) {
    fail(c, CMB_E_CUDA, "cmb_create: device %d is sm_%d%d; this library is built for sm_100a only", c->device, prop.major, prop.minor);
    return bail(CMB_E_CUDA);
  }
  c->sm_count = prop.multiProcessorCount;
  CREATE_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  for (auto& ev : c->ev) CREATE_TRY(cudaEventCreate(&ev));
  size_t offs[13];
  const size_t slab = batch_slab_bytes(c->cfg.batch_records, c->cfg.batch_intervals, offs);
  for (uint32_t i = 0; i < c->cfg.n_staging; ++i) {
    void* h = nullptr;
    CREATE_TRY(cudaHostAlloc(&h, slab, cudaHostAllocDefault));
    c->host_slab.push_back(h);
    cmb_read_batch hb;
    carve_batch(h, c->cfg.batch_records, c->cfg.batch_intervals, &hb);
    c->host_batch.push_back(hb);
    DevBatch db;
    CREATE_TRY(cudaMalloc(&db.slab, slab));
    carve_batch(db.slab, c->cfg.batch_records, c->cfg.batch_intervals, &db.ptr);
    c->dev_batch.push_back(db);
    cudaEvent_t ev;
    CREATE_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    c->batch_done.push_back(ev);
    c->batch_busy.push_back(false);
  }
  CREATE_TRY(cudaMalloc(&c->d_counters, 64));
  CREATE_TRY(cudaMemset(c->d_counters, 0, 64));
  c->block_minmax_capacity = 1u << 16;
  CREATE_TRY(cudaMalloc(&c->d_block_minmax, sizeof(int2) * (size_t)c->block_minmax_capacity));
  const char* env = getenv("CMB_CLEAN_AS_YOU_GO");
  if (env && env[0] == '0') c->clean_as_you_go = false;
  *out = c;
  return CMB_OK;
}

void cmb_destroy(cmb_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  free_reference(c);
  for (auto h : c->host_slab) cudaFreeHost(h);
  for (auto& d : c->dev_batch) cudaFree(d.slab);
  for (auto e : c->batch_done) cudaEventDestroy(e);
  for (auto& e : c->k1_events) {
    cudaEventDestroy(e.first);
    cudaEventDestroy(e.second);
  }
  for (auto e : c->ev)
    if (e) cudaEventDestroy(e);
  cudaFree(c->d_counters);
  cudaFree(c->d_block_minmax);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int cmb_set_reference(cmb_ctx* c, uint32_t n_contigs, const uint64_t* contig_len, uint32_t tid_begin, uint32_t tid_end) {
  if (!c || !contig_len || tid_begin > tid_end || tid_end > n_contigs) return fail(c, CMB_E_ARG, "cmb_set_reference: bad arguments");
  if (c->in_sample) return fail(c, CMB_E_ARG, "cmb_set_reference: a sample is in progress");
  CU_TRY(c, cudaSetDevice(c->device));
  free_reference(c);
  c->n_contigs = n_contigs;
  c->tid_begin = tid_begin;
  c->tid_end = tid_end;
  c->n_local = tid_end - tid_begin;
  std::vector<uint32_t> off_span(c->n_local + 1), len(c->n_local);
  uint64_t spans = 0;
  for (uint32_t i = 0; i < c->n_local; ++i) {
    const uint64_t L = contig_len[tid_begin + i];
    if (L > 0x7fffffffull) return fail(c, CMB_E_ARG, "cmb_set_reference: contig %u longer than 2^31-1", tid_begin + i);
    off_span[i] = (uint32_t)spans;
    len[i] = (uint32_t)L;
    spans += std::max<uint64_t>(1, (L + SPAN - 1) / SPAN);
    if (spans > 0xfffffff0ull) return fail(c, CMB_E_ARG, "cmb_set_reference: shard larger than 2^36 bases; use more shards");
  }
  off_span[c->n_local] = (uint32_t)spans;
  const uint64_t chunks = std::max<uint64_t>(1, (spans + CHUNK_SPANS - 1) / CHUNK_SPANS);
  c->n_chunks = (uint32_t)chunks;
  c->arena_elems = chunks * CHUNK;
  std::vector<uint32_t> chunk_first(c->n_chunks + 1);
  {
    uint32_t ci = 0;
    for (uint32_t k = 0; k < c->n_chunks; ++k) {
      const uint64_t s = (uint64_t)k * CHUNK_SPANS;
      while (ci + 1 < c->n_local && off_span[ci + 1] <= s) ++ci;
      chunk_first[k] = ci;
    }
    chunk_first[c->n_chunks] = c->n_local ? c->n_local - 1 : 0;
  }
  if (c->n_local == 0) {  // empty shard: nothing to allocate beyond the rows
    CU_TRY(c, cudaMalloc(&c->d_rows, sizeof(cmb_contig_stats) * std::max<size_t>(1, n_contigs)));
    return CMB_OK;
  }
  CU_TRY(c, cudaMalloc(&c->d_arena, c->arena_elems * 4));
  CU_TRY(c, cudaMalloc(&c->d_off_span, 4ull * (c->n_local + 1)));
  CU_TRY(c, cudaMalloc(&c->d_len, 4ull * c->n_local));
  CU_TRY(c, cudaMalloc(&c->d_chunk_first, 4ull * (c->n_chunks + 1)));
  CU_TRY(c, cudaMalloc(&c->d_tail_sum, 4ull * c->n_chunks));
  CU_TRY(c, cudaMalloc(&c->d_carry_in, 4ull * c->n_chunks));
  CU_TRY(c, cudaMalloc(&c->d_rows, sizeof(cmb_contig_stats) * (size_t)n_contigs));
  CU_TRY(c, cudaMemcpyAsync(c->d_off_span, off_span.data(), 4ull * (c->n_local + 1), cudaMemcpyHostToDevice, c->stream));
  CU_TRY(c, cudaMemcpyAsync(c->d_len, len.data(), 4ull * c->n_local, cudaMemcpyHostToDevice, c->stream));
  CU_TRY(c, cudaMemcpyAsync(c->d_chunk_first, chunk_first.data(), 4ull * (c->n_chunks + 1), cudaMemcpyHostToDevice, c->stream));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  // histogram record buffers: one 8 B record per 8 arena elements is far above anything a real sample produces
  c->rec_capacity = (uint32_t)std::min<uint64_t>(0xfffffff0ull, std::max<uint64_t>(1u << 20, c->arena_elems / 8));
  c->ovf_capacity = (uint32_t)std::min<uint64_t>(1u << 26, std::max<uint64_t>(1u << 20, c->arena_elems / 64));
  CU_TRY(c, cudaMalloc(&c->d_rec, 8ull * c->rec_capacity));
  CU_TRY(c, cudaMalloc(&c->d_warp_table, 8ull * c->n_chunks * K2_WARPS));
  CU_TRY(c, cudaMalloc(&c->d_ovf, 16ull * c->ovf_capacity));
  c->arena_dirty = true;
  // TMA descriptor: the arena as [rows][32] i32, box = one chunk (256 rows x 128 B), 128B swizzle
  PFN_encodeTiled encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CU_TRY(c, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
  if (!encode || qres != cudaDriverEntryPointSuccess) return fail(c, CMB_E_CUDA, "cuTensorMapEncodeTiled not available in this driver");
  cuuint64_t gdim[2] = {ROW_ELEMS, c->arena_elems / ROW_ELEMS};
  cuuint64_t gstride[1] = {ROW_ELEMS * 4};
  cuuint32_t box[2] = {ROW_ELEMS, CHUNK_ROWS};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(&c->tmap, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, c->d_arena, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(c, CMB_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return CMB_OK;
}

int cmb_set_params(cmb_ctx* c, const cmb_params* p, cmb_filter_mode* mode_out) {
  if (!c || !p) return fail(c, CMB_E_ARG, "cmb_set_params: null argument");
  if (c->in_sample) return fail(c, CMB_E_ARG, "cmb_set_params: a sample is in progress");
  c->params = *p;
  // filter.rs:48-61
  const bool single_initial = p->min_aligned_length_single > 0 || p->min_percent_identity_single > 0.0f || p->min_aligned_percent_single > 0.0f;
  const bool pairs_initial = p->min_aligned_length_pair > 0 || p->min_percent_identity_pair > 0.0f || p->min_aligned_percent_pair > 0.0f;
  const bool fs = single_initial || (!pairs_initial && p->min_mapq != 255);
  const bool fp = pairs_initial || ((!fs || !p->include_improper_pairs) && p->min_mapq != 255);
  c->mode.filter_single_reads = p->filtering ? fs : 0;
  c->mode.filter_pairs = p->filtering ? fp : 0;
  if (mode_out) *mode_out = c->mode;
  c->have_params = true;
  return CMB_OK;
}

int cmb_begin_sample(cmb_ctx* c) {
  if (!c) return CMB_E_ARG;
  if (!c->d_rows || !c->have_params) return fail(c, CMB_E_ARG, "cmb_begin_sample: set_reference and set_params first");
  if (c->in_sample) return fail(c, CMB_E_ARG, "cmb_begin_sample: previous sample not ended");
  CU_TRY(c, cudaSetDevice(c->device));
  c->timing = cmb_sample_timing{};
  c->k1_events_used = 0;
  c->block_minmax_used = 0;
  c->n_records = c->n_intervals = 0;
  CU_TRY(c, cudaEventRecord(c->ev[0], c->stream));
  if (c->n_local) {
    if (c->arena_dirty) CU_TRY(c, cudaMemsetAsync(c->d_arena, 0, c->arena_elems * 4, c->stream));
    CU_TRY(c, cudaMemsetAsync(c->d_tail_sum, 0, 4ull * c->n_chunks, c->stream));
  }
  CU_TRY(c, cudaMemsetAsync(c->d_rows, 0, sizeof(cmb_contig_stats) * (size_t)c->n_contigs, c->stream));
  CU_TRY(c, cudaMemsetAsync(c->d_counters, 0, 64, c->stream));
  CU_TRY(c, cudaEventRecord(c->ev[1], c->stream));
  c->arena_dirty = true;  // until K2 has cleaned it
  if ((c->params.want & CMB_WANT_HIST_CSR) && c->n_local) {
    const uint64_t want = std::max<uint64_t>(1u << 20, c->arena_elems / 16);
    if (c->pair_capacity < want) {
      cudaFree(c->d_pairs);
      c->d_pairs = nullptr;
      c->pair_capacity = 0;
      CU_TRY(c, cudaMalloc(&c->d_pairs, sizeof(cmb_hist_pair) * want));
      c->pair_capacity = want;
    }
  }
  c->in_sample = true;
  c->ended = false;
  c->acquired = -1;
  return CMB_OK;
}

int cmb_acquire_batch(cmb_ctx* c, cmb_read_batch* batch) {
  if (!c || !batch) return fail(c, CMB_E_ARG, "cmb_acquire_batch: null argument");
  if (!c->in_sample) return fail(c, CMB_E_ARG, "cmb_acquire_batch: no sample in progress");
  if (c->acquired >= 0) return fail(c, CMB_E_ARG, "cmb_acquire_batch: previous batch not submitted");
  const uint32_t i = c->next_batch;
  if (c->batch_busy[i]) {
    CU_TRY(c, cudaEventSynchronize(c->batch_done[i]));
    c->batch_busy[i] = false;
  }
  *batch = c->host_batch[i];
  c->acquired = (int)i;
  return CMB_OK;
}

int cmb_submit_batch(cmb_ctx* c, uint32_t n_records, uint32_t n_intervals) {
  if (!c) return CMB_E_ARG;
  if (!c->in_sample || c->acquired < 0) return fail(c, CMB_E_ARG, "cmb_submit_batch: no acquired batch");
  if (n_records > c->cfg.batch_records || n_intervals > c->cfg.batch_intervals) return fail(c, CMB_E_ARG, "cmb_submit_batch: batch exceeds capacity");
  const uint32_t i = (uint32_t)c->acquired;
  c->acquired = -1;
  c->next_batch = (i + 1) % c->cfg.n_staging;
  if (n_records == 0) return CMB_OK;
  if (c->n_local == 0) return CMB_OK;
  const cmb_read_batch& h = c->host_batch[i];
  const cmb_read_batch& d = c->dev_batch[i].ptr;
  CU_TRY(c, cudaSetDevice(c->device));
#define H2D(col, bytes) CU_TRY(c, cudaMemcpyAsync(d.col, h.col, (bytes), cudaMemcpyHostToDevice, c->stream))
  H2D(tid, 4ull * n_records);
  H2D(pos, 4ull * n_records);
  H2D(nm, 4ull * n_records);
  H2D(l_seq, 4ull * n_records);
  H2D(aligned, 4ull * n_records);
  H2D(del, 4ull * n_records);
  H2D(ins, 4ull * n_records);
  H2D(iv_begin, 4ull * (n_records + 1));
  if (n_intervals) {
    H2D(iv_start, 4ull * n_intervals);
    H2D(iv_len, 4ull * n_intervals);
  }
  H2D(flag, 2ull * n_records);
  H2D(mapq, 1ull * n_records);
  H2D(nm_state, 1ull * n_records);
#undef H2D
  int rc = launch_k1(c, d, n_records, n_intervals);
  if (rc) return rc;
  CU_TRY(c, cudaEventRecord(c->batch_done[i], c->stream));
  c->batch_busy[i] = true;
  return CMB_OK;
}

int cmb_submit_device_batch(cmb_ctx* c, const cmb_read_batch* dev, uint32_t n_records, uint32_t n_intervals) {
  if (!c || !dev) return fail(c, CMB_E_ARG, "cmb_submit_device_batch: null argument");
  if (!c->in_sample) return fail(c, CMB_E_ARG, "cmb_submit_device_batch: no sample in progress");
  if (c->n_local == 0) return CMB_OK;
  CU_TRY(c, cudaSetDevice(c->device));
  return launch_k1(c, *dev, n_records, n_intervals);
}

int cmb_end_sample_device(cmb_ctx* c, const cmb_contig_stats** dev_stats) {
  if (!c) return CMB_E_ARG;
  if (!c->in_sample) return fail(c, CMB_E_ARG, "cmb_end_sample: no sample in progress");
  if (c->acquired >= 0) return fail(c, CMB_E_ARG, "cmb_end_sample: an acquired batch was not submitted");
  CU_TRY(c, cudaSetDevice(c->device));
  c->in_sample = false;
  if (c->n_local) {
    int rc = run_end_of_sample(c);
    if (rc) return rc;
  } else {
    for (int i = 2; i <= 5; ++i) CU_TRY(c, cudaEventRecord(c->ev[i], c->stream));
  }
  uint32_t counters[6];
  int rc = collect_errors_and_timing(c, counters);
  if (rc) return rc;
  c->ended = true;
  if (dev_stats) *dev_stats = c->d_rows;
  return CMB_OK;
}

int cmb_end_sample(cmb_ctx* c, cmb_contig_stats* stats, cmb_hist_pair* pairs, uint64_t pairs_capacity, uint64_t* n_pairs) {
  if (!c || !stats) return fail(c, CMB_E_ARG, "cmb_end_sample: null argument");
  int rc = cmb_end_sample_device(c, nullptr);
  if (rc) return rc;
  CU_TRY(c, cudaMemcpyAsync(stats, c->d_rows, sizeof(cmb_contig_stats) * (size_t)c->n_contigs, cudaMemcpyDeviceToHost, c->stream));
  uint64_t np = 0;
  if ((c->params.want & CMB_WANT_HIST_CSR) && c->n_local) {
    unsigned long long cnt = 0;
    CU_TRY(c, cudaMemcpyAsync(&cnt, c->d_counters + 4, 8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    np = cnt;
    if (np > c->pair_capacity) return fail(c, CMB_E_CAPACITY, "device histogram pair buffer overflowed");
    if (pairs) {
      if (np > pairs_capacity) return fail(c, CMB_E_CAPACITY, "cmb_end_sample: caller's pair buffer too small (%llu needed)", (unsigned long long)np);
      if (np) CU_TRY(c, cudaMemcpyAsync(pairs, c->d_pairs, sizeof(cmb_hist_pair) * np, cudaMemcpyDeviceToHost, c->stream));
    }
  }
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  if (n_pairs) *n_pairs = np;
  return CMB_OK;
}

int cmb_fetch_pairs(cmb_ctx* c, cmb_hist_pair* pairs, uint64_t n_pairs) {
  if (!c || (!pairs && n_pairs)) return fail(c, CMB_E_ARG, "cmb_fetch_pairs: null argument");
  if (!c->ended) return fail(c, CMB_E_ARG, "cmb_fetch_pairs: no ended sample");
  if (n_pairs > c->pair_capacity) return fail(c, CMB_E_ARG, "cmb_fetch_pairs: more pairs requested than produced");
  if (n_pairs) {
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, cudaMemcpyAsync(pairs, c->d_pairs, sizeof(cmb_hist_pair) * n_pairs, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
  }
  return CMB_OK;
}

int cmb_get_timing(const cmb_ctx* c, cmb_sample_timing* out) {
  if (!c || !out) return CMB_E_ARG;
  *out = c->timing;
  return CMB_OK;
}

void* cmb_stream(cmb_ctx* c) { return c ? (void*)c->stream : nullptr; }

}  // extern "C"
