// K1b: running depth at the first element of every chunk (carry_in), from the per-chunk tail sums K1 accumulated.
//
//   x_0 = 0;   x_{k+1} = mid_k ? ((same_k ? x_k : 0) + tail_sum[k]) : 0
//   mid_k  = chunk k+1 starts in the middle of a contig,  same_k = that contig also owns the first span of chunk k.
// A segmented scan over ~L/8192 elements, done in two tiny launches: k1b_local scans 1024-chunk blocks (thread = 4
// consecutive chunks, warp shuffles, one shared-memory step) and leaves a "needs the block carry" flag in tail_sum;
// k1b_apply folds in the block carries.  With the carries known up front K2 needs no inter-CTA look-back at all.
#pragma once

constexpr uint32_t K1B_THREADS = 256;
constexpr uint32_t K1B_PER = 4;
constexpr uint32_t K1B_BLOCK = K1B_THREADS * K1B_PER;  // chunks per block

struct K1bElem {
  bool reset;
  int add;
};
__device__ __forceinline__ K1bElem k1b_elem(uint32_t k, uint32_t n_chunks, const int32_t* tail_sum, const uint32_t* chunk_first,
                                           const uint32_t* off_span) {
  K1bElem e;
  if (k + 1 >= n_chunks) {
    e.reset = true;
    e.add = 0;
    return e;
  }
  const uint32_t cn = chunk_first[k + 1];
  const bool mid = off_span[cn] < (k + 1) * CHUNK_SPANS;
  const bool same = chunk_first[k] == cn;
  e.reset = !(mid && same);
  e.add = mid ? tail_sum[k] : 0;
  return e;
}

__global__ void __launch_bounds__(K1B_THREADS) k1b_local(int32_t* tail_sum, const uint32_t* chunk_first, const uint32_t* off_span,
                                                        uint32_t n_chunks, int32_t* carry_in, int2* block_agg) {
  __shared__ int2 s_w[K1B_THREADS / 32];
  const uint32_t t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const uint32_t k0 = blockIdx.x * K1B_BLOCK + t * K1B_PER;
  K1bElem el[K1B_PER];
  int val = 0, flg = 0;
#pragma unroll
  for (uint32_t i = 0; i < K1B_PER; ++i) {
    const uint32_t k = k0 + i;
    if (k < n_chunks) {
      el[i] = k1b_elem(k, n_chunks, tail_sum, chunk_first, off_span);
    } else {
      el[i].reset = false;
      el[i].add = 0;
    }
    if (el[i].reset) {
      val = el[i].add;
      flg = 1;
    } else {
      val += el[i].add;
    }
  }
  // inclusive segmented scan of (flg, val) over the threads of the block
  int v = val, f = flg;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int ov = __shfl_up_sync(FULL, v, d), of = __shfl_up_sync(FULL, f, d);
    if ((int)lane >= d) {
      if (!f) v += ov;
      f |= of;
    }
  }
  int pv = __shfl_up_sync(FULL, v, 1), pf = __shfl_up_sync(FULL, f, 1);
  if (lane == 0) {
    pv = 0;
    pf = 0;
  }
  if (lane == 31) s_w[warp] = make_int2(v, f);
  __syncthreads();
  int wv = 0, wf = 0;  // exclusive over the preceding warps
  for (uint32_t w = 0; w < warp; ++w) {
    const int2 x = s_w[w];
    wv = x.y ? x.x : wv + x.x;
    wf |= x.y;
  }
  // exclusive prefix for this thread: preceding warps, then preceding lanes of this warp
  int ev = pf ? pv : wv + pv, ef = pf | wf;
  if (t == K1B_THREADS - 1) {
    const int bv = f ? v : wv + v, bf = f | wf;
    block_agg[blockIdx.x] = make_int2(bv, bf);
  }
  int x = ev, xf = ef;
#pragma unroll
  for (uint32_t i = 0; i < K1B_PER; ++i) {
    const uint32_t k = k0 + i;
    if (k < n_chunks) {
      carry_in[k] = x;         // exact if a reset precedes it inside the block, else missing the block carry
      tail_sum[k] = xf ? 0 : 1;  // 1 = "add the block carry"
    }
    if (el[i].reset) {
      x = el[i].add;
      xf = 1;
    } else {
      x += el[i].add;
    }
  }
}

__global__ void __launch_bounds__(K1B_THREADS) k1b_apply(const int32_t* needs, const int2* block_agg, uint32_t n_chunks, int32_t* carry_in) {
  __shared__ int s_carry;
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int j = (int)blockIdx.x - 1; j >= 0; --j) {  // walk back to the nearest block that contains a reset
      const int2 a = block_agg[j];
      acc += a.x;
      if (a.y) break;
    }
    s_carry = acc;
  }
  __syncthreads();
  const int bc = s_carry;
  if (bc == 0) return;
#pragma unroll
  for (uint32_t i = 0; i < K1B_PER; ++i) {
    const uint32_t k = blockIdx.x * K1B_BLOCK + threadIdx.x * K1B_PER + i;
    if (k < n_chunks && needs[k]) carry_in[k] += bc;
  }
}
