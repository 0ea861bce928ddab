// kd_inflate_g8: DEFLATE with FOUR streams per warp (8 lanes each), written warp-synchronously: every statement is
// executed by all 32 lanes and per-stream predicates select who takes effect, so one issued instruction advances four
// BGZF blocks.  Compared with kd_inflate (one stream per warp) the Huffman tables are smaller (9-bit literal/length root,
// 8-bit distance root, 3.3 KB per stream) so that 64 streams fit per SM instead of 32.  Same contract as kd_inflate:
// a block is either inflated and CRC-checked or declined (status != 0) for the host's zlib.
#pragma once

constexpr uint32_t G8 = 8;                      // lanes per stream
constexpr uint32_t G8_STREAMS = 32 / G8;        // streams per warp
constexpr uint32_t G8_WARPS = 8;                // warps per CTA
constexpr uint32_t G8_LROOT = 9, G8_DROOT = 8;
constexpr uint32_t G8_LIT_ENTRIES = (1u << G8_LROOT) + 384;  // zlib's ENOUGH(286, 9, 15) = 852
constexpr uint32_t G8_DST_ENTRIES = (1u << G8_DROOT) + 192;  // ENOUGH(30, 8, 15) <= 402
constexpr uint32_t G8_SUBQ = 64;

struct G8Smem {
  uint16_t lit[G8_LIT_ENTRIES];
  uint16_t dst[G8_DST_ENTRIES];
  uint8_t lens[320];
  uint32_t nc[16];
  uint32_t subq[G8_SUBQ];
  uint32_t win[16];  // compressed-byte window of the bit reader (two 32-byte lines)
  uint32_t overflow;
  uint32_t pad[3];
};
struct G8Consts {
  uint16_t len_base[32], dist_base[32];
  uint8_t len_extra[32], dist_extra[32];
};
constexpr uint32_t G8_SMEM_BYTES = G8_WARPS * G8_STREAMS * sizeof(G8Smem) + INF_CRC_TABLE_BYTES + sizeof(G8Consts);

// Per-stream bit reader.  The compressed bytes are staged in a 64-byte shared-memory window (two 32-byte lines) filled
// with cp.async.cg (L2 only: the copy engine is still writing the file while this kernel runs), one line ahead of use, so
// no register ever waits on a global load.  Every member is uniform across the 8 lanes of a stream.
__device__ __forceinline__ void g8_cp_async16(uint32_t smem_addr, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(g) : "memory");
}
__device__ __forceinline__ void g8_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void g8_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

struct G8Reader {
  const uint8_t* base;  // 32-byte aligned
  uint32_t* win;        // this stream's window: 16 words
  uint32_t widx;        // next word to take
  uint64_t buf;
  uint32_t cnt;
  // line `l` (32 bytes at base + 32 l) into half (l & 1) of the window: lanes 0 and 1 of the stream copy 16 bytes each
  __device__ __forceinline__ void fetch_line(uint32_t l, uint32_t glane) {
    if (glane < 2) g8_cp_async16(smem_u32(win + (l & 1) * 8 + glane * 4), base + (uint64_t)l * 32 + glane * 16);
    g8_cp_async_commit();
  }
  // executed by all lanes; stream `p` takes a word when it has room
  __device__ __forceinline__ void refill(bool p, uint32_t glane, uint32_t gmask) {
    const uint32_t w = win[widx & 15];
    if (p && cnt <= 32) {
      buf |= (uint64_t)w << cnt;
      cnt += 32;
      ++widx;
      if ((widx & 7) == 0) {  // entering a new line: it was requested a line ago; request the one after it
        g8_cp_async_wait_all();
        __syncwarp(gmask);
        fetch_line((widx >> 3) + 1, glane);
      }
    }
  }
  __device__ __forceinline__ void init(bool p, const uint8_t* ptr, uint32_t glane, uint32_t gmask) {
    uint32_t drop = 0;
    if (p) {
      const uintptr_t a = (uintptr_t)ptr;
      base = (const uint8_t*)(a & ~(uintptr_t)31);
      const uint32_t skip = (uint32_t)(a & 31);
      widx = skip >> 2;
      drop = (skip & 3) * 8;
      buf = 0;
      cnt = 0;
      __syncwarp(gmask);  // nobody of this stream still reads the old window
      fetch_line(0, glane);
      fetch_line(1, glane);
      g8_cp_async_wait_all();
      __syncwarp(gmask);
    }
    refill(p, glane, gmask);  // cnt == 0 -> takes one word
    if (p) {
      buf >>= drop;
      cnt -= drop;
    }
    refill(p, glane, gmask);
  }
  __device__ __forceinline__ void consume(uint32_t n) {
    buf >>= n;
    cnt -= n;
  }
  __device__ __forceinline__ uint32_t bits(uint32_t n) const { return (uint32_t)buf & ((1u << n) - 1); }
  __device__ __forceinline__ const uint8_t* byte_pos() const { return base + (((uint64_t)widx * 32 - cnt) >> 3); }
  __device__ __forceinline__ const uint8_t* byte_pos_ceil() const { return base + (((uint64_t)widx * 32 - cnt + 7) >> 3); }
};

// Canonical Huffman tables for the four streams of a warp at once.  Stream parameters: lens/n/tab (n == 0: this stream
// is not building).  Returns per stream: true = table built.
__device__ bool g8_build_table(G8Smem& S, const uint8_t* lens, uint32_t n, uint16_t* tab, uint32_t root, uint32_t n_entries, uint32_t lane) {
  const uint32_t glane = lane & (G8 - 1), gid = lane / G8;
  if (n)
    for (uint32_t i = glane; i < n_entries / 2; i += G8) reinterpret_cast<uint32_t*>(tab)[i] = 0;
  const uint32_t nmax = __reduce_max_sync(FULL, n);
  // lane g counts the codes of length g and g + 8
  uint32_t cnt0 = 0, cnt1 = 0;
  for (uint32_t s = 0; s < nmax; ++s) {
    const uint32_t v = s < n ? lens[s] : 0xffu;
    cnt0 += (v == glane) ? 1u : 0u;
    cnt1 += (v == glane + G8) ? 1u : 0u;
  }
  if (glane == 0) cnt0 = 0;
  uint32_t code = 0, f0 = 0, f1 = 0;
  int left = 1;
  bool over = false;
  for (uint32_t L = 1; L <= 15; ++L) {
    const uint32_t pa = __shfl_sync(FULL, cnt0, (L - 1) & 7, G8), pb = __shfl_sync(FULL, cnt1, (L - 1) & 7, G8);
    const uint32_t ca = __shfl_sync(FULL, cnt0, L & 7, G8), cb = __shfl_sync(FULL, cnt1, L & 7, G8);
    code = (code + ((L - 1) < 8 ? pa : pb)) << 1;
    if (glane == (L & 7)) {
      if (L < 8) f0 = code;
      else f1 = code;
    }
    left = (left << 1) - (int)(L < 8 ? ca : cb);
    if (left < 0) over = true;
  }
  const uint32_t pr = root + 1;
  const uint32_t P0a = __shfl_sync(FULL, f0, pr & 7, G8), P0b = __shfl_sync(FULL, f1, pr & 7, G8);
  const uint32_t P0 = (pr < 8 ? P0a : P0b) >> 1;
  S.nc[glane] = f0;
  S.nc[glane + G8] = f1;
  for (uint32_t q = glane; q < G8_SUBQ; q += G8) S.subq[q] = 0;
  if (glane == 0) S.overflow = 0;
  __syncwarp();
  const uint32_t root_size = 1u << root;
  bool any_long = false;
  for (uint32_t base = 0; base < nmax; base += G8) {
    const uint32_t s = base + glane;
    const uint32_t L = s < n ? lens[s] : 0;
    const uint32_t mask = __match_any_sync(FULL, L | (gid << 8));
    const uint32_t rank = __popc(mask & ((1u << lane) - 1));
    const uint32_t leader = __ffs(mask) - 1;
    const uint32_t c0 = S.nc[L & 15];
    __syncwarp();
    if (lane == leader && L) S.nc[L] = c0 + __popc(mask);
    __syncwarp();
    const uint32_t cd = c0 + rank;
    if (L && L <= root) {
      const uint32_t rev = __brev(cd) >> (32 - L);
      const uint16_t e = (uint16_t)((s << 4) | L);
      for (uint32_t i = rev; i < root_size; i += 1u << L) tab[i] = e;
    } else if (L > root) {
      const uint32_t q = (cd >> (L - root)) - P0;
      if (q < G8_SUBQ) atomicMax(&S.subq[q], L - root);
      else S.overflow = 1;
      any_long = true;
    }
  }
  __syncwarp();
  any_long = __any_sync(FULL, any_long);  // warp-uniform: some stream has codes longer than the root
  if (any_long) {
    // subtable sizes -> offsets: lane g owns q in [8g, 8g+8)
    uint32_t r8[8], mine = 0;
#pragma unroll
    for (uint32_t k = 0; k < 8; ++k) {
      r8[k] = S.subq[glane * 8 + k];
      mine += r8[k] ? (1u << r8[k]) : 0;
    }
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < (int)G8; d <<= 1) {
      const uint32_t o = __shfl_up_sync(FULL, incl, d, G8);
      if ((int)glane >= d) incl += o;
    }
    uint32_t off = incl - mine;
#pragma unroll
    for (uint32_t k = 0; k < 8; ++k) {
      const uint32_t r = r8[k];
      if (r) {
        const uint32_t sz = 1u << r;
        if (root_size + off + sz <= n_entries) {
          const uint32_t slot = __brev(P0 + glane * 8 + k) >> (32 - root);
          tab[slot] = (uint16_t)(0x8000u | ((root_size + off) << 4) | r);
        } else {
          S.overflow = 1;
        }
        S.subq[glane * 8 + k] = (off << 4) | r;
        off += sz;
      }
    }
    S.nc[glane] = f0;  // second pass: the same code assignment, now placing the long codes
    S.nc[glane + G8] = f1;
    __syncwarp();
    for (uint32_t base = 0; base < nmax; base += G8) {
      const uint32_t s = base + glane;
      const uint32_t L = s < n ? lens[s] : 0;
      const uint32_t mask = __match_any_sync(FULL, L | (gid << 8));
      const uint32_t rank = __popc(mask & ((1u << lane) - 1));
      const uint32_t leader = __ffs(mask) - 1;
      const uint32_t c0 = S.nc[L & 15];
      __syncwarp();
      if (lane == leader && L) S.nc[L] = c0 + __popc(mask);
      __syncwarp();
      if (L > root && S.overflow == 0) {
        const uint32_t cd = c0 + rank;
        const uint32_t q = (cd >> (L - root)) - P0;
        if (q < G8_SUBQ) {
          const uint32_t v = S.subq[q];
          const uint32_t r = v & 15, off = v >> 4, rem = L - root;
          const uint32_t rev = __brev(cd) >> (32 - L);
          const uint16_t e = (uint16_t)((s << 4) | rem);
          for (uint32_t i = rev >> root; i < (1u << r); i += 1u << rem) tab[root_size + off + i] = e;
        }
      }
    }
  }
  __syncwarp();
  const bool ok = !over && S.overflow == 0;
  __syncwarp();  // the next build resets S.overflow: every lane must have read it first (racecheck: WAR between two builds)
  return ok;
}

// CRC-32 of a stream's output by its 8 lanes (8 KB slices), combined as in warp_crc32.
__device__ uint32_t g8_crc32(bool p, const uint8_t* data, uint32_t n, const uint32_t* T, uint32_t glane) {
  constexpr uint32_t SL = 8192;
  uint32_t c = 0;
  if (p) {
    const uint32_t b0 = min(n, glane * SL), b1 = min(n, (glane + 1) * SL);
    if (b1 > b0) {
      const uint8_t* q = data + b0;
      const uint8_t* e = data + b1;
      c = 0xffffffffu;
      while (q < e && ((uintptr_t)q & 3)) c = T[(c ^ __ldcg(q++)) & 0xff] ^ (c >> 8);
      for (; q + 4 <= e; q += 4) {
        c ^= __ldcg(reinterpret_cast<const uint32_t*>(q));
        c = T[768 + (c & 0xff)] ^ T[512 + ((c >> 8) & 0xff)] ^ T[256 + ((c >> 16) & 0xff)] ^ T[c >> 24];
      }
      while (q < e) c = T[(c ^ __ldcg(q++)) & 0xff] ^ (c >> 8);
      c = ~c;
      c = gf2_mulmod(gf2_x_pow_8n(n - b1), c);
    }
  }
  c ^= __shfl_xor_sync(FULL, c, 1);
  c ^= __shfl_xor_sync(FULL, c, 2);
  c ^= __shfl_xor_sync(FULL, c, 4);
  return c;
}

#define G8_FAIL(cond, code)       \
  do {                            \
    if ((cond) && st == 0) st = (code); \
  } while (0)

__global__ void __launch_bounds__(G8_WARPS * 32, 2) kd_inflate_g8(const InflateArgs a) {
  extern __shared__ __align__(16) uint8_t g8_smem[];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t glane = lane & (G8 - 1), gid = lane / G8;
  const uint32_t gmask = 0xffu << (gid * G8);
  G8Smem& S = reinterpret_cast<G8Smem*>(g8_smem)[warp * G8_STREAMS + gid];
  uint32_t* crcT = reinterpret_cast<uint32_t*>(g8_smem + G8_WARPS * G8_STREAMS * sizeof(G8Smem));
  G8Consts& K = *reinterpret_cast<G8Consts*>(g8_smem + G8_WARPS * G8_STREAMS * sizeof(G8Smem) + INF_CRC_TABLE_BYTES);
  {
    uint32_t c = threadIdx.x;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ CRC_POLY : c >> 1;
    crcT[threadIdx.x] = c;
    if (threadIdx.x < 32) {
      K.len_base[threadIdx.x] = c_len_base[threadIdx.x];
      K.len_extra[threadIdx.x] = c_len_extra[threadIdx.x];
      K.dist_base[threadIdx.x] = c_dist_base[threadIdx.x];
      K.dist_extra[threadIdx.x] = c_dist_extra[threadIdx.x];
    }
  }
  __syncthreads();
  {
    uint32_t c = crcT[threadIdx.x];
    for (int k = 1; k < 4; ++k) {
      c = crcT[c & 0xff] ^ (c >> 8);
      crcT[k * 256 + threadIdx.x] = c;
    }
  }
  __syncthreads();

  for (;;) {
    // ---- one block per stream
    uint32_t b = 0;
    if (glane == 0) b = a.b0 + atomicAdd(a.ticket, 1u);
    b = __shfl_sync(FULL, b, 0, G8);
    const bool have = b < a.b1;
    if (!__any_sync(FULL, have)) break;
    bool arrived = true;
    if (a.ready) {
      const volatile uint32_t* flag = a.ready + (have ? a.block_window[b] : 0);
      uint32_t spins = 0;
      arrived = !have || *flag != 0;
      unsigned long long waited_ns = 0;
      while (!__all_sync(FULL, arrived) && waited_ns < 2000000000ull) {  // the whole warp waits (its four blocks start together)
        const uint32_t ns = 256u << min(spins, 4u);  // 0.25 .. 4 us: keeps the polls off the L2 without adding latency
        __nanosleep(ns);
        waited_ns += ns;
        ++spins;
        arrived = !have || *flag != 0;
      }
      __threadfence_system();  // the flag was written by the copy engine after the window's bytes: order the payload reads behind it
    }
    const uint32_t n_out = have ? a.isize[b] : 0;
    const bool run = have && arrived && n_out != 0;
    const uint8_t* in = a.comp + (have ? a.coff[b] : 0);
    const uint32_t in_len = have ? a.clen[b] : 0;
    const uint8_t* const in_end = in + in_len;
    // an idle stream still issues the rotated loop's unconditional loads: give it a real address (a.out is a biased base
    // pointer -- only a.out + uoff[b] of a block in range is backed by memory)
    uint8_t* const out = a.out + a.uoff[have ? b : a.b0];
    uint32_t st = (have && !arrived) ? 31u : 0u;  // 0 = fine so far; otherwise the check that declined the block
    uint32_t op = 0;
    G8Reader br;
    br.base = a.comp;
    br.win = S.win;
    br.widx = 0;
    br.buf = 0;
    br.cnt = 64;  // an idle stream never refills
    br.init(run, in, glane, gmask);
    bool done = !run;  // this stream has finished (or failed)

    while (__any_sync(FULL, !done)) {  // ---- deflate blocks
      bool act = !done;
      G8_FAIL(act && br.byte_pos_ceil() > in_end, 1u);
      act = act && st == 0;
      br.refill(act, glane, gmask);
      const uint32_t bfinal = br.bits(1), btype = ((uint32_t)br.buf >> 1) & 3;
      if (act) br.consume(3);
      G8_FAIL(act && btype == 3, 4u);
      // -- stored block
      const bool stored = act && btype == 0;
      if (__any_sync(FULL, stored)) {
        if (stored) br.consume(br.cnt & 7);
        br.refill(stored, glane, gmask);
        const uint32_t len = (uint32_t)br.buf & 0xffff, nlen = ((uint32_t)br.buf >> 16) & 0xffff;
        if (stored) br.consume(32);
        G8_FAIL(stored && (len ^ nlen) != 0xffff, 2u);
        const uint8_t* src = br.byte_pos();
        G8_FAIL(stored && (src + len > in_end || op + len > n_out), 3u);
        const bool cp = stored && st == 0;
        if (cp) {
          for (uint32_t i = glane; i < len; i += G8) out[op + i] = __ldcg(src + i);
          op += len;
        }
        br.init(cp, src + len, glane, gmask);
      }
      // -- Huffman block: code lengths
      const bool huff = act && st == 0 && (btype == 1 || btype == 2);
      const bool fixed = huff && btype == 1, dyn = huff && btype == 2;
      uint32_t hlit = 288, hdist = 32;
      if (__any_sync(FULL, fixed)) {
        if (fixed)
          for (uint32_t i = glane; i < 320; i += G8) S.lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : i < 288 ? 8 : 5;
        __syncwarp();
      }
      if (__any_sync(FULL, dyn)) {
        br.refill(dyn, glane, gmask);
        uint32_t hclen = 0;
        if (dyn) {
          hlit = br.bits(5) + 257;
          hdist = (((uint32_t)br.buf >> 5) & 31) + 1;
          hclen = (((uint32_t)br.buf >> 10) & 15) + 4;
          br.consume(14);
        }
        G8_FAIL(dyn && (hlit > 286 || hdist > 30), 5u);
        bool d2 = dyn && st == 0;
        if (d2)
          for (uint32_t i = glane; i < 19; i += G8) S.lens[i] = 0;
        __syncwarp();
        for (uint32_t i = 0; i < 19; ++i) {
          const bool pi = d2 && i < hclen;
          br.refill(pi, glane, gmask);
          if (pi) {
            if (glane == 0) S.lens[c_clen_order[i]] = (uint8_t)br.bits(3);
            br.consume(3);
          }
        }
        __syncwarp();
        const bool okp = g8_build_table(S, S.lens, d2 ? 19u : 0u, S.dst, 7, 128, lane);
        G8_FAIL(d2 && !okp, 6u);
        d2 = d2 && st == 0;
        const uint32_t total = hlit + hdist;
        uint32_t n = 0, prev = 0;
        while (__any_sync(FULL, d2 && n < total)) {
          const bool p = d2 && n < total;
          br.refill(p, glane, gmask);
          const uint32_t e = S.dst[(uint32_t)br.buf & 127];
          const uint32_t l = e & 15, sym = e >> 4;
          if (p) {
            if (!l) {
              G8_FAIL(true, 7u);
              d2 = false;
            } else {
              br.consume(l);
              if (sym < 16) {
                if (glane == 0) S.lens[n] = (uint8_t)sym;
                prev = sym;
                ++n;
              } else {
                uint32_t rep, val = 0;
                if (sym == 16) {
                  val = prev;
                  rep = 3 + br.bits(2);
                  br.consume(2);
                  if (n == 0) {
                    G8_FAIL(true, 8u);
                    d2 = false;
                  }
                } else if (sym == 17) {
                  rep = 3 + br.bits(3);
                  br.consume(3);
                } else {
                  rep = 11 + br.bits(7);
                  br.consume(7);
                }
                if (n + rep > total) {
                  G8_FAIL(true, 9u);
                  d2 = false;
                } else if (d2) {
                  for (uint32_t i = glane; i < rep; i += G8) S.lens[n + i] = (uint8_t)val;
                  prev = val;
                  n += rep;
                }
              }
            }
          }
        }
        __syncwarp();
        G8_FAIL(d2 && S.lens[256] == 0, 10u);
      }
      // -- tables
      const bool tb = huff && st == 0;
      if (__any_sync(FULL, tb)) {
        const bool okd = g8_build_table(S, S.lens + hlit, tb ? hdist : 0u, S.dst, G8_DROOT, G8_DST_ENTRIES, lane);
        G8_FAIL(tb && !okd, 11u);
        const bool tl = tb && st == 0;
        const bool okl = g8_build_table(S, S.lens, tl ? hlit : 0u, S.lit, G8_LROOT, G8_LIT_ENTRIES, lane);
        G8_FAIL(tl && !okl, 12u);
      }
      // -- symbols.  The loop is rotated: an iteration first issues the loads of the match decoded by the PREVIOUS
      //    iteration, then decodes the next symbol of every stream, then stores what it loaded -- the L2 latency of the
      //    LZ77 copy hides behind the decoding (loads cannot stay in flight across the loop's back edge).
      bool in_blk = huff && st == 0;
      uint32_t pm_len = 0, pm_dist = 0, pm_op = 0;  // this stream's pending match
      while (__any_sync(FULL, in_blk || pm_len != 0)) {
        // A: loads of the pending matches (all four streams together)
        __syncwarp();  // every earlier store of this stream is visible to the loads below
        uint8_t* const dstp = out + pm_op;
        const uint8_t* const srcp = dstp - pm_dist;
        const uint32_t bound = __reduce_max_sync(FULL, pm_len);
        bool cpy = false;
        uint8_t val = 0;
        if (bound != 0) {
          if (bound <= G8) {
            cpy = glane < pm_len;
            uint32_t j = glane;
            if (__any_sync(FULL, cpy && pm_dist < pm_len)) {
              if (cpy && pm_dist < pm_len) j = glane % pm_dist;
            }
            val = *(cpy ? srcp + j : out);  // unconditional load through a selected address
          } else {
            for (uint32_t i = glane; i < bound; i += G8)
              if (i < pm_len) dstp[i] = srcp[pm_dist >= pm_len ? i : i % pm_dist];
          }
        }
        pm_len = 0;
        // B: one symbol per stream
        br.refill(in_blk, glane, gmask);
        uint32_t e = S.lit[(uint32_t)br.buf & ((1u << G8_LROOT) - 1)];
        uint32_t used = 0;
        if (__any_sync(FULL, in_blk && (e & 0x8000))) {
          if (e & 0x8000) {
            used = G8_LROOT;
            e = S.lit[((e >> 4) & 0x7ff) + (((uint32_t)(br.buf >> G8_LROOT)) & ((1u << (e & 15)) - 1))];
          }
        }
        uint32_t l = e & 15;
        const uint32_t sym = (e >> 4) & 0x1ff;
        G8_FAIL(in_blk && (l == 0 || (e & 0x8000)), 13u);
        in_blk = in_blk && st == 0;
        if (in_blk) br.consume(used + l);
        const bool is_lit = in_blk && sym < 256;
        const bool is_len = in_blk && sym > 256;
        if (in_blk && sym == 256) in_blk = false;  // end of block
        if (is_lit) {
          if (op < n_out) {
            if (glane == 0) out[op] = (uint8_t)sym;
            ++op;
          } else {
            G8_FAIL(true, 14u);
          }
        }
        if (__any_sync(FULL, is_len)) {
          const uint32_t ls = is_len ? sym - 257 : 0;
          G8_FAIL(is_len && ls > 28, 15u);
          bool m = is_len && st == 0;
          const uint32_t lsc = m ? ls : 0;
          const uint32_t xl = K.len_extra[lsc];
          const uint32_t len = K.len_base[lsc] + br.bits(xl);
          if (m) br.consume(xl);
          br.refill(m, glane, gmask);
          e = S.dst[(uint32_t)br.buf & ((1u << G8_DROOT) - 1)];
          used = 0;
          if (__any_sync(FULL, m && (e & 0x8000))) {
            if (e & 0x8000) {
              used = G8_DROOT;
              e = S.dst[((e >> 4) & 0x7ff) + (((uint32_t)(br.buf >> G8_DROOT)) & ((1u << (e & 15)) - 1))];
            }
          }
          l = e & 15;
          const uint32_t dsym = (e >> 4) & 0x1ff;
          G8_FAIL(m && (l == 0 || (e & 0x8000)), 16u);
          G8_FAIL(m && dsym > 29, 17u);
          m = m && st == 0;
          if (m) br.consume(used + l);
          const uint32_t dsc = m ? dsym : 0;
          const uint32_t xd = K.dist_extra[dsc];
          const uint32_t dist = K.dist_base[dsc] + br.bits(xd);
          if (m) br.consume(xd);
          G8_FAIL(m && (dist > op || op + len > n_out), 18u);
          m = m && st == 0;
          if (m) {  // copied by the next iteration
            pm_len = len;
            pm_dist = dist;
            pm_op = op;
            op += len;
          }
        }
        // C: stores of the loads issued in A
        if (cpy) dstp[glane] = val;
        in_blk = in_blk && st == 0;
      }
      done = done || st != 0 || (act && bfinal != 0);
    }
    // ---- per-stream verdict
    G8_FAIL(run && op != n_out, 19u);
    G8_FAIL(run && br.byte_pos_ceil() > in_end, 20u);
    __syncwarp();
    {
      const bool ck = run && st == 0;
      const uint8_t* f = in_end;
      uint32_t want = 0;
      if (ck) want = (uint32_t)__ldcg(f) | ((uint32_t)__ldcg(f + 1) << 8) | ((uint32_t)__ldcg(f + 2) << 16) | ((uint32_t)__ldcg(f + 3) << 24);
      const uint32_t got = g8_crc32(ck, out, n_out, crcT, glane);
      G8_FAIL(ck && got != want, 30u);
    }
    if (have && glane == 0) {
      a.status[b] = st;
      if (st) atomicAdd(a.fail_count, 1u);
    }
    __syncwarp();
  }
}
