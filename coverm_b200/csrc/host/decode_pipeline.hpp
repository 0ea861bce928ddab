// Region-parallel BAM decode for libcoverm_b200's host side.
//
// The file is cut into work items of ~1 MB of uncompressed data (whole BGZF blocks).  Worker threads take items in
// order; each one inflates its item into a private buffer (so the data stays in that core's cache), waits for the
// record alignment handed over by its predecessor (the bytes of the record that straddles the item boundary), walks
// the block_size chain of its own records, hands the alignment on, and only then extracts its tuples straight into
// the pinned SoA staging batch of the device library.  The only sequential part is the block_size walk (one load
// per record on cache-hot data); inflate (zlib) and tuple extraction run on all cores.
// The calling thread is the coordinator: it alone talks to the (not thread-safe) device ABI — acquiring staging
// batches ahead of the workers and submitting each batch once every item assigned to it has been extracted.
//
// Replaces for this path: bam::Reader::read + htslib's BGZF thread pool (bam_generator.rs:103-134, 125-129).
#pragma once
#include <atomic>
#include <climits>
#include <exception>
#include <memory>
#include <mutex>
#include <thread>

#include "bam_source.hpp"

namespace cmbh {

struct BlockRef {
  size_t cdata, clen;  // compressed payload (BGZF) or raw slice [cdata, cdata+clen)
  uint32_t isize;      // uncompressed size
};

// All blocks of the input with their uncompressed offsets.
class BlockIndex {
 public:
  std::vector<BlockRef> blocks;
  std::vector<uint64_t> ustart;  // blocks.size()+1 cumulative uncompressed offsets
  bool bgzf = false;
  const uint8_t* p = nullptr;
  size_t n = 0;

  void build(const uint8_t* data, size_t size) {
    p = data;
    n = size;
    blocks.clear();
    bgzf = size >= 18 && data[0] == 0x1f && data[1] == 0x8b;
    if (bgzf) {
      size_t o = 0;
      while (o < size) {
        if (o + 18 > size || data[o] != 0x1f || data[o + 1] != 0x8b || data[o + 2] != 8 || !(data[o + 3] & 4))
          throw Panic("Error reading BAM record: corrupt BGZF block header");
        const size_t xlen = data[o + 10] | (data[o + 11] << 8);
        size_t x = o + 12;
        const size_t xend = x + xlen;
        if (xend > size) throw Panic("Error reading BAM record: truncated BGZF block");
        int bsize = -1;
        while (x + 4 <= xend) {
          const size_t slen = data[x + 2] | (data[x + 3] << 8);
          if (data[x] == 'B' && data[x + 1] == 'C' && slen == 2) bsize = data[x + 4] | (data[x + 5] << 8);
          x += 4 + slen;
        }
        if (bsize < 0) throw Panic("Error reading BAM record: gzip member without a BGZF block size");
        const size_t end = o + (size_t)bsize + 1;
        if (end > size || end < xend + 8) throw Panic("Error reading BAM record: truncated BGZF block");
        BlockRef b;
        b.cdata = xend;
        b.clen = end - 8 - xend;
        b.isize = data[end - 4] | (data[end - 3] << 8) | (data[end - 2] << 16) | ((uint32_t)data[end - 1] << 24);
        blocks.push_back(b);
        o = end;
      }
    } else {
      for (size_t o = 0; o < size; o += 65536) blocks.push_back({o, std::min<size_t>(65536, size - o), (uint32_t)std::min<size_t>(65536, size - o)});
    }
    ustart.assign(blocks.size() + 1, 0);
    for (size_t i = 0; i < blocks.size(); ++i) ustart[i + 1] = ustart[i] + blocks[i].isize;
  }

  // Inflate blocks [b0, b1) into dst (which has room for ustart[b1]-ustart[b0] bytes).
  void inflate(size_t b0, size_t b1, uint8_t* dst, z_stream* zs) const {
    for (size_t b = b0; b < b1; ++b) {
      const BlockRef& r = blocks[b];
      uint8_t* out = dst + (ustart[b] - ustart[b0]);
      if (!bgzf) {
        memcpy(out, p + r.cdata, r.clen);
        continue;
      }
      if (r.isize == 0) continue;
      inflateReset(zs);
      zs->next_in = const_cast<Bytef*>(p + r.cdata);
      zs->avail_in = (uInt)r.clen;
      zs->next_out = out;
      zs->avail_out = r.isize;
      if (::inflate(zs, Z_FINISH) != Z_STREAM_END || zs->avail_out != 0) throw Panic("Error reading BAM record: BGZF inflate failed");
    }
  }
};

// Decode the fixed fields, CIGAR summary and NM aux of one BAM record; M/=/X blocks (contig.rs:171-186) are written to
// ivs/ivl (room for n_cigar_op entries).  Returns the number of intervals written.
inline uint32_t decode_bam_record_into(const uint8_t* rec, Tuple& t, int32_t* ivs, int32_t* ivl) {
  const uint32_t block_size = rd_u32(rec);
  const uint8_t* o = rec + 4;
  const uint8_t* end = o + block_size;
  t.tid = (int32_t)rd_u32(o);
  t.pos = (int32_t)rd_u32(o + 4);
  const uint32_t l_read_name = o[8];
  t.mapq = o[9];
  const uint32_t n_cigar = rd_u16(o + 12);
  t.flag = rd_u16(o + 14);
  t.l_seq = rd_u32(o + 16);
  t.mtid = (int32_t)rd_u32(o + 20);
  const uint8_t* cig = o + 32 + l_read_name;
  uint32_t aligned = 0, del = 0, ins = 0, n_iv = 0;
  int64_t cursor = t.pos;
  for (uint32_t i = 0; i < n_cigar; ++i) {
    const uint32_t v = rd_u32(cig + 4 * i);
    const uint32_t op = v & 0xf, len = v >> 4;
    switch (op) {
      case 0: case 7: case 8:
        ivs[n_iv] = cursor < 0 ? -1 : (int32_t)std::min<int64_t>(cursor, INT32_MAX);
        ivl[n_iv] = (int32_t)len;
        ++n_iv;
        cursor += len;
        aligned += len;
        break;
      case 2: cursor += len; del += len; aligned += len; break;
      case 3: cursor += len; break;
      case 1: ins += len; aligned += len; break;
      default: break;
    }
  }
  t.aligned = aligned;
  t.del = del;
  t.ins = ins;
  t.n_iv = n_iv;
  const uint8_t* a = cig + 4 * (size_t)n_cigar + (t.l_seq + 1) / 2 + t.l_seq;
  t.nm_state = 0;
  t.nm = 0;
  while (a + 3 <= end) {
    const uint8_t t0 = a[0], t1 = a[1], ty = a[2];
    a += 3;
    size_t sz;
    switch (ty) {
      case 'A': case 'c': case 'C': sz = 1; break;
      case 's': case 'S': sz = 2; break;
      case 'i': case 'I': case 'f': sz = 4; break;
      case 'Z': case 'H': {
        const uint8_t* e = (const uint8_t*)memchr(a, 0, (size_t)(end - a));
        sz = e ? (size_t)(e - a) + 1 : (size_t)(end - a);
        break;
      }
      case 'B': {
        if (a + 5 > end) { sz = (size_t)(end - a); break; }
        const uint8_t sub = a[0];
        const uint32_t cnt = rd_u32(a + 1);
        const size_t es = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
        sz = 5 + es * (size_t)cnt;
        break;
      }
      default: throw Panic("Error reading BAM record: unknown aux type");
    }
    if (t0 == 'N' && t1 == 'M' && t.nm_state == 0) {
      if (ty == 'C') { t.nm_state = 1; t.nm = a[0]; }
      else if (ty == 'S') { t.nm_state = 1; t.nm = rd_u16(a); }
      else if (ty == 'I') { t.nm_state = 1; t.nm = rd_u32(a); }
      else t.nm_state = 2;
    }
    a += sz;
  }
  return n_iv;
}

struct PipelineCounts {
  uint64_t n_records = 0, primaries = 0;
};

// Runs the whole record stream (starting at uncompressed offset `records_at`) through the device context.
// acquire()/submit() are the caller's wrappers around cmb_acquire_batch / cmb_submit_batch (called only from this thread).
template <class Acquire, class Submit>
PipelineCounts run_decode_pipeline(const BlockIndex& bx, uint64_t records_at, int n_threads, uint32_t cap_r, uint32_t cap_i,
                                   uint32_t n_staging, Acquire acquire, Submit submit) {
  constexpr size_t ITEM_BYTES = 1u << 20;
  struct Item { size_t b0, b1; };
  std::vector<Item> items;
  size_t first_block = 0;
  while (first_block < bx.blocks.size() && bx.ustart[first_block + 1] <= records_at) ++first_block;
  const uint64_t skip0 = first_block < bx.blocks.size() ? records_at - bx.ustart[first_block] : 0;
  size_t max_item = 0;
  for (size_t b = first_block; b < bx.blocks.size();) {
    size_t e = b;
    uint64_t sz = 0;
    while (e < bx.blocks.size() && (sz == 0 || sz + bx.blocks[e].isize <= ITEM_BYTES)) sz += bx.blocks[e++].isize;
    items.push_back({b, e});
    max_item = std::max<size_t>(max_item, sz);
    b = e;
  }
  const size_t n_items = items.size();

  struct Chain {  // state handed from item i-1 to item i
    std::atomic<int> ready{0};
    std::vector<uint8_t> carry;  // bytes of the record that straddles the boundary
    uint64_t batch_seq = 0;
    uint32_t used_r = 0, used_i = 0;
    uint32_t items_in_batch = 0;  // items with records assigned to batch_seq so far
  };
  std::unique_ptr<Chain[]> chain(new Chain[n_items + 1]);
  chain[0].ready.store(1);

  struct Slot {  // one staging batch in flight
    std::atomic<int64_t> have_seq{-1};  // batch sequence number whose pointers are published here
    cmb_read_batch ptrs{};
    std::atomic<uint32_t> done{0};      // items of this batch fully extracted
    std::atomic<int> closed{0};
    uint32_t final_r = 0, final_i = 0, final_items = 0;
  };
  std::unique_ptr<Slot[]> slots(new Slot[n_staging]);
  std::atomic<int64_t> needed_seq{0};
  std::atomic<size_t> next_item{0};
  std::atomic<bool> abort{false};
  std::atomic<int> workers_left{0};
  std::exception_ptr first_error;
  std::mutex err_mu;
  std::atomic<uint64_t> tot_records{0}, tot_primaries{0};

  auto fail = [&](std::exception_ptr e) {
    std::lock_guard<std::mutex> g(err_mu);
    if (!first_error) first_error = e;
    abort = true;
  };

  auto worker = [&]() {
    try {
      z_stream zs;
      memset(&zs, 0, sizeof zs);
      if (inflateInit2(&zs, -15) != Z_OK) throw Panic("zlib init failed");
      std::unique_ptr<uint8_t[]> buf(new uint8_t[max_item + 8]);
      std::vector<uint32_t> offs;
      std::vector<uint8_t> stitched;
      uint64_t my_records = 0, my_primaries = 0;
      for (;;) {
        const size_t i = next_item.fetch_add(1);
        if (i >= n_items || abort) break;
        const Item& it = items[i];
        const size_t usize = (size_t)(bx.ustart[it.b1] - bx.ustart[it.b0]);
        bx.inflate(it.b0, it.b1, buf.get(), &zs);
        while (!chain[i].ready.load(std::memory_order_acquire)) {
          if (abort) break;
          std::this_thread::yield();
        }
        if (abort) break;
        // ---- alignment: finish the straddling record, then walk my records
        Chain& in = chain[i];
        Chain& out = chain[i + 1];
        size_t pos = i == 0 ? (size_t)skip0 : 0;
        bool have_stitched = false;
        offs.clear();
        uint64_t ub_iv = 0;
        bool swallowed = false;  // the whole item is the middle of one huge record
        if (!in.carry.empty()) {
          stitched = std::move(in.carry);
          while (stitched.size() < 4 && pos < usize) stitched.push_back(buf[pos++]);
          if (stitched.size() < 4) {
            swallowed = true;
          } else {
            const size_t need = 4 + (size_t)rd_u32(stitched.data());
            if (need < 36) throw Panic("Error reading BAM record: corrupt block_size");
            const size_t take = std::min(need - stitched.size(), usize - pos);
            stitched.insert(stitched.end(), buf.get() + pos, buf.get() + pos + take);
            pos += take;
            if (stitched.size() < need) swallowed = true;
            else {
              have_stitched = true;
              ub_iv += rd_u16(stitched.data() + 4 + 12);
            }
          }
        }
        if (swallowed) {
          out.carry = std::move(stitched);
        } else {
          while (pos + 4 <= usize) {
            const uint32_t bs = rd_u32(buf.get() + pos);
            if (bs < 32) throw Panic("Error reading BAM record: corrupt block_size");
            if (pos + 4 + (size_t)bs > usize) break;
            offs.push_back((uint32_t)pos);
            ub_iv += rd_u16(buf.get() + pos + 4 + 12);
            pos += 4 + (size_t)bs;
          }
          out.carry.assign(buf.get() + pos, buf.get() + usize);
        }
        const uint32_t n_rec = (uint32_t)offs.size() + (have_stitched ? 1u : 0u);
        if (n_rec > cap_r || ub_iv > cap_i) throw ExitError(1, "a decode work item holds more records than a device batch");
        // ---- batch assignment
        uint64_t seq = in.batch_seq;
        uint32_t used_r = in.used_r, used_i = in.used_i, items_in_batch = in.items_in_batch;
        auto wait_for_slot = [&](uint64_t q) {  // until the coordinator has handed staging memory to batch q
          while (slots[q % n_staging].have_seq.load(std::memory_order_acquire) != (int64_t)q) {
            if (abort) return false;
            std::this_thread::yield();
          }
          return true;
        };
        if (used_r + (uint64_t)n_rec > cap_r || used_i + ub_iv > cap_i) {  // close the current batch, open the next
          if (!wait_for_slot(seq)) break;
          Slot& s = slots[seq % n_staging];
          s.final_r = used_r;
          s.final_i = used_i;
          s.final_items = items_in_batch;
          s.closed.store(1, std::memory_order_release);
          ++seq;
          used_r = used_i = items_in_batch = 0;
          int64_t cur = needed_seq.load();
          while (cur < (int64_t)seq && !needed_seq.compare_exchange_weak(cur, (int64_t)seq)) {
          }
        }
        out.batch_seq = seq;
        out.used_r = used_r + n_rec;
        out.used_i = used_i + (uint32_t)ub_iv;
        out.items_in_batch = items_in_batch + (n_rec ? 1u : 0u);
        out.ready.store(1, std::memory_order_release);
        if (!n_rec) continue;
        // ---- extraction into the staging batch
        if (!wait_for_slot(seq)) break;
        Slot& slot = slots[seq % n_staging];
        const cmb_read_batch& b = slot.ptrs;
        uint32_t r = used_r, iv = used_i;
        Tuple t;
        auto put = [&](const uint8_t* rec) {
          const uint32_t n_iv = decode_bam_record_into(rec, t, b.iv_start + iv, b.iv_len + iv);
          b.tid[r] = t.tid; b.pos[r] = t.pos; b.flag[r] = t.flag; b.mapq[r] = t.mapq; b.nm_state[r] = t.nm_state;
          b.nm[r] = t.nm; b.l_seq[r] = t.l_seq; b.aligned[r] = t.aligned; b.del[r] = t.del; b.ins[r] = t.ins;
          b.iv_begin[r] = iv;
          iv += n_iv;
          ++r;
          if (!(t.flag & 0x900)) ++my_primaries;
        };
        if (have_stitched) put(stitched.data());
        for (uint32_t o : offs) put(buf.get() + o);
        for (const uint32_t end_iv = used_i + (uint32_t)ub_iv; iv < end_iv; ++iv) {  // unused part of my interval reservation
          b.iv_start[iv] = CMB_IV_PAD;
          b.iv_len[iv] = 0;
        }
        my_records += n_rec;
        slot.done.fetch_add(1, std::memory_order_acq_rel);
      }
      inflateEnd(&zs);
      tot_records += my_records;
      tot_primaries += my_primaries;
    } catch (...) {
      fail(std::current_exception());
    }
    workers_left.fetch_sub(1);
  };

  const int nt = std::max(1, std::min<int>(n_threads, (int)std::max<size_t>(1, n_items)));
  workers_left = nt;
  std::vector<std::thread> pool;
  for (int k = 0; k < nt; ++k) pool.emplace_back(worker);

  // ---- coordinator: the only thread that touches the device ABI
  int64_t acquired = 0, submitted = 0;
  try {
    for (;;) {
      bool progressed = false;
      while (acquired <= needed_seq.load() && acquired - submitted < (int64_t)n_staging) {
        Slot& s = slots[acquired % n_staging];
        s.closed.store(0);
        s.done.store(0);
        acquire(&s.ptrs);
        s.have_seq.store(acquired, std::memory_order_release);
        ++acquired;
        progressed = true;
      }
      if (submitted < acquired) {
        Slot& s = slots[submitted % n_staging];
        if (s.closed.load(std::memory_order_acquire) && s.done.load(std::memory_order_acquire) == s.final_items) {
          s.ptrs.iv_begin[s.final_r] = s.final_i;
          submit(s.final_r, s.final_i);
          ++submitted;
          progressed = true;
        }
      }
      if (abort) break;
      if (workers_left.load() == 0 && !progressed) {
        bool more = false;  // workers are done: is a closed batch still waiting, or one still to acquire?
        if (submitted < acquired && slots[submitted % n_staging].closed.load()) more = true;
        if (acquired <= needed_seq.load() && acquired - submitted < (int64_t)n_staging) more = true;
        if (!more) break;
      }
      if (!progressed) std::this_thread::yield();
    }
  } catch (...) {
    fail(std::current_exception());
  }
  for (auto& th : pool) th.join();
  if (first_error) {
    // hand back every acquired batch so that the context can be reused
    try {
      for (; submitted < acquired; ++submitted) submit(0, 0);
    } catch (...) {
    }
    std::rethrow_exception(first_error);
  }
  // ---- the last, still open batch (and any batch acquired ahead but never used)
  const Chain& fin = chain[n_items];
  for (; submitted < acquired; ++submitted) {
    Slot& s = slots[submitted % n_staging];
    if ((uint64_t)submitted == fin.batch_seq && !s.closed.load()) {
      s.ptrs.iv_begin[fin.used_r] = fin.used_i;
      submit(fin.used_r, fin.used_i);
    } else if (s.closed.load()) {
      s.ptrs.iv_begin[s.final_r] = s.final_i;
      submit(s.final_r, s.final_i);
    } else {
      submit(0, 0);
    }
  }
  PipelineCounts c;
  c.n_records = tot_records;
  c.primaries = tot_primaries;
  return c;
}

}  // namespace cmbh
