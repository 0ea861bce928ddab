#include "fast_inflate.hpp"
#include <zlib.h>
#include <vector>
#include <random>
#include <cstdio>
#include <chrono>
#include <string>
using namespace std;
static vector<uint8_t> deflate_raw(const vector<uint8_t>& in, int level, int strategy) {
  z_stream zs{}; deflateInit2(&zs, level, Z_DEFLATED, -15, 8, strategy);
  vector<uint8_t> out(deflateBound(&zs, in.size()) + 64);
  zs.next_in = (Bytef*)in.data(); zs.avail_in = in.size(); zs.next_out = out.data(); zs.avail_out = out.size();
  deflate(&zs, Z_FINISH); out.resize(zs.total_out); deflateEnd(&zs); return out;
}
int main(int argc, char** argv) {
  mt19937_64 rng(1);
  cmbh::FastInflate fi;
  int fails = 0, n = 0;
  for (int iter = 0; iter < 1500; ++iter) {
    size_t len = iter < 20 ? iter : rng() % 65536;
    vector<uint8_t> data(len);
    int kind = iter % 6;
    for (size_t i = 0; i < len; ++i) {
      switch (kind) {
        case 0: data[i] = rng(); break;
        case 1: data[i] = "ACGT"[rng() & 3]; break;
        case 2: data[i] = (i % 37) ^ (rng() % 100 == 0); break;
        case 3: data[i] = rng() % 8 ? 'A' : rng(); break;
        case 4: data[i] = i > 3 && rng() % 4 ? data[i - 1 - rng() % min<size_t>(i - 1, 5)] : rng(); break;
        case 5: { static const char* w = "the quick brown fox jumps over the lazy dog "; data[i] = w[(i + (rng() % 50 == 0)) % 44]; } break;
      }
    }
    int level = iter % 10; int strat = (iter / 10) % 4 == 3 ? Z_FIXED : (iter / 10) % 4 == 2 ? Z_HUFFMAN_ONLY : Z_DEFAULT_STRATEGY;
    auto c = deflate_raw(data, level, strat);
    size_t clen = c.size();
    c.resize(clen + 8, 0xAB);
    vector<uint8_t> out(len + 1);
    out[len] = 0x5a;
    bool ok = fi.run(c.data(), clen, out.data(), len);
    ++n;
    if (!ok || memcmp(out.data(), data.data(), len) || out[len] != 0x5a) { ++fails; printf("FAIL iter %d len %zu level %d strat %d ok %d\n", iter, len, level, strat, ok); }
    // corruption robustness: flip bytes; must not crash, and wrong size must return false
    if (clen > 4) {
      auto c2 = c; c2[rng() % clen] ^= 1 << (rng() % 8);
      vector<uint8_t> o2(len + 1); fi.run(c2.data(), clen, o2.data(), len);
      // truncated
      fi.run(c.data(), clen / 2, o2.data(), len);
      // wrong out_len
      if (len > 1 && fi.run(c.data(), clen, o2.data(), len - 1)) { ++fails; printf("FAIL short accepted\n"); }
    }
  }
  printf("%d tests, %d fails\n", n, fails);
  // speed
  for (int kind : {1, 3, 0}) {
    vector<uint8_t> data(65280);
    for (size_t i = 0; i < data.size(); ++i) data[i] = kind == 1 ? "ACGT"[rng() & 3] : kind == 3 ? (rng() % 8 ? 'A' + (i % 3) : rng()) : rng() & 63;
    for (int level : {1, 6}) {
      auto c = deflate_raw(data, level, Z_DEFAULT_STRATEGY); size_t clen = c.size(); c.resize(clen + 8);
      vector<uint8_t> out(data.size());
      auto t0 = chrono::steady_clock::now();
      int reps = 2000;
      for (int r = 0; r < reps; ++r) fi.run(c.data(), clen, out.data(), data.size());
      double t1 = chrono::duration<double>(chrono::steady_clock::now() - t0).count();
      z_stream zs{}; inflateInit2(&zs, -15);
      t0 = chrono::steady_clock::now();
      for (int r = 0; r < reps; ++r) { inflateReset(&zs); zs.next_in = c.data(); zs.avail_in = clen; zs.next_out = out.data(); zs.avail_out = out.size(); inflate(&zs, Z_FINISH); }
      double t2 = chrono::duration<double>(chrono::steady_clock::now() - t0).count();
      t0 = chrono::steady_clock::now();
      unsigned long cr = 0; for (int r = 0; r < reps; ++r) cr += crc32(0, out.data(), out.size());
      double t3 = chrono::duration<double>(chrono::steady_clock::now() - t0).count();
      printf("kind %d level %d ratio %.2f: fast %.0f MB/s zlib %.0f MB/s crc %.0f MB/s (%lu)\n", kind, level, (double)data.size() / clen, reps * data.size() / t1 / 1e6, reps * data.size() / t2 / 1e6, reps * data.size() / t3 / 1e6, cr & 1);
    }
  }
}
