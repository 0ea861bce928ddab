// C ABI of the host drivers (include/coverm_b200_host.h).
#include <memory>

#include "../../../include/coverm_b200_host.h"
#include "cli.hpp"

using namespace cmbh;

struct cmbh_session {
  std::unique_ptr<DeviceSession> dev;
};

namespace {
std::string g_err;
char* dup_text(const std::string& s) {
  char* p = (char*)malloc(s.size() + 1);
  if (p) {
    memcpy(p, s.data(), s.size());
    p[s.size()] = 0;
  }
  return p;
}
}  // namespace

extern "C" {

const char* cmbh_last_error(void) { return g_err.c_str(); }

int cmbh_session_create(int device, int threads, cmbh_session** out) {
  if (!out) return -2;
  *out = nullptr;
  try {
    auto s = std::make_unique<cmbh_session>();
    s->dev = std::make_unique<DeviceSession>(device, threads);
    *out = s.release();
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

void cmbh_session_destroy(cmbh_session* s) { delete s; }

int cmbh_session_set_shard(cmbh_session* s, uint32_t tid_begin, uint32_t tid_end) {
  if (!s) return -2;
  s->dev->set_shard(tid_begin, tid_end);
  return 0;
}

int cmbh_run(cmbh_session* s, int argc, const char* const* argv, const cmbh_mem_input* mem, int n_mem, cmbh_result* res) {
  if (!res || (argc > 0 && !argv)) return -2;
  memset(res, 0, sizeof *res);
  std::vector<std::string> args(argv, argv + argc);
  std::vector<InputSpec> inputs;
  for (int i = 0; i < n_mem; ++i) {
    InputSpec in;
    in.path = mem[i].path;
    in.data = mem[i].data;
    in.size = mem[i].size;
    inputs.push_back(in);
  }
  std::ostringstream out, err;
  const CliResult r = run_cli(args, inputs, out, err, s ? s->dev.get() : nullptr);
  res->status = r.status;
  const std::string so = out.str(), se = err.str();
  res->out = dup_text(so);
  res->out_len = so.size();
  res->err = dup_text(se);
  res->err_len = se.size();
  res->n_samples = (uint32_t)std::min<size_t>(CMBH_MAX_SAMPLES, r.timings.size());
  for (uint32_t i = 0; i < res->n_samples; ++i) {
    cmbh_sample_info& si = res->samples[i];
    if (i < r.reads_mapped.size()) {
      si.num_mapped_reads = r.reads_mapped[i].num_mapped_reads;
      si.num_reads = r.reads_mapped[i].num_reads;
    }
    si.n_records = r.record_counts[i];
    const SampleTiming& t = r.timings[i];
    si.total_s = t.total_s;
    si.decode_s = t.decode_s;
    si.submit_wait_s = t.submit_wait_s;
    si.end_sample_s = t.end_sample_s;
    si.k0_ms = t.device.ms_zero;
    si.k1_ms = t.device.ms_accumulate;
    si.k2_ms = t.device.ms_scan;
    si.k3_ms = t.device.ms_finalize;
    si.device_total_ms = t.device.ms_total;
    si.k1_launches = t.device.k1_launches;
    si.k2_launches = t.device.k2_launches;
    si.k3_launches = t.device.k3_launches;
    si.arena_elems = t.device.arena_elems;
    si.n_intervals = t.device.n_intervals;
  }
  return 0;
}

void cmbh_free_result(cmbh_result* res) {
  if (!res) return;
  free(res->out);
  free(res->err);
  res->out = res->err = nullptr;
}

int cmbh_main(int argc, char** argv) {
  std::ios::sync_with_stdio(false);
  std::vector<std::string> args(argv + (argc > 0 ? 1 : 0), argv + argc);
  const CliResult r = run_cli(args, {}, std::cout, std::cerr);
  std::cout.flush();
  return r.status;
}

}  // extern "C"
