/*
 * coverm_b200_host.h — C ABI of the host half of libcoverm_b200.so: the `coverm contig` / `coverm genome`
 * (--bam-files) drivers built on the device ABI of coverm_b200.h.  This is the call a non-Rust embedder makes;
 * it corresponds to run_contig / run_genome in the reference (src/bin/coverm.rs:2088-2131, 1539-1628), i.e.
 * contig_coverage (src/contig.rs:13-253) / mosdepth_genome_coverage* (src/genome.rs:17-322, 419-797) followed by
 * CoveragePrinter::finalise_printing (src/coverage_printer.rs:20-121).
 *
 * argv is the coverm command line without the program name, e.g.
 *   {"contig", "-m", "mean", "trimmed_mean", "covered_fraction", "-b", "sample.bam", "-t", "32"}
 * BAM inputs may be given as host memory buffers (`mem`): a `-b` path equal to mem[i].path is read from
 * mem[i].data instead of the filesystem.  The table is returned as text exactly as the reference prints it.
 */
#ifndef COVERM_B200_HOST_H
#define COVERM_B200_HOST_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMBH_MAX_SAMPLES 64

typedef struct cmbh_mem_input {
  const char* path;
  const uint8_t* data;
  size_t size;
} cmbh_mem_input;

typedef struct cmbh_sample_info {
  uint64_t num_mapped_reads; /* ReadsMapped (src/lib.rs:53-57) */
  uint64_t num_reads;
  uint64_t n_records;        /* records read from the file */
  double total_s, decode_s, submit_wait_s, end_sample_s;
  float k0_ms, k1_ms, k2_ms, k3_ms, device_total_ms;
  uint32_t k1_launches, k2_launches, k3_launches;
  uint64_t arena_elems, n_intervals;
  uint64_t h2d_bytes;        /* bytes copied host->device for this sample: tuples (host decode) or BGZF bytes + block table */
  uint32_t device_decode;    /* 1: the GPU inflated and parsed the BAM (cmb_submit_bgzf); 0: host decode pipeline */
  uint32_t decode_host_blocks; /* BGZF blocks the device declined (inflated with zlib by the library) */
  float decode_copy_inflate_ms, decode_chain_ms, decode_extract_ms; /* device decode stages (CUDA events) */
  uint32_t decode_launches;  /* kernels launched by the device decode */
  uint32_t group_ranks;      /* ranks that processed this sample together (1 = single GPU) */
  uint32_t shard_blocks, total_blocks; /* BGZF blocks this rank walked / blocks in the file */
  uint32_t range_probes;     /* blocks inflated on the host to find the rank's block range */
  uint32_t tid_begin, tid_end; /* contigs this rank owned for the sample */
  uint32_t decode_second_pass_blocks; /* blocks the first device inflate pass declined or timed out on */
  double gather_s;           /* summary exchange + table gather (wall clock) */
  float decode_copy_enqueue_wall_ms, decode_host_wall_ms; /* host wall clocks inside cmb_submit_bgzf (diagnostics) */
} cmbh_sample_info;

typedef struct cmbh_result {
  int32_t status;        /* process exit status the reference would give: 0, 1 (error!+exit), 2 (usage), 101 (panic) */
  char* out;             /* stdout text (malloc'd; free with cmbh_free_result) */
  size_t out_len;
  char* err;             /* stderr text */
  size_t err_len;
  uint32_t n_samples;
  cmbh_sample_info samples[CMBH_MAX_SAMPLES];
} cmbh_result;

typedef struct cmbh_session cmbh_session;

/* A session owns one GPU context (pinned staging + device arena) and a host thread pool; reuse it across runs. */
int cmbh_session_create(int device, int threads, cmbh_session** out);
void cmbh_session_destroy(cmbh_session* s);
const char* cmbh_last_error(void);
/* Restrict the session to contigs [tid_begin, tid_end) (multi-GPU contig sharding); rows outside come back zero. */
int cmbh_session_set_shard(cmbh_session* s, uint32_t tid_begin, uint32_t tid_end);

/* Multi-GPU, one sample at a time over all GPUs (the reference is single-process, src/contig.rs:22):
 * make this session rank `rank` of `n_ranks`.  Every following cmbh_run is then COLLECTIVE -- each rank calls it with the
 * same argv -- and per sample each rank owns a contig range (balanced by length), uploads and inflates only the BGZF
 * blocks holding it, and one gather completes the per-contig table on every rank, after which the drivers and printers run
 * as on one GPU (global scalars in entry order, src/contig.rs:70-72, src/coverage_printer.rs:457-465) on rank 0, whose
 * cmbh_run returns the text (see cmbh_session_set_group_output).
 *   nccl_id != NULL : the gather runs over NCCL inside the library (id from cmb_comm_unique_id on rank 0, shared by the
 *                     caller -- torch.distributed broadcast, MPI, a file);
 *   nccl_id == NULL : `allgather(user, send, bytes_per_rank, recv)` is the caller's own all-gather of host buffers
 *                     (MPI / gloo between hosts without NCCL; the CPU tests).  Returns 0 on success.
 * n_ranks == 1 leaves the group. */
typedef int (*cmbh_allgather_fn)(void* user, const void* send, size_t bytes_per_rank, void* recv);
int cmbh_session_set_group(cmbh_session* s, int rank, int n_ranks, const uint8_t* nccl_id, cmbh_allgather_fn allgather, void* user);
/* After the gather every rank holds the complete table; by default only rank 0 turns it into text (the other ranks' cmbh_run
 * returns an empty table, status 0).  every_rank_prints != 0 makes all of them print -- the tests use it to show that the
 * gathered table is complete on every rank. */
int cmbh_session_set_group_output(cmbh_session* s, int every_rank_prints);

/* The session's device context (a cmb_ctx* of coverm_b200.h), for callers that continue on the device ABI after a
 * cmbh_run -- e.g. re-running the kernels over the tuples the run left in HBM (cmb_last_bgzf_batch). */
void* cmbh_session_ctx(cmbh_session* s);

int cmbh_run(cmbh_session* s, int argc, const char* const* argv, const cmbh_mem_input* mem, int n_mem, cmbh_result* res);
/* The device parameters (cmb_params of coverm_b200.h: FlagFilter, filter thresholds, contig-end exclusion, trim bounds,
 * wanted statistics) that `coverm <argv>` hands to the device library -- EstimatorsAndTaker / FilterParameters
 * (src/bin/coverm.rs:1315-1504, 1648-1704) reduced to what the kernels need.  `params` points at a cmb_params.
 * For callers that drive the device ABI themselves (bench.py's device-resident timing). */
int cmbh_plan_params(int argc, const char* const* argv, void* params);
void cmbh_free_result(cmbh_result* res);

/* Whole-file tuple extraction on the host (no GPU): the SoA columns of cmb_read_batch for every record of a BAM/SAM
 * file (or memory buffer when data != NULL), for callers that stage tuples in HBM themselves and feed
 * cmb_submit_device_batch.  All arrays are malloc'd; release with cmbh_free_tuples. */
typedef struct cmbh_tuples {
  uint32_t n_contigs;
  uint64_t* contig_len;
  uint64_t n_records;
  uint64_t n_intervals;
  int32_t* tid;
  int32_t* pos;
  uint16_t* flag;
  uint8_t* mapq;
  uint8_t* nm_state;
  uint32_t* nm;
  uint32_t* l_seq;
  uint32_t* aligned;
  uint32_t* del;
  uint32_t* ins;
  uint32_t* iv_begin; /* n_records + 1 */
  int32_t* iv_start;
  int32_t* iv_len;
} cmbh_tuples;
int cmbh_extract_tuples(const char* path, const uint8_t* data, size_t size, int threads, cmbh_tuples* out);
void cmbh_free_tuples(cmbh_tuples* t);

/* The command-line entry point (what the `coverm` binary calls). */
int cmbh_main(int argc, char** argv);

#ifdef __cplusplus
}
#endif
#endif
