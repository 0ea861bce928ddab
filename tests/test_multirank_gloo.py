"""world_size-2 `gloo` test of the multi-rank host logic (runs on CPU):
 * sample-per-rank: each rank runs `coverm contig` on its own BAM, the tables are all-gathered;
 * contig sharding: both ranks read the same BAM but own half of the contigs each (cmbh_session_set_shard), the
   per-rank tables are gathered and merged row-wise by ownership.
The device half is the test-only CPU emulator (oracle/libcoverm_hostcheck.so); results are checked against the oracle.
On the GPU the same logic runs in bench.py with NCCL (one process per GPU)."""
import os
import socket
import subprocess
import sys

import pytest

from case_runner import DATA, ORACLE_BIN, ROOT

EMU_LIB = os.path.join(ROOT, "oracle", "libcoverm_hostcheck.so")
WORKER = r'''
import os, sys, json
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
import coverm_b200
rank, world, port, lib_path = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
bams = json.loads(sys.argv[6])
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
lib = coverm_b200.load_library(lib_path)
methods = ["mean", "trimmed_mean", "covered_fraction", "variance", "count"]
# ---- one sample per rank
sess = coverm_b200.Session(device=0, threads=2, lib=lib)
r = sess.run(["contig", "-m"] + methods + ["-b", bams[rank], "-t", "2"])
tables = [None] * world
dist.all_gather_object(tables, (r.status, r.out))
# ---- contig sharding of one sample
n_contigs = 7
cut = [0, 3, n_contigs]
sess.set_shard(cut[rank], cut[rank + 1])
s = sess.run(["contig", "-m"] + methods + ["-b", bams[0], "-t", "2"])
shards = [None] * world
dist.all_gather_object(shards, (s.status, s.out))
sess.close()
if rank == 0:
    print(json.dumps({"tables": tables, "shards": shards, "cut": cut}))
dist.destroy_process_group()
'''


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_ranks_gloo(tmp_path):
    if not os.path.exists(EMU_LIB):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    import json
    bams = [DATA + "/7seqs.reads_for_seq1_and_seq2.bam", DATA + "/7seqs.fnaVbad_read.bam"]
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(_free_port())
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(r), "2", port, EMU_LIB, json.dumps(bams)],
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-2000:]
    res = json.loads(outs[0][0].strip().splitlines()[-1])
    methods = ["mean", "trimmed_mean", "covered_fraction", "variance", "count"]
    want = [subprocess.run([ORACLE_BIN, "contig", "-m"] + methods + ["-b", b], capture_output=True, text=True,
                           check=True).stdout for b in bams]
    # sample-per-rank: every rank's table equals the oracle's for its sample
    for (status, out), w in zip(res["tables"], want):
        assert status == 0 and out == w
    # contig sharding: row i comes from the rank that owns contig i
    cut = res["cut"]
    rows = [out.splitlines() for _, out in res["shards"]]
    merged = [rows[0][0]]
    for i in range(7):
        owner = 0 if i < cut[1] else 1
        merged.append(rows[owner][1 + i])
    assert "\n".join(merged) + "\n" == want[0]
    # a rank prints zero rows for contigs it does not own
    assert rows[1][1 + 2].split("\t")[1:] == ["0"] * 5 and rows[0][1 + 5].split("\t")[1:] == ["0"] * 5
