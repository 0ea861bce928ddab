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


# ---------------------------------------------------------------------------------------------- group mode (block-range split)
GROUP_WORKER = r'''
import os, sys, json
sys.path.insert(0, sys.argv[1])
import torch
import torch.distributed as dist
import coverm_b200
rank, world, port, lib_path = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
runs = json.loads(sys.argv[6])
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
lib = coverm_b200.load_library(lib_path)

def allgather(send):  # the caller-supplied host all-gather of cmbh_session_set_group (here: gloo)
    mine = torch.frombuffer(bytearray(send), dtype=torch.uint8)
    outs = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(outs, mine)
    return b"".join(bytes(o.numpy()) for o in outs)

sess = coverm_b200.Session(device=0, threads=2, lib=lib)
sess.set_group(rank, world, allgather=allgather)
sess.set_group_output(every_rank_prints=(os.environ.get("RANK0_ONLY") != "1"))
results = []
for argv in runs:
    r = sess.run(argv + ["-t", "2", "--print-reads-mapped"])
    s = r.samples[0] if r.samples else {}
    results.append({"status": r.status, "out": r.out, "rm": [l for l in r.err.splitlines() if l.startswith("#reads_mapped")],
                    "err": r.err[-300:] if r.status else "", "device_decode": s.get("device_decode"), "ranks": s.get("group_ranks"),
                    "shard_blocks": s.get("shard_blocks"), "total_blocks": s.get("total_blocks"), "n_records": s.get("n_records")})
sess.close()
print(json.dumps(results))
dist.destroy_process_group()
'''


def _run_group(tmp_path, world, runs, env=None):
    if not os.path.exists(EMU_LIB):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    import json
    script = tmp_path / "group_worker.py"
    script.write_text(GROUP_WORKER)
    port = str(_free_port())
    e = dict(os.environ, **(env or {}))
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(r), str(world), port, EMU_LIB, json.dumps(runs)],
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=e) for r in range(world)]
    outs = [p.communicate(timeout=600) for p in procs]
    for p, (o, er) in zip(procs, outs):
        assert p.returncode == 0, er[-3000:]
    return [json.loads(o.strip().splitlines()[-1]) for o, _ in outs]


def _oracle(argv):
    p = subprocess.run([ORACLE_BIN] + argv + ["--print-reads-mapped"], capture_output=True, text=True)
    return p.returncode, p.stdout, [l for l in p.stderr.splitlines() if l.startswith("#reads_mapped")]


@pytest.fixture(scope="module")
def group_bams(tmp_path_factory):
    import coverm_b200
    d = tmp_path_factory.mktemp("group")
    small = str(d / "small.bam")
    subprocess.check_call([coverm_b200.BAMGEN_BIN, "--out", small, "--threads", "4", "--contigs", "3000", "--reads", "120000", "--seed", "21",
                           "--median-len", "2500", "--min-len", "200", "--max-len", "60000"], stdout=subprocess.DEVNULL)
    mags = str(d / "mags.bam")
    subprocess.check_call([coverm_b200.BAMGEN_BIN, "--out", mags, "--threads", "4", "--contigs", "1200", "--genomes", "30", "--reads", "90000",
                           "--seed", "22", "--median-len", "8000"], stdout=subprocess.DEVNULL)
    return {"small": small, "mags": mags}


GROUP_RUNS = lambda b: [
    ["contig", "-m", "mean", "trimmed_mean", "covered_fraction", "variance", "count", "rpkm", "tpm", "-b", b["small"]],
    ["contig", "-m", "mean", "trimmed_mean", "--min-read-percent-identity", "97", "--min-mapq", "20", "-b", b["small"], DATA + "/1.bam"][:-1],
    ["contig", "-m", "coverage_histogram", "-b", DATA + "/1.bam"],
    ["contig", "-m", "mean", "covered_bases", "-b", DATA + "/eg2.bam", DATA + "/7seqs.reads_for_seq1_and_seq2.bam"][:-1],
    ["genome", "-s", "~", "-m", "relative_abundance", "mean", "trimmed_mean", "variance", "--min-covered-fraction", "0", "-b", b["mags"]],
    ["contig", "-m", "mean", "variance", "--proper-pairs-only", "--min-read-aligned-length-pair", "250", "-b", b["small"]],  # host mate matching on every rank
    ["contig", "-m", "mean", "-b", DATA + "/2seqs.bad_read.1.bam", "--min-read-percent-identity", "0.5"][:5] + ["--min-read-percent-identity", "0.5"],
]


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("emu_bgzf", ["1", None], ids=["ranged-device-decode", "host-decode"])
def test_group_mode_matches_the_oracle(tmp_path, group_bams, world, emu_bgzf):
    """Every rank of a group returns the table the oracle prints for the whole file: contigs range-partitioned by length,
    (emulated) device decode of each rank's own BGZF block range, rank summaries + table gather through the caller's
    all-gather (gloo), whole-file counters summed over the ranks' owned records."""
    runs = GROUP_RUNS(group_bams)
    res = _run_group(tmp_path, world, runs, env={"CMB_EMU_BGZF": emu_bgzf} if emu_bgzf else {"CMB_EMU_BGZF": ""})
    for i, argv in enumerate(runs):
        rc, out, rm = _oracle(argv)
        for r in range(world):
            got = res[r][i]
            assert got["status"] == rc, (argv, r, got["err"])
            assert got["out"] == out, (argv, r)
            assert got["rm"] == rm, (argv, r, got["rm"], rm)
            assert got["ranks"] == world
    if emu_bgzf:
        first = [res[r][0] for r in range(world)]
        assert all(x["device_decode"] == 1 for x in first)
        # each rank walked only its share of the blocks (plus the one-block overlap), and together they cover the file
        assert all(x["shard_blocks"] < x["total_blocks"] for x in first), first
        assert sum(x["shard_blocks"] for x in first) >= first[0]["total_blocks"] - 2


def test_group_mode_eight_ranks(tmp_path, group_bams):
    """World size 8 (`coverm --gpus 8`, the driver's 8-GPU scaling run): more ranks than some inputs have contigs or BGZF blocks,
    so several ranks own nothing -- their (empty) ranges, summaries and gathers must still line up."""
    runs = [GROUP_RUNS(group_bams)[i] for i in (0, 2, 3, 4, 5)]
    res = _run_group(tmp_path, 8, runs, env={"CMB_EMU_BGZF": "1"})
    for i, argv in enumerate(runs):
        rc, out, rm = _oracle(argv)
        for r in range(8):
            got = res[r][i]
            assert got["status"] == rc, (argv, r, got["err"])
            assert got["out"] == out, (argv, r)
            assert got["rm"] == rm, (argv, r, got["rm"], rm)
            assert got["ranks"] == 8


def test_group_mode_errors_are_collective(tmp_path):
    """A failure on any rank (here: a record without NM where the reference calls nm(), an unsorted file) reaches every rank
    as the error the reference raises."""
    runs = [["contig", "-m", "mean", "-b", DATA + "/7seqs.reads_for_seq1_and_seq2.bam", "--min-read-percent-identity", "0.9"],
            ["contig", "-m", "mean", "-b", DATA + "/2seqs.reads_for_seq1.with_unmapped.bam"]]
    unsorted = [c for c in __import__("reference_cases").CLI_CASES if c.get("status") == 101]
    res = _run_group(tmp_path, 2, runs, env={"CMB_EMU_BGZF": "1"})
    for i, argv in enumerate(runs):
        rc, out, rm = _oracle(argv)
        for r in range(2):
            assert res[r][i]["status"] == rc and res[r][i]["out"] == out, (argv, r, res[r][i]["err"])


def test_group_mode_only_rank0_prints_by_default(tmp_path, group_bams):
    runs = GROUP_RUNS(group_bams)[:2]
    res = _run_group(tmp_path, 2, runs, env={"CMB_EMU_BGZF": "1", "RANK0_ONLY": "1"})
    for i, argv in enumerate(runs):
        rc, out, rm = _oracle(argv)
        assert res[0][i]["status"] == rc and res[0][i]["out"] == out and res[0][i]["rm"] == rm
        assert res[1][i]["status"] == 0 and res[1][i]["out"] == ""
