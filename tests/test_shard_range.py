"""The multi-GPU block-range search (coverm_b200/csrc/host/shard_range.hpp) against a brute-force walk of the whole file
(tests/native/shard_range_check.cpp): every block's speculative "first record that starts here", every rank's block range at
2/3/4/7/8 ranks, and the rule that a start hint — right, wrong or random — never changes a result."""
import os
import random
import subprocess

import pytest

import bam_writer as bw
import coverm_b200
from case_runner import ROOT as REPO

SRC = os.path.join(REPO, "tests", "native", "shard_range_check.cpp")


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("shard") / "shard_range_check")
    inc = os.path.join(REPO, "coverm_b200", "csrc")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", inc, "-I", os.path.join(inc, "host"), SRC, "-lz", "-lpthread", "-o", exe], check=True)
    return exe


def _check(exe, path):
    for hint_min_blocks in ("0", "1024"):  # 0: hints steer the search even in small files
        r = subprocess.run([exe, path, hint_min_blocks], capture_output=True, text=True)
        assert r.returncode == 0 and r.stdout.startswith("ok "), (path, hint_min_blocks, r.stdout, r.stderr[-800:])


@pytest.mark.parametrize("args", [
    ["--contigs", "3000", "--reads", "200000", "--seed", "5", "--median-len", "2500", "--min-len", "200", "--max-len", "60000"],
    ["--contigs", "12", "--reads", "100000", "--seed", "6", "--median-len", "900000", "--sigma", "0.6", "--min-len", "50000", "--max-len", "5000000"],
    ["--contigs", "20000", "--reads", "400000", "--seed", "9", "--median-len", "4000", "--sigma", "0.8", "--min-len", "1000", "--max-len", "2000000"],
], ids=["many-small-contigs", "few-long-contigs", "wide"])
def test_block_ranges_on_generated_bams(checker, tmp_path, args):
    p = str(tmp_path / "x.bam")
    subprocess.run([coverm_b200.BAMGEN_BIN, "--out", p, "--threads", "4"] + args, check=True, stdout=subprocess.DEVNULL)
    _check(checker, p)


def _skewed(block_sizes, long_reads=False):
    """Reads only on the last third of the contigs behind a large empty one (a length-based hint is far off), every record ending
    in `NM:C:0` -- the four bytes in front of a record then read as a 4.4 MB block_size, the stray-header case of the search --
    and unplaced records at the end."""
    rng = random.Random(3)
    contigs = [("big", 1_500_000)] + [(f"c{i}", rng.randint(500, 4000)) for i in range(900)]
    recs = []
    for i in range(600, 901):
        L = contigs[i][1]
        rows = []
        for k in range(rng.randint(0, 60)):
            rl = rng.randint(40, 140)
            if long_reads and k % 17 == 5:
                rl = min(L - 1, 3000)
            pos = rng.randint(0, max(0, L - rl - 1))
            rows.append((pos, bw.record(i, pos, [("M", rl)], qname=f"read{i}_{k}_{rng.randint(0, 10 ** rng.randint(1, 6))}",
                                        flag=rng.choice([0, 16, 99, 147]), tags=(("XS", "i", -k), ("NM", "C", 0)), rng=rng)))
        rows.sort(key=lambda x: x[0])
        recs += [r for _, r in rows]
    for k in range(50):
        recs.append(bw.record(-1, -1, [], flag=4, qname=f"u{k}", l_seq=50, rng=rng))
    return bw.bgzf(bw.bam_stream(contigs, recs), level=6, block_sizes=block_sizes, seed=4)


@pytest.mark.parametrize("block_sizes,long_reads", [(9000, False), (None, False), ((2000, 30000), True), (1500, True), ((300, 2500), True)],
                         ids=["9k-blocks", "64k-blocks", "ragged-blocks-long-reads", "blocks-shorter-than-reads", "tiny-ragged-blocks"])
def test_block_ranges_with_stray_headers_and_a_bad_hint(checker, tmp_path, block_sizes, long_reads):
    p = str(tmp_path / "skew.bam")
    open(p, "wb").write(_skewed(block_sizes, long_reads))
    _check(checker, p)
