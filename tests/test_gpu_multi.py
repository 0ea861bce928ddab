"""Contig-partitioned multi-GPU runs on real hardware (skipped with fewer than 2 GPUs): `coverm --gpus N` drives N GPUs from
one process (cmb_comm_init_local), every rank decodes only its BGZF block range, the in-library NCCL gather completes the
per-contig table, and the printed table must equal the oracle's for the whole file."""
import os
import subprocess

import pytest

import coverm_b200
from case_runner import DATA, ORACLE_BIN

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except Exception:
        return 0


needs2 = pytest.mark.skipif(_n_gpus() < 2, reason="needs at least 2 GPUs")


@pytest.fixture(scope="module")
def bams(tmp_path_factory):
    d = tmp_path_factory.mktemp("multi")
    out = {}
    for name, args in (("small", ["--contigs", "3000", "--reads", "200000", "--seed", "41", "--median-len", "2500", "--min-len", "200", "--max-len", "60000"]),
                       ("wide", ["--contigs", "60000", "--reads", "1500000", "--seed", "42"]),
                       ("mags", ["--contigs", "2500", "--genomes", "60", "--reads", "250000", "--seed", "43", "--median-len", "8000"])):
        p = str(d / f"{name}.bam")
        subprocess.check_call([coverm_b200.BAMGEN_BIN, "--out", p, "--threads", "8"] + args, stdout=subprocess.DEVNULL)
        out[name] = p
    return out


RUNS = [
    ("small", ["contig", "-m", "mean", "trimmed_mean", "covered_fraction", "variance", "count", "rpkm", "tpm"]),
    ("wide", ["contig", "-m", "mean", "trimmed_mean", "covered_fraction"]),
    ("wide", ["contig", "-m", "mean", "variance", "--min-read-percent-identity", "97", "--min-mapq", "20", "--output-format", "sparse"]),
    ("small", ["contig", "-m", "coverage_histogram"]),
    ("mags", ["genome", "-s", "~", "-m", "relative_abundance", "mean", "trimmed_mean", "variance", "--min-covered-fraction", "0"]),
    ("small", ["contig", "-m", "mean", "variance", "--proper-pairs-only", "--min-read-aligned-length-pair", "250"]),
]


@needs2
@pytest.mark.parametrize("which,argv", RUNS, ids=[f"{w}:{' '.join(a[:5])}#{i}" for i, (w, a) in enumerate(RUNS)])
@pytest.mark.parametrize("gpus", [2, 4, 8])
def test_coverm_gpus_matches_the_oracle(bams, which, argv, gpus):
    if _n_gpus() < gpus:
        pytest.skip(f"needs {gpus} GPUs")
    args = argv + ["-b", bams[which], "-t", "8", "--print-reads-mapped"]
    g = subprocess.run([coverm_b200.COVERM_BIN] + args + ["--gpus", str(gpus), "--timing"], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, CMB_PIPELINE_STATS="1"))
    o = subprocess.run([ORACLE_BIN] + args, capture_output=True, text=True, timeout=600)
    assert g.returncode == o.returncode == 0, g.stderr[-1500:]
    assert g.stdout == o.stdout
    rm = lambda p: [l for l in p.stderr.splitlines() if l.startswith("#reads_mapped")]
    assert rm(g) == rm(o)
    if "--proper-pairs-only" not in argv:  # every rank decoded its own block range on the device
        assert g.stderr.count("#device_decode\tblocks=") >= gpus, g.stderr[-1500:]


@needs2
def test_coverm_gpus_errors_like_one_gpu():
    for bam in ("7seqs.reads_for_seq1_and_seq2.bam", "2seqs.reads_for_seq1.with_unmapped.bam"):
        args = ["contig", "-m", "mean", "-b", os.path.join(DATA, bam), "--min-read-percent-identity", "0.9"]
        g = subprocess.run([coverm_b200.COVERM_BIN] + args + ["--gpus", "2"], capture_output=True, text=True, timeout=300)
        o = subprocess.run([ORACLE_BIN] + args, capture_output=True, text=True, timeout=300)
        assert g.returncode == o.returncode and g.stdout == o.stdout, (bam, g.stderr[-500:])
