"""Per-gene coverage (`--gff`, src/genes.rs): random gene sets -- overlapping, nested, single-base, running past the contig end,
on unknown contigs, without ids, GTF-style attributes, filtered by feature type -- over synthetic BAMs, product vs the oracle
(whose gene driver is pinned by the reference's own tests in tests/golden/reference_cases.py GENE_CASES).
CPU: the product's host code with the device emulator; GPU: the CUDA library (cmb_set_genes: blocks clipped per gene in K1)."""
import os
import random
import struct
import subprocess
import zlib

import pytest

import coverm_b200
from case_runner import ORACLE_BIN, ROOT

HOSTCHECK = os.path.join(ROOT, "oracle", "coverm_hostcheck")
METHODS = ["mean", "trimmed_mean", "covered_fraction", "covered_bases", "variance", "length", "count", "reads_per_base", "rpkm", "tpm", "anir"]


def bam_header(path):
    """(name, length) of every reference sequence: inflate BGZF members until the header is complete."""
    raw = open(path, "rb").read()
    data, o = b"", 0
    while o < len(raw):
        bsize = struct.unpack_from("<H", raw, o + 16)[0] + 1
        data += zlib.decompress(raw[o + 18:o + bsize - 8], -15)
        o += bsize
        if len(data) >= 12:
            l_text = struct.unpack_from("<I", data, 4)[0]
            if len(data) >= 12 + l_text:
                n_ref = struct.unpack_from("<I", data, 8 + l_text)[0]
                p, refs = 12 + l_text, []
                try:
                    for _ in range(n_ref):
                        l_name = struct.unpack_from("<I", data, p)[0]
                        name = data[p + 4:p + 4 + l_name - 1].decode()
                        refs.append((name, struct.unpack_from("<I", data, p + 4 + l_name)[0]))
                        p += 8 + l_name
                    return refs
                except struct.error:
                    continue
    raise AssertionError("no BAM header")


def write_gff(path, refs, n_genes, seed):
    rng = random.Random(seed)
    lines = ["##gff-version 3"]
    for i in range(n_genes):
        name, L = rng.choice(refs)
        kind = rng.random()
        if kind < 0.6:
            a = rng.randint(1, L)
            b = min(L + 50, a + rng.randint(0, 3000))  # some run past the contig end
        elif kind < 0.7:
            a = b = rng.randint(1, L)  # single base
        elif kind < 0.8:
            a, b = 1, L  # the whole contig
        elif kind < 0.85:
            a, b = L + 5, L + 100  # entirely outside: dropped
        else:
            a = rng.randint(1, max(1, L // 2))
            b = a + rng.randint(1, max(1, L // 2))  # long, overlaps many others
        ftype = rng.choice(["gene", "gene", "CDS", "tRNA"])
        attr = rng.choice([f"ID=g{i};Name=n{i}", f"locus_tag=lt{i}", f'gene_id "gtf{i}"; transcript_id "t{i}"', "Note=no id here", "", f"ID=;Name=named{i}", f"Parent=p{i}"])
        if rng.random() < 0.03:
            name = "not_in_the_reference"
        cols = [name, "test", ftype, str(a), str(b), ".", rng.choice("+-"), ".", attr]
        if attr == "" and rng.random() < 0.5:
            cols = cols[:8]
        lines.append("\t".join(cols))
        if rng.random() < 0.02:
            lines.append("# a comment")
        if rng.random() < 0.02:
            lines.append("malformed\tline")
        if rng.random() < 0.02:
            lines.append("\t".join([name, "test", "gene", "0", "10", ".", "+", ".", "ID=zero_start"]))
        if rng.random() < 0.02:
            lines.append("\t".join([name, "test", "gene", "x", "10", ".", "+", ".", "ID=bad_start"]))
    open(path, "w").write("\n".join(lines) + "\n")


@pytest.fixture(scope="module")
def gene_inputs(tmp_path_factory):
    d = tmp_path_factory.mktemp("genes")
    out = {}
    for name, args, n_genes in (("small", ["--contigs", "400", "--reads", "60000", "--seed", "51", "--median-len", "3000", "--min-len", "200", "--max-len", "40000"], 1500),
                                ("long", ["--contigs", "6", "--reads", "80000", "--seed", "52", "--median-len", "300000", "--sigma", "0.5", "--min-len", "50000", "--max-len", "900000"], 800),
                                ("mags", ["--contigs", "300", "--genomes", "12", "--reads", "50000", "--seed", "53", "--median-len", "6000"], 900)):
        bam = str(d / f"{name}.bam")
        subprocess.check_call([coverm_b200.BAMGEN_BIN, "--out", bam, "--threads", "4"] + args, stdout=subprocess.DEVNULL)
        gff = str(d / f"{name}.gff")
        write_gff(gff, bam_header(bam), n_genes, seed=len(name))
        out[name] = (bam, gff)
    return out


RUNS = [
    ("small", ["contig", "-m"] + METHODS),
    ("small", ["contig", "-m", "mean", "trimmed_mean", "variance", "--contig-end-exclusion", "0", "--no-zeros", "--output-format", "sparse"]),
    ("small", ["contig", "-m", "mean", "count", "--gff-feature-type", "CDS", "--min-read-percent-identity", "97"]),
    ("small", ["contig", "-m", "coverage_histogram"]),
    ("long", ["contig", "-m"] + METHODS + ["--contig-end-exclusion", "10"]),
    ("long", ["contig", "-m", "mean", "covered_bases", "--exclude-supplementary", "--include-secondary", "--min-mapq", "20"]),
    ("mags", ["genome", "-s", "~", "-m", "mean", "trimmed_mean", "count", "relative_abundance", "--min-covered-fraction", "0"]),
    ("mags", ["genome", "--single-genome", "-m", "mean", "variance", "--min-covered-fraction", "0", "--output-format", "sparse"]),
    ("small", ["contig", "-m", "mean", "--proper-pairs-only", "--min-read-aligned-length-pair", "250"]),
]


def _same(binary, argv, bam, gff, env=None):
    args = argv + ["-b", bam, "--gff", gff, "-t", "3", "--print-reads-mapped"]
    g = subprocess.run([binary] + args, capture_output=True, text=True, timeout=600, env=dict(os.environ, **(env or {})))
    o = subprocess.run([ORACLE_BIN] + args, capture_output=True, text=True, timeout=600)
    assert g.returncode == o.returncode, (argv, g.stderr[-800:], o.stderr[-400:])
    if g.stdout != o.stdout:
        gl, ol = g.stdout.splitlines(), o.stdout.splitlines()
        diff = [(i, a, b) for i, (a, b) in enumerate(zip(gl, ol)) if a != b][:6]
        raise AssertionError(f"{argv}: {len(gl)} vs {len(ol)} lines; first differences {diff}")
    rm = lambda p: [l for l in p.stderr.splitlines() if l.startswith("#reads_mapped")]
    assert rm(g) == rm(o), (rm(g), rm(o))
    assert o.stdout.count("\n") > 3
    return g


@pytest.mark.parametrize("which,argv", RUNS, ids=[f"{w}:{' '.join(a[:6])}#{i}" for i, (w, a) in enumerate(RUNS)])
@pytest.mark.parametrize("emu_bgzf", [None, "1"], ids=["host-decode", "device-decode-branch"])
def test_gene_coverage_host_code_matches_oracle(gene_inputs, which, argv, emu_bgzf):
    if not os.path.exists(HOSTCHECK):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    bam, gff = gene_inputs[which]
    _same(HOSTCHECK, argv, bam, gff, env={"CMB_EMU_BGZF": emu_bgzf} if emu_bgzf else {"CMB_EMU_BGZF": ""})


@pytest.mark.gpu
@pytest.mark.parametrize("which,argv", RUNS, ids=[f"{w}:{' '.join(a[:6])}#{i}" for i, (w, a) in enumerate(RUNS)])
def test_gene_coverage_cuda_matches_oracle(gene_inputs, which, argv):
    bam, gff = gene_inputs[which]
    _same(coverm_b200.COVERM_BIN, argv, bam, gff)
    _same(coverm_b200.COVERM_BIN, argv, bam, gff, env={"CMB_HOST_DECODE": "1"})
