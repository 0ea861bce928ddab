"""Decoder edge cases on purpose-built BAM files (tests/bam_writer.py): stored (level-0) deflate blocks, maximum
compression with overlapping LZ77 copies, tiny and random BGZF block sizes (records straddling many blocks), 120 kb reads
(blocks with no record start at all), empty blocks, a missing EOF marker, every aux type in front of NM, zero records,
a truncated file and a corrupted block.  The CPU half runs the product's host decoder (with the test-only device emulator)
against the oracle; the GPU half runs the CUDA path twice — device-side decode and CMB_HOST_DECODE=1 — against the oracle."""
import os
import struct
import subprocess

import pytest

import bam_writer as bw
from case_runner import ORACLE_BIN, ROOT

HOSTCHECK = os.path.join(ROOT, "oracle", "coverm_hostcheck")
CONTIGS = [("ctgA", 5000), ("ctgB", 300000), ("ctgC", 64), ("ctgD", 1500000), ("ctgE", 90000)]
METHODS = ["mean", "trimmed_mean", "covered_fraction", "variance", "count", "length", "reads_per_base"]


def _files(tmp_path_factory):
    d = tmp_path_factory.mktemp("edge")
    out = {}

    def put(name, data):
        p = str(d / (name + ".bam"))
        with open(p, "wb") as f:
            f.write(data)
        out[name] = p

    base = bw.bam_stream(CONTIGS, bw.random_records(CONTIGS, 6000, seed=1))
    put("level0_stored", bw.bgzf(base, level=0))
    put("level1", bw.bgzf(base, level=1))
    put("level9_random_blocks", bw.bgzf(base, level=9, block_sizes=(1, 40000), seed=3))
    put("tiny_blocks", bw.bgzf(bw.bam_stream(CONTIGS, bw.random_records(CONTIGS, 1500, seed=9)), level=6, block_sizes=97))
    put("cut_mid_record", bw.bgzf(base[:400000], level=6, block_sizes=(100, 9000), seed=10))  # intact BGZF, last record cut short
    put("no_eof_empty_blocks", bw.bgzf(base, level=6, block_sizes=(2000, 65000), eof=False, empty_block_every=7, seed=4))
    homo = bw.bam_stream(CONTIGS, bw.random_records(CONTIGS, 4000, seed=2, homopolymer=True))
    put("homopolymer_level9", bw.bgzf(homo, level=9))
    longr = bw.bam_stream(CONTIGS, bw.random_records(CONTIGS, 600, seed=5, long_every=5, long_len=400000))
    put("long_reads", bw.bgzf(longr, level=1, block_sizes=(20000, 65000), seed=6))
    put("long_reads_level0", bw.bgzf(longr, level=0))
    rich = bw.bam_stream(CONTIGS, bw.random_records(CONTIGS, 3000, seed=7, rich_tags=True))
    put("rich_tags", bw.bgzf(rich, level=6, block_sizes=(500, 30000), seed=8))
    put("no_records", bw.bgzf(bw.bam_stream(CONTIGS, []), level=6))
    one = bw.bam_stream(CONTIGS, [bw.record(1, 10, [("M", 100)], qname="only")])
    put("one_record", bw.bgzf(one, level=6))
    full = bw.bgzf(base, level=6)
    put("truncated", full[: len(full) * 2 // 3])
    bad = bytearray(full)
    bad[len(bad) // 2] ^= 0x5A
    put("corrupt_block", bytes(bad))
    signed_nm = bw.bam_stream(CONTIGS, [bw.record(1, 10, [("M", 100)], tags=[("NM", "c", 1)])])
    put("nm_signed_type", bw.bgzf(signed_nm, level=6))
    # fixed-size fields that overrun block_size (htslib's bam_read1 fails, the reference panics): an inflated n_cigar_op
    # and an inflated l_read_name, each in the middle of an otherwise healthy stream
    for name, field_off, fmt, value in (("bad_n_cigar", 4 + 12, "<H", 40000), ("bad_l_read_name", 4 + 8, "<B", 250)):
        recs = bw.random_records(CONTIGS, 800, seed=31)
        victim = bytearray(bw.record(1, 100, [("M", 60)], qname="v"))
        victim[field_off:field_off + struct.calcsize(fmt)] = struct.pack(fmt, value)
        recs = sorted(recs[:400], key=lambda r: struct.unpack_from("<ii", r, 4)) + [bytes(victim)] + recs[400:]
        put(name, bw.bgzf(bw.bam_stream(CONTIGS, recs), level=6, block_sizes=(3000, 20000), seed=12))
    put("long_cigar_cg", bw.bgzf(bw.bam_stream(CONTIGS, _long_cigar_records()), level=6))
    return out


LONG_OPS = 70001  # > 65535: stored as `<l_seq>S<reflen>N` + CG:B,I (SAMv1 4.2.2); htslib restores it when reading


def _long_cigar_records():
    ops = []
    for k in range(LONG_OPS // 2):
        ops += [("M", 3), ("I", 1)]
    ops.append(("M", 3))
    l_seq = sum(n for c, n in ops if c in "MI")
    reflen = sum(n for c, n in ops if c == "M")
    cg = [(n << 4) | bw.CIGAR_OPS.index(c) for c, n in ops]
    recs = [bw.record(0, 5, [("M", 100)], qname="before"),
            bw.record(3, 1000, [("S", l_seq), ("N", reflen)], qname="ultralong", tags=[("NM", "I", LONG_OPS // 2), ("CG", "B", ("I", cg))]),
            # the same shape WITHOUT a CG tag is just a soft-clipped read with a skip: no coverage
            bw.record(3, 2000, [("S", 50), ("N", 500)], qname="plain_placeholder_shape"),
            bw.record(4, 10, [("M", 100)], qname="after")]
    return recs


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    return _files(tmp_path_factory)


NAMES = ["level0_stored", "level1", "level9_random_blocks", "tiny_blocks", "no_eof_empty_blocks", "homopolymer_level9", "long_reads",
         "long_reads_level0", "rich_tags", "no_records", "one_record", "truncated", "corrupt_block", "nm_signed_type", "cut_mid_record",
         "bad_n_cigar", "bad_l_read_name", "long_cigar_cg"]
FAILING = {"truncated", "corrupt_block", "nm_signed_type", "cut_mid_record", "bad_n_cigar", "bad_l_read_name"}


def _run(binary, path, env=None, threads="4"):
    return subprocess.run([binary, "contig", "-m"] + METHODS + ["--min-covered-fraction", "0", "-b", path, "-t", threads, "--print-reads-mapped"],
                          capture_output=True, text=True, timeout=600, env=dict(os.environ, **(env or {})))


def _same(a, o, name):
    if name in FAILING:  # the reference panics / errors out: no table, non-zero status
        assert a.returncode != 0 and o.returncode != 0, (name, a.returncode, o.returncode, a.stderr[-300:])
        return
    assert a.returncode == o.returncode == 0, (name, a.returncode, o.returncode, a.stderr[-500:], o.stderr[-300:])
    assert a.stdout == o.stdout, name
    rm = lambda p: [l for l in p.stderr.splitlines() if l.startswith("#reads_mapped")]
    assert rm(a) == rm(o), name


@pytest.mark.parametrize("name", NAMES)
def test_host_decoder_edge_cases(files, name):
    if not os.path.exists(HOSTCHECK):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    _same(_run(HOSTCHECK, files[name], threads="3"), _run(ORACLE_BIN, files[name]), name)


def test_long_cigar_is_restored_from_the_cg_tag(files):
    """htslib (bam_tag2cigar) swaps the CG:B,I array in for the placeholder CIGAR: the 35 001 M blocks of 3 bases must be counted."""
    for binary in (ORACLE_BIN, HOSTCHECK):
        p = subprocess.run([binary, "contig", "-m", "covered_bases", "count", "--contig-end-exclusion", "0", "-b", files["long_cigar_cg"], "-t", "2"],
                           capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, p.stderr[-500:]
        rows = dict(l.split("\t", 1) for l in p.stdout.splitlines()[1:])
        assert rows["ctgD"] == "%d\t2" % (3 * (LONG_OPS // 2 + 1)), (binary, rows)
        assert rows["ctgA"] == "100\t1" and rows["ctgE"] == "100\t1"


def test_malformed_record_layout_is_a_read_error(files):
    for name in ("bad_n_cigar", "bad_l_read_name"):
        for binary in (ORACLE_BIN, HOSTCHECK):
            p = _run(binary, files[name], threads="3")
            assert p.returncode == 101 and "Error reading BAM record" in p.stderr, (name, binary, p.returncode, p.stderr[-300:])


@pytest.mark.parametrize("name", NAMES)
def test_device_decode_branch_edge_cases(files, name):
    """The host's device-decode branch with the emulator as the decoder (CMB_EMU_BGZF=1): declines must fall back cleanly."""
    _same(_run(HOSTCHECK, files[name], env={"CMB_EMU_BGZF": "1", "CMB_PIPELINE_STATS": "1"}, threads="3"), _run(ORACLE_BIN, files[name]), name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_device_decoder_edge_cases(files, name):
    import coverm_b200
    o = _run(ORACLE_BIN, files[name])
    g = _run(coverm_b200.COVERM_BIN, files[name], env={"CMB_PIPELINE_STATS": "1", "CMB_DECODE_VERIFY": "1"})
    _same(g, o, name)
    if name not in FAILING and name != "no_records":
        st = [l for l in g.stderr.splitlines() if l.startswith("#decode_") or l.startswith("#device_decode")]
        assert any(l.startswith("#decode_verify\t0 of ") for l in st), st  # device-inflated bytes == zlib's, block by block
    _same(_run(coverm_b200.COVERM_BIN, files[name], env={"CMB_HOST_DECODE": "1"}), o, name)


# ---- many-contig inputs share the header's names with the cached taker; a second sample with a different reference must
#      still fail exactly like the reference ("Found a difference amongst the reference sets", coverage_takers.rs:138-149)
@pytest.fixture(scope="module")
def many_contig_pair(tmp_path_factory):
    d = tmp_path_factory.mktemp("refs")
    a = [("c%05d" % i, 400 + i % 300) for i in range(2500)]
    b = list(a)
    b[1700] = ("different", b[1700][1])
    paths = []
    for k, contigs in enumerate((a, a, b)):
        recs = bw.random_records(contigs, 3000, seed=20 + k, read_len=(30, 120))
        p = str(d / f"s{k}.bam")
        with open(p, "wb") as f:
            f.write(bw.bgzf(bw.bam_stream(contigs, recs, text="@HD\tVN:1.6\n@PG\tID:run%d\n" % k), level=6))
        paths.append(p)
    return paths


def _multi(binary, paths, extra=(), env=None):
    return subprocess.run([binary, "contig", "-m", "mean", "covered_fraction", "-t", "3", "-b"] + paths + list(extra), capture_output=True,
                          text=True, timeout=600, env=dict(os.environ, **(env or {})))


def test_shared_names_across_samples_host(many_contig_pair):
    s0, s1, sdiff = many_contig_pair
    for paths, extra in (([s0, s1], []), ([s0, s1], ["--no-zeros"]), ([s0, sdiff], []), ([s0, s1, sdiff], ["--output-format", "sparse"])):
        a, o = _multi(HOSTCHECK, paths, extra), _multi(ORACLE_BIN, paths, extra)
        assert a.returncode == o.returncode, (paths, a.stderr[-300:], o.stderr[-300:])
        assert a.stdout == o.stdout
        if o.returncode != 0:
            assert "Found a difference amongst the reference sets" in a.stderr and "different" in a.stderr


@pytest.mark.gpu
def test_shared_names_across_samples_gpu(many_contig_pair):
    import coverm_b200
    s0, s1, sdiff = many_contig_pair
    for paths in ([s0, s1], [s0, sdiff]):
        a, o = _multi(coverm_b200.COVERM_BIN, paths), _multi(ORACLE_BIN, paths)
        assert a.returncode == o.returncode and a.stdout == o.stdout, (paths, a.stderr[-300:])


# ---- record-level oddities the filters and the delta accumulation must treat exactly like the reference
def _odd_records():
    contigs = [("one", 1), ("short", 40), ("mid", 3000), ("unused", 500)]
    R = bw.record
    recs = [
        R(0, 0, [("M", 1)], qname="on_len1"),
        R(0, 0, [("M", 30)], qname="overhang_len1"),                       # runs off a 1-base contig
        R(1, 39, [("M", 20)], qname="last_base"),                          # starts on the last base
        R(1, 5, [("S", 10), ("M", 10), ("H", 3)], qname="clips", tags=[("NM", "C", 0)]),
        R(2, 10, [("M", 50)], l_seq=0, qname="star_seq", tags=[("NM", "C", 2)]),   # SEQ '*': l_seq 0 -> aligned / 0 in the filters
        R(2, 20, [("I", 30)], qname="insert_only", tags=[("NM", "C", 30)]),        # no reference-consuming op at all
        R(2, 30, [("M", 10), ("D", 2990)], qname="deletion_to_end", tags=[("NM", "S", 2990)]),
        R(2, 40, [("M", 100)], mapq=255, qname="mapq255"),
        R(2, 50, [("M", 100)], mapq=0, qname="mapq0", tags=[("NM", "I", 70000)]),  # NM larger than the alignment
        R(2, 60, [("=", 40), ("X", 1), ("=", 59)], flag=99, mtid=2, mpos=200, tlen=240, qname="pair1", tags=[("NM", "C", 1)]),
        R(2, 200, [("M", 100)], flag=147, mtid=2, mpos=60, tlen=-240, qname="pair1", tags=[("NM", "C", 0)]),
        R(2, 300, [("M", 100)], flag=0x4 | 0x1 | 0x40, qname="unmapped_placed"),   # flag says unmapped but it carries a position
        R(2, 310, [], flag=0x4, l_seq=20, qname="unmapped_no_cigar"),
        R(2, 400, [("M", 100)], flag=0x100, qname="secondary"),
        R(2, 500, [("M", 100)], flag=0x800, qname="supplementary"),
        R(2, 2950, [("M", 100)], qname="overhang_mid"),
        R(-1, -1, [], flag=0x4 | 0x1 | 0x40, l_seq=30, qname="unmapped_tail1"),
        R(-1, -1, [], flag=0x4 | 0x1 | 0x80, l_seq=30, qname="unmapped_tail2"),
    ]
    return contigs, recs


ODD_ARGS = [
    [],
    ["--min-read-percent-identity", "95"],
    ["--min-read-aligned-percent", "50", "--min-read-aligned-length", "20"],
    ["--min-mapq", "1"],
    ["--proper-pairs-only"],
    ["--include-secondary", "--exclude-supplementary"],
    ["--contig-end-exclusion", "0", "--trim-min", "0", "--trim-max", "1"],
    ["--contig-end-exclusion", "20"],
]


@pytest.fixture(scope="module")
def odd_bam(tmp_path_factory):
    contigs, recs = _odd_records()
    p = str(tmp_path_factory.mktemp("odd") / "odd.bam")
    with open(p, "wb") as f:
        f.write(bw.bgzf(bw.bam_stream(contigs, recs), level=6, block_sizes=(60, 700), seed=2))
    return p


def _odd(binary, path, extra, env=None):
    return subprocess.run([binary, "contig", "-m", "mean", "trimmed_mean", "covered_fraction", "covered_bases", "variance", "rpkm", "tpm",
                           "--min-covered-fraction", "0", "-b", path, "-t", "2", "--print-reads-mapped"] + extra,
                          capture_output=True, text=True, timeout=300, env=dict(os.environ, **(env or {})))


@pytest.mark.parametrize("extra", ODD_ARGS, ids=[" ".join(a) or "defaults" for a in ODD_ARGS])
def test_record_oddities_host(odd_bam, extra):
    a, o = _odd(HOSTCHECK, odd_bam, extra), _odd(ORACLE_BIN, odd_bam, extra)
    assert a.returncode == o.returncode, (a.stderr[-400:], o.stderr[-400:])
    assert a.stdout == o.stdout
    assert [l for l in a.stderr.splitlines() if l.startswith("#reads")] == [l for l in o.stderr.splitlines() if l.startswith("#reads")]


@pytest.mark.gpu
@pytest.mark.parametrize("extra", ODD_ARGS, ids=[" ".join(a) or "defaults" for a in ODD_ARGS])
def test_record_oddities_gpu(odd_bam, extra):
    import coverm_b200
    o = _odd(ORACLE_BIN, odd_bam, extra)
    for env in ({}, {"CMB_HOST_DECODE": "1"}):
        g = _odd(coverm_b200.COVERM_BIN, odd_bam, extra, env=env)
        assert g.returncode == o.returncode, (env, g.stderr[-400:], o.stderr[-400:])
        assert g.stdout == o.stdout, env
        assert [l for l in g.stderr.splitlines() if l.startswith("#reads")] == [l for l in o.stderr.splitlines() if l.startswith("#reads")]
