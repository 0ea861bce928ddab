"""The product's HOST code (BGZF/BAM/SAM decode, batching, mate matching, driver replay, estimator
finalisation, printers, CLI) against the reference's golden vectors, with the device half replaced by the
test-only CPU emulator (oracle/device_emulator.cpp -> oracle/coverm_hostcheck).  Runs without a GPU.
The CUDA kernels themselves are checked on the GPU in tests/test_gpu_parity.py."""
import os
import subprocess

import pytest

from case_runner import ROOT, check_case, run_case
from reference_cases import CASES, CLI_CASES

HOSTCHECK = os.path.join(ROOT, "oracle", "coverm_hostcheck")
from reference_cases import FILTER_CASES  # noqa: E402

HOST_CASES = [c for c in CASES + CLI_CASES + FILTER_CASES if c["sub"] in ("contig", "genome", "filter-names")]


@pytest.fixture(scope="session", autouse=True)
def _built():
    if not os.path.exists(HOSTCHECK):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])


@pytest.mark.parametrize("case", HOST_CASES, ids=[f"{c['sub']}@{c['ref']}" for c in HOST_CASES])
@pytest.mark.parametrize("threads", ["1", "3"])
def test_host_code_matches_reference_golden(case, threads):
    check_case(case, run_case(HOSTCHECK, case, extra_args=["-t", threads]))


# Many-contig inputs take the host's table-building path (estimator maths on all threads, taker fed in tid order):
# compare with the oracle on the reference's 54 579-contig fixture.
@pytest.mark.parametrize("extra", [[], ["--no-zeros"], ["--output-format", "sparse"], ["--min-covered-fraction", "0.3"]],
                         ids=["dense", "no-zeros", "sparse", "min-covered"])
@pytest.mark.parametrize("threads", ["1", "5"])
def test_many_contig_table_path_matches_oracle(extra, threads):
    from case_runner import DATA, ORACLE_BIN
    argv = ["contig", "-m", "mean", "trimmed_mean", "covered_fraction", "variance", "rpkm", "tpm", "-b",
            os.path.join(DATA, "eg2.bam")] + extra
    a = subprocess.run([HOSTCHECK] + argv + ["-t", threads, "--print-reads-mapped"], capture_output=True, text=True)
    b = subprocess.run([ORACLE_BIN] + argv + ["--print-reads-mapped"], capture_output=True, text=True)
    assert a.returncode == b.returncode == 0, a.stderr[-500:]
    assert a.stdout == b.stdout
    rm = lambda p: [l for l in p.stderr.splitlines() if l.startswith("#reads_mapped")]
    assert rm(a) == rm(b)


def test_in_memory_sink_bulk_path_through_the_c_abi():
    """cmbh_run's in-memory output sink takes formatted row blocks in parallel (more than one 4096-row block): the text
    must equal the CLI's and the oracle's.  Uses the emulator build of the same ABI (no GPU)."""
    import numpy as np

    import coverm_b200
    from case_runner import DATA, ORACLE_BIN
    lib = coverm_b200.load_library(os.path.join(ROOT, "oracle", "libcoverm_hostcheck.so"))
    sess = coverm_b200.Session(device=0, threads=4, lib=lib)
    bam = os.path.join(DATA, "eg2.bam")
    argv = ["contig", "-m", "mean", "trimmed_mean", "covered_fraction", "-b", bam, "-t", "4"]
    r1 = sess.run(argv)
    r2 = sess.run(argv, memory_inputs={bam: np.fromfile(bam, dtype=np.uint8)})  # second sample reuses the parsed reference list
    sess.close()
    want = subprocess.run([ORACLE_BIN] + argv, capture_output=True, text=True, check=True).stdout
    assert r1.status == 0 and r1.out == want and r2.out == want
    assert r1.out_len == len(want.encode())


# The host's device-decode branch (DeviceSession::process handing the BGZF bytes to cmb_submit_bgzf and falling back when it
# declines) with the emulator playing the device decoder (CMB_EMU_BGZF=1): same goldens, no GPU.
@pytest.mark.parametrize("case", HOST_CASES, ids=[f"{c['sub']}@{c['ref']}" for c in HOST_CASES])
def test_device_decode_branch_matches_reference_golden(case):
    check_case(case, run_case(HOSTCHECK, case, extra_args=["-t", "2"], env={"CMB_EMU_BGZF": "1"}))


def test_coverm_filter_writes_the_filtered_bam(tmp_path):
    """`coverm filter` through the host's filter loop (no GPU): output header == input header, records == the oracle's
    filter-names sequence, byte-identical to the input's records, file readable as BGZF with an EOF marker."""
    import struct
    import zlib
    from case_runner import DATA, ORACLE_BIN

    def records(path):
        raw = open(path, "rb").read()
        assert raw.endswith(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
        data, o = bytearray(), 0
        while o < len(raw):
            bsize = struct.unpack_from("<H", raw, o + 16)[0] + 1
            data += zlib.decompress(raw[o + 18:o + bsize - 8], -15)
            o += bsize
        l_text = struct.unpack_from("<I", data, 4)[0]
        n_ref = struct.unpack_from("<I", data, 8 + l_text)[0]
        p = 12 + l_text
        for _ in range(n_ref):
            p += 8 + struct.unpack_from("<I", data, p)[0]
        header, recs = bytes(data[:p]), []
        while p < len(data):
            bs = struct.unpack_from("<I", data, p)[0]
            recs.append(bytes(data[p:p + 4 + bs]))
            p += 4 + bs
        return header, recs

    src = os.path.join(DATA, "1.bam")
    for extra in (["--min-read-percent-identity", "95"], ["--proper-pairs-only", "--min-read-aligned-length-pair", "150"],
                  ["--proper-pairs-only", "--min-read-aligned-length-pair", "150", "--inverse"], []):
        out = str(tmp_path / "f.bam")
        p = subprocess.run([HOSTCHECK, "filter", "-b", src, "-o", out, "-t", "3"] + extra, capture_output=True, text=True)
        assert p.returncode == 0, p.stderr[-500:]
        want = subprocess.run([ORACLE_BIN, "filter-names", "-b", src] + extra, capture_output=True, text=True, check=True).stdout.split("\n")[:-1]
        in_header, in_recs = records(src)
        header, recs = records(out)
        assert header == in_header
        assert [r[36:36 + r[12] - 1].decode() for r in recs] == want
        assert set(recs) <= set(in_recs)
