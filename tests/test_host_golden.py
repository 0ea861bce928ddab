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
HOST_CASES = [c for c in CASES + CLI_CASES if c["sub"] in ("contig", "genome")]


@pytest.fixture(scope="session", autouse=True)
def _built():
    if not os.path.exists(HOSTCHECK):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])


@pytest.mark.parametrize("case", HOST_CASES, ids=[f"{c['sub']}@{c['ref']}" for c in HOST_CASES])
@pytest.mark.parametrize("threads", ["1", "3"])
def test_host_code_matches_reference_golden(case, threads):
    check_case(case, run_case(HOSTCHECK, case, extra_args=["-t", threads]))
