"""The host's BGZF block decoder (coverm_b200/csrc/host/fast_inflate.hpp) against zlib: round trips over compression
levels 0-9, default / Huffman-only / fixed strategies and six data shapes, plus bit flips, truncation and wrong sizes
(must be rejected or fall through to zlib, never crash)."""
import os
import subprocess

import pytest

from case_runner import ROOT as REPO, DATA

PRODUCT_HOSTCHECK = os.path.join(REPO, "oracle", "coverm_hostcheck")

SRC = os.path.join(REPO, "tests", "native", "fast_inflate_check.cpp")


def test_fast_inflate_matches_zlib(tmp_path):
    exe = str(tmp_path / "fast_inflate_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(REPO, "coverm_b200", "csrc", "host"), SRC, "-lz", "-o", exe],
                   check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert "1500 tests, 0 fails" in out, out


def test_corrupt_bgzf_block_is_rejected(tmp_path):
    """A flipped bit inside a block's deflate payload must not produce a coverage table (htslib fails the CRC32)."""
    src = os.path.join(DATA, "tpm_test.bam")
    raw = bytearray(open(src, "rb").read())
    raw[len(raw) // 2] ^= 0x10
    bad = str(tmp_path / "bad.bam")
    open(bad, "wb").write(raw)
    r = subprocess.run([PRODUCT_HOSTCHECK, "contig", "-b", bad, "-m", "mean"], capture_output=True, text=True)
    assert r.returncode != 0
    good = subprocess.run([PRODUCT_HOSTCHECK, "contig", "-b", src, "-m", "mean"], capture_output=True, text=True)
    assert good.returncode == 0


def test_t1_inflate_logic_matches_zlib(tmp_path):
    """kd_inflate_t1's decoder (one thread per BGZF block: canonical-Huffman decode by limit comparison, no lookup tables),
    compiled as plain C++ for one emulated thread, against zlib over 1500 streams + corrupted copies (ASan/UBSan build)."""
    src = os.path.join(REPO, "tests", "native", "t1_inflate_check.cpp")
    exe = str(tmp_path / "t1_inflate_check")
    subprocess.run(["g++", "-O1", "-std=c++17", "-fsanitize=address,undefined", "-I", os.path.join(REPO, "coverm_b200", "csrc"), src, "-lz",
                    "-o", exe], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert "1500 tests, 0 fails" in out, out
