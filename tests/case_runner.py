"""Shared runner for the golden cases (tests/golden/reference_cases.py): runs a
coverm-compatible binary (the CPU oracle or the CUDA product CLI) on one case
and checks it the way the reference test does."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "tests", "golden", "data")
ORACLE_BIN = os.path.join(ROOT, "oracle", "coverm_oracle")
PRODUCT_BIN = os.path.join(ROOT, "coverm_b200", "bin", "coverm")


def run_case(binary, case, extra_args=(), timeout=300, env=None):
    with tempfile.TemporaryDirectory() as td:
        defpath = os.path.join(td, "genome.definition")
        if "definition" in case:
            with open(defpath, "w") as f:
                f.write(case["definition"])
        argv = [a.replace("{D}", DATA).replace("{DEF}", defpath) for a in case["argv"]]
        p = subprocess.run([binary, case["sub"]] + argv + list(extra_args), capture_output=True, text=True,
                           timeout=timeout, env=dict(os.environ, **(env or {})))
    return p


def check_case(case, p):
    want_status = case.get("status", 0)
    assert p.returncode == want_status, f"{case['ref']}: exit {p.returncode} != {want_status}\nstderr: {p.stderr[-2000:]}"
    if "stderr_contains" in case:
        assert case["stderr_contains"] in p.stderr, f"{case['ref']}: stderr lacks {case['stderr_contains']!r}"
    if "stdout" in case:
        assert p.stdout == case["stdout"], f"{case['ref']}:\nexpected {case['stdout']!r}\nobserved {p.stdout!r}"
    if "stdout_prefix" in case:
        assert p.stdout.startswith(case["stdout_prefix"]), \
            f"{case['ref']}:\nexpected prefix {case['stdout_prefix']!r}\nobserved {p.stdout!r}"
    if "line_count" in case:
        n = p.stdout.count("\n")
        assert n == case["line_count"], f"{case['ref']}: {n} lines != {case['line_count']}"
    for s in case.get("contains", []):
        assert s in p.stdout, f"{case['ref']}:\nexpected to contain {s!r}\nobserved {p.stdout!r}"
    if "table" in case:  # assert_equal_table, tests/test_cmdline.rs:17-31
        e = case["table"].splitlines()
        o = p.stdout.splitlines()
        assert e[:1] == o[:1], f"{case['ref']}: header {o[:1]} != {e[:1]}"
        assert sorted(e[1:]) == sorted(o[1:]), f"{case['ref']}:\nexpected {sorted(e[1:])}\nobserved {sorted(o[1:])}"
    if "reads_mapped" in case:
        got = [[int(x) for x in l.split("\t")[2:4]] for l in p.stderr.splitlines() if l.startswith("#reads_mapped")]
        assert got == case["reads_mapped"], f"{case['ref']}: reads_mapped {got} != {case['reads_mapped']}"
