#!/usr/bin/env python
"""Stall-sample and instruction budget of one kernel from an `ncu --set full --import-source on` report.

usage: python profiles/stall_breakdown.py gpurun_out/r2_k2_scan_reduce.ncu-rep [elements_per_warp_iteration] > profiles/r2_k2_stalls.md
Reads the report's SASS page (`ncu -i ... --page source --csv`); needs the ncu CLI only, no GPU."""
import csv
import io
import subprocess
import sys


def main():
    path = sys.argv[1]
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    name = rows[0][1] if rows and len(rows[0]) > 1 else "?"
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    n_samples = sum(int(r[ix["# Samples"]]) for r in data)
    n_inst = sum(int(r[ix["Instructions Executed"]]) for r in data)
    print(f"# {name}\n\nsource: `{path}` (SASS page); {n_samples} stall samples, {n_inst} warp instructions executed\n")
    agg = sorted(((sum(int(r[ix[s]]) for r in data), s) for s in stalls), reverse=True)
    print("| stall reason | samples | share |\n|---|---:|---:|")
    for v, s in agg[:10]:
        print(f"| {s} | {v} | {100.0 * v / max(1, n_samples):.1f} % |")
    print("\nTop instructions by samples:\n\n| samples | executed | SASS | main stall |\n|---:|---:|---|---|")
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:12]:
        st = max(((int(r[ix[s]]), s) for s in stalls), default=(0, ""))
        print(f"| {r[ix['# Samples']]} | {r[ix['Instructions Executed']]} | `{r[1].strip()[:70]}` | {st[1]} |")
    # the hottest loop: instructions executed as often as the most frequent barrier / try-wait
    loop = max((int(r[ix["Instructions Executed"]]) for r in data if "BAR.SYNC" in r[1] or "TRYWAIT" in r[1]), default=0)
    if loop:
        per_iter = n_inst / loop
        print(f"\nwarp iterations of the main loop (executions of its barrier): {loop}; warp instructions per iteration: {per_iter:.0f}")
        if len(sys.argv) > 2:
            e = float(sys.argv[2])
            print(f"= {per_iter / e:.3f} warp instructions per element ({e:.0f} elements per warp iteration)")


if __name__ == "__main__":
    main()
