#!/usr/bin/env python
"""Shared-memory wavefronts of one kernel by SASS instruction cluster, from an ncu report.
usage: ncu_smem_breakdown.py <report.ncu-rep>
Reads `ncu -i <rep> --page source --csv --print-source sass` and groups consecutive shared-memory
instructions of one opcode (the unrolled bodies of a source loop) with their wavefronts, the ideal count
and the instruction count; prints the stall-sample split as well."""
import csv, subprocess, sys
from collections import defaultdict

rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def f(r, k):
    try:
        return float(r[ix[k]])
    except (ValueError, IndexError, KeyError):
        return 0.0


def opcode(r):
    src = r[ix["Source"]].strip().split()
    return src[1] if src and src[0].startswith("@") else (src[0] if src else "?")


tw = sum(f(r, "L1 Wavefronts Shared") for r in data)
ti = sum(f(r, "L1 Wavefronts Shared Ideal") for r in data)
print(f"shared-memory wavefronts {tw / 1e6:.1f} M, ideal {ti / 1e6:.1f} M ({tw / max(ti, 1):.2f}x)")
by_op = defaultdict(lambda: [0.0, 0.0, 0.0])
clusters = []
for n, r in enumerate(data):
    w = f(r, "L1 Wavefronts Shared")
    if w <= 0:
        continue
    op = opcode(r)
    by_op[op][0] += w
    by_op[op][1] += f(r, "L1 Wavefronts Shared Ideal")
    by_op[op][2] += f(r, "Instructions Executed")
    if clusters and clusters[-1]["op"] == op and n - clusters[-1]["last"] < 40:
        c = clusters[-1]
    else:
        c = {"op": op, "first": n, "w": 0.0, "i": 0.0, "inst": 0.0, "lines": 0}
        clusters.append(c)
    c["last"] = n
    c["w"] += w
    c["i"] += f(r, "L1 Wavefronts Shared Ideal")
    c["inst"] += f(r, "Instructions Executed")
    c["lines"] += 1
print("\n| opcode | wavefronts (M) | ideal (M) | ratio | warp instructions (M) |\n|---|---|---|---|---|")
for op, v in sorted(by_op.items(), key=lambda kv: -kv[1][0]):
    print(f"| {op} | {v[0] / 1e6:.1f} | {v[1] / 1e6:.1f} | {v[0] / max(v[1], 1):.2f} | {v[2] / 1e6:.1f} |")
print("\n| cluster (SASS lines) | opcode | instructions in cluster | warp instructions (M) | wavefronts (M) | ideal (M) | ratio |\n|---|---|---|---|---|---|---|")
for c in clusters:
    if c["w"] > 0.003 * tw:
        print(f"| {c['first']}..{c['last']} | {c['op']} | {c['lines']} | {c['inst'] / 1e6:.1f} | {c['w'] / 1e6:.1f} | {c['i'] / 1e6:.1f} | {c['w'] / max(c['i'], 1):.2f} |")
samples = sum(f(r, "# Samples") for r in data)
print(f"\nstall samples ({samples:.0f}):", ", ".join(
    f"{k[6:]} {100 * sum(f(r, k) for r in data) / max(samples, 1):.1f}%" for k in
    ["stall_selected", "stall_wait", "stall_short_sb", "stall_barrier", "stall_branch_resolving", "stall_not_selected",
     "stall_mio", "stall_no_inst", "stall_long_sb", "stall_math", "stall_dispatch"]))
