#!/usr/bin/env python
"""Regenerates the numbers tables of DESIGN.md §7 from profiles/r2_bench_*.json (round 2)."""
import json, os, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(name):
    return json.loads([l for l in open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines() if l.startswith("{")][-1])


names = {"T": "T 1M truss", "B": "B 2M beam", "P": "P 4M plate", "M": "M 10M mixed (default)"}
r1 = {"T": (0.125, 0.239), "B": (0.672, 0.359), "P": (2.936, 0.563), "M": (3.921, 0.438)}   # round-1 step ms, whole-step frac
M = load("r2_bench_M.json")
peak = M["roofline"]["peak"]
L = [f"| config | elements | step ms (round 1) | G elem/s | assembly kernel ms | records ms | kernel: algorithmic GB/s, of measured copy peak ({peak:.0f} GB/s) | whole step of peak (round 1) | e2e, non-zero CSR read-back | e2e, structural values | K separation |",
     "|---|---|---|---|---|---|---|---|---|---|---|"]
for c in "TBPM":
    d = load(f"r2_bench_{c}.json"); r = d["roofline"]; e = d["e2e"]; st = e["structural_readback"]; s = d["separation"]
    L.append(f"| {names[c]} | {d['config']['elements'] / 1e6:.1f} M | {d['ms_per_step']:.3f} ({r1[c][0]:.3f}) | {d['value'] / 1e9:.2f} | {r['kernel_ms']:.3f} | "
             f"{r['prep_ms']:.3f} | {r['achieved']:.0f}, {100 * r['frac']:.1f} % | **{100 * r['whole_step_frac']:.1f} %** ({100 * r1[c][1]:.1f} %) | "
             f"{e['value'] / 1e6:.1f} M elem/s ({e['seconds_per_step']:.3f} s, {e['d2h_bytes_per_step'] / 1e9:.2f} GB) | "
             f"{st['value'] / 1e6:.1f} M elem/s ({st['seconds_per_step']:.3f} s, {st['d2h_bytes_per_step'] / 1e9:.2f} GB) | {s['ms']:.1f} ms |")
L += ["", "General-orientation variants (`profiles/r2_bench_variants.txt`): M-jitter 3.665 ms, M-x0 4.406 ms, P-jitter 2.865 ms, P-x0 4.016 ms, "
          "B-jitter 0.579 ms, T-jitter 0.127 ms.", ""]
one = M["value"]
rows = []
for n in (2, 4, 8):
    f = f"r2_bench_M_n{n}.json"
    if not os.path.exists(os.path.join(ROOT, "profiles", f)):
        continue
    d = load(f); r = d["roofline"]; e = d.get("e2e") or {}; w = d.get("weak") or {}
    rows.append(f"| {n} | {d['ms_per_step']:.3f} | {d['value'] / 1e9:.2f} | **{d['value'] / (n * one):.3f}** | {r['kernel_ms']:.3f} ({r['kernel_launches_per_step']} launches) | {r['prep_ms']:.3f} | "
                f"{r['exchange_ms']:.3f} | {w.get('value', 0) / 1e9:.2f} G elem/s on {w.get('elements', 0) / 1e6:.0f} M elements ({w.get('ms_per_step', 0):.3f} ms, "
                f"{w.get('value', 0) / (n * one):.3f}) | {e.get('value', 0) / 1e6:.1f} M elem/s ({e.get('seconds_per_step', 0):.3f} s) |")
if rows:
    L += ["Config 5 as the north star states it — ONE 10M-element mesh partitioned into row strips over N B200 (strong scaling; `profiles/r2_bench_M_n*.json`, "
          "rank 0's kernel / record / exchange times; efficiency = value ÷ (N × the single-GPU value above)):", "",
          "| N | step ms | G elem/s | strong-scaling efficiency | assembly kernel ms | records ms | exchange ms (rank 0) | weak scaling in the same run (efficiency) | e2e, non-zero read-back |",
          "|---|---|---|---|---|---|---|---|---|"] + rows + ["",
          "What limits the strong scaling (per-rank CUDA-event times, `FEMGPU_BENCH_DEBUG=1`, `profiles/README.md`): the parts of a pass that do not shrink with "
          "1/N. At N = 8 a pass is 0.536 ms against 3.603 / 8 = 0.450: the assembly kernel takes 0.421 instead of 0.389 ms (two launches — ghost slabs, then the rest — "
          "each with its ramp and tail, at ~105 slabs per CTA), the three record kernels 0.074 instead of 0.061 (launch and fork / join latency of ~20 µs kernels), and a "
          "receiving rank adds ~0.025 for the apply kernel with its system-scope fences and the join of the second stream; rank 0 (sends only) and the last rank (receives "
          "only) finish ~0.03 ms before the middle ranks. The exchange itself is hidden (0.003 ms on the sender). Weak scaling: 0.986-0.993. "
          "The N = 2 and N = 4 lines are runs of the final tree (session 3: faster host ingest, registered staging); the N = 8 line is the session-2 tree — the "
          "same numeric pass, its e2e column predates the faster ingest. e2e does not scale past two ranks on one host: at N = 4 the four ranks' read-backs share the "
          "host's memory and PCIe complex (rank 0: 1.70 GB in 0.081 s) and its 16 cores (the ranks divide them: `host_threads_share`).", ""]
er = M["separation"]["analysis"]["element_results"]; an = M["separation"]["analysis"]
ref = load("r2_bench_reference.json"); cpu = M["cpu_baseline"]; e = M["e2e"]; ph = e["phases_last_step"]; fp = M["fp64"]
tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["M"]     # the final tree's captures (profiles/r2_fp64_counts.md, r2_assemble_M.md)
flops, kms, ab = tj["fp64_flops_per_launch"], M["roofline"]["kernel_ms"] * 1e-3, M["roofline"]["algorithmic_bytes_per_launch"]
L += [f"FP64: {flops / 1e9:.2f} GFLOP per `assemble_kernel` launch on M (ncu instruction counts of the final tree) = {flops / kms / 1e12:.2f} TFLOP/s = "
      f"{100 * flops / kms / 1e12 / fp['measured_peak_TFLOPs']:.1f} % of the FMA peak measured in the bench run ({fp['measured_peak_TFLOPs']:.1f} TFLOP/s); arithmetic intensity "
      f"{flops / ab:.2f} FLOP/B, ridge {fp['ridge_flop_per_byte']:.1f}. DRAM traffic of the kernel (ncu): {tj['dram_bytes_per_launch'] / 1e9:.2f} GB = "
      f"{tj['dram_bytes_per_launch'] / ab:.2f} × algorithmic.",
      "",
      f"Downstream of K on config M (`r2_bench_M.json` → `separation.analysis`; K_aa: {M['separation']['n_aa'] / 1e6:.1f} M rows, "
      f"{M['separation']['nnz_aa_ab_ba_bb'][0] / 1e6:.0f} M stored entries; unchanged from round 1):",
      "",
      "| step | time | algorithmic GB/s | of measured copy peak |", "|---|---|---|---|",
      f"| K separation (`femgpu_separate_sparse`) | {M['separation']['ms']:.1f} ms | {M['separation']['achieved_GBps']:.0f} | {100 * M['separation']['frac_of_hbm_peak']:.1f} % |",
      f"| Jacobi PCG, one iteration | {an['pcg_jacobi']['ms_per_iteration']:.2f} ms | {an['pcg_jacobi']['achieved_GBps']:.0f} | {100 * an['pcg_jacobi']['frac_of_hbm_peak']:.1f} % |",
      f"| block-Jacobi PCG, one iteration | {an['pcg_block_jacobi']['ms_per_iteration']:.2f} ms | {an['pcg_block_jacobi']['achieved_GBps']:.0f} | {100 * an['pcg_block_jacobi']['frac_of_hbm_peak']:.1f} % |",
      f"| element results, {er['truss']['elements'] / 1e6:.0f} M trusses / {er['beam']['elements'] / 1e6:.0f} M beams / {er['plate']['elements'] / 1e6:.0f} M plates | "
      f"{er['truss']['ms_wall']:.2f} / {er['beam']['ms_wall']:.2f} / {er['plate']['ms_wall']:.2f} ms | {er['truss']['algorithmic_GBps']:.0f} / {er['beam']['algorithmic_GBps']:.0f} / {er['plate']['algorithmic_GBps']:.0f} | — |",
      "",
      f"Reference arm (`bench.py --impl reference`, faithful single-thread port on the box's host): {ref['value'] / 1e3:.1f} k elem/s on the mixed sample; "
      f"the optimised multi-core CPU port reaches {cpu['optimized_multicore_port']['value'] / 1e6:.1f} M elem/s on {cpu['optimized_multicore_port']['cores']} threads "
      f"({cpu['optimized_multicore_port']['sample']}). e2e on M is host bookkeeping of `add_*` ({ph['reset_add_nodes_add_elements_s']:.2f} s for 14 M items) + symbolic "
      f"({ph['symbolic_s']:.3f} s) + numeric ({ph['numeric_s']:.3f} s) + compaction and D2H of the non-zero CSR ({ph['matrix_d2h_s']:.2f} s for {e['d2h_bytes_per_step'] / 1e9:.1f} GB)."]
p = os.path.join(ROOT, "DESIGN.md")
s = open(p).read()
s = re.sub(r"(<!-- NUMBERS:BEGIN[^\n]*-->\n).*?(<!-- NUMBERS:END -->)", lambda m: m.group(1) + "\n".join(L) + "\n" + m.group(2), s, flags=re.S)
s = re.sub(r"## 7\. Numbers[^\n]*", "## 7. Numbers (round 2 final, `profiles/r2_bench_*.json`)", s)
open(p, "w").write(s)
print("\n".join(L))
