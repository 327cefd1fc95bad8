#!/usr/bin/env python
"""Host side of an e2e step on a staging-only handle (no GPU needed): femgpu_reset + the batched add_* of a bench
configuration, repeated on one handle like bench.py's e2e steps. FEMGPU_HOST_TIMING=1 prints the phases of every add
call, FEMGPU_HOST_THREADS=n caps the worker threads.
usage: host_ingest_bench.py [M|P|B|T] [repetitions]"""
import os
import sys
import time
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from finite_element_method_b200 import FEM

config = sys.argv[1] if len(sys.argv) > 1 else "M"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
args = types.SimpleNamespace(config=config, variant="flat", nx=None, ny=None)
local, name, n_nodes, n_el, _, _ = bench.workload(args, "strong", 0, 1)
fem = FEM(local["rel_tol"], local["abs_tol"], n_nodes, device=-1)
ts = []
for _ in range(reps):
    t0 = time.perf_counter()
    fem.reset(n_nodes)
    fem.load_mesh(local, cache=True)
    ts.append(time.perf_counter() - t0)
steady = sorted(ts[1:]) or ts
print(f"{name}: {n_nodes} nodes + {n_el} elements; reset + add_*: first {ts[0]:.4f} s, then min {steady[0]:.4f} s, "
      f"median {steady[len(steady) // 2]:.4f} s ({os.cpu_count()} cores)")
