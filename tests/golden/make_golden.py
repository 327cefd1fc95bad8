"""Generates tests/golden/element_golden.json from the oracle (the reference crate is Rust and cannot
be run here, so these vectors pin the ORACLE bit-for-bit across machines/compilers; the reference's
own known answers are checked separately in test_oracle_golden.py)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import oracle as O  # noqa: E402

rng = np.random.default_rng(20241017)
gold = {"truss": [], "beam": [], "plate": []}
for i in range(6):
    p1, p2 = rng.uniform(-2, 2, 3), rng.uniform(-2, 2, 3)
    if i == 0:
        p1, p2 = np.zeros(3), np.array([30.0, 0, 0])
    if i == 1:
        p1, p2 = np.zeros(3), np.array([0.0, 0, 2.5])
    A2 = None if i % 2 == 0 else 2e-4
    q, kl, kg = O.truss(p1, p2, 2.1e11, 1e-4, A2)
    gold["truss"].append({"p1": p1.tolist(), "p2": p2.tolist(), "E": 2.1e11, "A": 1e-4, "A2": A2, "kg": kg.tolist()})
for i in range(6):
    p1, p2 = rng.uniform(-2, 2, 3), rng.uniform(-2, 2, 3)
    axis = [0.1, 0.2, 1.0]
    if i == 0:
        p1, p2, axis = np.zeros(3), np.array([2.0, 0, 0]), [0, 0, 1.0]
    if i == 1:
        p1, p2, axis = np.zeros(3), np.array([0.0, 3.0, 0]), [1.0, 0, 0]   # vertical-member branch (c_xz == 0)
    props = [2.1e11, 0.3, 1e-2, 8e-6, 4e-6, 1e-6 if i >= 4 else 0.0, 1e-5, 5 / 6]
    q, pr, kl, kg = O.beam(p1, p2, *props, axis)
    gold["beam"].append({"p1": p1.tolist(), "p2": p2.tolist(), "props": props, "axis": axis, "q": q.tolist(), "kg": kg.tolist()})
base = np.array([[1, 0.75, 0], [0, 0.75, 0], [0, 0, 0], [1, 0, 0]], float)
for i in range(6):
    p = base.copy()
    if i >= 1:
        p[:, :2] += rng.uniform(-0.1, 0.1, (4, 2))
    if i == 3:
        p = np.stack([np.zeros(4), p[:, 0], p[:, 1]], axis=1)
    if i == 4:
        p = np.stack([p[:, 1], np.full(4, 2.0), p[:, 0]], axis=1)
    props = [2.1e11, 0.3, 0.01 * (1 + 0.1 * i), 5 / 6]
    q, kl, kg = O.plate(*p, *props)
    gold["plate"].append({"p": p.tolist(), "props": props, "kg": kg.tolist()})
with open(os.path.join(os.path.dirname(__file__), "element_golden.json"), "w") as f:
    json.dump(gold, f)
print("written")

# ---- element result recovery: oracle outputs for fixed displacement vectors ----------------
from finite_element_method_b200 import meshes  # noqa: E402

res = {}
for name, mesh in (("truss_cube", meshes.truss_cube(3)), ("beam_frame_jitter", meshes.beam_frame(3, 10 ** 9, jitter=True)),
                   ("plate_x0", meshes.plate_grid(3, 2, "x0")), ("mixed", meshes.mixed_structure(3, 2))):
    u = np.random.default_rng(20241018).normal(size=6 * len(mesh["x"])) * 1e-3
    ot, ob, op = O.element_results(mesh, u)
    res[name] = {"truss": ot.tolist(), "beam": ob.tolist(), "plate": op.tolist()}
with open(os.path.join(os.path.dirname(__file__), "element_results_golden.json"), "w") as f:
    json.dump(res, f)
print("element results written")
