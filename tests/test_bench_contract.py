"""bench.py's reference arm runs on CPU (it is the oracle port): one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "T", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "elements/s" and d["higher_is_better"] is True
    assert d["metric"] == "elements assembled/s (FP64)" and d["value"] > 0 and d["dtype"] == "f64"
    assert d["e2e"] == {"value": d["value"], "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["gpu_launches"] == 0


def test_other_ranks_of_the_reference_arm_exit_without_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                       text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_bench_workload_partitions_one_mesh_over_the_ranks():
    """bench.py --gpus N (strong scaling, the default): every element of the configured mesh lands on exactly one rank,
    every rank gets its own rows' nodes plus the halo; weak scaling multiplies the strip count instead."""
    import types
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    from finite_element_method_b200 import meshes
    for config in ("M", "P", "B", "T"):
        for world in (1, 2, 4, 8):
            args = types.SimpleNamespace(config=config, variant="flat", nx=24 if config in "MP" else 6, ny=16)
            total, covered, n_el_total = 0, 0, None
            for rank in range(world):
                local, name, n_nodes, n_el, begin, end = bench.workload(args, "strong", rank, world)
                n_el_total = n_el if n_el_total is None else n_el_total
                assert n_el == n_el_total
                total += meshes.n_elements(local)
                covered += end - begin
                w0 = int(local.get("node_window_begin", 0))
                assert w0 == (begin if world > 1 else 0) and w0 + len(local["x"]) >= end
                if world > 1:
                    assert local["nodes_number"] == n_nodes and len(local["x"]) < n_nodes
            assert total == n_el_total and covered == n_nodes
        if config in "MP":
            sizes = [bench.workload(args, "weak", 0, w)[3] for w in (1, 2, 4)]
            assert sizes[1] == 2 * sizes[0] and sizes[2] == 4 * sizes[0]
