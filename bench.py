#!/usr/bin/env python
"""Benchmark of the stiffness-assembly hot path (BASELINE.json metric: elements assembled/s, FP64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config M|P|B|T] [--impl ours|reference]

A "step" is one numeric pass (element records + deterministic assembly into the prebuilt CSR) over
the whole synthetic mesh. Default workload: config M, the 10M-element mixed truss/beam/plate
structure of BASELINE.json on one B200; with N GPUs (torchrun, one rank per GPU) that ONE mesh is
partitioned into N contiguous row strips (strong scaling — BASELINE.json configs 4 and 5); the weak-scaling
figure (every rank a strip of the full size) is measured in the same run and reported under the key "weak".

Prints ONE JSON line (rank 0). Keys beyond the base contract:
  roofline      dominant kernel (assemble_kernel) against the measured HBM copy peak
  cpu_baseline  the oracle (CPU restatement of the reference) timed on this box's host cores
  e2e           same metric through the public API with host buffers (H2D + symbolic + numeric + D2H)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (builder kwargs, description)
    "M": "mixed truss+beam+plate structure, 2000x2000 plate grid + 4M beams + 2M trusses = 10M elements",
    "P": "plate mesh 2000x2000 quad plate elements (4M elements)",
    "B": "3D space frame 88^3 nodes, 2M beam elements",
    "T": "3D truss lattice 64^3 nodes, 1M truss elements",
}


def build_mesh(config: str, scale_y: int = 1, nx: int | None = None, ny: int | None = None):
    from finite_element_method_b200 import meshes
    if config == "M":
        return meshes.mixed_structure(nx or 2000, (ny or 2000) * scale_y), (nx or 2000) + 1
    if config == "P":
        return meshes.plate_grid(nx or 2000, (ny or 2000) * scale_y, "flat"), (nx or 2000) + 1
    if config == "B":
        return meshes.beam_frame(nx or 88, 2_000_000 if nx is None else 10 ** 9), None
    if config == "T":
        return meshes.truss_lattice(nx or 64, 1_000_000 if nx is None else 10 ** 9), None
    raise ValueError(config)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clock / throttle sampling during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


SAMPLE = {"M": (300, 200), "P": (300, 200), "B": (34, None), "T": (40, None)}   # bounded CPU samples (faithful port)
# the multi-core port is ~60x faster: it runs the full configuration (one timed pass for M / P: 10M / 4M elements)
SAMPLE_FAST = {"M": (None, None), "P": (None, None), "B": (None, None), "T": (None, None)}


def _cpu_sample(config: str, table=None):
    from finite_element_method_b200 import meshes
    nx, ny = (table or SAMPLE)[config]
    mesh, _ = build_mesh(config, 1, nx, ny)
    return mesh, meshes.n_elements(mesh)


def _fast_port(config: str, cores: int):
    """the optimised multi-core CPU port (not the reference's algorithm) on its own, larger sample"""
    from oracle import oracle as O
    mesh, n_el = _cpu_sample(config, SAMPLE_FAST)
    sec = O.fast_assemble(mesh, n_threads=cores, repeats=1 if config in ("M", "P") else 2)["seconds"]
    return {"value": n_el / sec, "unit": "elements/s", "cores": cores, "sample": f"{mesh['name']} ({n_el} elements, the full configuration)",
            "note": "owner-computes OpenMP port, not the reference's algorithm"}


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm for this path. The crate is Rust (no
    rustc/cargo in the image, un-vendored dependencies), so this times the oracle port: the faithful
    restatement of FEM::add_plate / add_beam / add_truss — dense R^T k R per element, accumulation into
    a position-keyed map — which is single-threaded like the crate (FEM<V> has no parallelism at all,
    so "all the host threads it can use" is one). Each step is one pass over a bounded sample of the
    configured workload. The multi-core optimised CPU port (not the reference's algorithm) is
    reported next to it for context."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    mesh, n_el = _cpu_sample(args.config)
    for _ in range(min(args.warmup, 1)):
        O.faithful_time(mesh)
    t = [O.faithful_time(mesh) for _ in range(args.steps)]
    sec = float(np.mean(t))
    val = n_el / sec
    cores = os.cpu_count() or 1
    sample = (f"{mesh['name']}: {n_el} elements per step (bounded sample of config {args.config}); faithful "
              f"operation-by-operation port of the crate's add_* path, 1 thread (the crate is single-threaded), "
              f"duplicate scans disabled")
    line = {
        "impl": "reference", "metric": "elements assembled/s (FP64)", "value": val, "unit": "elements/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": CONFIGS[args.config], "config": args.config, "sample": sample},
        "cpu_baseline": {"value": val, "unit": "elements/s", "cores": 1, "kind": "port", "sample": sample,
                         "optimized_multicore_port": _fast_port(args.config, cores)},
        "e2e": {"value": val, "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(config: str):
    """Bounded CPU sample (about 10-30 s): the faithful single-thread port (the reference's algorithm; the crate
    is single-threaded) is the baseline value; the multi-core optimised port is reported next to it."""
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    mesh, n_el = _cpu_sample(config)
    reps = 3 if config in ("M", "P") else 2
    t_faithful = float(np.mean([O.faithful_time(mesh) for _ in range(reps)]))
    return {
        "value": n_el / t_faithful, "unit": "elements/s", "cores": 1, "kind": "port",
        "sample": f"{mesh['name']} ({n_el} elements, same shape as config {config}), {reps} passes; faithful "
                  f"operation-by-operation port of the crate's add_* path (dense R^T k R, position-keyed global K, "
                  f"duplicate scans disabled), single thread like the crate",
        "optimized_multicore_port": _fast_port(config, cores),
    }


_T0 = time.perf_counter()


def log(msg: str) -> None:
    """Per-phase progress on stderr (flushed), so that a run that stops somewhere says where."""
    rank = os.environ.get("RANK", "0")
    print(f"[bench rank {rank} +{time.perf_counter() - _T0:7.2f}s] {msg}", file=sys.stderr, flush=True)


def watchdog(seconds: int) -> None:
    """(Re-)arm the hang watchdog: if the process is still in the same phase after `seconds`, dump every thread's
    stack to stderr and exit — a stuck rank must not spin on its GPU until the launcher's own timeout."""
    import faulthandler
    faulthandler.cancel_dump_traceback_later()
    if seconds > 0:
        faulthandler.dump_traceback_later(seconds, exit=True)


def workload(args, scaling, rank, world):
    """(local mesh of this rank, whole-mesh name, n_nodes, n_el_total, begin, end): with several ranks only
    this rank's strip of elements is built (identical to local_part() of the whole mesh, tests/test_host_logic.py)."""
    from finite_element_method_b200 import meshes
    variant = args.variant
    if args.config in ("M", "P"):
        nx = args.nx or 2000
        ny = (args.ny or 2000) * (world if scaling == "weak" else 1)
        grid_w, n_nodes = nx + 1, (nx + 1) * (ny + 1)
        begin, end = meshes.partition_rows({"x": np.empty(n_nodes, np.uint8)}, world, grid_w)[rank]
        rows = (begin // grid_w, end // grid_w) if world > 1 else None
        local = (meshes.mixed_structure(nx, ny, rows=rows, variant=variant) if args.config == "M"
                 else meshes.plate_grid(nx, ny, variant, rows=rows))
        n_el_total = nx * ny * (1 if args.config == "P" else 2) + (len(range(0, nx, 2)) * ny if args.config == "M" else 0)
        if world > 1:      # the rank is handed its own rows' nodes plus the halo its elements touch, not every node
            local = meshes.with_node_window(local, begin, end)
        return local, local["name"], n_nodes, n_el_total, begin, end
    jitter = variant == "jitter"
    if args.config == "B":
        mesh = meshes.beam_frame(args.nx or 88, 2_000_000 if args.nx is None else 10 ** 9, jitter=jitter)
    else:
        mesh = meshes.truss_lattice(args.nx or 64, 1_000_000 if args.nx is None else 10 ** 9, jitter=jitter)
    n_nodes = len(mesh["x"])
    begin, end = meshes.partition_rows(mesh, world, None)[rank]
    local = meshes.with_node_window(meshes.local_part(mesh, begin, end), begin, end) if world > 1 else mesh
    return local, mesh["name"], n_nodes, meshes.n_elements(mesh), begin, end


def measure_numeric(args, scaling, ctx, sample_clocks):
    """Build this rank's part of the workload, run W warm-up + K timed numeric passes; returns the numbers of the
    JSON line that depend on the scaling mode. Collective over the ranks."""
    import torch
    from finite_element_method_b200 import FEM, meshes
    rank, world, local_rank, dist = ctx["rank"], ctx["world"], ctx["local_rank"], ctx["dist"]

    def barrier():
        # this handle's stream first: the torch collective must never be in flight together with work of the
        # library's own communicator / peer windows
        fem.synchronize()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    log(f"{scaling}: building the mesh")
    local, mesh_name, n_nodes, n_el_total, begin, end = workload(args, scaling, rank, world)
    n_el_local = meshes.n_elements(local)
    fem = FEM(local["rel_tol"], local["abs_tol"], n_nodes, device=local_rank)
    if world > 1:
        uid = [FEM.dist_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        fem.dist_init(rank, world, uid[0])
        fem.dist_set_ownership(begin, end)
    log(f"{scaling}: add_* ({n_el_local} elements, {n_nodes} nodes)")
    t0 = time.perf_counter()
    fem.load_mesh(local)
    t_load = time.perf_counter() - t0
    log(f"{scaling}: symbolic")
    t0 = time.perf_counter()
    n_rows, nnz_local = fem.symbolic()
    t_sym = time.perf_counter() - t0
    stream = torch.cuda.ExternalStream(fem.stream(), device=torch.device("cuda", local_rank))
    p2p = fem.dist_info()[0] if world > 1 else None

    log(f"{scaling}: warm-up ({args.warmup} passes)")
    fem.launch_count(reset=True)
    for _ in range(args.warmup):
        fem.numeric()
    barrier()
    fem.launch_count(reset=True)
    sampler = ClockSampler(local_rank) if sample_clocks else None
    if sampler:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    log(f"{scaling}: timed region ({args.steps} passes)")
    ev0.record(stream)
    for _ in range(args.steps):
        fem.numeric()
    ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = fem.launch_count()
    hist = [fem.numeric_ms_history(i) for i in range(min(args.steps, 64))]   # the timed passes
    kern = [fem.numeric_kernel_ms(i) for i in range(min(args.steps, 64))]
    if os.environ.get("FEMGPU_BENCH_DEBUG"):      # every timed pass, oldest first: [total, records, assembly, exchange] ms
        log(f"{scaling}: per-pass ms " + " | ".join(" ".join(f"{x:.3f}" for x in h) for h in reversed(hist)))
    if dist is not None:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        cnt = torch.tensor([float(n_el_local), float(launches)], device="cuda", dtype=torch.float64)
        dist.all_reduce(cnt)
        torch.cuda.synchronize()
        assert int(cnt[0].item()) == n_el_total, "partition lost or duplicated elements"
        launches = int(cnt[1].item())
    ms_step = ms_total / args.steps
    clocks = None
    if sampler:
        # The timed region lasts ~0.1 s, shorter than nvidia-smi's polling period: keep the identical load running
        # (untimed) for about a second more so the clock sample covers several polls under load. The number of
        # extra passes is derived from the all-reduced step time — the SAME on every rank (the passes exchange
        # ghost rows between neighbours: a rank-local wall-clock bound would leave the ranks with unequal pass
        # counts, i.e. an exchange that never completes).
        extra = int(min(2000, max(10, 1000.0 / max(ms_step, 0.05))))
        log(f"{scaling}: {extra} more passes under the clock sampler")
        for i in range(extra):
            fem.numeric()
            if i % 16 == 15:
                fem.synchronize()
        barrier()
        clocks = sampler.stop()
        clocks["window"] = f"timed steps + {extra} more of the same passes, untimed (nvidia-smi polls every 100 ms)"
    # the dominant kernel: sum of the assemble_kernel launches of a pass (one per slab range; CUDA events on the
    # library's stream around every launch). prep_ms = the element-record time left ON the critical path (the first
    # range's records; the other ranges' record kernels run on a second stream under the assembly kernel)
    asm_ms = float(np.mean([k[0] for k in kern]))
    asm_launches = int(kern[0][1])
    prep_ms = float(np.mean([h[1] for h in hist]))
    xchg_ms = float(np.mean([h[3] for h in hist]))
    value = n_el_total / (ms_step * 1e-3)

    # ---------------------------------------------------------------- roofline (dominant kernel)
    ab = meshes.algorithmic_bytes(local) if n_el_local <= 2_000_000 else None
    if ab is None:
        # closed form for the grid configs (identical to meshes.algorithmic_bytes; that one needs a
        # numpy unique over ~100M keys): structural nnz is what the symbolic pass reports
        nt, nb = len(local["t_n1"]), len(local["b_n1"])
        npl = np.asarray(local["p_n"]).reshape(4, -1).shape[1]
        ab = {"nnz": nnz_local, "total_bytes": 24 * nt + 96 * nb + 48 * npl + 24 * (end - begin) + 8 * nnz_local}
    peak, peak_src = measured_peaks()
    achieved = ab["total_bytes"] / (asm_ms * 1e-3) / 1e9
    traffic = None          # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture (profiles/)
    traffic_src = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        key = args.config if args.variant == "flat" else f"{args.config}-{args.variant}"
        if world == 1 and not (args.nx or args.ny) and key in tj:
            traffic = tj[key].get("dram_bytes_per_launch")
            traffic_src = "profiles/traffic.json (" + tj[key].get("source", "ncu --set full capture") + "), not measured in this run"
    except OSError:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "kernel": "assemble_kernel", "kernel_ms": asm_ms,
                "kernel_launches_per_step": asm_launches,
                "prep_ms": prep_ms, "exchange_ms": xchg_ms, "algorithmic_bytes_per_launch": ab["total_bytes"],
                "peak_source": peak_src, "per": "rank 0",
                "whole_step_frac": ab["total_bytes"] / (ms_step * 1e-3) / 1e9 / peak}
    out = {"value": value, "ms_per_step": ms_step, "launches": launches, "roofline": roofline, "clocks": clocks,
           "mesh_name": mesh_name, "n_el_total": n_el_total, "n_el_local": n_el_local, "n_nodes": n_nodes,
           "nnz_local": nnz_local, "symbolic_s": t_sym, "load_s": t_load, "begin": begin, "end": end,
           "exchange": None if world == 1 else ("peer windows over NVLink (CUDA IPC), flags in HBM" if p2p
                                                else "ncclSend/ncclRecv")}
    return fem, local, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="M", choices=list(CONFIGS))
    ap.add_argument("--variant", default="flat", choices=["flat", "jitter", "x0"],
                    help="P / M: flat (Q == I), jitter (in-plane), x0 (the mesh in the x = 0 plane, Q != I); B / T: jitter")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="N > 1: strong = ONE mesh of the configured size split over the ranks (BASELINE.json configs 4-5, "
                         "default); weak = every rank a strip of the configured size")
    ap.add_argument("--no-weak", action="store_true", help="N > 1: skip the additional weak-scaling measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-separation", action="store_true")
    ap.add_argument("--nx", type=int, default=None, help="override grid size (testing)")
    ap.add_argument("--ny", type=int, default=None)
    ap.add_argument("--phase-timeout", type=int, default=420, help="seconds a phase may take before the run aborts")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = max(args.warmup, 1)
    if args.variant == "x0" and args.config in ("B", "T"):
        ap.error("--variant x0 applies to the plate configurations")

    if args.impl == "reference":
        run_reference(args)
        return

    watchdog(args.phase_timeout)
    import datetime

    import torch
    from finite_element_method_b200 import meshes

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (femgpu has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        log("init_process_group")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank),
                                timeout=datetime.timedelta(seconds=180))
    ctx = {"rank": rank, "world": world, "local_rank": local_rank, "dist": dist}
    scaling = args.scaling if world > 1 else "strong"

    fem, local, m = measure_numeric(args, scaling, ctx, sample_clocks=True)
    n_nodes, n_el_total, n_el_local = m["n_nodes"], m["n_el_total"], m["n_el_local"]
    roofline, peak = m["roofline"], m["roofline"]["peak"]
    nnz_local = m["nnz_local"]

    # ---------------------------------------------------------------- FP64 pipe, N = 1
    fp64 = None
    if world == 1:
        watchdog(args.phase_timeout)
        log("FP64 FMA micro-benchmark")
        peak_tf = fem.fp64_fma_peak()
        # FP64 operations of one assemble_kernel launch: thread-level DFMA (x2) + DMUL + DADD instruction counts of the
        # ncu capture named in profiles/traffic.json (exact for this workload: the kernel's work does not vary)
        flops = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)
            key = args.config if args.variant == "flat" else f"{args.config}-{args.variant}"
            if not (args.nx or args.ny):
                flops = tj.get(key, {}).get("fp64_flops_per_launch")
        except OSError:
            pass
        kms = roofline["kernel_ms"] * 1e-3
        fp64 = {"measured_peak_TFLOPs": peak_tf,
                "how": "femgpu_fp64_fma_peak: 8 independent DFMA chains per thread, 8 x 256 threads per SM, best of 5",
                "kernel_flops_per_launch": flops,
                "flops_source": "ncu smsp__sass_thread_inst_executed_op_{dfma x2, dmul, dadd}_pred_on of the capture in profiles/traffic.json",
                "achieved_TFLOPs": None if flops is None else flops / kms / 1e12,
                "fp64_pipe_frac": None if flops is None else flops / kms / 1e12 / peak_tf,
                "ridge_flop_per_byte": peak_tf * 1e3 / peak,
                "arithmetic_intensity_flop_per_byte": None if flops is None else flops / roofline["algorithmic_bytes_per_launch"]}
        roofline["fp64_pipe_frac"] = fp64["fp64_pipe_frac"]

    # ---------------------------------------------------------------- next row, reported aside
    separation = None
    if world == 1 and not args.no_separation:
        watchdog(args.phase_timeout)
        log("separation / PCG / element results")
        separation = measure_downstream(args, fem, local, n_nodes, nnz_local, peak)
    if dist is not None:
        fem.synchronize()
        torch.cuda.synchronize()
        dist.barrier()
    fem.close()          # one library communicator at a time: the next measurement creates its own

    # ---------------------------------------------------------------- the other scaling mode (N > 1)
    other = None
    if world > 1 and not args.no_weak and args.config in ("M", "P"):
        watchdog(args.phase_timeout)
        mode2 = "weak" if scaling == "strong" else "strong"
        fem2, _, m2 = measure_numeric(args, mode2, ctx, sample_clocks=False)
        fem2.synchronize()
        torch.cuda.synchronize()
        dist.barrier()
        fem2.close()
        other = {"scaling": mode2, "value": m2["value"], "unit": "elements/s", "ms_per_step": m2["ms_per_step"],
                 "elements": m2["n_el_total"], "mesh": m2["mesh_name"], "kernel_ms": m2["roofline"]["kernel_ms"],
                 "prep_ms": m2["roofline"]["prep_ms"], "exchange_ms": m2["roofline"]["exchange_ms"],
                 "frac": m2["roofline"]["frac"], "whole_step_frac": m2["roofline"]["whole_step_frac"]}

    # ---------------------------------------------------------------- e2e through the public API
    e2e = None
    if not args.no_e2e:
        watchdog(args.phase_timeout)
        log("e2e")
        e2e = measure_e2e(args, local, n_nodes, local_rank, n_el_local, n_el_total, dist, rank, world, m["begin"], m["end"])

    watchdog(args.phase_timeout)
    if rank == 0:
        # the CPU legs are a single-GPU matter (rank 0 at N = 1 only): with several ranks the others would only wait
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            log("cpu baseline")
            cpu = cpu_baseline(args.config)
        line = {
            "metric": "elements assembled/s (FP64)", "value": m["value"], "unit": "elements/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["ms_per_step"], "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": CONFIGS[args.config], "config": args.config, "variant": args.variant,
                       "mesh": m["mesh_name"], "elements": n_el_total, "nodes": n_nodes, "nnz_rank0": nnz_local,
                       "parallelism": (f"{world} contiguous row strips of one mesh, ghost rows exchanged through "
                                       f"{m['exchange']}" if world > 1 and scaling == "strong" else
                                       f"row-strips x{world} ({m['exchange']})" if world > 1 else "single GPU"),
                       "l2": "working set (>= 0.2 GB of CSR values rewritten per step) is larger than the 126 MB L2; no flush needed",
                       "symbolic_s": m["symbolic_s"], "load_s": m["load_s"]},
            "roofline": roofline, "fp64": fp64, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": m["launches"],
            "clocks": m["clocks"], "separation": separation,
        }
        if other is not None:
            line[other["scaling"]] = other
        print(json.dumps(line), flush=True)
    watchdog(0)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def measure_downstream(args, fem, local, n_nodes, nnz_local, peak):
    """SURVEY §8f on the assembled matrix of the timed handle (N = 1): distributed loads, K separation, a fixed number
    of PCG iterations, element results."""
    # clamp the first 64 nodes (translations only for the truss lattice: it has no rotational stiffness)
    ndof = 3 if args.config == "T" else 6
    nodes = np.repeat(np.arange(1, 65, dtype=np.uint32), ndof)
    fem.add_displacement(nodes, np.tile(np.arange(ndof, dtype=np.int32), 64), np.zeros(len(nodes)))
    fem.add_concentrated_load(n_nodes, 0, 1.0e3)
    # §8f rank 2: a uniform load on every beam and every plate -> nodal loads, evaluated on the device
    n_pl = np.asarray(local["p_n"]).reshape(4, -1).shape[1]
    n_bm = len(local["b_n1"])
    loads = None
    if n_pl + n_bm:
        if n_bm:
            fem.add_uniformly_distributed_line_load(np.arange(1, n_bm + 1, dtype=np.uint32),
                                                    np.full(n_bm, 2, np.int32), np.full(n_bm, -2.0e3))
        if n_pl:
            fem.add_uniformly_distributed_surface_load(np.arange(1, n_pl + 1, dtype=np.uint32),
                                                       np.full(n_pl, 2, np.int32), np.full(n_pl, -1.0e3))
        fem.synchronize()
        t0 = time.perf_counter()
        fem.forces_vector(copy_out=False)
        fem.synchronize()
        loads = {"op": "uniformly distributed line/surface loads -> forces vector, on the device (upload of the load "
                       "list, nodal loads, stable sort by DOF, in-order sums)", "n_loads": int(n_pl + n_bm),
                 "ms_wall": (time.perf_counter() - t0) * 1e3}
    fem.separate_stiffness_matrix_sparse_iterative(copy_out=False)          # warm-up (allocations)
    n_aa, n_bb, q_nnz, sep_ms = fem.separate_stiffness_matrix_sparse_iterative(copy_out=False)
    sep_bytes = 12 * nnz_local + 12 * sum(q_nnz)      # col_idx + values read once, compacted copies written once
    separation = {"op": "separate_stiffness_matrix_sparse_iterative + b = R_a - K_ab u_b, on the device",
                  "ms": sep_ms, "n_aa": n_aa, "n_bb": n_bb, "nnz_aa_ab_ba_bb": q_nnz,
                  "algorithmic_bytes": sep_bytes, "achieved_GBps": sep_bytes / (sep_ms * 1e-3) / 1e9,
                  "frac_of_hbm_peak": sep_bytes / (sep_ms * 1e-3) / 1e9 / peak, "distributed_loads": loads}
    # §8f ranks 3-4: a fixed number of PCG iterations on K_aa (the model is far from converged after that:
    # only the per-iteration cost is reported) and the element result recovery from a displacement vector
    from finite_element_method_b200 import FemError
    analysis = {}
    for name, solve in (("pcg_jacobi", fem.find_ua_vector_iterative_pcg_jacobi_sparse),
                        ("pcg_block_jacobi", fem.find_ua_vector_iterative_pcg_block_jacobi_sparse)):
        iters = 0
        for max_iter in (3, 25):          # the first call allocates the work vectors
            try:
                iters = solve(max_iter, copy_out=False)[1]
            except FemError:
                iters = fem.solve_info()[0]
        it_ms = fem.solve_info()[2] / max(1, iters)
        it_bytes = 12 * q_nnz[0] + 8 * n_aa * (14 if name == "pcg_jacobi" else 22)   # K_aa once + ~14 (22) vector passes of 8 B per row
        analysis[name] = {"iterations_timed": iters, "ms_per_iteration": it_ms, "algorithmic_bytes_per_iteration": it_bytes,
                          "achieved_GBps": it_bytes / (it_ms * 1e-3) / 1e9,
                          "frac_of_hbm_peak": it_bytes / (it_ms * 1e-3) / 1e9 / peak}
    fem.set_displacements_vector(np.random.default_rng(7).normal(size=6 * n_nodes) * 1e-3)
    res_ms = {}
    for fam, fname, n_f, comps, rec in ((0, "truss", len(local["t_n1"]), 1, 24), (1, "beam", n_bm, 10, 96), (2, "plate", n_pl, 8, 48)):
        if not n_f:
            continue
        fem.element_results(fam, copy_out=False)
        fem.synchronize()
        t0 = time.perf_counter()
        fem.element_results(fam, copy_out=False)
        dt = (time.perf_counter() - t0) * 1e3
        nn = 4 if fam == 2 else 2
        b = n_f * (rec + 8 * comps + nn * (24 + 48))      # record + results + gathered coordinates / displacements
        res_ms[fname] = {"elements": int(n_f), "ms_wall": dt, "algorithmic_GBps": b / (dt * 1e-3) / 1e9}
    analysis["element_results"] = res_ms
    separation["analysis"] = analysis
    return separation


def measure_e2e(args, local, n_nodes, local_rank, n_el_local, n_el_total, dist, rank, world, begin, end):
    """One full pass per step through the public API with HOST buffers: reset -> add_* (host checks, host->device
    copies) -> symbolic -> numeric -> the assembled matrix back in (pinned) host memory. Measured twice, with the two
    read-backs the API offers:
      nonzero     femgpu_get_nonzero_csr — row_ptr, col_idx, values of the entries != 0.0, compacted on the device: the
                  set the reference's position-keyed map holds (its add_* skip exact zeros). THE e2e number.
      structural  femgpu_get_csr, values only, on the structural block pattern (what round 1 reported)."""
    import torch
    from finite_element_method_b200 import FEM
    steps = 3 if n_el_total > 2_000_000 else 5
    h2d = (sum(np.asarray(local[k]).nbytes for k in ("x", "y", "z", "t_n1", "t_n2", "t_E", "t_A", "b_n1", "b_n2",
                                                      "b_props", "b_axis", "p_n", "p_props")))
    # One handle (and, with several GPUs, one NCCL communicator) for the whole measurement, re-used
    # through FEM::reset (fem.rs:155) like a long-lived reference instance would be; every timed step
    # starts from an empty model.
    fem = FEM(local["rel_tol"], local["abs_tol"], n_nodes, device=local_rank)
    if world > 1:
        u = [FEM.dist_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(u, src=0)
        fem.dist_init(rank, world, u[0])
    results = {}
    for mode in ("nonzero", "structural"):
        bufs = None
        times, phases, d2h = [], None, 0
        for it in range(steps + 1):
            log(f"e2e[{mode}]: step {it} of {steps} (+1 warm-up)")
            fem.synchronize()
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            fem.reset(n_nodes)
            if world > 1:
                fem.dist_set_ownership(begin, end)
            fem.load_mesh(local, cache=True)    # the caller's label arrays are input, built once; every step passes them again
            t1 = time.perf_counter()
            n_rows, nnz = fem.symbolic()
            t2 = time.perf_counter()
            fem.numeric()
            fem.synchronize()
            t3 = time.perf_counter()
            t_pin = 0.0
            if bufs is None:                # one-time pinned allocations are not part of a step (sized by the warm-up step)
                tp = time.perf_counter()
                pin = lambda n, dt: torch.empty(n, dtype=dt).pin_memory().numpy()
                if mode == "nonzero":
                    bufs = (pin(n_rows + 1, torch.int64), pin(nnz, torch.int32), pin(nnz, torch.float64))
                else:
                    bufs = pin(nnz, torch.float64)
                t_pin = time.perf_counter() - tp
            if mode == "nonzero":
                rp, ci, v = fem.nonzero_csr(out=bufs)
                d2h = rp.nbytes + ci.nbytes + v.nbytes
            else:
                fem.csr(values_only=True, out=bufs)
                d2h = bufs.nbytes
            t4 = time.perf_counter()
            if it > 0:
                times.append(t4 - t0 - t_pin)
                phases = {"reset_add_nodes_add_elements_s": t1 - t0, "symbolic_s": t2 - t1, "numeric_s": t3 - t2,
                          "matrix_d2h_s": t4 - t3 - t_pin}
        sec = float(np.median(times))       # median of the timed steps (a step now and then pays for pool growth)
        if dist is not None:
            t = torch.tensor([sec, float(d2h)], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t[0].item())
            torch.cuda.synchronize()
        results[mode] = {"value": n_el_total / sec, "unit": "elements/s", "h2d_bytes_per_step": int(h2d),
                         "d2h_bytes_per_step": int(d2h), "seconds_per_step": sec, "seconds_per_step_is": "median of the timed steps",
                         "steps": steps, "step_seconds": times,
                         "phases_last_step": phases}
        del bufs
    fem.synchronize()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()      # no rank may still be in a pass when the joined handles go
    fem.close()
    out = dict(results["nonzero"])
    out["readback"] = ("femgpu_get_nonzero_csr: row_ptr + col_idx + values of the entries != 0.0 (the reference's stored set), "
                       "compacted on the device, into pinned host memory" + (" — per rank: its own rows" if world > 1 else ""))
    out["structural_readback"] = dict(results["structural"], readback="femgpu_get_csr: all values of the structural block pattern")
    out["includes"] = ("femgpu_reset, add_nodes/add_* host validation + H2D, symbolic pass, numeric pass, device-side compaction, "
                       "D2H of the matrix (handle and NCCL communicator created once, outside the timed steps; the label arrays "
                       "the caller passes — node / element numbers — are built once per mesh, every step passes all arrays again "
                       "from pageable host memory; a re-used handle DMAs from its own CUDA-registered staging copy)")
    return out


if __name__ == "__main__":
    main()
