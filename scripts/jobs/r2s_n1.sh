#!/bin/bash
# round 2, session 3: GPU test-suite + the default bench line with the faster host ingest; host / symbolic stage timings of an e2e step
out=gpurun_out/r2s_n1; mkdir -p $out
nproc > $out/host.txt; lscpu | grep -i "model name\|^CPU(s)\|socket\|numa" >> $out/host.txt; free -g | head -2 >> $out/host.txt; cat $out/host.txt
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_dist_gpu.py > $out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -4 $out/pytest_gpu.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-separation > $out/bench_M.json 2> $out/bench_M.err; echo "bench rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/r2s_n1/bench_M.json"):
    if l.startswith("{"):
        d = json.loads(l); e = d["e2e"]
        print("step %.3f ms; e2e %.2f M elem/s, %.4f s/step" % (d["ms_per_step"], e["value"] / 1e6, e["seconds_per_step"]), e["step_seconds"], e["phases_last_step"])
        print("structural", e["structural_readback"]["value"] / 1e6, e["structural_readback"]["phases_last_step"])
PY
FEMGPU_HOST_TIMING=1 FEMGPU_SYM_TIMING=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-separation > /dev/null 2> $out/timing.err; grep "femgpu add\|femgpu symbolic" $out/timing.err | tail -44
