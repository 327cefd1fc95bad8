#!/bin/bash
# round 2, session 3: full GPU test-suite + the default bench line after the cbase runs
out=gpurun_out/r2w_n1; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_dist_gpu.py > $out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -4 $out/pytest_gpu.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-separation > $out/bench_M.json 2> $out/bench_M.err; echo "bench rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/r2w_n1/bench_M.json"):
    if l.startswith("{"):
        d = json.loads(l); e = d["e2e"]
        print("step %.3f ms; e2e %.2f M elem/s, %.4f s/step" % (d["ms_per_step"], e["value"] / 1e6, e["seconds_per_step"]), e["step_seconds"], e["phases_last_step"])
PY
