#!/bin/bash
# round 2, session 3: registered host staging + early uploads, one-pass separation: the GPU tests, then the bench line
out=gpurun_out/r2t_n1; mkdir -p $out
timeout 600 python -m pytest tests/test_separation.py tests/test_reuse_gpu.py -m gpu -q -x > $out/pytest_new.txt 2>&1; echo "new tests rc=$?"; tail -15 $out/pytest_new.txt
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_dist_gpu.py --deselect tests/test_separation.py --deselect tests/test_reuse_gpu.py > $out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -4 $out/pytest_gpu.txt
FEMGPU_ASM_INFO=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $out/bench_M.json 2> $out/bench_M.err; echo "bench rc=$?"; grep "femgpu separate" $out/bench_M.err | head -3
python - <<'PY'
import json
for l in open("gpurun_out/r2t_n1/bench_M.json"):
    if l.startswith("{"):
        d = json.loads(l); e = d["e2e"]; s = d["separation"]
        print("step %.3f ms; e2e %.2f M elem/s, %.4f s/step" % (d["ms_per_step"], e["value"] / 1e6, e["seconds_per_step"]), e["step_seconds"], e["phases_last_step"])
        print("separation %.3f ms frac %.3f" % (s["ms"], s["frac_of_hbm_peak"]), s["nnz_aa_ab_ba_bb"])
PY
