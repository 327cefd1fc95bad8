#!/bin/bash
# round 2, job i (1 GPU): coalesced record stores — tests, sanitizer, the four configurations
out=gpurun_out/r2i_n1; mkdir -p $out
B="--steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-separation"
summ='import sys,json
for l in sys.stdin:
    if not l.startswith("{"): continue
    d=json.loads(l); r=d["roofline"]
    print(d["config"]["mesh"], "step %.3f ms  kernel %.3f (x%d)  prep %.3f  frac %.3f  step_frac %.3f" % (d["ms_per_step"], r["kernel_ms"], r.get("kernel_launches_per_step",1), r["prep_ms"], r["frac"], r["whole_step_frac"]))'
for c in M B P T; do timeout 300 python bench.py --config $c $B 2>/dev/null | python -c "$summ"; done
timeout 300 python bench.py --config M --variant x0 $B 2>/dev/null | python -c "$summ"
timeout 300 python bench.py --config B --variant jitter $B 2>/dev/null | python -c "$summ"
timeout 1700 python -m pytest tests -m gpu -q --deselect tests/test_dist_gpu.py > $out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_gpu.txt
for tool in memcheck racecheck; do timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/gpu_sanitize.py > $out/sanitizer_$tool.txt 2>&1; echo "$tool rc=$?"; tail -4 $out/sanitizer_$tool.txt; done
