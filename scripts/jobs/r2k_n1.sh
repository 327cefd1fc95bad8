#!/bin/bash
# round 2, job k (1 GPU): beam identity-rotation path — GPU tests, all configurations and variants
out=gpurun_out/r2k_n1; mkdir -p $out
B="--steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-separation"
summ='import sys,json
for l in sys.stdin:
    if not l.startswith("{"): continue
    d=json.loads(l); r=d["roofline"]
    print(d["config"]["mesh"], "step %.3f ms  kernel %.3f (x%d)  prep %.3f  frac %.3f  step_frac %.3f" % (d["ms_per_step"], r["kernel_ms"], r.get("kernel_launches_per_step",1), r["prep_ms"], r["frac"], r["whole_step_frac"]))'
for c in M B P T; do timeout 300 python bench.py --config $c $B 2>/dev/null | tee $out/bench_$c.json | python -c "$summ"; done
for v in x0 jitter; do timeout 300 python bench.py --config M --variant $v $B 2>/dev/null | tee $out/bench_M_$v.json | python -c "$summ"; done
for v in x0 jitter; do timeout 300 python bench.py --config P --variant $v $B 2>/dev/null | tee $out/bench_P_$v.json | python -c "$summ"; done
for c in B T; do timeout 300 python bench.py --config $c --variant jitter $B 2>/dev/null | tee $out/bench_${c}_jitter.json | python -c "$summ"; done
timeout 1700 python -m pytest tests -m gpu -q --deselect tests/test_dist_gpu.py > $out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_gpu.txt
