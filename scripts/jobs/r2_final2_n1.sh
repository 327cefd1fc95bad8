#!/bin/bash
# round 2, session 3: single-GPU evidence run of the final tree: smoke, GPU test-suite, the bench lines of every configuration, the reference arm
out=gpurun_out/r2_final2_n1; mkdir -p $out
summ='import sys,json
for l in sys.stdin:
    if not l.startswith("{"): continue
    d=json.loads(l); r=d["roofline"]; e=d.get("e2e") or {}
    print(d["config"]["mesh"], "step %.3f ms  kernel %.3f  prep %.3f  frac %.3f  step_frac %.3f  e2e %.2f M/s (%.4f s)" % (d["ms_per_step"], r["kernel_ms"], r["prep_ms"], r["frac"], r["whole_step_frac"], e.get("value", 0) / 1e6, e.get("seconds_per_step", 0)), e.get("phases_last_step"))'
timeout 300 python __graft_entry__.py smoke > $out/smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 $out/smoke.txt
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_dist_gpu.py > $out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_gpu.txt
timeout 600 python bench.py --steps 20 --warmup 3 > $out/bench_M.json 2> $out/bench_M.err; echo "bench M rc=$?"; python -c "$summ" < $out/bench_M.json
for c in P B T; do timeout 400 python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline > $out/bench_$c.json 2> $out/bench_$c.err; echo "bench $c rc=$?"; python -c "$summ" < $out/bench_$c.json; done
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference.json 2>/dev/null; echo "reference arm rc=$?"; cut -c1-400 $out/bench_reference.json
