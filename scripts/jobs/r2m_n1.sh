#!/bin/bash
# round 2, job m (1 GPU): per-handle trig table — the four configurations (+ jitter) and the GPU parity tests
out=gpurun_out/r2m_n1; mkdir -p $out
B="--steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-separation"
summ='import sys,json
for l in sys.stdin:
    if not l.startswith("{"): continue
    d=json.loads(l); r=d["roofline"]
    print(d["config"]["mesh"], "step %.3f ms  kernel %.3f  prep %.3f  frac %.3f  step_frac %.3f" % (d["ms_per_step"], r["kernel_ms"], r["prep_ms"], r["frac"], r["whole_step_frac"]))'
for c in M B P T; do timeout 300 python bench.py --config $c $B 2>/dev/null | python -c "$summ"; done
for c in M B; do timeout 300 python bench.py --config $c --variant jitter $B 2>/dev/null | python -c "$summ"; done
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_analysis.py tests/test_separation.py -m gpu -q > $out/pytest.txt 2>&1; echo "pytest rc=$?"; tail -2 $out/pytest.txt
