#!/bin/bash
# round 2, job h (1 GPU): slab-range pipeline again, now that the record kernels are ~2x cheaper
out=gpurun_out/r2h_n1; mkdir -p $out
B="--steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-separation"
summ='import sys,json
for l in sys.stdin:
    if not l.startswith("{"): continue
    d=json.loads(l); r=d["roofline"]
    print(d["config"]["mesh"], "step %.3f ms  kernel %.3f (x%d)  prep %.3f  frac %.3f  step_frac %.3f" % (d["ms_per_step"], r["kernel_ms"], r.get("kernel_launches_per_step",1), r["prep_ms"], r["frac"], r["whole_step_frac"]))'
for c in M B P; do for r in 1 2 3 4 6 8; do echo "config $c RANGES=$r"; FEMGPU_NUMERIC_RANGES=$r timeout 300 python bench.py --config $c $B 2>/dev/null | python -c "$summ"; done; done
