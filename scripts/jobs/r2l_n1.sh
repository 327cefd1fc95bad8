#!/bin/bash
# round 2, job l (1 GPU): balancing weights of the work lists after the beam path got cheaper; slab quota
out=gpurun_out/r2l_n1; mkdir -p $out
B="--steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-separation"
summ='import sys,json
for l in sys.stdin:
    if not l.startswith("{"): continue
    d=json.loads(l); r=d["roofline"]
    print(d["config"]["mesh"], "step %.3f ms  kernel %.3f  prep %.3f  frac %.3f  step_frac %.3f" % (d["ms_per_step"], r["kernel_ms"], r["prep_ms"], r["frac"], r["whole_step_frac"]))'
for cb in 1 2 4 6 8 12; do echo "COST_B=$cb"; FEMGPU_COST_B=$cb timeout 300 python bench.py $B 2>/dev/null | python -c "$summ"; done
for cp in 8 12 24 32; do echo "COST_P=$cp COST_B=4"; FEMGPU_COST_P=$cp FEMGPU_COST_B=4 timeout 300 python bench.py $B 2>/dev/null | python -c "$summ"; done
for q in 60 66 78 84; do echo "QUOTA=$q"; FEMGPU_SLAB_QUOTA=$q timeout 300 python bench.py $B 2>/dev/null | python -c "$summ"; done
for cb in 2 4; do echo "M-jitter COST_B=$cb"; FEMGPU_COST_B=$cb timeout 300 python bench.py --variant jitter $B 2>/dev/null | python -c "$summ"; done
