#!/bin/bash
# round 2, job c (1 GPU): GPU test-suite, the default bench line, slab-range sweep, ablation builds, other configs
out=gpurun_out/r2c_n1; mkdir -p $out
B="--steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-separation"
summ='import sys,json
for l in sys.stdin:
    if not l.startswith("{"): continue
    d=json.loads(l); r=d["roofline"]
    print(d["config"]["mesh"], "step %.3f ms  kernel %.3f (x%d)  prep %.3f  frac %.3f  step_frac %.3f" % (d["ms_per_step"], r["kernel_ms"], r.get("kernel_launches_per_step",1), r["prep_ms"], r["frac"], r["whole_step_frac"]))'
timeout 1500 python -m pytest tests -m gpu -x -q -s --deselect tests/test_dist_gpu.py > $out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_gpu.txt
FEMGPU_NUMERIC_RANGES=1 timeout 600 python bench.py --steps 20 --warmup 3 > $out/bench_M.json 2> $out/bench_M.err; echo "bench rc=$?"; python -c "$summ" < $out/bench_M.json
for r in 2 3 4 8; do echo "RANGES=$r"; FEMGPU_BENCH_DEBUG=1 FEMGPU_NUMERIC_RANGES=$r timeout 300 python bench.py $B 2> $out/ranges_$r.err | tee $out/ranges_$r.json | python -c "$summ"; grep per-pass $out/ranges_$r.err | cut -c1-260; done
for v in 1 2 4 8 16 32 3; do echo "ABL=$v"; FEMGPU_NUMERIC_RANGES=1 FEMGPU_LIB=$PWD/finite_element_method_b200/libfemgpu_abl$v.so timeout 300 python bench.py $B 2>/dev/null | python -c "$summ"; done
for c in P B T; do FEMGPU_NUMERIC_RANGES=1 timeout 300 python bench.py --config $c $B 2>/dev/null | tee $out/bench_$c.json | python -c "$summ"; done
for v in x0 jitter; do FEMGPU_NUMERIC_RANGES=1 timeout 300 python bench.py --config P --variant $v $B 2>/dev/null | tee $out/bench_P_$v.json | python -c "$summ"; done
FEMGPU_NUMERIC_RANGES=1 timeout 300 python bench.py --config M --variant x0 $B 2>/dev/null | tee $out/bench_M_x0.json | python -c "$summ"
for c in B T; do FEMGPU_NUMERIC_RANGES=1 timeout 300 python bench.py --config $c --variant jitter $B 2>/dev/null | tee $out/bench_${c}_jitter.json | python -c "$summ"; done
