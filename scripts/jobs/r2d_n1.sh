#!/bin/bash
# round 2, job d (1 GPU): whole GPU test-suite, store-path micro-benchmark, record-kernel occupancy A/B
out=gpurun_out/r2d_n1; mkdir -p $out
B="--steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-separation"
summ='import sys,json
for l in sys.stdin:
    if not l.startswith("{"): continue
    d=json.loads(l); r=d["roofline"]
    print(d["config"]["mesh"], "step %.3f ms  kernel %.3f (x%d)  prep %.3f  frac %.3f  step_frac %.3f" % (d["ms_per_step"], r["kernel_ms"], r.get("kernel_launches_per_step",1), r["prep_ms"], r["frac"], r["whole_step_frac"]))'
S=scripts/micro/store_bench
echo "== store path micro-benchmark"
for cfg in "0 22464 64 4 1 0" "0 22464 64 4 1 16" "0 22464 64 5 1 0" "0 22464 64 8 1 0" "0 22464 64 4 2 0" "0 22464 64 2 4 0" "0 44928 64 4 1 0" "0 11264 64 8 1 0" "0 65536 64 3 1 0" "0 22464 32 5 1 0" \
           "2 22464 64 4 1 0" "2 22464 64 4 2 0" "2 22464 256 4 1 0" "1 22464 64 4 1 0" "1 22464 256 4 1 0" "1 22464 256 8 1 0" "1 22464 1024 2 1 0"; do $S $cfg; done 2>&1 | tee $out/store_bench.txt
echo "== record kernels: 64 registers (4 blocks/SM, default) vs 85 (3 blocks/SM)"
for c in M B P; do python bench.py --config $c $B 2>/dev/null | python -c "$summ"; FEMGPU_LIB=$PWD/finite_element_method_b200/libfemgpu_prep3.so python bench.py --config $c $B 2>/dev/null | python -c "$summ"; done
echo "== pytest -m gpu"
timeout 1700 python -m pytest tests -m gpu -q -s --deselect tests/test_dist_gpu.py > $out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -5 $out/pytest_gpu.txt
