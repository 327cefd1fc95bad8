#!/bin/bash
# round 2, job p (1 GPU): sanity of the final tree (GPU tests, default bench line) + K separation rows-per-warp A/B
out=gpurun_out/r2p_n1; mkdir -p $out
timeout 1700 python -m pytest tests -m gpu -q --deselect tests/test_dist_gpu.py > $out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -2 $out/pytest_gpu.txt
sepsum='import sys,json
for l in sys.stdin:
    if not l.startswith("{"): continue
    d=json.loads(l); s=d["separation"]; r=d["roofline"]
    print(d["config"]["mesh"], "step %.3f kernel %.3f prep %.3f | separation %.2f ms (%.1f %% of peak)" % (d["ms_per_step"], r["kernel_ms"], r["prep_ms"], s["ms"], 100*s["frac_of_hbm_peak"]))'
for lib in "" sep6 sep8; do echo "lib=$lib"; for c in M B; do if [ -n "$lib" ]; then export FEMGPU_LIB=$PWD/finite_element_method_b200/libfemgpu_$lib.so; else unset FEMGPU_LIB; fi; timeout 300 python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "$sepsum"; done; done
unset FEMGPU_LIB
if [ -f finite_element_method_b200/libfemgpu_sep8.so ]; then FEMGPU_LIB=$PWD/finite_element_method_b200/libfemgpu_sep8.so timeout 600 python -m pytest tests/test_separation.py tests/test_analysis.py -m gpu -q 2>&1 | tail -2; fi
