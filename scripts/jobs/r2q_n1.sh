#!/bin/bash
# round 2, job q (1 GPU): K separation, rows per warp 2 / 3 / 4
sepsum='import sys,json
for l in sys.stdin:
    if not l.startswith("{"): continue
    d=json.loads(l); s=d["separation"]
    print(d["config"]["mesh"], "separation %.2f ms (%.1f %% of peak)" % (s["ms"], 100*s["frac_of_hbm_peak"]))'
for lib in sep2 sep3; do echo "lib=$lib"; for c in M B; do FEMGPU_LIB=$PWD/finite_element_method_b200/libfemgpu_$lib.so timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "$sepsum"; done; done
FEMGPU_LIB=$PWD/finite_element_method_b200/libfemgpu_sep2.so timeout 600 python -m pytest tests/test_separation.py tests/test_analysis.py -m gpu -q 2>&1 | tail -2
