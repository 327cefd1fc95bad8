#!/bin/bash
# round 2, session 3: one-pass separation with 128-row tiles: tests + the separation leg of the bench
out=gpurun_out/r2u_n1; mkdir -p $out
timeout 600 python -m pytest tests/test_separation.py -m gpu -q -x > $out/pytest_sep.txt 2>&1; echo "separation tests rc=$?"; tail -5 $out/pytest_sep.txt
for tp in 0 1; do
  if [ $tp = 1 ]; then export FEMGPU_SEP_TWO_PASS=1; fi
  FEMGPU_ASM_INFO=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $out/bench_M_$tp.json 2> $out/bench_M_$tp.err; echo "bench rc=$?"; grep "femgpu separate" $out/bench_M_$tp.err | head -3
  python - $out/bench_M_$tp.json <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); s = d["separation"]
        print("step %.3f ms; separation %.3f ms frac %.3f" % (d["ms_per_step"], s["ms"], s["frac_of_hbm_peak"]), s["nnz_aa_ab_ba_bb"])
PY
done
