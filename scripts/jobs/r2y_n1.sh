#!/bin/bash
# round 2, session 3: ncu --set full of the separation kernels (count, fill, one-pass) on config M, one application run
out=gpurun_out/r2y_n1; mkdir -p $out
timeout 400 ncu --set full --import-source on --clock-control none -k regex:quadrant -s 2 -c 3 -o $out/sep_kernels -f python scripts/gpu_sep_profile.py > $out/ncu.log 2>&1; echo "ncu rc=$?"; tail -5 $out/ncu.log
ls -la $out
