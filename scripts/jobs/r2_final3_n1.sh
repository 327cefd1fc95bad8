#!/bin/bash
# round 2, session 3: last check of the final tree — smoke + the whole GPU test-suite the way the driver runs it
out=gpurun_out/r2_final3_n1; mkdir -p $out
timeout 300 python __graft_entry__.py smoke > $out/smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 $out/smoke.txt
timeout 900 python -m pytest tests/ -x -q -m gpu > $out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_gpu.txt
