#!/bin/bash
# round 2, session 3: one-pass separation after the __grid_constant__ fix, tiles of 128 / 64 / 32 rows (library variants)
out=gpurun_out/r2z_n1; mkdir -p $out
for v in "" _sepit2 _sepit1; do
  echo "lib=libfemgpu$v.so"
  FEMGPU_LIB=$PWD/finite_element_method_b200/libfemgpu$v.so timeout 150 python scripts/gpu_sep_profile.py 2>&1 | grep "rep=1" | tee -a $out/sep_times.txt
done
