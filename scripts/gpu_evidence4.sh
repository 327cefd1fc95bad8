# Final round-1 evidence (session 4, kernel with pair merge + forms built ahead by warp 1): everything of gpu_evidence3.sh + the per-phase cycle dump
set -x
TAG=${TAG:-r1h}
bash scripts/gpu_evidence3.sh
FEMGPU_LIB=$PWD/finite_element_method_b200/libfemgpu_prof.so FEMGPU_PHASE_DUMP=1 timeout 600 python bench.py --config M --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-separation 2>&1 | grep -a "femgpu phases" | tail -12 > gpurun_out/${TAG}_phases_M.txt
cat gpurun_out/${TAG}_phases_M.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
