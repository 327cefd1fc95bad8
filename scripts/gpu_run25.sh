# session 4: global analysis (PCG, reactions, element results) tests + full gpu suite + default bench
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1s4d}
timeout 900 python -m pytest tests/test_analysis.py -m gpu -x -q 2>&1 | tail -25
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6
FEMGPU_ASM_INFO=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_M.json 2> gpurun_out/${TAG}_bench_M.err
grep -a "femgpu asm" gpurun_out/${TAG}_bench_M.err | head -1
tail -c 1500 gpurun_out/${TAG}_bench_M.json
