# session 4: the two-warp shape (pair merge, forms ahead) on the plate-only and truss-only configs
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1s4o}
export FEMGPU_ASM_THREADS=64
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -2
for c in P T; do
FEMGPU_ASM_INFO=1 timeout 600 python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-separation > gpurun_out/${TAG}_bench_${c}64.json 2> gpurun_out/${TAG}_bench_${c}64.err
grep -a "femgpu asm" gpurun_out/${TAG}_bench_${c}64.err | head -1 | cut -c1-120
python -c "
import json;d=json.loads(open('gpurun_out/${TAG}_bench_${c}64.json').read().strip().splitlines()[-1]);print('RESULT','${c}64',d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['prep_ms'],d['roofline']['frac'])"
done
