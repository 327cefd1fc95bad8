# session 4: A/B of staging variants (bulk record copies, compact work items) on top of striped chunks
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1s4b}
for v in "" _v2 _v3 _v23; do
export FEMGPU_LIB=$PWD/finite_element_method_b200/libfemgpu$v.so
echo "=== variant '$v'"
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -2
for c in M P B T; do
FEMGPU_ASM_INFO=1 timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-separation > gpurun_out/${TAG}${v}_bench_$c.json 2> gpurun_out/${TAG}${v}_bench_$c.err
grep -a "femgpu asm" gpurun_out/${TAG}${v}_bench_$c.err | head -1
python -c "
import json;d=json.loads(open('gpurun_out/${TAG}${v}_bench_$c.json').read().strip().splitlines()[-1]);print('RESULT','$v','$c',d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac'])"
done
done
