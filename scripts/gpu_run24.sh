# session 4: 168-register two-warp shape (3 warps per SM sub-partition) + 4-byte work items + 16-byte smem alignment
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1s4c}
for v in _v4_168 _v4_160; do
export FEMGPU_LIB=$PWD/finite_element_method_b200/libfemgpu$v.so
echo "=== variant '$v'"
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -2
for c in M B P T; do
FEMGPU_ASM_INFO=1 timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-separation > gpurun_out/${TAG}${v}_bench_$c.json 2> gpurun_out/${TAG}${v}_bench_$c.err
grep -a "femgpu asm" gpurun_out/${TAG}${v}_bench_$c.err | head -1
python -c "
import json;d=json.loads(open('gpurun_out/${TAG}${v}_bench_$c.json').read().strip().splitlines()[-1]);print('RESULT','$v','$c',d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac'])"
done
for c in P T; do
FEMGPU_ASM_THREADS=64 FEMGPU_ASM_INFO=1 timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-separation > gpurun_out/${TAG}${v}_bench_${c}64.json 2> gpurun_out/${TAG}${v}_bench_${c}64.err
grep -a "femgpu asm" gpurun_out/${TAG}${v}_bench_${c}64.err | head -1
python -c "
import json;d=json.loads(open('gpurun_out/${TAG}${v}_bench_${c}64.json').read().strip().splitlines()[-1]);print('RESULT','$v','${c}64',d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac'])"
done
done
