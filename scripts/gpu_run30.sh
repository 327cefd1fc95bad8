# session 4: run-wise TMA bulk staging of element records (FEMGPU_ASM_BULK auto / 0 / 1)
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1s4i}
for b in 1 0; do
echo "=== forced bulk $b"
FEMGPU_ASM_BULK=$b timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q 2>&1 | tail -3
done
for b in auto 0; do
if [ $b = auto ]; then unset FEMGPU_ASM_BULK; else export FEMGPU_ASM_BULK=$b; fi
echo "=== bench bulk $b"
for c in M P B T; do
FEMGPU_ASM_INFO=1 timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-separation > gpurun_out/${TAG}_b${b}_bench_$c.json 2> gpurun_out/${TAG}_b${b}_bench_$c.err
grep -a "femgpu asm" gpurun_out/${TAG}_b${b}_bench_$c.err | head -1 | cut -c1-60
python -c "
import json;d=json.loads(open('gpurun_out/${TAG}_b${b}_bench_$c.json').read().strip().splitlines()[-1]);print('RESULT','bulk=$b','$c',d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac'])"
done
done
unset FEMGPU_ASM_BULK
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3
