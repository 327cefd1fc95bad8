# session 4: pair-exchange merge (both) and warp 1 building the next slab's plate forms (default) vs not (noahead)
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1s4l}
for v in "" _noahead; do
export FEMGPU_LIB=$PWD/finite_element_method_b200/libfemgpu$v.so
echo "=== variant '$v'"
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q 2>&1 | tail -3
for c in M B P T; do
FEMGPU_ASM_INFO=1 timeout 600 python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-separation > gpurun_out/${TAG}${v}_bench_$c.json 2> gpurun_out/${TAG}${v}_bench_$c.err
grep -a "femgpu asm" gpurun_out/${TAG}${v}_bench_$c.err | head -1 | cut -c1-70
python -c "
import json;d=json.loads(open('gpurun_out/${TAG}${v}_bench_$c.json').read().strip().splitlines()[-1]);print('RESULT','$v','$c',d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['prep_ms'],d['roofline']['frac'])"
done
done
