# session 4: family-striped chunks of split blocks. parity subset + quick bench lines + phase dump
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1s4a}
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_separation.py -m gpu -x -q 2>&1 | tail -6
for c in M B P T; do
timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-separation > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err
tail -2 gpurun_out/${TAG}_bench_$c.err
python -c "
import json;d=json.loads(open('gpurun_out/${TAG}_bench_$c.json').read().strip().splitlines()[-1]);print('$c',d['value'],d['ms_per_step'],d['roofline'])"
done
FEMGPU_LIB=$PWD/finite_element_method_b200/libfemgpu_prof.so FEMGPU_PHASE_DUMP=1 timeout 600 python bench.py --config M --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-separation 2>&1 | grep -a "femgpu phases" | tail -14
