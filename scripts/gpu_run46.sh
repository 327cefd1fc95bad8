set -x
mkdir -p gpurun_out
FEMGPU_SYM_TIMING=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-separation > gpurun_out/r1n_bench_M.json 2> gpurun_out/r1n_bench_M.err
python -c "
import json;d=json.loads(open('gpurun_out/r1n_bench_M.json').read().strip().splitlines()[-1]);print('RESULT',d['value'],d['e2e']['value'],d['e2e']['step_seconds'],d['e2e']['phases_last_step'])"
grep -a "femgpu symbolic" gpurun_out/r1n_bench_M.err | tail -14
