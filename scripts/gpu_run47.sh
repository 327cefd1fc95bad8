set -x
mkdir -p gpurun_out
timeout 100 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r1o_pytest_gpu.txt
timeout 100 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-separation > gpurun_out/r1o_bench_M.json 2> gpurun_out/r1o_bench_M.err
python -c "
import json;d=json.loads(open('gpurun_out/r1o_bench_M.json').read().strip().splitlines()[-1]);print('RESULT',d['value'],d['e2e']['value'],d['e2e']['step_seconds'],d['e2e']['phases_last_step'])"
