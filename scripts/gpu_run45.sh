# final: analysis tests with the 4-lane SpMV + the default bench line
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1m}
timeout 300 python -m pytest tests/test_analysis.py -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_M.json 2> gpurun_out/${TAG}_bench_M.err
python -c "
import json;d=json.loads(open('gpurun_out/${TAG}_bench_M.json').read().strip().splitlines()[-1]);a=d['separation']['analysis'];print('RESULT',d['value'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['e2e']['value'],a['pcg_jacobi'],a['pcg_block_jacobi'])"
