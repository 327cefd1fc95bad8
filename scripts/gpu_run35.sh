# session 4: TMA issue work moved to warp 1 (two-warp shape)
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1s4m}
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q 2>&1 | tail -3
for c in M B; do
timeout 600 python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-separation > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err
python -c "
import json;d=json.loads(open('gpurun_out/${TAG}_bench_$c.json').read().strip().splitlines()[-1]);print('RESULT','$c',d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['prep_ms'],d['roofline']['frac'])"
done
