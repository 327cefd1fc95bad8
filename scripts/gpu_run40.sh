# session 4: host bookkeeping of a re-used instance (storage kept across reset, appends on all cores) — e2e
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1j}
timeout 900 python -m pytest tests/test_host_logic.py tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -2
for c in M P B T; do
timeout 900 python bench.py --config $c --steps 20 --warmup 3 --no-separation > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err
python -c "
import json;d=json.loads(open('gpurun_out/${TAG}_bench_$c.json').read().strip().splitlines()[-1]);print('RESULT','$c',d['value'],d['roofline']['kernel_ms'],d['e2e']['value'],d['e2e']['seconds_per_step'],d['e2e']['phases_last_step'])"
done
