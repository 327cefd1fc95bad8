# session 4: bank-spreading lane order v2 (greedy on the wavefront count) A/B
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1s4h}
for sb in 1 0; do
export FEMGPU_SPREAD_BANKS=$sb
echo "=== spread $sb"
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -2
for c in M P B T; do
timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-separation > gpurun_out/${TAG}_sb${sb}_bench_$c.json 2> gpurun_out/${TAG}_sb${sb}_bench_$c.err
python -c "
import json;d=json.loads(open('gpurun_out/${TAG}_sb${sb}_bench_$c.json').read().strip().splitlines()[-1]);print('RESULT','sb$sb','$c',d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['config']['symbolic_s'])"
done
done
