set -x
mkdir -p gpurun_out
TAG=${TAG:-r1s3o}
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for c in M B; do
python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err
tail -2 gpurun_out/${TAG}_bench_$c.err
python -c "
import json;d=json.loads(open('gpurun_out/${TAG}_bench_$c.json').read().strip().splitlines()[-1]);print(d['value'],d['roofline']['kernel_ms'],d['separation'])"
done
