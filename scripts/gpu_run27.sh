# session 4: ncu captures of the new default kernel (M, P) + bench M with the analysis leg
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1s4f}
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_M.json 2> gpurun_out/${TAG}_bench_M.err
tail -3 gpurun_out/${TAG}_bench_M.err
python -c "
import json;d=json.loads(open('gpurun_out/${TAG}_bench_M.json').read().strip().splitlines()[-1]);print(d['roofline']['kernel_ms'], json.dumps(d['separation']['analysis'], indent=1))"
for c in M P; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:assemble_kernel -s 1 -c 1 -o gpurun_out/${TAG}_prof_$c -f python bench.py --config $c --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-separation > gpurun_out/${TAG}_ncu_$c.log 2>&1
done
ls -la gpurun_out | tail -5
