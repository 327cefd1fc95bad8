# final bench lines (all legs) after the host-side change; kernels unchanged since the r1i captures
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1k}
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_M.json 2> gpurun_out/${TAG}_bench_M.err
for c in P B T; do
python bench.py --config $c --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r1k_bench_*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
    print(f, 'value=%.4g ms=%.3f'%(d['value'],d['ms_per_step']), r['kernel_ms'], r['frac'], 'e2e', d['e2e']['value'], d['e2e']['seconds_per_step'], 'sep', d['separation']['ms'])
PY
