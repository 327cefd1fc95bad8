# separation feature: gpu tests (all), bench M/P/B/T incl. e2e (pinned uploads) and separation timing
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1s3i}
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${TAG}_pytest_gpu.txt
cat gpurun_out/${TAG}_pytest_gpu.txt
for c in M P B T; do
FEMGPU_SYM_TIMING=1 python bench.py --config $c --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err
tail -3 gpurun_out/${TAG}_bench_$c.err
done
python - <<'PY'
import json,glob,os
tag=os.environ.get('TAG','r1s3i')
for f in sorted(glob.glob(f'gpurun_out/{tag}_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d['roofline']
        print(f, 'Gelem/s=%.3f step_ms=%.3f asm_ms=%.3f prep_ms=%.3f frac=%.3f'%(d['value']/1e9,d['ms_per_step'],r['kernel_ms'],r['prep_ms'],r['frac']))
        print('   e2e', d['e2e']['value'], d['e2e']['phases_last_step'])
        print('   sep', d['separation'])
        print('   clocks', d['clocks'])
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-2000:])
PY
