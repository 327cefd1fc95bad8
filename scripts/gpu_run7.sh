# session 3 re-entry: sanity (tests + default bench), slab-quota sweep, symbolic stage timing
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1s3}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${TAG}_pytest_gpu.txt
cat gpurun_out/${TAG}_pytest_gpu.txt
FEMGPU_SYM_TIMING=1 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_M.json 2> gpurun_out/${TAG}_bench_M.err
for q in 36 54 108 144; do
for c in M P; do
FEMGPU_SLAB_QUOTA=$q python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_${c}_q$q.json 2> gpurun_out/${TAG}_bench_${c}_q$q.err
done
done
python - <<'PY'
import json,glob,os
tag=os.environ.get('TAG','r1s3')
for f in sorted(glob.glob(f'gpurun_out/{tag}_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d['roofline']
        print(f, 'Gelem/s=%.3f step_ms=%.3f asm_ms=%.3f prep_ms=%.3f frac=%.3f'%(d['value']/1e9,d['ms_per_step'],r['kernel_ms'],r['prep_ms'],r['frac']), 'e2e', (d.get('e2e') or {}).get('value'))
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-2000:])
PY
grep "femgpu" gpurun_out/${TAG}_bench_M.err | tail -60
