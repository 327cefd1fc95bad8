set -x
mkdir -p gpurun_out
TAG=${TAG:-r1q}
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -3
for c in M P B T; do
python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err
done
FEMGPU_ASM_THREADS=64 python bench.py --config P --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_P64.json 2> gpurun_out/${TAG}_bench_P64.err
FEMGPU_ASM_THREADS=32 python bench.py --config M --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_M32.json 2> gpurun_out/${TAG}_bench_M32.err
FEMGPU_ASM_THREADS=64 python bench.py --config T --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_T64.json 2> gpurun_out/${TAG}_bench_T64.err
FEMGPU_ASM_THREADS=32 python bench.py --config B --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_B32.json 2> gpurun_out/${TAG}_bench_B32.err
python - <<'PY'
import json,glob,os
tag=os.environ.get('TAG','r1q')
for f in sorted(glob.glob(f'gpurun_out/{tag}_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d['roofline']
        print(f, 'Gelem/s=%.3f step_ms=%.3f asm_ms=%.3f prep_ms=%.3f frac=%.3f'%(d['value']/1e9,d['ms_per_step'],r['kernel_ms'],r['prep_ms'],r['frac']))
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-2000:])
PY
