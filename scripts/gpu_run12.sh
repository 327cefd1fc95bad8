# stage cap fix (B), cost-weight sweep on M
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1s3f}
for c in B M; do
python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_${c}.json 2> gpurun_out/${TAG}_bench_${c}.err
done
FEMGPU_ASM_THREADS=32 python bench.py --config B --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_B32.json 2> gpurun_out/${TAG}_bench_B32.err
for v in c12_2 c12_4 c16_4 c16_8 c10_3 c6_1; do for c in M; do
FEMGPU_LIB=$PWD/finite_element_method_b200/libfemgpu_$v.so python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_${c}_$v.json 2> gpurun_out/${TAG}_bench_${c}_$v.err
done; done
python - <<'PY'
import json,glob,os
tag=os.environ.get('TAG','r1s3f')
for f in sorted(glob.glob(f'gpurun_out/{tag}_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d['roofline']
        print(f, 'Gelem/s=%.3f step_ms=%.3f asm_ms=%.3f prep_ms=%.3f frac=%.3f'%(d['value']/1e9,d['ms_per_step'],r['kernel_ms'],r['prep_ms'],r['frac']))
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-2000:])
PY
