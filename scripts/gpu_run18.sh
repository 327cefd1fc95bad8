# work-balancing variants on M (and the mixed parity tests on the unit-cost one)
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1s3l}
for v in u2 u4 w2 h2 h3; do
FEMGPU_LIB=$PWD/finite_element_method_b200/libfemgpu_$v.so python bench.py --config M --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-separation > gpurun_out/${TAG}_bench_M_$v.json 2> gpurun_out/${TAG}_bench_M_$v.err
done
FEMGPU_LIB=$PWD/finite_element_method_b200/libfemgpu_u2.so python bench.py --config B --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-separation > gpurun_out/${TAG}_bench_B_u2.json 2> gpurun_out/${TAG}_bench_B_u2.err
FEMGPU_LIB=$PWD/finite_element_method_b200/libfemgpu_u2.so timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -3
python - <<'PY'
import json,glob,os
tag=os.environ.get('TAG','r1s3l')
for f in sorted(glob.glob(f'gpurun_out/{tag}_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d['roofline']
        print(f, 'Gelem/s=%.3f step_ms=%.3f asm_ms=%.3f prep_ms=%.3f frac=%.3f'%(d['value']/1e9,d['ms_per_step'],r['kernel_ms'],r['prep_ms'],r['frac']))
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-2000:])
PY
