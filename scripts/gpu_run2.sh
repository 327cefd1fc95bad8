set -x
mkdir -p gpurun_out
TAG=${TAG:-r1f}
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --config P --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_P.json 2> gpurun_out/${TAG}_bench_P.err
python bench.py --config M --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_M.json 2> gpurun_out/${TAG}_bench_M.err
python bench.py --config B --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_B.json 2> gpurun_out/${TAG}_bench_B.err
python bench.py --config T --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_T.json 2> gpurun_out/${TAG}_bench_T.err
python - <<'PY'
import json,glob,os
tag=os.environ.get('TAG','r1f')
for f in sorted(glob.glob(f'gpurun_out/{tag}_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d['roofline']
        print(f, 'Gelem/s=%.3f step_ms=%.3f asm_ms=%.3f prep_ms=%.3f frac=%.3f'%(d['value']/1e9,d['ms_per_step'],r['kernel_ms'],r['prep_ms'],r['frac']))
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-2000:])
PY
if [ -n "$NCU" ]; then
ncu --set full --clock-control none --import-source on -k regex:assemble_kernel -s 1 -c 1 -o gpurun_out/prof_P_${TAG} -f python bench.py --config P --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_P_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:assemble_kernel -s 1 -c 1 -o gpurun_out/prof_M_${TAG} -f python bench.py --config M --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_M_${TAG}.log 2>&1
fi
