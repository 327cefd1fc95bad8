# pooled allocations + faster separation: gpu tests, bench M (x2: run-to-run spread) P B T
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1s3j}
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${TAG}_pytest_gpu.txt
cat gpurun_out/${TAG}_pytest_gpu.txt
for c in M M2 P B T; do
python bench.py --config ${c:0:1} --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err
tail -3 gpurun_out/${TAG}_bench_$c.err
done
python - <<'PY'
import json,glob,os
tag=os.environ.get('TAG','r1s3j')
for f in sorted(glob.glob(f'gpurun_out/{tag}_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d['roofline']
        print(f, 'Gelem/s=%.3f step_ms=%.3f asm_ms=%.3f prep_ms=%.3f frac=%.3f'%(d['value']/1e9,d['ms_per_step'],r['kernel_ms'],r['prep_ms'],r['frac']), 'sym_s %.3f load_s %.3f'%(d['config']['symbolic_s'], d['config']['load_s']))
        print('   e2e', d['e2e']['value'], d['e2e']['step_seconds'], d['e2e']['phases_last_step'])
        s=d['separation']; print('   sep ms %.2f GB/s %.0f frac %.3f'%(s['ms'], s['achieved_GBps'], s['frac_of_hbm_peak']), s['nnz_aa_ab_ba_bb'])
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-2000:])
PY
