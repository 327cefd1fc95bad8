# Round evidence run on one B200 (session 4): tests, bench lines, reference arm, ncu launch list + full captures. Outputs -> gpurun_out/
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1g}
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/${TAG}_pytest_gpu.txt
cat gpurun_out/${TAG}_pytest_gpu.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_M.json 2> gpurun_out/${TAG}_bench_M.err
for c in P B T; do
python bench.py --config $c --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err
done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_M.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-separation > gpurun_out/${TAG}_launches_M.log 2>&1
for c in M P; do
ncu --set full --clock-control none --import-source on -k regex:assemble_kernel -s 1 -c 1 -o gpurun_out/${TAG}_prof_$c -f python bench.py --config $c --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-separation > gpurun_out/${TAG}_ncu_$c.log 2>&1
done
ncu --set full --clock-control none -k regex:"spmv_dot_kernel|update_kernel|direction_kernel" -s 6 -c 3 -o gpurun_out/${TAG}_prof_pcg_M -f python bench.py --config M --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_ncu_pcg.log 2>&1
python - <<'PY'
import json,glob,os
tag=os.environ.get('TAG','r1g')
for f in sorted(glob.glob(f'gpurun_out/{tag}_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get('roofline') or {}
        print(f, 'value=%.4g ms=%.3f'%(d['value'],d['ms_per_step']), {k:r.get(k) for k in ('kernel_ms','prep_ms','frac')}, 'e2e', (d.get('e2e') or {}).get('value'), 'sep', (d.get('separation') or {}).get('ms'))
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-2000:])
PY
ls -la gpurun_out | tail -30
