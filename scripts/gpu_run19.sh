# shuffle merge of split-block chunks: parity + bench + phases
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1s3m}
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for c in M P B T; do
python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-separation > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err
done
FEMGPU_ASM_THREADS=64 python bench.py --config P --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-separation > gpurun_out/${TAG}_bench_P64.json 2> gpurun_out/${TAG}_bench_P64.err
python - <<'PY'
import json,glob,os
tag=os.environ.get('TAG','r1s3m')
for f in sorted(glob.glob(f'gpurun_out/{tag}_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d['roofline']
        print(f, 'Gelem/s=%.3f step_ms=%.3f asm_ms=%.3f prep_ms=%.3f frac=%.3f'%(d['value']/1e9,d['ms_per_step'],r['kernel_ms'],r['prep_ms'],r['frac']))
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-2000:])
PY
export FEMGPU_LIB=$PWD/finite_element_method_b200/libfemgpu_prof.so
for c in M; do
FEMGPU_PHASE_DUMP=1 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-separation > gpurun_out/${TAG}_phase_${c}.json 2> gpurun_out/${TAG}_phase_${c}.err
grep "femgpu phases" gpurun_out/${TAG}_phase_${c}.err | tail -11
done
