# v7 candidates: packed pair entry prefetched a trip ahead, block metadata hoisted, beam global-frame
# record, late image-free wait. Parity + bench (+ beam cost weight variants) + phases
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1s3e}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in M P B T; do
python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_${c}.json 2> gpurun_out/${TAG}_bench_${c}.err
done
for v in cb4 cb5; do for c in M B; do
FEMGPU_LIB=$PWD/finite_element_method_b200/libfemgpu_$v.so python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_${c}_$v.json 2> gpurun_out/${TAG}_bench_${c}_$v.err
done; done
FEMGPU_ASM_THREADS=32 python bench.py --config B --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_B32.json 2> gpurun_out/${TAG}_bench_B32.err
python - <<'PY'
import json,glob,os
tag=os.environ.get('TAG','r1s3e')
for f in sorted(glob.glob(f'gpurun_out/{tag}_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d['roofline']
        print(f, 'Gelem/s=%.3f step_ms=%.3f asm_ms=%.3f prep_ms=%.3f frac=%.3f'%(d['value']/1e9,d['ms_per_step'],r['kernel_ms'],r['prep_ms'],r['frac']))
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-2000:])
PY
export FEMGPU_LIB=$PWD/finite_element_method_b200/libfemgpu_prof.so
for c in M P; do
FEMGPU_PHASE_DUMP=1 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_phase_${c}.json 2> gpurun_out/${TAG}_phase_${c}.err
grep "femgpu phases" gpurun_out/${TAG}_phase_${c}.err | tail -11
done
