# direct-store variant (no slab image): parity + bench + phases
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1s3d}
FEMGPU_ASM_DIRECT=1 timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -3
for c in M P; do for t in 32 64; do
FEMGPU_ASM_DIRECT=1 FEMGPU_ASM_THREADS=$t python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_${c}${t}d.json 2> gpurun_out/${TAG}_bench_${c}${t}d.err
done; done
for q in 108; do for c in M P; do for t in 32 64; do
FEMGPU_SLAB_QUOTA=$q FEMGPU_ASM_DIRECT=1 FEMGPU_ASM_THREADS=$t python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_${c}${t}d_q$q.json 2> gpurun_out/${TAG}_bench_${c}${t}d_q$q.err
done; done; done
for c in B T; do
FEMGPU_ASM_DIRECT=1 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_${c}d.json 2> gpurun_out/${TAG}_bench_${c}d.err
done
python - <<'PY'
import json,glob,os
tag=os.environ.get('TAG','r1s3d')
for f in sorted(glob.glob(f'gpurun_out/{tag}_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d['roofline']
        print(f, 'Gelem/s=%.3f step_ms=%.3f asm_ms=%.3f prep_ms=%.3f frac=%.3f'%(d['value']/1e9,d['ms_per_step'],r['kernel_ms'],r['prep_ms'],r['frac']))
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-2000:])
PY
export FEMGPU_LIB=$PWD/finite_element_method_b200/libfemgpu_prof.so
for c in M P; do for t in 32 64; do
FEMGPU_ASM_DIRECT=1 FEMGPU_PHASE_DUMP=1 FEMGPU_ASM_THREADS=$t python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_phase_${c}$t.json 2> gpurun_out/${TAG}_phase_${c}$t.err
grep "femgpu phases" gpurun_out/${TAG}_phase_${c}$t.err | tail -11
done; done
