# session 4: SpMV of the PCG — row-pointer prefetch (all), 4 / 8 / 16 lanes per row
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1s4r}
for v in _sp2 _sp1 _sp4; do
export FEMGPU_LIB=$PWD/finite_element_method_b200/libfemgpu$v.so
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}${v}_bench_M.json 2> gpurun_out/${TAG}${v}_bench_M.err
python -c "
import json;d=json.loads(open('gpurun_out/${TAG}${v}_bench_M.json').read().strip().splitlines()[-1]);a=d['separation']['analysis'];print('RESULT','$v',a['pcg_jacobi']['ms_per_iteration'],a['pcg_block_jacobi']['ms_per_iteration'])"
done
