# session 4: ncu of the bank-spreading lane order (P and M) to read the flush wavefronts directly
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1s4g}
export FEMGPU_SPREAD_BANKS=1
for c in P M; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:assemble_kernel -s 1 -c 1 -o gpurun_out/${TAG}_prof_sb1_$c -f python bench.py --config $c --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-separation > gpurun_out/${TAG}_ncu_$c.log 2>&1
done
ls -la gpurun_out | tail -3
