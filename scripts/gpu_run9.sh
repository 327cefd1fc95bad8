# ncu full captures (with source) of assemble_kernel on P and M
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1s3c}
for c in P M; do
ncu --set full --clock-control none --import-source on -k regex:assemble_kernel -s 1 -c 1 -o gpurun_out/${TAG}_prof_$c -f python bench.py --config $c --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_ncu_$c.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_$c.log
done
ls -la gpurun_out/
