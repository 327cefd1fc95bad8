set -x
mkdir -p gpurun_out
python bench.py --config P --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r1e_bench_P.json 2> gpurun_out/r1e_bench_P.err
python bench.py --config M --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r1e_bench_M.json 2> gpurun_out/r1e_bench_M.err
ncu --set full --clock-control none --import-source on -k regex:assemble_kernel -s 1 -c 1 -o gpurun_out/prof_P_r1e -f python bench.py --config P --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_P_r1e.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:assemble_kernel -s 1 -c 1 -o gpurun_out/prof_M_r1e -f python bench.py --config M --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_M_r1e.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_M_r1e.csv python bench.py --config M --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/launches_M_r1e.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
