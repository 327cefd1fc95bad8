#!/bin/bash
# round 2, job g (1 GPU): GPU test-suite, default bench line, ncu evidence (launch list, --set full of the assembly and the
# record kernels, FP64 instruction counts)
out=gpurun_out/r2g_n1; mkdir -p $out
timeout 1700 python -m pytest tests -m gpu -q -s --deselect tests/test_dist_gpu.py > $out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_gpu.txt
timeout 600 python bench.py --steps 20 --warmup 3 > $out/bench_M.json 2> $out/bench_M.err; echo "bench rc=$?"; cut -c1-400 $out/bench_M.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference.json 2>/dev/null; cut -c1-300 $out/bench_reference.json
A="--steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-separation"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_M.csv python bench.py $A > $out/launches_M.log 2>&1; echo "launch list rc=$?"
for c in M P; do
timeout 900 ncu --set full --import-source on --clock-control none -k regex:assemble_kernel -s 1 -c 1 -o $out/asm_$c -f python bench.py --config $c $A > $out/ncu_asm_$c.log 2>&1; echo "ncu asm $c rc=$?"
timeout 900 ncu --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__inst_executed.sum --clock-control none -k regex:assemble_kernel -s 1 -c 1 --csv --log-file $out/fp64_$c.csv python bench.py --config $c $A > /dev/null 2>&1; echo "fp64 counts $c rc=$?"
done
timeout 900 ncu --set full --import-source on --clock-control none -k regex:prep_kernel -s 3 -c 3 -o $out/prep_M -f python bench.py $A > $out/ncu_prep_M.log 2>&1; echo "ncu prep rc=$?"
ls -la $out
