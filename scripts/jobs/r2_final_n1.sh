#!/bin/bash
# round 2, final single-GPU evidence run: GPU test-suite, sanitizer, the bench lines of every configuration / variant,
# the reference arm, ncu launch list + --set full captures + FP64 instruction counts, per-phase cycle accounting
out=gpurun_out/r2_final_n1; mkdir -p $out
summ='import sys,json
for l in sys.stdin:
    if not l.startswith("{"): continue
    d=json.loads(l); r=d["roofline"]
    print(d["config"]["mesh"], "step %.3f ms  kernel %.3f  prep %.3f  frac %.3f  step_frac %.3f" % (d["ms_per_step"], r["kernel_ms"], r["prep_ms"], r["frac"], r["whole_step_frac"]))'
timeout 1700 python -m pytest tests -m gpu -q -s --deselect tests/test_dist_gpu.py > $out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_gpu.txt
for c in M P B T; do timeout 600 python bench.py --config $c --steps 20 --warmup 3 > $out/bench_$c.json 2> $out/bench_$c.err; echo "bench $c rc=$?"; python -c "$summ" < $out/bench_$c.json; done
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference.json 2>/dev/null; echo "reference arm rc=$?"
B="--steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-separation"
for v in x0 jitter; do for c in M P; do timeout 300 python bench.py --config $c --variant $v $B 2>/dev/null | tee $out/bench_${c}_$v.json | python -c "$summ"; done; done
for c in B T; do timeout 300 python bench.py --config $c --variant jitter $B 2>/dev/null | tee $out/bench_${c}_jitter.json | python -c "$summ"; done
A="--steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-separation"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_M.csv python bench.py $A > $out/launches_M.log 2>&1; echo "launch list rc=$?"
for c in M P; do
timeout 900 ncu --set full --import-source on --clock-control none -k regex:assemble_kernel -s 1 -c 1 -o $out/asm_$c -f python bench.py --config $c $A > $out/ncu_asm_$c.log 2>&1; echo "ncu asm $c rc=$?"
timeout 900 ncu --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__inst_executed.sum --clock-control none -k regex:assemble_kernel -s 1 -c 1 --csv --log-file $out/fp64_$c.csv python bench.py --config $c $A > /dev/null 2>&1; echo "fp64 counts $c rc=$?"
done
timeout 900 ncu --set full --import-source on --clock-control none -k regex:prep_kernel -s 3 -c 3 -o $out/prep_M -f python bench.py $A > $out/ncu_prep_M.log 2>&1; echo "ncu prep rc=$?"
FEMGPU_LIB=$PWD/finite_element_method_b200/libfemgpu_prof.so FEMGPU_PHASE_DUMP=1 timeout 300 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-separation 2>&1 | grep "femgpu phases" | tail -13 > $out/phases_M.txt; cat $out/phases_M.txt | tail -4
for tool in memcheck racecheck; do timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/gpu_sanitize.py > $out/sanitizer_$tool.txt 2>&1; echo "$tool rc=$?"; tail -2 $out/sanitizer_$tool.txt; done
