#!/bin/bash
# One parameterised GPU job (replaces the one-off scripts/gpu_run*.sh of round 1). Runs from the repo root on the box.
#   scripts/gpu_job.sh <name> <steps...>     steps: tests | dist | bench:<args> | benchn:<N>:<args> | ncu:<kernel-regex>:<args>
# Everything is written under gpurun_out/<name>/.
set -u
name=$1; shift
out=gpurun_out/$name
mkdir -p "$out"
nvidia-smi --query-gpu=index,name,memory.total --format=csv > "$out/gpus.txt" 2>&1
i=0
for step in "$@"; do
  i=$((i+1))
  kind=${step%%:*}
  rest=${step#*:}
  case $kind in
    tests)
      timeout 1500 python -m pytest tests -m gpu -x -q -s --deselect tests/test_dist_gpu.py > "$out/${i}_pytest_gpu.txt" 2>&1
      echo "[$name] tests rc=$?"; tail -3 "$out/${i}_pytest_gpu.txt" ;;
    dist)
      FEMGPU_DIST_INFO=1 timeout 1500 python -m pytest tests/test_dist_gpu.py -x -q -s > "$out/${i}_pytest_dist.txt" 2>&1
      echo "[$name] dist rc=$?"; tail -5 "$out/${i}_pytest_dist.txt" ;;
    bench)
      timeout 900 python bench.py $rest > "$out/${i}_bench.json" 2> "$out/${i}_bench.err"
      echo "[$name] bench $rest rc=$?"; cut -c1-600 "$out/${i}_bench.json"; tail -3 "$out/${i}_bench.err" ;;
    benchn)
      n=${rest%%:*}; args=${rest#*:}
      FEMGPU_DIST_INFO=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port 29544 \
        bench.py --gpus "$n" $args > "$out/${i}_bench_n$n.json" 2> "$out/${i}_bench_n$n.err"
      echo "[$name] bench N=$n $args rc=$?"; grep '^{' "$out/${i}_bench_n$n.json" | cut -c1-800; tail -4 "$out/${i}_bench_n$n.err" ;;
    ncu)
      k=${rest%%:*}; args=${rest#*:}
      timeout 1200 ncu --set full --import-source on --clock-control none -k "regex:$k" -c 2 -o "$out/${i}_ncu" -f python bench.py $args > "$out/${i}_ncu.log" 2>&1
      echo "[$name] ncu $k rc=$?" ;;
    launches)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/${i}_launches.csv" python bench.py $rest > "$out/${i}_launches.log" 2>&1
      echo "[$name] launches rc=$?" ;;
    sh)
      timeout 1200 bash -c "$rest" > "$out/${i}_sh.txt" 2>&1
      echo "[$name] sh rc=$?"; tail -5 "$out/${i}_sh.txt" ;;
  esac
done
