#!/bin/bash
# round 2, job e (8 GPUs): the multi-rank tests that use every GPU of the box, then the default bench line at N = 8
out=gpurun_out/r2e_n8; mkdir -p $out
nvidia-smi --query-gpu=index,name --format=csv,noheader > $out/gpus.txt
FEMGPU_DIST_INFO=1 timeout 900 python -m pytest tests/test_dist_gpu.py -q -s -k "mixed or full_size" > $out/pytest_dist_n8.txt 2>&1; echo "dist tests rc=$?"; grep -E "DIST_OK|passed|failed" $out/pytest_dist_n8.txt | cut -c1-300
FEMGPU_DIST_INFO=1 timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 \
  bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > $out/bench_n8.json 2> $out/bench_n8.err; echo "bench N=8 rc=$?"
grep '^{' $out/bench_n8.json | python -c '
import sys,json
d=json.loads(sys.stdin.readline()); r=d["roofline"]
print("N=8", d["scaling"], "value %.3f G elem/s  step %.3f ms  kernel %.3f prep %.3f xchg %.3f" % (d["value"]/1e9, d["ms_per_step"], r["kernel_ms"], r["prep_ms"], r["exchange_ms"]))
print("weak", d.get("weak")); print("e2e", d.get("e2e")); print(d["config"]["parallelism"])'
tail -5 $out/bench_n8.err
