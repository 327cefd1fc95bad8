#!/bin/bash
# round 2, session 3, 4 GPUs: the multi-rank mixed test and the default bench line at N = 4 with the final tree
out=gpurun_out/r2x_n4; mkdir -p $out
FEMGPU_DIST_INFO=1 timeout 300 python -m pytest tests/test_dist_gpu.py -q -s -k "matches_single_gpu and mixed" > $out/pytest_dist.txt 2>&1
echo "dist tests rc=$?"; grep -E "DIST_OK|passed|failed|Error|error" $out/pytest_dist.txt | cut -c1-300 | head
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 \
  bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu-baseline > $out/bench_n4.json 2> $out/bench_n4.err; echo "bench N=4 rc=$?"
grep '^{' $out/bench_n4.json | python -c '
import sys,json
d=json.loads(sys.stdin.readline()); r=d["roofline"]
print("N=%d" % d["n_gpus"], d["scaling"], "value %.3f G elem/s  step %.3f ms  kernel %.3f prep %.3f" % (d["value"]/1e9, d["ms_per_step"], r["kernel_ms"], r["prep_ms"]))
e=d["e2e"]; print("e2e nonzero %.1f M/s %.3f s d2h %.2f GB" % (e["value"]/1e6, e["seconds_per_step"], e["d2h_bytes_per_step"]/1e9), e["phases_last_step"])'
tail -2 $out/bench_n4.err
