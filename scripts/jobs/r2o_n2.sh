#!/bin/bash
# round 2, job o (2 GPUs): single-launch ghost-first (rotated slab order + gate kernel): tests, then the three modes
out=gpurun_out/r2o_n2; mkdir -p $out
FEMGPU_DIST_INFO=1 timeout 900 python -m pytest tests/test_dist_gpu.py -q -s > $out/pytest_dist.txt 2>&1; echo "dist tests rc=$?"; grep -E "DIST_OK|passed|failed|Error" $out/pytest_dist.txt | cut -c1-300
B="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-weak"
for gf in 1 2 0; do echo "GHOST_FIRST=$gf"; FEMGPU_BENCH_DEBUG=1 FEMGPU_DIST_GHOST_FIRST=$gf timeout 300 $B 2> $out/gf$gf.err | python -c '
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]; print("step %.3f kernel %.3f (x%d) prep %.3f xchg %.3f value %.3f" % (d["ms_per_step"], r["kernel_ms"], r["kernel_launches_per_step"], r["prep_ms"], r["exchange_ms"], d["value"]/1e9))'
grep "per-pass" $out/gf$gf.err | sort | cut -c1-150; done
