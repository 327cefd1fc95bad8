#!/bin/bash
# round 2, job f (2 GPUs): ghost-first exchange — multi-rank tests, then the N = 2 bench with and without it
out=gpurun_out/r2f_n2; mkdir -p $out
FEMGPU_DIST_INFO=1 timeout 900 python -m pytest tests/test_dist_gpu.py -q -s > $out/pytest_dist.txt 2>&1; echo "dist tests rc=$?"; grep -E "DIST_OK|passed|failed|Error" $out/pytest_dist.txt | cut -c1-300
B="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-weak"
summ='import sys,json
for l in sys.stdin:
    if not l.startswith("{"): continue
    d=json.loads(l); r=d["roofline"]
    print(d["config"]["mesh"], "N=%d step %.3f ms  kernel %.3f (x%d)  prep %.3f xchg %.3f  value %.3f G/s" % (d["n_gpus"], d["ms_per_step"], r["kernel_ms"], r.get("kernel_launches_per_step",1), r["prep_ms"], r["exchange_ms"], d["value"]/1e9))'
FEMGPU_BENCH_DEBUG=1 timeout 300 $B 2> $out/bench_gf.err | tee $out/bench_gf.json | python -c "$summ"; grep "per-pass" $out/bench_gf.err | cut -c1-200
FEMGPU_DIST_GHOST_FIRST=0 timeout 300 $B 2> $out/bench_nogf.err | tee $out/bench_nogf.json | python -c "$summ"
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-separation 2>/dev/null | python -c "$summ"
