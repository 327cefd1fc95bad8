#!/bin/bash
# round 2, session 3: the final bench.py at N = 2 exactly as the driver launches it (defaults)
out=gpurun_out/r2_final5_n2; mkdir -p $out
t0=$(date +%s)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 \
  bench.py --gpus 2 --steps 20 --warmup 3 > $out/bench_n2.json 2> $out/bench_n2.err; echo "bench N=2 rc=$? wall $(( $(date +%s) - t0 )) s"
grep '^{' $out/bench_n2.json | python -c '
import sys,json
d=json.loads(sys.stdin.readline()); r=d["roofline"]
print("N=%d" % d["n_gpus"], d["scaling"], "value %.3f G elem/s  step %.3f ms; cpu_baseline" % (d["value"]/1e9, d["ms_per_step"]), d["cpu_baseline"])
e=d["e2e"]; print("e2e nonzero %.1f M/s %.3f s" % (e["value"]/1e6, e["seconds_per_step"]), e["phases_last_step"])'
tail -2 $out/bench_n2.err
