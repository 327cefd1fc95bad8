#!/bin/bash
# round 2, final multi-GPU evidence run on N GPUs: the multi-rank tests, then the default bench line at N
N=$1
out=gpurun_out/r2_final_n$N; mkdir -p $out
nvidia-smi --query-gpu=index,name --format=csv,noheader > $out/gpus.txt
timeout 300 python __graft_entry__.py smoke > $out/smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 $out/smoke.txt
if [ "$N" = "2" ]; then K=""; else K='-k mixed_or_full_size'; K='-k "mixed or full_size"'; fi
if [ "$N" = "2" ]; then
  FEMGPU_DIST_INFO=1 timeout 900 python -m pytest tests/test_dist_gpu.py -q -s > $out/pytest_dist.txt 2>&1
else
  FEMGPU_DIST_INFO=1 timeout 900 python -m pytest tests/test_dist_gpu.py -q -s -k "mixed or full_size" > $out/pytest_dist.txt 2>&1
fi
echo "dist tests rc=$?"; grep -E "DIST_OK|passed|failed" $out/pytest_dist.txt | cut -c1-300
FEMGPU_DIST_INFO=1 timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 \
  bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > $out/bench_n$N.json 2> $out/bench_n$N.err; echo "bench N=$N rc=$?"
grep '^{' $out/bench_n$N.json | python -c '
import sys,json
d=json.loads(sys.stdin.readline()); r=d["roofline"]
print("N=%d" % d["n_gpus"], d["scaling"], "value %.3f G elem/s  step %.3f ms  kernel %.3f (x%d) prep %.3f xchg %.3f" % (d["value"]/1e9, d["ms_per_step"], r["kernel_ms"], r["kernel_launches_per_step"], r["prep_ms"], r["exchange_ms"]))
w=d.get("weak"); print("weak value %.3f G/s step %.3f" % (w["value"]/1e9, w["ms_per_step"]))
e=d["e2e"]; print("e2e nonzero %.1f M/s %.3f s d2h %.2f GB" % (e["value"]/1e6, e["seconds_per_step"], e["d2h_bytes_per_step"]/1e9), e["phases_last_step"])
e=e["structural_readback"]; print("e2e structural %.1f M/s %.3f s d2h %.2f GB" % (e["value"]/1e6, e["seconds_per_step"], e["d2h_bytes_per_step"]/1e9))'
if [ "$N" = "2" ]; then
  for c in B P; do timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 \
    bench.py --gpus $N --config $c --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $out/bench_${c}_n$N.json 2>/dev/null; grep '^{' $out/bench_${c}_n$N.json | python -c '
import sys,json
d=json.loads(sys.stdin.readline()); print(d["config"]["mesh"], "N=%d value %.3f G/s step %.3f ms" % (d["n_gpus"], d["value"]/1e9, d["ms_per_step"]))'; done
fi
tail -2 $out/bench_n$N.err
