#!/bin/bash
# round 2, job j (4 GPUs): multi-rank tests on 4 ranks, bench at N = 4 (strong + weak + e2e with both read-backs), and the
# new single-GPU tests (non-zero CSR read-back, load order)
out=gpurun_out/r2j_n4; mkdir -p $out
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_separation.py -m gpu -q -k "nonzero_csr or call_order" > $out/pytest_new.txt 2>&1; echo "new tests rc=$?"; tail -3 $out/pytest_new.txt
FEMGPU_DIST_INFO=1 timeout 900 python -m pytest tests/test_dist_gpu.py -q -s -k "mixed or full_size" > $out/pytest_dist_n4.txt 2>&1; echo "dist tests rc=$?"; grep -E "DIST_OK|passed|failed" $out/pytest_dist_n4.txt | cut -c1-300
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 \
  bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu-baseline > $out/bench_n4.json 2> $out/bench_n4.err; echo "bench N=4 rc=$?"
grep '^{' $out/bench_n4.json | python -c '
import sys,json
d=json.loads(sys.stdin.readline()); r=d["roofline"]
print("N=4", d["scaling"], "value %.3f G elem/s  step %.3f ms  kernel %.3f (x%d) prep %.3f xchg %.3f" % (d["value"]/1e9, d["ms_per_step"], r["kernel_ms"], r["kernel_launches_per_step"], r["prep_ms"], r["exchange_ms"]))
w=d.get("weak"); print("weak value %.3f G/s step %.3f" % (w["value"]/1e9, w["ms_per_step"]))
e=d["e2e"]; print("e2e nonzero %.1f M/s %.3f s d2h %.2f GB" % (e["value"]/1e6, e["seconds_per_step"], e["d2h_bytes_per_step"]/1e9), e["phases_last_step"])
e=e["structural_readback"]; print("e2e structural %.1f M/s %.3f s d2h %.2f GB" % (e["value"]/1e6, e["seconds_per_step"], e["d2h_bytes_per_step"]/1e9), e["phases_last_step"])'
tail -3 $out/bench_n4.err
