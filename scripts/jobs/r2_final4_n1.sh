#!/bin/bash
# round 2, session 3: the default bench line with the multi-core CPU port at full size (wall time of the whole command noted)
out=gpurun_out/r2_final4_n1; mkdir -p $out
t0=$(date +%s)
timeout 600 python bench.py > $out/bench_M.json 2> $out/bench_M.err; echo "bench M (defaults) rc=$? wall $(( $(date +%s) - t0 )) s"
python - <<'PY'
import json
for l in open("gpurun_out/r2_final4_n1/bench_M.json"):
    if l.startswith("{"):
        d = json.loads(l); e = d["e2e"]; c = d["cpu_baseline"]
        print("step %.3f ms frac %.3f step_frac %.3f; e2e %.2f M elem/s (%.4f s)" % (d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["whole_step_frac"], e["value"] / 1e6, e["seconds_per_step"]))
        print("cpu:", c["value"], c["optimized_multicore_port"])
PY
grep "cpu baseline\|done\|e2e\[" $out/bench_M.err | tail -4
