#!/bin/bash
# Round-2 experiments behind the tables of profiles/README.md, one sub-command each (run from the repo root on a GPU box,
# e.g. `gpurun -- 'bash scripts/jobs/r2_experiments.sh ranges'`). The evidence runs are r2_final_n1.sh / r2_final_dist.sh.
B="--steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-separation"
summ='import sys,json
for l in sys.stdin:
    if not l.startswith("{"): continue
    d=json.loads(l); r=d["roofline"]
    print(d["config"]["mesh"], "N=%d step %.3f ms  kernel %.3f (x%d)  records %.3f  exchange %.3f  frac %.3f  step_frac %.3f" % (d["n_gpus"], d["ms_per_step"], r["kernel_ms"], r.get("kernel_launches_per_step",1), r["prep_ms"], r["exchange_ms"], r["frac"], r["whole_step_frac"]))'
case "$1" in
  ranges)      # slab-range pipeline: element records of range r+1 under the assembly of range r
    for c in M B P; do for r in 1 2 3 4 6 8; do echo "config $c RANGES=$r"; FEMGPU_NUMERIC_RANGES=$r python bench.py --config $c $B 2>/dev/null | python -c "$summ"; done; done ;;
  ablation)    # leave parts of the assembly kernel out (results wrong, time only): make variant NAME=abl$v DEFS=-DFEMGPU_ABL=$v first
    for v in 1 2 8 16 32 3; do echo "ABL=$v"; FEMGPU_LIB=$PWD/finite_element_method_b200/libfemgpu_abl$v.so python bench.py $B 2>/dev/null | python -c "$summ"; done ;;
  store)       # scripts/micro/store_bench.cu: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/micro/store_bench scripts/micro/store_bench.cu
    for cfg in "0 22464 64 4 1 0" "0 22464 64 4 1 16" "0 22464 64 5 1 0" "0 22464 64 8 1 0" "0 22464 64 4 2 0" "0 22464 64 2 4 0" "0 44928 64 4 1 0" "0 11264 64 8 1 0" \
               "0 65536 64 3 1 0" "0 22464 32 5 1 0" "2 22464 64 4 1 0" "2 22464 64 4 2 0" "2 22464 256 4 1 0" "1 22464 64 4 1 0" "1 22464 256 4 1 0" "1 22464 256 8 1 0" "1 22464 1024 2 1 0"; do
      scripts/micro/store_bench $cfg; done ;;
  weights)     # balancing weights of the per-lane work lists, slab quota
    for cb in 1 2 4 6 8 12; do echo "COST_B=$cb"; FEMGPU_COST_B=$cb python bench.py $B 2>/dev/null | python -c "$summ"; done
    for q in 60 66 78 84; do echo "QUOTA=$q"; FEMGPU_SLAB_QUOTA=$q python bench.py $B 2>/dev/null | python -c "$summ"; done ;;
  ghostfirst)  # needs N GPUs: ghost slabs first (1) or one launch with pack + apply behind it (0), per-rank per-pass times on stderr
    N=${2:-4}
    for gf in 1 0; do echo "GHOST_FIRST=$gf"; FEMGPU_BENCH_DEBUG=1 FEMGPU_DIST_GHOST_FIRST=$gf python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 29544 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-weak 2> /tmp/gf$gf.err | python -c "$summ"; grep per-pass /tmp/gf$gf.err | sort | cut -c1-150; done ;;
  p2p)         # needs 2 GPUs: peer windows vs the ncclSend/ncclRecv fallback
    for p in 1 0; do echo "P2P=$p"; FEMGPU_DIST_P2P=$p python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 \
      bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-weak 2>/dev/null | python -c "$summ"; done ;;
  seprows)     # K separation, rows per warp: make variant NAME=sep$r DEFS=-DFEMGPU_SEP_ROWS=$r first
    for r in 2 3 6 8; do echo "rows per warp $r"; FEMGPU_LIB=$PWD/finite_element_method_b200/libfemgpu_sep$r.so python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | \
      python -c 'import sys,json; d=json.loads(sys.stdin.readline()); print("separation %.2f ms" % d["separation"]["ms"])'; done ;;
  *) echo "usage: $0 ranges|ablation|store|weights|ghostfirst [N]|p2p|seprows" ;;
esac
