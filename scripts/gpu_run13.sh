# 4 GPUs: dist tests + weak-scaling bench lines (M, P) + strong M
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1s3g}
nvidia-smi -L
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -x -q 2>&1 | tail -3
N=${N:-4}
for c in M P; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --config $c --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_${c}_n$N.json 2> gpurun_out/${TAG}_bench_${c}_n$N.err
tail -c 1500 gpurun_out/${TAG}_bench_${c}_n$N.json
tail -5 gpurun_out/${TAG}_bench_${c}_n$N.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --config M --scaling strong --steps 10 --warmup 3 --no-e2e > gpurun_out/${TAG}_bench_M_strong_n$N.json 2> gpurun_out/${TAG}_bench_M_strong_n$N.err
tail -c 1200 gpurun_out/${TAG}_bench_M_strong_n$N.json
