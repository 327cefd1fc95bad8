# 8 GPUs: dist test (mixed) + weak-scaling M bench line incl. e2e
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1s3h}
N=${N:-8}
timeout 300 python -m pytest tests/test_dist_gpu.py -m gpu -x -q -k mixed 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --config M --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_M_n$N.json 2> gpurun_out/${TAG}_bench_M_n$N.err
tail -c 2500 gpurun_out/${TAG}_bench_M_n$N.json
tail -5 gpurun_out/${TAG}_bench_M_n$N.err
free -g | head -2
