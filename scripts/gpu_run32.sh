# session 4: two-GPU check of the new kernel (dist tests + weak-scaling bench line)
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1g}
timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_M_n2.json 2> gpurun_out/${TAG}_bench_M_n2.err
tail -3 gpurun_out/${TAG}_bench_M_n2.err
python -c "
import json;d=json.loads(open('gpurun_out/${TAG}_bench_M_n2.json').read().strip().splitlines()[-1]);print('RESULT n2',d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['exchange_ms'],(d.get('e2e') or {}).get('value'))"
