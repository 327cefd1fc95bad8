# session 4: new parity tests + compute-sanitizer (memcheck, racecheck) over small models
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1s4p}
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "staging or shuffled" 2>&1 | tail -3
timeout 500 compute-sanitizer --tool memcheck --print-limit 5 python scripts/gpu_sanitize.py > gpurun_out/${TAG}_memcheck.txt 2>&1
tail -8 gpurun_out/${TAG}_memcheck.txt
timeout 500 compute-sanitizer --tool racecheck --print-limit 5 python scripts/gpu_sanitize.py > gpurun_out/${TAG}_racecheck.txt 2>&1
tail -8 gpurun_out/${TAG}_racecheck.txt
