# session 4: direct solve (skyline LDL^T) tests + the whole gpu suite + smoke on the final tree
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1l}
timeout 600 python -m pytest tests/test_separation.py -m gpu -x -q 2>&1 | tail -8
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/${TAG}_pytest_gpu.txt
cat gpurun_out/${TAG}_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
