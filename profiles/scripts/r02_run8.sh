set -x
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_dist.py 2>&1 | tail -30 > gpurun_out/r2_t8.log
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_b8.json 2> gpurun_out/r2_b8.err
tail -12 gpurun_out/r2_t8.log; tail -3 gpurun_out/r2_b8.err
