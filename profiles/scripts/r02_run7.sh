set -x
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_contrastive.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2_t7.log
python profiles/r02_kernels.py synth > gpurun_out/r2_k7.log 2>&1
tail -4 gpurun_out/r2_t7.log; cat gpurun_out/r2_k7.log
