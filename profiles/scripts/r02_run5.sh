set -x
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_dist.py -k "gram or distance or pairwise or virtual or c2 or dropin" 2>&1 | tail -8 > gpurun_out/r2_t5.log
python profiles/r02_kernels.py gram5 > gpurun_out/r2_k5.log 2>&1
python profiles/r02_kernels.py gram5 3124 >> gpurun_out/r2_k5.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --skip_extra > gpurun_out/r2_b5.json 2> gpurun_out/r2_b5.err
ncu --set full --clock-control none --import-source on -k regex:gram_l2_s8_2cta -s 2 -c 1 -o gpurun_out/r2_gram5_full python profiles/r02_kernels.py gram5 3124 > /dev/null 2>&1
tail -4 gpurun_out/r2_t5.log; cat gpurun_out/r2_k5.log; tail -3 gpurun_out/r2_b5.err
