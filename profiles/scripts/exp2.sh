cd /root/repo
python -m pytest tests/test_gpu_classic.py -x -q -m gpu 2>&1 | tail -3
python profiles/r02_kernels.py filter1 16000
python profiles/r02_kernels.py filter1 20000
python profiles/r02_kernels.py gramsym 40000
python profiles/r02_kernels.py gramsym 100000
AVTEX_GRAM_ST=1 AVTEX_GRAM_HINT=2 python profiles/r02_kernels.py gramsym 100000
AVTEX_GRAM_GROUP=8 python profiles/r02_kernels.py gramsym 100000
ncu --set full --clock-control none -k regex:gram_l2 -c 1 -o gpurun_out/exp2_gramsym100k python profiles/r02_kernels.py gramsym 100000 > gpurun_out/exp2_ncu.log 2>&1
