cd /root/repo
python -m pytest tests/test_gpu_classic.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -4
python profiles/r02_kernels.py filter1 16000
python profiles/r02_kernels.py filter1s 16000
python profiles/r02_kernels.py filter4s
