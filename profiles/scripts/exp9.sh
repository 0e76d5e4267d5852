cd /root/repo
python -m pytest tests/test_gpu_classic.py -x -q -m gpu -k "filter" 2>&1 | tail -2
python profiles/r02_kernels.py filter1s 16000
python profiles/r02_kernels.py filter1s 20000
python profiles/r02_kernels.py filter4s
