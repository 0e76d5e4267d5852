cd /root/repo
python -m pytest tests/test_gpu_classic.py tests/test_gpu_virtual.py -x -q -m gpu 2>&1 | tail -5
