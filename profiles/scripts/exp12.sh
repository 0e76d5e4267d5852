cd /root/repo
python -m pytest tests/test_gpu_classic.py tests/test_gpu_virtual.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -3
python profiles/r02_kernels.py fc 20000
python profiles/r02_kernels.py fc 25000
AVTEX_FC_STAGED=0 python profiles/r02_kernels.py fc 25000
python profiles/scripts/probe_fc_csr.py fc5
