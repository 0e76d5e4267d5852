cd /root/repo
python profiles/r02_kernels.py fc 20000
