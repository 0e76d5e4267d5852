cd /root/repo
ncu --set full --import-source on --clock-control none -k regex:diag_filter -s 3 -c 1 -o gpurun_out/exp8_filter1s python profiles/r02_kernels.py filter1s 16000 > gpurun_out/exp8_ncu.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:diag_filter -s 3 -c 1 -o gpurun_out/exp8_filter1 python profiles/r02_kernels.py filter1 16000 >> gpurun_out/exp8_ncu.log 2>&1
