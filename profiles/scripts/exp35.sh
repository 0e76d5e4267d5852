cd /root/repo
python profiles/r02_kernels.py gramfused
ncu --set full --clock-control none -k regex:gram_l2_s8_2cta -s 9 -c 1 -o gpurun_out/r02f_gram_c2_fused_norms python profiles/r02_kernels.py gramfused > /dev/null 2>&1
ls -la gpurun_out/r02f_gram_c2_fused_norms.ncu-rep
