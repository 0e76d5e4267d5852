set -x
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rs --deselect tests/test_gpu_dist.py 2>&1 | tail -40 > gpurun_out/r2_t3.log
python profiles/r02_kernels.py norms > gpurun_out/r2_k3.log 2>&1
python profiles/r02_kernels.py filter1 >> gpurun_out/r2_k3.log 2>&1
python profiles/r02_kernels.py filter1 20000 >> gpurun_out/r2_k3.log 2>&1
AVTEX_FILTER_S1=0 python profiles/r02_kernels.py filter1 >> gpurun_out/r2_k3.log 2>&1
python profiles/r02_kernels.py filter4 >> gpurun_out/r2_k3.log 2>&1
python profiles/r02_kernels.py synth >> gpurun_out/r2_k3.log 2>&1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__cycles_elapsed.avg.per_second,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum"
ncu --metrics $M --clock-control none -k regex:diag_filter --csv --log-file gpurun_out/r2_f1b.csv python profiles/r02_kernels.py filter1 > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:"synthesis_step|frame_norms" --csv --log-file gpurun_out/r2_sy.csv python profiles/r02_kernels.py synth > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:diag_filter_s1 -s 2 -c 1 -o gpurun_out/r2_f1b_full python profiles/r02_kernels.py filter1 > /dev/null 2>&1
tail -8 gpurun_out/r2_t3.log; cat gpurun_out/r2_k3.log
