set -x
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_dist.py -k "filter or dropin or pipeline or virtual or c2" 2>&1 | tail -8 > gpurun_out/r2_t6.log
python profiles/r02_kernels.py filter1 > gpurun_out/r2_k6.log 2>&1
python profiles/r02_kernels.py filter1 20000 >> gpurun_out/r2_k6.log 2>&1
AVTEX_FILTER_S1=1 python profiles/r02_kernels.py filter1 >> gpurun_out/r2_k6.log 2>&1
AVTEX_FILTER_S1=0 python profiles/r02_kernels.py filter1 >> gpurun_out/r2_k6.log 2>&1
python profiles/r02_kernels.py gram5 3124 >> gpurun_out/r2_k6.log 2>&1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__cycles_elapsed.avg.per_second,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum"
ncu --metrics $M --clock-control none -k regex:"diag_filter|gram_l2" --csv --log-file gpurun_out/r2_f1c.csv python profiles/r02_kernels.py filter1 > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:"gram_l2" --csv --log-file gpurun_out/r2_g5c.csv python profiles/r02_kernels.py gram5 3124 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:diag_filter_s1ws -s 2 -c 1 -o gpurun_out/r2_f1ws_full python profiles/r02_kernels.py filter1 > /dev/null 2>&1
tail -4 gpurun_out/r2_t6.log; cat gpurun_out/r2_k6.log
