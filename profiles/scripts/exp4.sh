cd /root/repo
profiles/scripts/bin/cluster_occ
python -m pytest tests/test_gpu_classic.py -x -q -m gpu -k "filter or dropin" 2>&1 | tail -2
AVTEX_FILTER_R=32 python -m pytest tests/test_gpu_classic.py -x -q -m gpu -k "filter or dropin" 2>&1 | tail -2
python profiles/r02_kernels.py filter1 16000
AVTEX_FILTER_R=32 python profiles/r02_kernels.py filter1 16000
M="gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second,sm__warps_active.avg.pct_of_peak_sustained_active"
ncu --metrics $M --clock-control none -k regex:diag_filter -c 2 --csv --log-file gpurun_out/exp4_f16.csv python profiles/r02_kernels.py filter1 16000 > /dev/null 2>&1
AVTEX_FILTER_R=32 ncu --metrics $M --clock-control none -k regex:diag_filter -c 2 --csv --log-file gpurun_out/exp4_f32.csv python profiles/r02_kernels.py filter1 16000 > /dev/null 2>&1
grep diag_filter gpurun_out/exp4_f16.csv | awk -F'","' '{print $(NF-2), $NF}' | head -5
grep diag_filter gpurun_out/exp4_f32.csv | awk -F'","' '{print $(NF-2), $NF}' | head -5
python profiles/r02_kernels.py gramjobs 0
python profiles/r02_kernels.py gramjobs 3
