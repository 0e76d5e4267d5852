set -x
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_dist.py -k "virtual or future_cost or fused or dropin or pipeline" 2>&1 | tail -8 > gpurun_out/r2_t4.log
python profiles/r02_kernels.py fc > gpurun_out/r2_k4.log 2>&1
python profiles/r02_kernels.py synth >> gpurun_out/r2_k4.log 2>&1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__cycles_elapsed.avg.per_second,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,l1tex__t_sector_hit_rate.pct"
ncu --metrics $M --clock-control none -k regex:"future_cost_fused" --csv --log-file gpurun_out/r2_fc.csv python profiles/r02_kernels.py fc > /dev/null 2>&1
tail -4 gpurun_out/r2_t4.log; cat gpurun_out/r2_k4.log
