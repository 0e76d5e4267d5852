cd /root/repo
profiles/scripts/bin/cluster_occ
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.avg.per_second,lts__t_bytes.sum,sm__pipe_tensor_subpipe_imma_cycles_active_realtime.avg"
i=0
for cfg in "X=0" "AVTEX_GRAM_ST=1" "AVTEX_GRAM_ST=1 AVTEX_GRAM_HINT=1" "AVTEX_GRAM_ST=1 AVTEX_GRAM_HINT=2" "AVTEX_GRAM_ST=1 AVTEX_GRAM_HINT=1 AVTEX_GRAM_GROUP=8" "AVTEX_GRAM_ST=1 AVTEX_GRAM_HINT=2 AVTEX_GRAM_GROUP=8" "AVTEX_GRAM_ST=1 AVTEX_GRAM_HINT=1 AVTEX_GRAM_GROUP=6"; do
i=$((i+1))
echo "=== $cfg"
env $cfg ncu --metrics $M --clock-control none -k regex:gram_l2 -c 1 --csv --log-file gpurun_out/exp3_$i.csv python profiles/r02_kernels.py gramsym 100000 > /dev/null 2>&1
grep -E "gram_l2" gpurun_out/exp3_$i.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done
