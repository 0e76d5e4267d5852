set -x
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
M3="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
ncu --metrics $M3 --clock-control none --csv --log-file gpurun_out/r02_launches_c2.csv python profiles/run_stage.py c2 2 > /dev/null 2>&1
ncu --metrics $M3 --clock-control none --csv --log-file gpurun_out/r02_launches_hbm.csv python profiles/run_stage.py hbm 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gram_l2_s8_2cta -s 1 -c 1 -o gpurun_out/r02_gram_c2 python profiles/run_stage.py c2 2 > /dev/null 2>&1
ncu --metrics $M3,lts__t_bytes.sum,sm__cycles_elapsed.avg.per_second --clock-control none -k regex:"synthesis_step|cosine|select" --csv --log-file gpurun_out/r02_launches_synth.csv python profiles/r02_kernels.py synth > gpurun_out/r02_synth_under_ncu.log 2>&1
timeout 900 compute-sanitizer --tool memcheck python profiles/r02_sanitize.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1
timeout 900 compute-sanitizer --tool racecheck python profiles/r02_sanitize.py > gpurun_out/r02_sanitizer_racecheck.log 2>&1
timeout 600 compute-sanitizer --tool synccheck python profiles/r02_sanitize.py > gpurun_out/r02_sanitizer_synccheck.log 2>&1
tail -5 gpurun_out/r02_sanitizer_memcheck.log gpurun_out/r02_sanitizer_racecheck.log gpurun_out/r02_sanitizer_synccheck.log
python profiles/r02_kernels.py synth
