set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}; mkdir -p gpurun_out
M3="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
ncu --metrics $M3 --clock-control none --csv --log-file gpurun_out/r02f_launches_c2.csv python profiles/run_stage.py c2 2 > /dev/null 2>&1
AVTEX_STAGE_RESIDUES=1 ncu --metrics $M3 --clock-control none --csv --log-file gpurun_out/r02f_launches_c2_residues.csv python profiles/run_stage.py c2 2 > /dev/null 2>&1
ncu --metrics $M3 --clock-control none --csv --log-file gpurun_out/r02f_launches_hbm.csv python profiles/run_stage.py hbm 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gram_l2_s8_2cta -s 1 -c 1 -o gpurun_out/r02f_gram_c2 python profiles/run_stage.py c2 2 > /dev/null 2>&1
AVTEX_STAGE_RESIDUES=1 ncu --set full --clock-control none --import-source on -k regex:gram_l2_s8_2cta -s 1 -c 1 -o gpurun_out/r02f_gram_c2_residues python profiles/run_stage.py c2 2 > /dev/null 2>&1
timeout 900 compute-sanitizer --tool memcheck python profiles/r02_sanitize.py > gpurun_out/r02f_sanitizer_memcheck.log 2>&1
timeout 900 compute-sanitizer --tool racecheck python profiles/r02_sanitize.py > gpurun_out/r02f_sanitizer_racecheck.log 2>&1
timeout 600 compute-sanitizer --tool synccheck python profiles/r02_sanitize.py > gpurun_out/r02f_sanitizer_synccheck.log 2>&1
tail -4 gpurun_out/r02f_sanitizer_memcheck.log gpurun_out/r02f_sanitizer_racecheck.log gpurun_out/r02f_sanitizer_synccheck.log
python -c "import __graft_entry__ as g; g.smoke()"
