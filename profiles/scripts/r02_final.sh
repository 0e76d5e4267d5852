set -x
cd ${GRAFT_REPO_ROOT:-/root/repo}; mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err
timeout 900 compute-sanitizer --tool memcheck python profiles/r02_sanitize.py > gpurun_out/r02f_sanitizer_memcheck.log 2>&1
timeout 900 compute-sanitizer --tool racecheck python profiles/r02_sanitize.py > gpurun_out/r02f_sanitizer_racecheck.log 2>&1
timeout 600 compute-sanitizer --tool synccheck python profiles/r02_sanitize.py > gpurun_out/r02f_sanitizer_synccheck.log 2>&1
for f in memcheck racecheck synccheck; do tail -n 2 gpurun_out/r02f_sanitizer_$f.log; done
M3="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
ncu --metrics $M3 --clock-control none --csv --log-file gpurun_out/r02f_launches_c2.csv python profiles/run_stage.py c2 2 > /dev/null 2>&1
python -c "import __graft_entry__ as g; g.smoke()"
