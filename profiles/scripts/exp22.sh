cd /root/repo
timeout 900 compute-sanitizer --tool racecheck python profiles/r02_sanitize.py > gpurun_out/r02f_sanitizer_racecheck.log 2>&1
tail -3 gpurun_out/r02f_sanitizer_racecheck.log
python -m pytest tests/test_gpu_classic.py -x -q -m gpu 2>&1 | tail -2
for mode in fresh big_first big_freed; do python profiles/scripts/probe_fc_place.py $mode 24991; done
python profiles/scripts/probe_fc_place.py fresh 24961
python profiles/scripts/probe_fc_place.py fresh 25024
