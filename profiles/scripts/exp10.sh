cd /root/repo
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py > gpurun_out/r2_b10.json 2> gpurun_out/r2_b10.err; tail -c 300 gpurun_out/r2_b10.json
