cd /root/repo
python -m pytest tests/test_gpu_classic.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -2
for mb in 0 6 7; do
echo "MINB=$mb"
AVTEX_FILTER_MINB=$mb python profiles/r02_kernels.py filter1s 16000
done
AVTEX_FILTER_MINB=6 python profiles/r02_kernels.py filter1 16000
python bench.py > gpurun_out/r2_b15.json 2> gpurun_out/r2_b15.err; python - <<'P'
import json
d=json.load(open('gpurun_out/r2_b15.json'))
print('ms/step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'csr',d['extra']['synth']['classic_walk_c2']['survivor_csr_ms'])
print(d['extra']['c5']['ms_per_step'], d['extra']['c5']['stages_ms_max_over_ranks'])
P
