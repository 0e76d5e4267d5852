cd /root/repo
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py > gpurun_out/r2_b20.json 2> gpurun_out/r2_b20.err; tail -2 gpurun_out/r2_b20.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_b20.json'))
print('ms/step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'])
print('residue',d['extra'].get('residue_pipeline'))
c=d['extra']['c5']; print('c5',c.get('ms_per_step'),c.get('stages_ms_max_over_ranks')); print('c5 residue',c.get('residue_pipeline'))
for r in d['roofline_hbm']: print(r['kernel'][:70], round(r['frac'],3), round(r['ms'],3))
P
