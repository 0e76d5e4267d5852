cd /root/repo
python bench.py > gpurun_out/r2_b25.json 2> gpurun_out/r2_b25.err; tail -2 gpurun_out/r2_b25.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_b25.json'))
print('ms/step',d['ms_per_step'],'value',d['value'],'e2e',d['e2e']['ms_per_step'], d['config'].get('launch'))
print('eager',d['extra'].get('eager'))
r=d['extra'].get('residue_pipeline'); print('residue',{k:r[k] for k in r if k!='note'})
c=d['extra']['c5']; print('c5',c.get('ms_per_step'),c.get('stages_ms_max_over_ranks')); r=c.get('residue_pipeline'); print('c5 residue',{k:r[k] for k in r if k!='note'})
print(d['roofline']['frac'], d['roofline']['share_of_step'], d['gpu_launches'], d['wall_s'])
P
