cd /root/repo
for i in 1 2; do
python bench.py > gpurun_out/r2_b32_$i.json 2> /dev/null
python - <<P
import json
d=json.load(open('gpurun_out/r2_b32_$i.json'))
print('run $i ms/step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'eager',d['extra']['eager']['ms_per_step'],'c5',d['extra']['c5']['ms_per_step'],'c5res',d['extra']['c5']['residue_pipeline']['ms_per_step'],'res',d['extra']['residue_pipeline']['ms_per_step'], d['clocks']['reasons'])
P
done
