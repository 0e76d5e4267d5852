cd /root/repo
for flag in "--no_graph" ""; do
python bench.py --skip_extra $flag > gpurun_out/r2_b29.json 2> /dev/null
python - <<P
import json
d=json.load(open('gpurun_out/r2_b29.json'))
print('$flag', 'ms/step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'])
P
done
python profiles/scripts/probe_e2e_tail.py 2>&1 | tail -4
