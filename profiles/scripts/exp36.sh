cd /root/repo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r2_b36_g4.json 2> gpurun_out/r2_b36_g4.err
tail -2 gpurun_out/r2_b36_g4.err | cut -c1-300
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_b36_g4.json').read().strip().splitlines()[-1])
print('ms/step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'parity',d.get('parity'), d['extra']['parity'].get('residue_class_shards'))
print('residue',d['extra']['residue_pipeline'].get('ms_per_step'))
c=d['extra']['c5']; print('c5',c.get('ms_per_step'),'e2e',c['e2e']['ms_per_step'],'c5 residue',c['residue_pipeline'].get('ms_per_step'))
P
