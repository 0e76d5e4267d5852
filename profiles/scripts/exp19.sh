cd /root/repo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_b19_g2.json 2> gpurun_out/r2_b19_g2.err
tail -5 gpurun_out/r2_b19_g2.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_b19_g2.json').read().strip().splitlines()[-1])
print('ms/step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'parity',d.get('parity'))
print('stages',d.get('stages'))
print('residue',d['extra'].get('residue_pipeline'))
print('parity rec',d['extra'].get('parity'))
c=d['extra']['c5']; print('c5',c.get('ms_per_step'),c.get('stages_ms_max_over_ranks')); print('c5 residue',c.get('residue_pipeline'))
P
