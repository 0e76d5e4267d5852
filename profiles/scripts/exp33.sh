cd /root/repo
python -m pytest tests/test_gpu_dist.py -x -q -m gpu 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_b33_g2.json 2> gpurun_out/r2_b33_g2.err
tail -2 gpurun_out/r2_b33_g2.err | cut -c1-300
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_b33_g2.json').read().strip().splitlines()[-1])
print('ms/step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'parity',d.get('parity'))
c=d['extra']['c5']; print('c5',c['ms_per_step'],'e2e',c['e2e']['ms_per_step'])
P
