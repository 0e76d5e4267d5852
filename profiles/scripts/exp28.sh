cd /root/repo
python bench.py --skip_extra > gpurun_out/r2_b28.json 2> gpurun_out/r2_b28.err; tail -1 gpurun_out/r2_b28.err | cut -c1-200
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 5 --skip_extra > gpurun_out/r2_b28_g2.json 2> gpurun_out/r2_b28_g2.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_b28.json'))
print('N=1 ms/step',d['ms_per_step'],'eager',d['extra']['eager']['ms_per_step'],'gram',d['roofline']['ms'],'frac',d['roofline']['frac'])
d=json.loads(open('gpurun_out/r2_b28_g2.json').read().strip().splitlines()[-1])
print('N=2 ms/step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'stages',d.get('stages'),d.get('parity'))
P
