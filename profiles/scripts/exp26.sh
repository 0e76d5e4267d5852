cd /root/repo
python -m pytest tests/test_gpu_classic.py -x -q -m gpu 2>&1 | tail -4
python bench.py --skip_extra > gpurun_out/r2_b26.json 2> gpurun_out/r2_b26.err; tail -2 gpurun_out/r2_b26.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_b26.json'))
print('ms/step',d['ms_per_step'],'value',d['value'], d['config'].get('launch'))
print('eager',d['extra'].get('eager'))
P
python - <<'P'
import torch, numpy as np
from audio_video_textures_b200 import engine
from audio_video_textures_b200.synth import synth_video_cuda
frames = synth_video_cuda(5000,224,224,seed=0)
flush = torch.empty(256<<20, dtype=torch.uint8, device="cuda")
for res in (False, True):
    g = engine.PipelineGraph(5000, 224*224*3, 40, 4, residues=res)
    g.frames.copy_(frames.reshape(5000,-1))
    ms=[]
    for _ in range(12):
        flush.fill_(1); torch.cuda.synchronize()
        e=[torch.cuda.Event(enable_timing=True) for _ in range(2)]
        e[0].record(); g(); e[1].record(); torch.cuda.synchronize(); ms.append(e[0].elapsed_time(e[1]))
    print("graph", g.how, "ms/step", float(np.mean(ms[2:])), "sweeps", g.n_sweeps)
P
