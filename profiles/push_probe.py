"""NVLink push probe (torchrun, world >= 2): the off-diagonal Gram jobs of a sharded step (every tile pushes its
transpose to a peer), timed with the transposed destinations in PEER memory and in LOCAL memory.
    torchrun --nproc-per-node 2 profiles/push_probe.py [n_frames]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from audio_video_textures_b200 import dist as D
from audio_video_textures_b200 import engine
from audio_video_textures_b200.synth import synth_video_cuda

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
fs, stride = 40, 4
ws = D.SymmetricShardWorkspace(n, fs, stride, rank, world, torch.device("cuda"))
frames = synth_video_cuda(n, 64, 64, seed=0)
pf = engine.pack_frames(frames)
variants = {
    "peer": ws.d1_ptrs,
    "local": [ws.d1_ptrs[rank]] * world,
}
for name, ptrs in variants.items():
    jobs = D.symmetric_jobs(ws.plans, rank, ptrs, ws.ld, stride)
    for sel, jl in (("all", jobs), ("offdiag", jobs[1:]), ("diag", jobs[:1])):
        ms = []
        for it in range(6):
            ws.barrier(channel=0)
            torch.cuda.synchronize()
            dist.barrier()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
            engine.gram_l2_jobs(pf, jl)
            ev[1].record()
            ws.barrier(channel=1)
            ev[2].record()
            torch.cuda.synchronize()
            ms.append((ev[0].elapsed_time(ev[1]), ev[0].elapsed_time(ev[2])))
        k, kb = sorted(m[0] for m in ms[2:])[len(ms[2:]) // 2], sorted(m[1] for m in ms[2:])[len(ms[2:]) // 2]
        pushed = sum(j["rows"] * j["cols"] * 4 for j in jl if not j["symmetric"]) / 1e9
        print(f"rank {rank}/{world} n={n} dest={name} jobs={sel}: kernel {k:.3f} ms, with barrier {kb:.3f} ms, "
              f"pushed {pushed:.2f} GB = {pushed / k * 1e3:.0f} GB/s", flush=True)
dist.destroy_process_group()
