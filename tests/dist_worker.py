"""torchrun worker for tests/test_gpu_dist.py: row-sharded classic++ vs the single-GPU result."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_video_textures_b200 import dist as avd  # noqa: E402
from audio_video_textures_b200 import engine, selfcheck  # noqa: E402
from audio_video_textures_b200.classic.video_textures import texture_walk  # noqa: E402
from audio_video_textures_b200.synth import synth_video  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, h, w, fs, stride = (int(v) for v in sys.argv[1:6])
    th, f = 0.08, 4.5
    frames = synth_video(n, h, w, seed=2).cuda()
    mode = sys.argv[6] if len(sys.argv) > 6 else "rows"
    ws = avd.SymmetricShardWorkspace(n, fs, stride, rank, world, frames.device) if mode == "sym" else None
    for _ in range(3 if mode == "sym" else 1):          # repeated calls reuse the symmetric buffers (parities, flag epochs)
        res = avd.classic_sharded(frames, fs, stride, rank, world, sigma_factor=f, threshold=th, workspace=ws)
    rowptr, colidx = avd.gather_survivors(res)
    p_m = res.plan.m
    single = selfcheck.single_gpu_pipeline(frames, fs, stride, f, th)
    pfs = avd.pack_frames_sharded(frames, rank, world)
    assert torch.equal(pfs.sqnorm, single["pf"].sqnorm) and pfs.exact_ok == single["pf"].exact_ok
    ok = selfcheck.shard_equals_single(res, single)
    rp1, ci1 = engine.csr_from_matrix(single["P3n"], single["counts"])
    ok["csr"] = np.array_equal(rowptr, rp1) and np.array_equal(colidx, ci1)
    np.random.seed(0)
    a = texture_walk((rowptr, colidx), 1, 30, 5, stride, fs)
    np.random.seed(0)
    b = texture_walk((rp1, ci1), 1, 30, 5, stride, fs)
    ok["walk"] = a == b
    if ws is not None:                                   # clip replication by NVLink pushes under the PCIe copy
        host = frames.cpu().pin_memory()
        for _ in range(2):
            got = avd.load_frames_pushed(host, ws, chunks=3)
            torch.cuda.synchronize()
            ok["pushed_frames"] = ok.get("pushed_frames", True) and torch.equal(got, frames.reshape(n, -1))
        res2 = avd.classic_sharded(got, fs, stride, rank, world, sigma_factor=f, threshold=th, workspace=ws)
        ok["pushed_pipeline"] = torch.equal(res2.D3_new, res.D3_new)
    if ws is not None:                                   # one-sided lazy rows out of the peers' symmetric P3_new shards
        ws.barrier(2)
        if rank == 0:
            rows = avd.sharded_survivor_rows(res, ws)
            np.random.seed(0)
            c = texture_walk(rows, 1, 30, 5, stride, fs)
            ok["lazy_walk"] = c == b and all(np.array_equal(rows[i], ci1[rp1[i]:rp1[i + 1]]) for i in (0, p_m // 2, p_m - 1))
        dist.barrier()
    flag = torch.tensor([int(all(ok.values()))], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    p = res.plan
    if rank == 0:
        print("DIST_CHECK", "PASS" if int(flag) == 1 else "FAIL", mode, ok, "world", world, "M", p.m, "sweeps", res.n_sweeps)
    else:
        if not all(ok.values()):
            print("rank", rank, ok)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
