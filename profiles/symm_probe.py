import os, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import torch.distributed._symmetric_memory as symm_mem
t = symm_mem.empty((1024, 1024), dtype=torch.float32, device=f"cuda:{local}")
hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
print(rank, "ptrs", [hex(p) for p in hdl.buffer_ptrs], "rank", hdl.rank, "world", hdl.world_size, flush=True)
t.fill_(rank + 1)
hdl.barrier()
peer = hdl.get_buffer((rank + 1) % world, (1024, 1024), torch.float32)
print(rank, "peer value", float(peer[0, 0]), "peer device", peer.device, flush=True)
peer[5, 5] = 100 + rank
hdl.barrier()
torch.cuda.synchronize()
print(rank, "mine[5,5]", float(t[5, 5]), flush=True)
dist.destroy_process_group()
