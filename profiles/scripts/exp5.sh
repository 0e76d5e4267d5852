cd /root/repo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 profiles/push_probe.py 50000 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$"
