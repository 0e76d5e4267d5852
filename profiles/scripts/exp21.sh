cd /root/repo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_b21_g8.json 2> gpurun_out/r2_b21_g8.err
tail -3 gpurun_out/r2_b21_g8.err
