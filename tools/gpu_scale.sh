#!/bin/bash
# 8-GPU box: the driver's scaling command at N=8 (shortened side legs), then N=4 for the fixed-global-batch line
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 8 --steps 5 --warmup 3 --sample-steps-multi 40 --latent-steps 10 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "rc=$?"; grep "^{" gpurun_out/bench_n8.json | head -c 500; echo; grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/bench_n8.err | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 4 --steps 5 --warmup 3 --sample-steps 0 --cfg3-steps 2 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "rc=$?"; grep "^{" gpurun_out/bench_n4.json | head -c 400; echo; grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/bench_n4.err | tail -5
