#!/bin/bash
# 2-GPU box: N = 2 == N = 1 test, then the driver's command at N = 2 (shortened side legs)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 400 python -m pytest tests/test_dp_nccl_gpu.py -m gpu -q 2>&1 | tail -3 | tee gpurun_out/pytest_dp.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 2 --steps 5 --warmup 3 --sample-steps-multi 40 --latent-steps 10 --cfg3-steps 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?"; grep "^{" gpurun_out/bench_n2.json | head -c 600; echo; grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/bench_n2.err | tail -5
