#!/bin/bash
# 2-GPU box: NCCL data-parallel parity test + bench at N=2 (graph with captured NCCL, then the two-graph fallback)
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_dp_nccl_gpu.py -m gpu -q -s 2>&1 | tail -30 > gpurun_out/pytest_dp.log; tail -15 gpurun_out/pytest_dp.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 --sample-steps-multi 30 --latent-steps 10 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?"; head -c 700 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
TSD_DP_OVERLAP=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 5 --warmup 3 --sample-steps 0 --cfg3-steps 0 > gpurun_out/bench_n2_nooverlap.json 2> gpurun_out/bench_n2_nooverlap.err; echo "rc=$?"; head -c 400 gpurun_out/bench_n2_nooverlap.json; tail -3 gpurun_out/bench_n2_nooverlap.err
