#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 tests/helpers/dp_worker.py > gpurun_out/dp_worker.log 2>&1; echo "rc=$?"; grep -v "^W1018\|OMP_NUM\|^\*\*\*" gpurun_out/dp_worker.log | grep -B2 -A12 "Traceback" | head -60; tail -3 gpurun_out/dp_worker.log
