#!/bin/bash
mkdir -p gpurun_out
for v in 1 0; do TSD_IN_CONV_TF32=$v python tools/head_bench.py 256; done 2>&1 | tee gpurun_out/head_bench.txt
timeout 400 python -m pytest tests/test_unet_gpu.py tests/test_golden_gpu.py tests/test_training_loop_gpu.py tests/test_edge_cases_gpu.py -m gpu -q -x 2>&1 | tail -4
