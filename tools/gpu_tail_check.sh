#!/bin/bash
# tail-conv kernels: TSD_TAIL_Y=2 (default, two 128-thread CTAs per SM) / 1 (one 256-thread CTA, two buffers) / 0 (tap-by-tap)
mkdir -p gpurun_out
for v in 2 1 0; do TSD_TAIL_Y=$v python tools/tail_bench.py 256; done 2>&1 | tee gpurun_out/tail_bench.txt
timeout 300 python -m pytest tests/test_unet_gpu.py tests/test_golden_gpu.py tests/test_sampling_loop_gpu.py tests/test_edge_cases_gpu.py -m gpu -q -x 2>&1 | tail -4
