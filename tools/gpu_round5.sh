#!/bin/bash
mkdir -p gpurun_out
python tools/cublas_ref.py 256 > gpurun_out/cublas_ref.log 2>&1; cat gpurun_out/cublas_ref.log
timeout 1200 python -m pytest tests -m gpu -q -s 2>&1 | tail -60 > gpurun_out/pytest.log
grep -E "codec rel|end-to-end|eps rel-L2|grad cosine|passed|failed|FAILED" gpurun_out/pytest.log
python tools/ncu_step.py 256 train 2>&1 | tail -1
