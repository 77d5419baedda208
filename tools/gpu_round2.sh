#!/bin/bash
mkdir -p gpurun_out
for s in 0 1; do TSD_ATTN_BWD_TC_SHARED=$s python tools/attn_bwd_check.py 64,4096,128 2>&1 | tail -6; done > gpurun_out/attn_ab.log 2>&1
cat gpurun_out/attn_ab.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log
python bench.py --no-cpu --sample-steps 50 --latent-steps 0 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"; head -c 600 gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick.err
