#!/bin/bash
mkdir -p gpurun_out
for s in 0 1; do TSD_ATTN_BWD_TC_PIPE=$s timeout 300 python tools/attn_bwd_check.py 64,4096,128 2>&1 | tail -6; done > gpurun_out/attn_ab.log 2>&1
cat gpurun_out/attn_ab.log
timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 > gpurun_out/pytest.log
grep -E "codec rel|end-to-end|eps rel-L2|grad cosine|passed|failed" gpurun_out/pytest.log
timeout 600 python bench.py --no-cpu --sample-steps 50 --latent-steps 0 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"; head -c 400 gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick.err
