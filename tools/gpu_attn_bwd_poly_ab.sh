#!/bin/bash
# A/B: every n-th pair of exponentials of attn_bwd_tc_kernel on the FMA pipe (TSD_ATTN_BWD_TC_POLY = 0 / 8 / 4 / 2).  One GPU.
mkdir -p gpurun_out
for v in 0 8 4 2 0 4; do
  echo "== TSD_ATTN_BWD_TC_POLY=$v"; TSD_ATTN_BWD_TC_POLY=$v python tools/attn_bwd_check.py 64,4096,128 2>&1 | tail -5
done 2>&1 | tee gpurun_out/attn_bwd_poly_ab.txt
