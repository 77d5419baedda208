#!/bin/bash
# Final-build evidence: launch lists of the replayed training / sampling steps, ncu --set full of the dominant kernels,
# and the attention-backward PIPE late-free A/B.  One GPU.
mkdir -p gpurun_out
for v in "" "TSD_ATTN_BWD_TC_PIPE=1" "TSD_ATTN_BWD_TC_PIPE=1 TSD_ATTN_BWD_TC_DBG=32"; do
  echo "== $v"; env $v python tools/attn_bwd_check.py 64,4096,128 2>&1 | tail -5
done > gpurun_out/attn_bwd_pipe_ab.txt 2>&1
cat gpurun_out/attn_bwd_pipe_ab.txt
python tools/ncu_step.py 256 train 2>&1 | tail -1 | tee gpurun_out/step_plain.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_train.csv python tools/ncu_step.py 256 train > gpurun_out/ncu_train.log 2>&1; tail -1 gpurun_out/ncu_train.log
TSD_NCU_SAMPLE_STEPS=40 python tools/ncu_step.py 256 sample 2>&1 | tail -1 | tee gpurun_out/sample_plain.log
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_sample.csv python tools/ncu_step.py 256 sample > gpurun_out/ncu_sample.log 2>&1; tail -1 gpurun_out/ncu_sample.log
python tools/ncu_kernels.py 256 64 > /dev/null 2>&1 && timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|attn_fwd_tc_kernel|attn_bwd_tc" -f -o gpurun_out/r02_kernels python tools/ncu_kernels.py 256 64 > gpurun_out/ncu_kernels.log 2>&1; tail -2 gpurun_out/ncu_kernels.log; ls -la gpurun_out/*.ncu-rep
