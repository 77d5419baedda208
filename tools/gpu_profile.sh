#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s 2>&1 | tail -60 > gpurun_out/pytest.log
grep -E "codec rel|end-to-end|eps rel-L2|grad cosine|passed|failed|FAILED" gpurun_out/pytest.log
timeout 900 python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; echo "bench rc=$?"; head -c 300 gpurun_out/bench_r02.json; echo; tail -2 gpurun_out/bench_r02.err
python tools/ncu_step.py 256 train > gpurun_out/step_plain.log 2>&1; cat gpurun_out/step_plain.log | tail -1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_train.csv python tools/ncu_step.py 256 train > gpurun_out/ncu_train.log 2>&1; tail -1 gpurun_out/ncu_train.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_sample.csv python tools/ncu_step.py 256 sample > gpurun_out/ncu_sample.log 2>&1; tail -1 gpurun_out/ncu_sample.log
python tools/ncu_kernels.py 256 64 > /dev/null 2>&1 && timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|attn_fwd_tc_kernel|attn_bwd_tc_kernel" -f -o gpurun_out/r02_kernels python tools/ncu_kernels.py 256 64 > gpurun_out/ncu_kernels.log 2>&1; tail -2 gpurun_out/ncu_kernels.log; ls -la gpurun_out/*.ncu-rep
