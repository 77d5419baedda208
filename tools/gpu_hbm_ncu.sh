#!/bin/bash
# ncu --set full of the HBM-bound kernels (one launch each at the benchmark shapes) + their CUDA-event rates on the same box.
mkdir -p gpurun_out
python tools/ncu_hbm_kernels.py 256 > gpurun_out/hbm_rows.json 2> gpurun_out/hbm_plain.err; echo "plain rc=$?"; tail -c 300 gpurun_out/hbm_plain.err
python tools/norm_bench.py 256 > gpurun_out/r02_norm_bench.txt 2>&1; cat gpurun_out/r02_norm_bench.txt
timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:"gn_apply|gn_bwd|ln_fwd|ln_bwd|adamw_clip|tail_conv_y|in_conv_tf32" -f -o gpurun_out/r02_hbm \
  python tools/ncu_hbm_kernels.py 256 > gpurun_out/ncu_hbm.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_hbm.log; ls -la gpurun_out/r02_hbm.ncu-rep
