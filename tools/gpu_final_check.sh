#!/bin/bash
# Round-end check of HEAD on one GPU: the whole -m gpu suite, smoke, the replayed training step (bench.py optional: arg "bench").
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/pytest_final.log; cat gpurun_out/pytest_final.log
python __graft_entry__.py smoke 2>&1 | tail -1
python tools/ncu_step.py 256 train 2>&1 | tail -1 | tee gpurun_out/step_plain.log
if [ "$1" = "bench" ]; then
  timeout 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; head -c 330 gpurun_out/bench_final.json; echo
fi
