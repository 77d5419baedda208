#!/bin/bash
# One GPU-box visit: tests, bench, launch list.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 > gpurun_out/pytest.log
tail -5 gpurun_out/pytest.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
./tools/ex2_bench > gpurun_out/ex2.log 2>&1; cat gpurun_out/ex2.log
