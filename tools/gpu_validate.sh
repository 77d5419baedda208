#!/bin/bash
# Round-end validation on one GPU: the whole -m gpu suite, smoke, the default bench line, launch lists of the replayed steps.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest.log; cat gpurun_out/pytest.log
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_r02.err; head -c 400 gpurun_out/bench_r02.json; echo
python tools/ncu_step.py 256 train 2>&1 | tail -1 | tee gpurun_out/step_plain.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_train.csv python tools/ncu_step.py 256 train > gpurun_out/ncu_train.log 2>&1; tail -1 gpurun_out/ncu_train.log
TSD_NCU_SAMPLE_STEPS=40 python tools/ncu_step.py 256 sample 2>&1 | tail -1 | tee gpurun_out/sample_plain.log
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_sample.csv python tools/ncu_step.py 256 sample > gpurun_out/ncu_sample.log 2>&1; tail -1 gpurun_out/ncu_sample.log
TSD_PROFILE=1 python tools/shape_profile.py 256 train > gpurun_out/r02_shape_profile_train_b256.txt 2>&1; head -3 gpurun_out/r02_shape_profile_train_b256.txt
