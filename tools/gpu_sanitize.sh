#!/bin/bash
# compute-sanitizer over one small pass of every kernel family (tools/sanitize_step.py).  One GPU.
# NOTE: on this project's GPU pool compute-sanitizer is closed (the wrapper answers rc 86 without running anything); the
# plain pass below still exercises every entry point at batch 2, and the script is what to run where the tool is open.
mkdir -p gpurun_out
python tools/sanitize_step.py > gpurun_out/sanitize_plain.log 2>&1; echo "plain rc=$?"; tail -3 gpurun_out/sanitize_plain.log
for tool in memcheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 3 --print-limit 30 python tools/sanitize_step.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|SANITIZE_PASS_DONE|Invalid|Error|hazard" gpurun_out/sanitize_$tool.log | head -12
done
