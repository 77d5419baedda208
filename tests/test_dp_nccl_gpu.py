"""Data-parallel path on real GPUs over NCCL (world size 2): N-rank result == 1-rank result on the same global batch."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_ranks_match_one(cuda):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under `gpurun --gpus 2`)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tests", "helpers", "dp_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    if r.returncode != 0 or "DP_NCCL_OK" not in r.stdout:
        # the worker's own traceback sits far above torchrun's summary: keep the whole log and show the relevant lines
        out_dir = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, "dp_worker_failure.log"), "w") as f:
            f.write(r.stdout + "\n==== stderr ====\n" + r.stderr)
        keep = [ln for ln in r.stderr.splitlines() if any(k in ln for k in ("Error", "assert", "dp_worker.py", "rank"))]
        pytest.fail("dp_worker failed:\n" + "\n".join(keep[-40:]))
