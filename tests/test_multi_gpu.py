"""Multi-GPU parity as a test (was tools/check_multi_gpu.py): two ranks under torchrun + NCCL on
one box.  Skipped when fewer than two CUDA devices are visible (the round-end 1-GPU box); run with
`gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu -q` -- a clean run is committed under profiles/.

What the worker (tests/tools/multi_gpu_worker.py) asserts, every rank against its own single-process references:
  * local-Dice / averaged gradients == mean of the shard gradients, global-Dice / summed == ONE process on the
    concatenated batch, SyncBN variant incl. running statistics;
  * sliding-window inference sharded over the ranks == one process: label maps BIT-IDENTICAL;
  * the Trainer's data-parallel path (equal shard lengths with an odd case count, broadcast split / parameters, reduced
    validation loss): identical parameters on all ranks after an epoch.
"""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two CUDA devices")
def test_two_rank_training_and_inference_parity():
    port = 29500 + os.getpid() % 400
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "tools", "multi_gpu_worker.py")]
    res = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    print(res.stdout[-6000:])
    assert res.returncode == 0, res.stdout[-3000:]
    assert "MISMATCH" not in res.stdout and res.stdout.strip().splitlines()[-1].strip() == "OK"
