"""Real runs of the multi-rank path: one GPU per rank over NCCL where the box has the GPUs, and -- on ANY box with a
GPU, the driver's single-GPU test box included -- two ranks sharing the GPU(s) with gloo collectives (same kernels, same
shards, same exchange + merge logic)."""
import os
import subprocess
import sys

import pytest
import torch

from helpers import ROOT

pytestmark = pytest.mark.gpu


def _run(world, backend, port):
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
           '--master-addr', '127.0.0.1', '--master-port', str(port), os.path.join(ROOT, 'tests', 'multi_gpu_check.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, GR_CHECK_BACKEND=backend))
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count('multi-gpu check ok') == 2, out.stdout[-2000:]


def test_two_ranks_sharing_the_gpu_match_single_rank():
    _run(2, 'gloo', 29519)


def test_sharded_path_matches_single_gpu():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs at least 2 GPUs')
    _run(2 if n < 4 else 4, 'nccl', 29517)
