"""Data-parallel path on real GPUs (skipped unless >= 2 are visible): spawns tests/dp_check.py under torchrun."""
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs >= 2 GPUs')
def test_two_gpu_data_parallel_matches_oracle():
    script = Path(__file__).resolve().parent / 'dp_check.py'
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
                        '127.0.0.1', '--master-port', '29533', str(script)], capture_output=True, text=True, timeout=400)
    assert 'DP_CHECK_PASS' in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    assert 'DP_TRAIN_CHECK_PASS' in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
