"""Domain-decomposed run on 2 GPUs (NCCL halo exchange / all-reduce) against the
single-rank oracle.  Needs two devices: skipped on a one-GPU box."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("key", ["3d-cylinder", "2d-cylinder"])
def test_two_gpus_match_oracle(pkg, key):
    if pkg.device_lib().nsb_device_count() < 2:
        pytest.skip("needs two GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541",
                        os.path.join(ROOT, "tests", "mgpu_worker.py"), key],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(" ok") == 2
