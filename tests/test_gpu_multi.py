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


def test_facade_driver_is_rank_count_invariant(pkg, tmp_path):
    """The C++ driver under a torchrun-style launcher on 2 GPUs writes the same
    forces_vs_time.csv (to solver tolerance) as on 1 GPU."""
    if pkg.device_lib().nsb_device_count() < 2:
        pytest.skip("needs two GPUs")
    import numpy as np
    from conftest import PKG_NAME
    pkgdir = os.path.join(ROOT, PKG_NAME)
    subprocess.check_call(["make", "-C", ROOT, "drivers"], stdout=subprocess.DEVNULL)
    rows = {}
    for n in (1, 2):
        base = tmp_path / f"run{n}"
        for d in ("build", "output", "cache", "mesh"):
            os.makedirs(base / d)
        subprocess.check_call([os.path.join(pkgdir, "drivers", "make_mesh"), "3d-cylinder", "0.1",
                               str(base / "mesh" / "domain.msh")])
        exe = os.path.join(pkgdir, "drivers", "d3_test_01")
        if n == 1:
            cmd = [exe, "../mesh/domain.msh", "0.03"]
        else:
            cmd = [sys.executable, "-m", "torch.distributed.run", "--no-python", "--nnodes=1", "--nproc-per-node=2",
                   "--master-addr", "127.0.0.1", "--master-port", "29547", exe, "../mesh/domain.msh", "0.03"]
        r = subprocess.run(cmd, cwd=str(base / "build"), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        lines = open(base / "build" / "forces_vs_time.csv").read().strip().split("\n")[1:]
        rows[n] = np.array([[float(x) for x in l.split(",")] for l in lines])
    assert rows[1].shape == rows[2].shape == (3, 9)
    assert np.allclose(rows[1][:, 7], rows[2][:, 7], rtol=2e-4), (rows[1][:, 7], rows[2][:, 7])
