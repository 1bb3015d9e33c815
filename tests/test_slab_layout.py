"""Host-side check of the slab (windowed sliced-ELL) storage of F_s (csrc/slab.cuh): the
layout built from a node pattern must reproduce the CSR product for every window capacity,
including degenerate ones (one row per slab, rows cut into chunks, ragged last slab)."""
import numpy as np
import pytest
import scipy.sparse as sp



def _pattern(pkg, name, h):
    prob = pkg.Problem.generate(name, h).build(expand_a00=False)
    return prob.array("nodes.rowptr").copy(), prob.array("nodes.colind").copy()


@pytest.mark.parametrize("name,h,dim", [("2d-cylinder", 0.05, 2), ("3d-cylinder", 0.1, 3), ("3d-cylinder", 0.06, 3)])
@pytest.mark.parametrize("cap", [1408, 300, 90])
def test_slab_product_matches_csr(pkg, name, h, dim, cap):
    rp, ci = _pattern(pkg, name, h)
    n = rp.size - 1
    rng = np.random.default_rng(7)
    val = rng.standard_normal(ci.size)
    x = rng.standard_normal(dim * n)
    y, st = pkg.device.slab_host_check(dim, rp, ci, val, x, window_cap=cap)
    F = sp.csr_matrix((val, ci.astype(np.int64), rp), shape=(n, n))
    ref = (F @ x.reshape(n, dim)).ravel()
    assert np.max(np.abs(y - ref)) <= 1e-12 * np.max(np.abs(ref))
    assert st["nnz"] == ci.size and st["padded"] >= st["nnz"] and st["padded"] % 32 == 0
    assert st["max_window"] <= cap
    if cap == 1408:
        # the layout must stay compact: little padding, a few window nodes per row
        assert st["padded"] <= 1.25 * st["nnz"], st
        assert st["window_total"] <= 6 * n, st


def test_slab_rejects_rows_longer_than_the_window(pkg):
    rp, ci = _pattern(pkg, "3d-cylinder", 0.1)
    n = rp.size - 1
    with pytest.raises(pkg.DeviceError):
        pkg.device.slab_host_check(3, rp, ci, np.ones(ci.size), np.ones(3 * n), window_cap=16)


def test_slab_rectangular_pattern_with_ghost_columns(pkg):
    """Owned rows x (owned + ghost) columns, as on a rank of a multi-GPU run."""
    rp, ci = _pattern(pkg, "3d-cylinder", 0.1)
    n = rp.size - 1
    n_own = n // 2
    rp2, ci2 = rp[: n_own + 1].copy(), ci[: rp[n_own]].copy()
    rng = np.random.default_rng(3)
    val = rng.standard_normal(ci2.size)
    x = rng.standard_normal(3 * n)
    y, _ = pkg.device.slab_host_check(3, rp2, ci2, val, x, n_cols=n)
    F = sp.csr_matrix((val, ci2.astype(np.int64), rp2), shape=(n_own, n))
    ref = (F @ x.reshape(n, 3)).ravel()
    assert np.max(np.abs(y - ref)) <= 1e-12 * np.max(np.abs(ref))


@pytest.mark.parametrize("name,h,dim", [("2d-cylinder", 0.05, 2), ("3d-cylinder", 0.08, 3)])
@pytest.mark.parametrize("cap", [1408, 120])
def test_a01_slab_product_matches_csr(pkg, name, h, dim, cap):
    prob = pkg.Problem.generate(name, h).build(expand_a00=False)
    rp, ci = prob.array("nodes.rowptr"), prob.array("nodes.colind")
    rp01, ci01 = prob.array("a01.rowptr"), prob.array("a01.colind")
    n, n_p = rp.size - 1, prob.sizes()["n_p"]
    rng = np.random.default_rng(11)
    val = rng.standard_normal(ci01.size)
    xp = rng.standard_normal(n_p)
    y, st = pkg.device.gslab_host_check(dim, rp, ci, rp01, ci01, val, xp, window_cap=cap)
    B = sp.csr_matrix((val, ci01.astype(np.int64), rp01), shape=(dim * n, n_p))
    ref = B @ xp
    assert np.max(np.abs(y - ref)) <= 1e-12 * np.max(np.abs(ref))
    assert st["nnz"] == ci01.size and st["padded"] % 32 == 0
    if cap == 1408:
        assert st["padded"] <= 1.25 * st["nnz"], st
