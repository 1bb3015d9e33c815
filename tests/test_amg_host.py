"""Host logic of the Schur hierarchy (no GPU): the greedy aggregation of csrc/amg.cuh (coarsen), which replaces the
ILU(0) factorisation of S in PreconditionASIMPLE::initialize / vmult (reference src/NavierStokes.cpp:958-959,
986-989).  Checked on the Schur complement of a small oracle system: a valid partition, the owner constraint of the
multi-GPU fine level, the strength measures, and agreement with the Python emulation the CPU studies use
(tests/amg_emul.py)."""
import numpy as np
import pytest
import scipy.sparse as sp

from amg_emul import coarsen as coarsen_py


@pytest.fixture(scope="module")
def schur(pkg, oracle_mod):
    prob = pkg.Problem.generate("2d-cylinder", 0.08).build(inlet=(pkg.INLET_PARABOLIC, 0.3, 0.41, 0))
    orc = oracle_mod.Oracle(2, prob.array("xyz"), prob.array("cells"), prob.array("bfaces"), prob.array("bids"))
    orc.set_inlet(0, 0.3, 0.41, 0)
    orc.set_params(0.01, 1e-3)
    orc.assemble(0.01)
    B = orc.scipy_blocks()
    D = B["a00"].diagonal()
    S = (B["a10"] @ sp.diags(1.0 / D) @ B["a01"]).tocsr()
    S.sort_indices()
    return S


@pytest.mark.parametrize("measure,theta", [(0, 0.35), (0, 0.2), (1, 0.08), (2, 0.35)])
def test_aggregates_partition_the_rows_and_match_the_emulation(pkg, schur, measure, theta):
    S = schur
    n = S.shape[0]
    agg, nc, nnzc = pkg.device.amg_coarsen(S.indptr, S.indices, S.data, theta, 8, measure)
    assert agg.min() == 0 and agg.max() == nc - 1 and np.unique(agg).size == nc  # every aggregate is used
    assert nc < 0.7 * n  # it coarsens
    # Galerkin pattern = pattern of P^T S P
    P = sp.csr_matrix((np.ones(n), (np.arange(n), agg.astype(np.int64))), shape=(n, nc))
    G = (P.T @ abs(S) @ P).tocsr()
    assert G.nnz == nnzc
    # strong-coupling rule: a non-root member was put into its aggregate through a strong coupling
    Pe = coarsen_py(S, theta, 8, signed=(measure == 0), rel=(measure != 1))
    agg_py = np.asarray(Pe.argmax(axis=1)).ravel()
    assert Pe.shape[1] == nc and np.array_equal(agg_py, agg.astype(np.int64))


def test_owner_constraint_keeps_aggregates_inside_one_rank(pkg, schur):
    S = schur
    n = S.shape[0]
    owner = (np.arange(n) * 3 // n).astype(np.int32)  # three contiguous owners
    agg, nc, _ = pkg.device.amg_coarsen(S.indptr, S.indices, S.data, 0.08, 8, 1, owner)
    for I in range(nc):
        assert np.unique(owner[agg == I]).size == 1
    # every owner's aggregates form a contiguous range of coarse ids (what the all-gather of the coarse rhs needs)
    first = [agg[owner == r].min() for r in range(3)]
    last = [agg[owner == r].max() for r in range(3)]
    assert first[0] == 0 and all(first[r + 1] == last[r] + 1 for r in range(2)) and last[2] == nc - 1


def test_relative_measure_ignores_positive_couplings(pkg):
    """measure 0 aggregates along negative couplings only (a fifth of the off-diagonal entries of B D^-1 Bt of
    Taylor-Hood are positive)."""
    n = 6
    A = np.eye(n) * 4.0
    for i in range(n - 1):
        A[i, i + 1] = A[i + 1, i] = 1.0 if i % 2 == 0 else -1.0  # +,-,+,-,+
    S = sp.csr_matrix(A)
    agg0, nc0, _ = pkg.device.amg_coarsen(S.indptr, S.indices, S.data, 0.35, 8, 0)
    assert nc0 == 4 and agg0[1] == agg0[2] and agg0[3] == agg0[4] and agg0[0] != agg0[1]
    agg1, nc1, _ = pkg.device.amg_coarsen(S.indptr, S.indices, S.data, 0.08, 8, 1)
    assert nc1 < nc0  # the absolute measure follows the positive couplings as well
