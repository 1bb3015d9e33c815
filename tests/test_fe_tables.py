"""Host logic of the assembly (no GPU): the reference-cell contraction tables the sm_100a assembly kernel works from
(csrc/fe_tables.h; they stand in for the FEValues evaluation of reference src/NavierStokes.cpp:141-146, 177-254)
against exact integrals of the P2/P1 bases in barycentric coordinates -- no quadrature and no table shared with the
product or the oracle.  With the deal.II >= 9.4 rules (degree 5 in 2D and 3D) every table is exact to round-off; with
the deal.II 9.3 rule in 3D (10 points, degree 3) the stiffness and divergence tables are still exact, mass and
convection are under-integrated -- both facts are asserted."""
import numpy as np
import pytest

from test_oracle_kat import _barycentric_p2, _poly_dl, _poly_int, _poly_mul

# deal.II local order of the P2 line dofs -> (i, j) vertex pairs; _barycentric_p2 lists edges as (i, j), i < j, lexicographic
LINES = {2: [(0, 1), (1, 2), (2, 0)], 3: [(0, 1), (1, 2), (2, 0), (0, 3), (1, 3), (2, 3)]}


def _to_dealii_order(dim):
    nb = dim + 1
    mine = [(i, j) for i in range(nb) for j in range(i + 1, nb)]
    return list(range(nb)) + [nb + mine.index(tuple(sorted(e))) for e in LINES[dim]]


def _exact_tables(dim):
    nb = dim + 1
    _, p2, p1, G = _barycentric_p2(dim)
    perm = _to_dealii_order(dim)
    p2 = [p2[i] for i in perm]
    nn = len(p2)

    def grad(p):
        return [[(a * G[k][c], ex) for k in range(nb) for a, ex in _poly_dl(p, k)] for c in range(dim)]

    g = [grad(p) for p in p2]
    m = np.array([[_poly_int(_poly_mul(p2[a], p2[b]), dim) for b in range(nn)] for a in range(nn)])
    k = np.zeros((dim, dim, nn, nn))
    ch = np.zeros((nn, dim, nn, nn))
    dh = np.zeros((nn, nb, dim))
    for a in range(nn):
        for b in range(nn):
            for d in range(dim):
                for e in range(dim):
                    k[d, e, a, b] = _poly_int(_poly_mul(g[a][d], g[b][e]), dim)
            for n in range(nn):
                pan = _poly_mul(p2[a], p2[n])
                for d in range(dim):
                    ch[n, d, a, b] = _poly_int(_poly_mul(pan, g[b][d]), dim)
        for kk in range(nb):
            for d in range(dim):
                dh[a, kk, d] = _poly_int(_poly_mul(g[a][d], p1[kk]), dim)
    return m, k, ch, dh


@pytest.mark.parametrize("dim", [2, 3])
def test_tables_of_the_degree_5_rules_are_exact(pkg, dim):
    T = pkg.device.fe_tables(dim, pkg.device.QUAD_DEALII95)
    m, k, ch, dh = _exact_tables(dim)
    for name, ref in (("mhat", m), ("khat", k), ("chat", ch), ("dhat", dh)):
        err = np.abs(T[name] - ref).max()
        assert err < 2e-15 * max(1.0, np.abs(ref).max()) + 1e-16, (name, err)


def test_tables_of_the_dealii_93_rule_in_3d(pkg):
    """10 points, exact to degree 3: gradients x gradients (degree 2) and gradients x P1 (degree 2) are exact,
    P2 x P2 (degree 4) and P2 x P2 x gradient (degree 5) are not -- the under-integration a deal.II 9.3 build of the
    reference has, which NSB_QUAD_DEALII93 reproduces."""
    T = pkg.device.fe_tables(3, pkg.device.QUAD_DEALII93)
    m, k, ch, dh = _exact_tables(3)
    assert np.abs(T["khat"] - k).max() < 1e-12 and np.abs(T["dhat"] - dh).max() < 1e-12  # 13-digit tables
    assert abs(T["mhat"].sum() - 1.0 / 6.0) < 1e-12       # the total mass is still exact
    assert np.abs(T["mhat"] - m).max() > 1e-4             # single entries are not
    assert np.abs(T["chat"] - ch).max() > 1e-5
