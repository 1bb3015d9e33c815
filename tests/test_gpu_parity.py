"""Parity of the sm_100a path (through the C ABI, include/nsb.h) against the
CPU oracle on the same seeded inputs.  Tolerances (BASELINE.json north star):
matrix / rhs entries 1e-10 relative to the block's largest entry, pattern and
numbering bit-exact, Cd/Cl 1e-6."""
import numpy as np
import pytest

from conftest import CASES, make_case, seeded_state

pytestmark = pytest.mark.gpu

ENTRY_TOL = 1e-10


def _device(pkg, prob, dim, nu, quad_rule=1, node_pattern=False):
    dev = pkg.Device(dim).load_problem(prob, quad_rule=quad_rule, node_pattern=node_pattern)
    dev.set_params(0.01, nu)
    return dev


def _rel(a, b):
    scale = max(np.max(np.abs(b)), 1e-300)
    return np.max(np.abs(a - b)) / scale


@pytest.mark.parametrize("key,rule", [("2d-cylinder", 1), ("2d-cylinder", 0), ("3d-square", 1), ("3d-square", 0),
                                      ("3d-cylinder", 1), ("naca2412", 1)])
def test_assembled_system_matches_oracle(pkg, oracle_mod, key, rule):
    prob, orc, dim, nu, um = make_case(pkg, oracle_mod, key, quad_rule=rule)
    dev = _device(pkg, prob, dim, nu, quad_rule=rule)
    x = seeded_state(orc)
    orc.set_solution(x)
    dev.set_solution(x)
    orc.assemble(0.01)
    dev.assemble(0.01)
    for blk, name in ((pkg.device.A00, "a00"), (pkg.device.A01, "a01"), (pkg.device.A10, "a10")):
        rp_o, ci_o = orc.pattern(name)
        rp_d, ci_d = dev.pattern(blk)
        assert np.array_equal(rp_o, rp_d) and np.array_equal(ci_o, ci_d), f"{name} pattern differs"
        assert _rel(dev.values(blk), orc.values(name)) < ENTRY_TOL, name
    assert _rel(dev.rhs(), orc.rhs()) < ENTRY_TOL
    # Dirichlet values land in the solution vector (apply_boundary_values)
    assert np.array_equal(dev.solution(), orc.solution())


def test_node_pattern_expansion_is_canonical(pkg, oracle_mod):
    prob, orc, dim, nu, um = make_case(pkg, oracle_mod, "3d-cylinder")
    dev = _device(pkg, prob, dim, nu, node_pattern=True)
    rp_o, ci_o = orc.pattern("a00")
    rp_d, ci_d = dev.pattern(pkg.device.A00)
    assert np.array_equal(rp_o, rp_d) and np.array_equal(ci_o, ci_d)


@pytest.mark.parametrize("mode", [0, 1])
def test_bc_diag_modes(pkg, oracle_mod, mode):
    prob, orc, dim, nu, um = make_case(pkg, oracle_mod, "2d-cylinder")
    dev = _device(pkg, prob, dim, nu)
    orc.set_bc_diag_mode(mode)
    dev.set_bc_diag_mode(mode)
    x = seeded_state(orc)
    orc.set_solution(x)
    dev.set_solution(x)
    orc.assemble(0.01)
    dev.assemble(0.01)
    assert _rel(dev.values(pkg.device.A00), orc.values("a00")) < ENTRY_TOL
    assert _rel(dev.rhs(), orc.rhs()) < ENTRY_TOL


@pytest.mark.parametrize("key", ["2d-cylinder", "3d-cylinder"])
def test_block_spmv_matches_oracle(pkg, oracle_mod, key):
    prob, orc, dim, nu, um = make_case(pkg, oracle_mod, key)
    dev = _device(pkg, prob, dim, nu)
    x = seeded_state(orc)
    orc.set_solution(x)
    dev.set_solution(x)
    orc.assemble(0.01)
    dev.assemble(0.01)
    v = np.sin(np.arange(orc.N, dtype=np.float64))
    assert _rel(dev.vmult(v), orc.vmult(v)) < 1e-12


@pytest.mark.parametrize("key", ["2d-cylinder", "3d-square"])
def test_schur_complement_matches_oracle(pkg, oracle_mod, key):
    """S = B diag(1/diag F) Bt, reference NavierStokes.cpp:948-956."""
    prob, orc, dim, nu, um = make_case(pkg, oracle_mod, key)
    dev = _device(pkg, prob, dim, nu)
    x = seeded_state(orc)
    for o in (orc, dev):
        o.set_solution(x)
        o.assemble(0.01)
    orc.set_solver(1e-6, 30, 10000, 1e-2)
    orc.solve_time_step()
    dev.solve_time_step()
    rp_o, ci_o = orc.pattern("s")
    rp_d, ci_d = dev.pattern(pkg.device.S)
    assert np.array_equal(rp_o, rp_d) and np.array_equal(ci_o, ci_d)
    assert _rel(dev.values(pkg.device.S), orc.values("s")) < ENTRY_TOL


@pytest.mark.parametrize("key", ["2d-cylinder", "3d-cylinder"])
def test_time_steps_match_oracle_at_tight_tolerance(pkg, oracle_mod, key):
    """Both solvers run to 1e-12 so that the (different) preconditioners do not
    show in the result: solution 1e-8 relative, Cd/Cl 1e-6.  The
    oracle's inner solves are tightened to 1e-10 as well: with the reference's
    1e-2 its preconditioner is a non-linear operator inside non-flexible GMRES
     and the recurrence residual is not the true one."""
    prob, orc, dim, nu, um = make_case(pkg, oracle_mod, key)
    dev = _device(pkg, prob, dim, nu)
    orc.set_solver(1e-12, 30, 10000, 1e-10)
    dev.set_solver(gmres_rtol=1e-12, restart=60)
    t = 0.0
    for step in range(3):
        t += 0.01
        orc.assemble(t)
        dev.assemble(t)
        rc, it_o, _, _ = orc.solve_time_step()
        it_d, _, _ = dev.solve_time_step()
        assert rc == 0 and it_d > 0
        f_o = orc.compute_forces(t)
        f_d = dev.compute_forces(prob.mean_velocity(t))
        xo, xd = orc.solution(), dev.solution()
        assert np.linalg.norm(xd - xo) / np.linalg.norm(xo) < 1e-8, f"step {step}"
        assert abs(f_d[2] - f_o[2]) < 1e-6 * max(1.0, abs(f_o[2])), (f_d, f_o)
        assert abs(f_d[3] - f_o[3]) < 1e-6 * max(1.0, abs(f_o[3])), (f_d, f_o)
        # keep both trajectories on the same state for entry-level parity of the next assembly
        dev.set_solution(xo)


def test_forces_match_oracle_on_seeded_state(pkg, oracle_mod):
    for key, rule in (("2d-cylinder", 0), ("2d-cylinder", 1), ("3d-square", 0), ("3d-cylinder", 1)):
        prob, orc, dim, nu, um = make_case(pkg, oracle_mod, key, quad_rule=rule)
        dev = _device(pkg, prob, dim, nu, quad_rule=rule)
        x = seeded_state(orc)
        orc.set_solution(x)
        dev.set_solution(x)
        orc.assemble(0.01)   # sets the inlet time used by get_mean_vel
        dev.assemble(0.01)
        orc.set_solution(x)
        dev.set_solution(x)
        f_o = orc.compute_forces(0.01)
        f_d = dev.compute_forces(prob.mean_velocity(0.01))
        assert np.allclose(f_d, f_o, rtol=1e-11, atol=1e-13), (key, rule, f_d, f_o)


def test_default_tolerance_step_converges_like_the_reference(pkg, oracle_mod):
    """Reference stopping rule (1e-6 ||rhs||, restart 28): the true residual of
    the GPU solution is small and the iteration count is of the oracle's order."""
    prob, orc, dim, nu, um = make_case(pkg, oracle_mod, "3d-cylinder")
    dev = _device(pkg, prob, dim, nu)
    t = 0.0
    for _ in range(2):
        t += 0.01
        orc.assemble(t)
        dev.assemble(t)
        rc, it_o, _, _ = orc.solve_time_step()
        it_d, _, _ = dev.solve_time_step()
        x = dev.solution()
        r = dev.rhs() - dev.vmult(x)
        assert np.linalg.norm(r) / np.linalg.norm(dev.rhs()) < 1e-4
        assert 0 < it_d < 20 * max(it_o, 5)
        assert np.linalg.norm(x - orc.solution()) / np.linalg.norm(x) < 1e-3


def test_errors_are_reported_not_thrown(pkg):
    dev = pkg.Device(2)
    with pytest.raises(pkg.DeviceError):
        dev.assemble(0.0)  # setup incomplete
    bad = np.arange(15, dtype=np.uint32)
    prob = pkg.Problem.generate("channel2d", 0.2).build()
    import ctypes as C
    xyz, cells = prob.array("xyz"), prob.array("cells")
    L = dev.L
    assert L.nsb_set_mesh(dev.h, prob.sizes()["n_verts"], xyz.ctypes.data_as(C.POINTER(C.c_double)),
                          prob.sizes()["n_cells"], cells.ctypes.data_as(C.POINTER(C.c_uint32))) == 0
    scr = np.array(prob.array("cell_dofs"))
    scr[0], scr[1] = scr[1], scr[0]  # break the dim*node+c structure
    rc = L.nsb_set_dofs(dev.h, prob.sizes()["n_u"], prob.sizes()["n_p"], scr.ctypes.data_as(C.POINTER(C.c_uint32)))
    assert rc == -3


@pytest.mark.parametrize("key", ["2d-cylinder", "3d-cylinder"])
def test_lumped_mass_inverse_matches_oracle(pkg, oracle_mod, key):
    """deltat_lumped_mass_inv (reference :232-236, 252, 284-290), velocity block."""
    prob, orc, dim, nu, um = make_case(pkg, oracle_mod, key)
    dev = _device(pkg, prob, dim, nu)
    orc.assemble(0.01)
    dev.assemble(0.01)
    ref = orc.lumped()[: orc.n_u]
    got = dev.lumped_mass_inv(orc.n_u)
    assert np.all(np.isfinite(ref)) and _rel(got, ref) < 1e-12


@pytest.mark.parametrize("key", ["2d-cylinder", "3d-cylinder"])
def test_yosida_preconditioner(pkg, oracle_mod, key):
    """PreconditionAYosida (reference :998-1051): S = B (deltat M_l^-1) Bt entry-wise against the
    product of the oracle's blocks, and the solution of the preconditioned solve against the oracle's
    (aSIMPLE) solution with both run to 1e-12 (a preconditioner does not change the solution)."""
    import scipy.sparse as sp
    prob, orc, dim, nu, um = make_case(pkg, oracle_mod, key)
    dev = _device(pkg, prob, dim, nu)
    orc.set_solver(1e-12, 30, 10000, 1e-10)
    dev.set_solver(gmres_rtol=1e-12, restart=60, preconditioner=pkg.device.PREC_AYOSIDA)
    x = seeded_state(orc)
    orc.set_solution(x)
    dev.set_solution(x)
    orc.assemble(0.01)
    dev.assemble(0.01)
    rc, it_o, _, _ = orc.solve_time_step()
    it_d, _, _ = dev.solve_time_step()
    assert rc == 0 and 0 < it_d < 400
    xo, xd = orc.solution(), dev.solution()
    assert np.linalg.norm(xd - xo) / np.linalg.norm(xo) < 1e-8
    n_u, n_p = orc.n_u, orc.n_p
    rp01, ci01 = orc.pattern("a01")
    rp10, ci10 = orc.pattern("a10")
    Bt = sp.csr_matrix((orc.values("a01"), ci01.astype(np.int64), rp01), shape=(n_u, n_p))
    B = sp.csr_matrix((orc.values("a10"), ci10.astype(np.int64), rp10), shape=(n_p, n_u))
    S_ref = (B @ sp.diags(orc.lumped()[:n_u]) @ Bt).tocsr()
    rp_s, ci_s = dev.pattern(pkg.device.S)
    S_dev = sp.csr_matrix((dev.values(pkg.device.S), ci_s.astype(np.int64), rp_s), shape=(n_p, n_p))
    assert abs(S_dev - S_ref).max() < ENTRY_TOL * abs(S_ref).max()
