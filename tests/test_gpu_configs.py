"""Time-step parity of the sm_100a path on the reference's other driver configurations
(BASELINE.json configs C2-C4 and the *_03 drivers), through the C ABI,
against the CPU oracle on the same mesh: solution 1e-8 relative, Cd/Cl 1e-6, both solvers run to
1e-12 (the preconditioners differ by mandate, the solutions must not).

Also the per-entry form of the 1e-10 matrix tolerance: every stored entry is compared relative to
max(|reference entry|, 1e-3 * largest entry of its row), i.e. small entries next to the M/dt
diagonal are checked against their own magnitude, not against the block maximum.  A third floor,
1e-5 * largest entry of the block, covers entries that vanish or nearly vanish by cancellation (e.g. the
d/dx coupling of an edge node whose cells are symmetric in x): there both sides hold the rounding noise
of summands of the block's scale, a few 1e-16 of it (measured: 3.6e-16), so the statement that can be
made is absolute: |error| <= 1e-15 * largest entry of the block."""
import math

import numpy as np
import pytest

from conftest import make_configured_case, seeded_state

pytestmark = pytest.mark.gpu

# name -> mesh, h, inlet kind, U_m, Re (None: keep nu), nu, deltat, sin(pi t/8) inlet, steps
CONFIGS = {
    # tests/2D/test_02/src/test_02.cpp:15,24,41,57-58 (Schaefer-Turek 2D-2 parameters)
    "C2-2d-cylinder-Re100": dict(mesh="2d-cylinder", h=0.05, uniform=False, um=1.5, re=100, dt=0.02, sin=False),
    # tests/3D/test_01/src/test_01.cpp:15,57-58 on the square obstacle of mesh/domain3D.geo
    "C3-3d-square-Re20": dict(mesh="3d-square", h=0.1, uniform=False, um=0.45, re=20, dt=0.01, sin=False),
    # tests/3D/test_02/src/test_02.cpp:15,57-58
    "3d-cylinder-Re100": dict(mesh="3d-cylinder", h=0.1, uniform=False, um=2.25, re=100, dt=0.01, sin=False),
    # tests/2D/test_naca/src/test_03.cpp:15,24,41,57 (uniform inflow, default nu = 1e-3, NavierStokes.hpp:254)
    "C4-naca2412": dict(mesh="naca2412", h=0.1, uniform=True, um=1.0, re=None, dt=0.01, sin=False),
    # tests/2D/test_naca/run_test.sh:7-9: NACA 2408 contour, chord 0.4, 10 degrees angle of attack (mesh/test.py)
    "naca2408-aoa10-run_test": dict(mesh="airfoil:2408:0.4:10", h=0.03, uniform=True, um=1.0, re=None, dt=0.01, sin=False),
    # tests/2D/test_03/src/test_03.cpp:24-25,43,59-60: inlet x sin(pi t/8); set_re_number at t = 0 gives nu = 0
    "2d-test_03-sin-inlet-nu0": dict(mesh="2d-cylinder", h=0.05, uniform=False, um=1.5, re=100, dt=0.01, sin=True),
    # tests/3D/test_03/src/test_03.cpp:15,25,43,59-60
    "3d-test_03-sin-inlet-nu0": dict(mesh="3d-cylinder", h=0.1, uniform=False, um=2.25, re=100, dt=0.01, sin=True),
}


def _device(pkg, prob, dim, dt, nu):
    dev = pkg.Device(dim).load_problem(prob)
    dev.set_params(dt, nu)
    return dev


@pytest.mark.parametrize("name", list(CONFIGS))
def test_time_steps_match_oracle_on_driver_configs(pkg, oracle_mod, name):
    cfg = CONFIGS[name]
    prob, orc, dim, nu = make_configured_case(pkg, oracle_mod, **cfg)
    if cfg["sin"]:
        assert nu == 0.0  # quirk B8: get_mean_vel() is 0 at t = 0
    dev = _device(pkg, prob, dim, cfg["dt"], nu)
    orc.set_solver(1e-12, 30, 10000, 1e-10)
    dev.set_solver(gmres_rtol=1e-12, restart=60)
    # the time-dependent inlet enters the device through nsb_scale_dirichlet: the dof set and the spatial
    # profile are fixed (reference :297-324), only the factor sin(pi t/8) moves
    base_dofs, base_vals = None, None
    if cfg["sin"]:
        steady = pkg.Problem.generate(cfg["mesh"], cfg["h"]).build(
            inlet=(pkg.INLET_PARABOLIC, cfg["um"], 0.41, 0))
        base_dofs, base_vals = np.array(steady.array("bc.dofs")), np.array(steady.array("bc.values"))
        dev.set_dirichlet(base_dofs, base_vals)
    t = 0.0
    for step in range(3):
        t += cfg["dt"]
        if cfg["sin"]:
            f = math.sin(math.pi * t / 8.0)
            assert f == prob.inlet_time_factor(t)
            dev.scale_dirichlet(f)
        orc.assemble(t)
        dev.assemble(t)
        if cfg["sin"]:  # the scaled list equals the oracle's interpolated boundary values
            d_o, v_o = orc.bc()
            assert np.array_equal(d_o, base_dofs) and np.allclose(v_o, f * base_vals, rtol=1e-14, atol=0)
        ref, got = orc.rhs(), dev.rhs()
        assert np.max(np.abs(got - ref)) <= 1e-10 * np.max(np.abs(ref)), f"rhs, step {step}"
        rc, it_o, _, _ = orc.solve_time_step()
        it_d, _, _ = dev.solve_time_step()
        assert rc == 0 and it_d > 0
        f_o = orc.compute_forces(t)
        f_d = dev.compute_forces(prob.mean_velocity(t))
        xo, xd = orc.solution(), dev.solution()
        assert np.linalg.norm(xd - xo) / np.linalg.norm(xo) < 1e-8, f"{name} step {step}"
        assert abs(f_d[2] - f_o[2]) < 1e-6 * max(1.0, abs(f_o[2])), (name, step, f_d, f_o)
        assert abs(f_d[3] - f_o[3]) < 1e-6 * max(1.0, abs(f_o[3])), (name, step, f_d, f_o)
        dev.set_solution(xo)  # same state for the next assembly


def _per_entry_worst(rowptr, got, ref, floor=1e-3, block_floor=1e-5):
    """max over entries of |got - ref| / max(|ref|, floor * max|row of ref|, block_floor * max|ref|)."""
    n = rowptr.size - 1
    lens = np.diff(rowptr)
    rowmax = np.maximum.reduceat(np.abs(ref), rowptr[:-1][lens > 0])
    full = np.zeros(n)
    full[lens > 0] = rowmax
    scale = np.maximum(np.abs(ref), floor * np.repeat(full, lens))
    scale = np.where(scale > 0, np.maximum(scale, block_floor * np.max(np.abs(ref))), 0.0)  # cleared rows stay exact
    ok = scale > 0
    assert np.all(got[~ok] == 0.0)  # rows cleared by the boundary conditions are exactly zero on both sides
    return float(np.max(np.abs(got - ref)[ok] / scale[ok])) if ok.any() else 0.0


@pytest.mark.parametrize("key,rule", [("2d-cylinder", 1), ("3d-square", 0), ("3d-cylinder", 1), ("naca2412", 1)])
def test_entries_match_oracle_per_entry(pkg, oracle_mod, key, rule):
    from conftest import make_case
    prob, orc, dim, nu, um = make_case(pkg, oracle_mod, key, quad_rule=rule)
    dev = pkg.Device(dim).load_problem(prob, quad_rule=rule)
    dev.set_params(0.01, nu)
    x = seeded_state(orc)
    orc.set_solution(x)
    dev.set_solution(x)
    orc.assemble(0.01)
    dev.assemble(0.01)
    worst = {}
    for blk, nm in ((pkg.device.A00, "a00"), (pkg.device.A01, "a01"), (pkg.device.A10, "a10")):
        rp, _ = orc.pattern(nm)
        worst[nm] = _per_entry_worst(rp, dev.values(blk), orc.values(nm))
    print(f"{key} rule {rule}: worst per-entry ratio {worst}")
    assert max(worst.values()) < 1e-10, worst


def test_inner_solves_adapt_to_the_convective_regime(pkg, oracle_mod):
    """Round-2 fix, kept under test at the reference's own tolerance: on the NACA 2408 / 10 degrees mesh of
    run_test.sh the spectrum of D^-1 F leaves the real axis as soon as the flow has started.  The device must
    (a) measure an imaginary extent of order one there (and none on the first, symmetric step), and (b) converge
    in a number of outer iterations comparable to the ILU-preconditioned oracle (the first GPU run of this case
    did not converge at all; with the round-1 aggregation measure it took 590-810 iterations)."""
    cfg = CONFIGS["naca2408-aoa10-run_test"]
    prob, orc, dim, nu = make_configured_case(pkg, oracle_mod, **cfg)
    dev = _device(pkg, prob, dim, cfg["dt"], nu)
    orc.set_solver(1e-6, 30, 10000, 1e-2)
    dev.set_solver(gmres_rtol=1e-6, restart=28, max_it=2000)
    t, its_d, its_o, imag = 0.0, [], [], []
    for step in range(3):
        t += cfg["dt"]
        orc.assemble(t)
        dev.assemble(t)
        its_o.append(orc.solve_time_step()[1])
        its_d.append(dev.solve_time_step()[0])
        imag.append(dev.inner_params()["imag_F"])
    assert imag[0] < 0.2 and min(imag[1:]) > 1.0, imag
    assert max(its_d) <= 2 * max(its_o), (its_d, its_o)


@pytest.mark.parametrize("measure,theta,decay", [(0, 0.35, 1.0), (1, 0.08, 0.5), (2, 0.35, 0.5)])
def test_schur_strength_measures_change_the_work_not_the_solution(pkg, oracle_mod, measure, theta, decay):
    """nsb_set_schur_strength selects how the aggregates of the Schur hierarchy are formed; every choice is a valid
    preconditioner, so a tight solve ends at the same solution (2d-cylinder, C2 parameters)."""
    cfg = CONFIGS["C2-2d-cylinder-Re100"]
    prob, orc, dim, nu = make_configured_case(pkg, oracle_mod, **cfg)
    orc.set_solver(1e-12, 30, 10000, 1e-10)
    dev = _device(pkg, prob, dim, cfg["dt"], nu)
    dev.set_schur_strength(measure, theta, decay)
    dev.set_solver(gmres_rtol=1e-12, restart=60)
    t = 0.0
    for step in range(2):
        t += cfg["dt"]
        orc.assemble(t)
        dev.assemble(t)
        rc, _, _, _ = orc.solve_time_step()
        dev.solve_time_step()
        assert rc == 0
    xo, xd = orc.solution(), dev.solution()
    assert np.linalg.norm(xd - xo) <= 1e-8 * np.linalg.norm(xo)
    assert dev.info()["schur_levels"] >= 2
