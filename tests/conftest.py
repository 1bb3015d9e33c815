import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG_NAME = "navierstokes-capoferri_cecchettini_untila_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 via gpurun)")


@pytest.fixture(scope="session")
def pkg():
    """The product package (host + device bindings); builds the host library and
    the oracle when missing (no GPU needed for either)."""
    import subprocess
    subprocess.check_call(["make", "-C", ROOT, "host", "oracle"], stdout=subprocess.DEVNULL)
    return importlib.import_module(PKG_NAME)


@pytest.fixture(scope="session")
def oracle_mod(pkg):
    from oracle import ns_oracle
    return ns_oracle


# (name, h, inlet U_m) of the small cases the oracle finishes in seconds
CASES = {
    "2d-cylinder": ("2d-cylinder", 0.05, 0.3),
    "3d-square": ("3d-square", 0.1, 0.45),
    "3d-cylinder": ("3d-cylinder", 0.1, 0.45),
    "naca2412": ("naca2412", 0.1, 1.0),
}


def make_case(pkg, oracle_mod, key, quad_rule=1, h=None):
    """Host problem + oracle on the same mesh, Re = 20 parameters of the drivers
    (tests/2D/test_01/src/test_01.cpp:57-58, tests/3D/test_01/src/test_01.cpp:57-58)."""
    name, h0, um = CASES[key]
    kind = pkg.INLET_UNIFORM if key.startswith("naca") else pkg.INLET_PARABOLIC
    prob = pkg.Problem.generate(name, h or h0).build(inlet=(kind, um, 0.41, 0))
    dim = prob.sizes()["dim"]
    orc = oracle_mod.Oracle(dim, prob.array("xyz"), prob.array("cells"), prob.array("bfaces"), prob.array("bids"),
                            quad_rule)
    orc.set_inlet(kind, um, 0.41, 0)
    orc.set_params(0.01, 1e-3)
    nu = orc.set_re_number(20) if not key.startswith("naca") else 1e-3
    orc.set_threads(min(8, os.cpu_count() or 1))
    return prob, orc, dim, nu, um


def make_configured_case(pkg, oracle_mod, mesh, h, uniform, um, re, dt, sin, nu=1e-3, quad_rule=1):
    """Host problem + oracle with the parameters of one reference driver:
    inlet profile, U_m, optional set_re_number(re) (evaluated at t = 0 like the drivers do), deltat and
    the sin(pi t/8) inlet factor of the *_03 drivers.  Returns (prob, orc, dim, nu)."""
    kind = pkg.INLET_UNIFORM if uniform else pkg.INLET_PARABOLIC
    if mesh.startswith("airfoil:"):  # "airfoil:<naca4>:<chord>:<angle of attack>": run_test.sh / mesh/test.py pre-processing
        _, naca4, chord, aoa = mesh.split(":")
        prob = pkg.Problem.generate_airfoil(h, naca4=int(naca4), chord=float(chord), aoa_deg=float(aoa))
    else:
        prob = pkg.Problem.generate(mesh, h)
    prob.build(inlet=(kind, um, 0.41, 1 if sin else 0))
    dim = prob.sizes()["dim"]
    orc = oracle_mod.Oracle(dim, prob.array("xyz"), prob.array("cells"), prob.array("bfaces"), prob.array("bids"),
                            quad_rule)
    orc.set_inlet(kind, um, 0.41, 1 if sin else 0)
    orc.set_params(dt, nu)
    if re is not None:
        nu = orc.set_re_number(re)
    orc.set_threads(min(8, os.cpu_count() or 1))
    return prob, orc, dim, nu


def seeded_state(orc, seed=1234):
    """A smooth-ish seeded velocity/pressure state so that the convective term is
    non-zero."""
    rng = np.random.default_rng(seed)
    x = np.zeros(orc.N)
    x[: orc.n_u] = 0.3 * rng.standard_normal(orc.n_u)
    x[orc.n_u:] = 0.1 * rng.standard_normal(orc.n_p)
    return x
