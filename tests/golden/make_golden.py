"""Writes tests/golden/oracle_golden.json: sizes, checksums and force values of
the oracle on small seeded cases.  The reference ships no golden vectors and
cannot be built here, so these pin the oracle against itself
(drift guard) -- they do NOT pin it against deal.II/Trilinos.

    python tests/golden/make_golden.py
"""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import make_case  # noqa: E402

pkg = importlib.import_module("navierstokes-capoferri_cecchettini_untila_b200")
from oracle import ns_oracle  # noqa: E402

cases = []
for key, rule in (("2d-cylinder", 0), ("2d-cylinder", 1), ("3d-square", 0), ("3d-cylinder", 1), ("naca2412", 1)):
    prob, orc, dim, nu, um = make_case(pkg, ns_oracle, key, quad_rule=rule)
    x = np.zeros(orc.N)
    x[: orc.n_u] = 0.1 * np.sin(np.arange(orc.n_u))
    orc.set_solution(x)
    orc.assemble(0.01)
    cases.append({"key": key, "rule": rule, "expect": {
        "n_u": orc.n_u, "n_p": orc.n_p, "nnz_a00": orc.sizes()["nnz_a00"],
        "sum_a00": float(orc.values("a00").sum()), "abs_a01": float(np.abs(orc.values("a01")).sum()),
        "rhs_norm": float(np.linalg.norm(orc.rhs())), "forces": [float(v) for v in orc.compute_forces(0.01)]}})
with open(os.path.join(ROOT, "tests", "golden", "oracle_golden.json"), "w") as f:
    json.dump({"generator": "tests/golden/make_golden.py", "cases": cases}, f, indent=1)
print("wrote", len(cases), "cases")
