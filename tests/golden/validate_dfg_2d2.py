#!/usr/bin/env python
"""Offline validation of the oracle against the unsteady Schaefer-Turek / DFG 2D-2 benchmark (Re = 100):
Strouhal number of the vortex shedding, 0.295 <= St <= 0.305 in the literature.  Too slow for the test
suite (~10-20 min on 8 cores); run by hand, the result is recorded in tests/golden/dfg_2d2.json.

    python tests/golden/validate_dfg_2d2.py [h] [dt] [T]
"""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("navierstokes-capoferri_cecchettini_untila_b200")
from oracle.ns_oracle import Oracle  # noqa: E402

h = float(sys.argv[1]) if len(sys.argv) > 1 else 0.025
dt = float(sys.argv[2]) if len(sys.argv) > 2 else 0.005
T = float(sys.argv[3]) if len(sys.argv) > 3 else 8.0
prob = pkg.Problem.generate("2d-cylinder", h).build(inlet=(pkg.INLET_PARABOLIC, 1.5, 0.41, 0))
orc = Oracle(2, prob.array("xyz"), prob.array("cells"), prob.array("bfaces"), prob.array("bids"))
orc.set_inlet(0, 1.5, 0.41, 0)   # U_max = 1.5 -> U_mean = 1.0 (tests/2D/test_02)
orc.set_params(dt, 1e-3)         # true Re = U_mean * 0.1 / nu = 100
orc.set_threads(os.cpu_count() or 1)
t, t0, lift = 0.0, time.time(), []
n = int(round(T / dt))
for k in range(n):
    t += dt
    orc.assemble(t)
    rc, it, _, _ = orc.solve_time_step()
    assert rc == 0
    lift.append(orc.compute_forces(t)[1])
    if k % 100 == 99:
        print(f"t={t:.2f} its={it} lift={lift[-1]:+.5f} wall={time.time() - t0:.0f}s", flush=True)
lift = np.array(lift)
# frequency from the upward zero crossings of the (mean-free) lift signal over the last third of the run
tail = lift[2 * n // 3:] - lift[2 * n // 3:].mean()
tt = dt * (np.arange(tail.size) + 2 * n // 3 + 1)
up = np.where((tail[:-1] < 0) & (tail[1:] >= 0))[0]
cross = tt[up] + dt * (-tail[up]) / (tail[up + 1] - tail[up])
freq = (len(cross) - 1) / (cross[-1] - cross[0]) if len(cross) > 2 else float("nan")
St = freq * 0.1 / 1.0
out = {"h": h, "dt": dt, "T": T, "n_dofs": int(orc.N), "strouhal": float(St), "periods_used": int(len(cross) - 1),
       "lift_amplitude_reference_formula": float(0.5 * (tail.max() - tail.min())), "literature": [0.295, 0.305]}
print(json.dumps(out))
# append the run to the committed record
path = os.path.join(ROOT, "tests", "golden", "dfg_2d2.json")
rec = json.load(open(path)) if os.path.exists(path) else {"runs": []}
out.pop("literature", None)
rec.setdefault("runs", []).append(out)
with open(path, "w") as f:
    json.dump(rec, f, indent=1)
