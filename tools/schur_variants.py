#!/usr/bin/env python
"""Experiment driver (GPU): outer iterations and time per step of the default bench workload for several
strength-of-connection measures of the Schur hierarchy (nsb_set_schur_strength) and with / without the ellipse
form of the F polynomial (NSB_SKEW).  One setup, a few time steps per variant."""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("navierstokes-capoferri_cecchettini_untila_b200")
h = float(sys.argv[1]) if len(sys.argv) > 1 else 0.011
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
prob = pkg.Problem.generate("3d-cylinder", h).build(inlet=(0, 0.45, 0.41, 0), expand_a00=False)
dev = pkg.Device(3, 0).load_problem(prob, node_pattern=True)
dev.set_params(0.01, prob.mean_velocity(0.0) * 0.4 / 20)
dev.set_solver(1e-6, 28, 10000, 0.5)
N = dev.info()["n_u"] + dev.info()["n_p"]
variants = [(1, 0.08, 0.5), (0, 0.35, 1.0), (0, 0.25, 1.0), (0, 0.35, 0.5), (0, 0.2, 0.5), (2, 0.35, 0.5), (0, 0.5, 0.5)]
if len(sys.argv) > 3:
    variants = [tuple(float(x) for x in v.split(":")) for v in sys.argv[3].split(",")]
for measure, theta, decay in variants:
    dev.set_schur_strength(int(measure), theta, decay)
    dev.set_solution(np.zeros(N))
    its, ms = [], []
    for s in range(nsteps):
        dev.assemble(0.01 * (s + 1))
        dev.timer_start()
        it, _, _ = dev.solve_time_step()
        ms.append(dev.timer_stop())
        its.append(it)
    print(json.dumps({"measure": measure, "theta": theta, "decay": decay, "levels": dev.info()["schur_levels"], "its": its,
                      "solve_ms": [round(m, 1) for m in ms], "inner": dev.inner_params()}), flush=True)
