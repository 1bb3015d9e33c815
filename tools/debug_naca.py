#!/usr/bin/env python
"""Experiment driver (GPU): iteration counts of the NACA 2408 / 10 degrees case (tests/2D/test_naca/run_test.sh) with
the inner-F-polynomial parameters in effect, at the reference tolerance and at the parity tolerance."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("navierstokes-capoferri_cecchettini_untila_b200")

cfg = dict(mesh="airfoil:2408:0.4:10", h=0.03, uniform=True, um=1.0, re=None, dt=0.01, sin=False)
for tol, restart in ((1e-6, 28), (1e-12, 60)):
    prob = pkg.Problem.generate_airfoil(cfg["h"], naca4=2408, chord=0.4, aoa_deg=10.0)
    prob.build(inlet=(pkg.INLET_UNIFORM, cfg["um"], 0.41, 0))
    dim, nu = 2, 1e-3
    dev = pkg.Device(dim).load_problem(prob)
    dev.set_params(cfg["dt"], nu)
    dev.set_solver(gmres_rtol=tol, restart=restart, max_it=3000)
    t = 0.0
    for step in range(4):
        t += cfg["dt"]
        dev.assemble(t)
        try:
            it, _, _ = dev.solve_time_step()
        except Exception as e:
            it = str(e)[:60]
        print(f"tol {tol} step {step}: its {it} {dev.inner_params()}", flush=True)
