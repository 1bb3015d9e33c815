#!/usr/bin/env python
"""Driver for `ncu --profile-from-start off`: sets up the default bench workload, takes one time step, then runs each
solver kernel ONCE between cuProfilerStart / cuProfilerStop (through nsb_bench_kernel), so that a --set full capture
holds exactly the kernels of one outer GMRES iteration on the resident 9.7 M-DoF system.

    ncu --set full --clock-control none --import-source on --profile-from-start off -o out python tools/profile_kernels.py
"""
import ctypes
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("navierstokes-capoferri_cecchettini_untila_b200")
h = float(sys.argv[1]) if len(sys.argv) > 1 else 0.011
prob = pkg.Problem.generate("3d-cylinder", h).build(inlet=(0, 0.45, 0.41, 0), expand_a00=False)
dev = pkg.Device(3, 0).load_problem(prob, node_pattern=True)
dev.set_params(0.01, prob.mean_velocity(0.0) * 0.4 / 20)
dev.set_solver(1e-6, 28, 10000, 0.5)
for s in range(2):
    dev.assemble(0.01 * (s + 1))
    it, _, _ = dev.solve_time_step()
dev.assemble(0.03)
cu = ctypes.CDLL("libcuda.so.1")
assert cu.cuProfilerStart() == 0
# 4: sweep on F, 5: block product, 7: Di .* (Bt d1), 8: src1 - B vec0, 6: sweep on S, 13: Gram-Schmidt against 14 vectors,
# 1: assembly, 3: Schur outer products
out = {w: dev.bench_kernel(w, 1) for w in (4, 5, 7, 8, 6, 13, 1, 3)}
assert cu.cuProfilerStop() == 0
print(it, out)
