#!/usr/bin/env python
"""Experiment driver: per-kernel CUDA-event timings on the default bench workload."""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("navierstokes-capoferri_cecchettini_untila_b200")
h = float(sys.argv[1]) if len(sys.argv) > 1 else 0.011
prob = pkg.Problem.generate("3d-cylinder", h).build(inlet=(0, 0.45, 0.41, 0), expand_a00=False)
dev = pkg.Device(3, 0).load_problem(prob, node_pattern=True)
dev.set_params(0.01, prob.mean_velocity(0.0) * 0.4 / 20)
dev.set_solver(1e-6, 28, 10000, 0.5)
dev.assemble(0.01)
if os.environ.get("NSB_TIME_ONLY"):
    names = os.environ["NSB_TIME_ONLY"].split(",")
    it = -1
else:
    names = None
    it, _, _ = dev.solve_time_step()
out = {"its": it}
for name, which in (("sweep_F", 4), ("block_spmv", 5), ("prec_apply", 2), ("g_apply", 7), ("a10_spmv", 8), ("sweep_S", 6),
                    ("gram_schmidt_k14", 13), ("canonical_spmv", 0)):
    if names is not None and name not in names:
        continue
    try:
        out[name + "_ms"] = round(dev.bench_kernel(which, 30), 4)
    except Exception as e:
        out[name + "_ms"] = str(e)[:40]
print(json.dumps(out))
