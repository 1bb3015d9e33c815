"""Thread scaling of the CPU baseline (oracle/ns_baseline.cpp) on BASELINE config C3: the reference algorithm as
`mpirun -n P` runs it, P = 1 ... all cores, 1 warm-up + 2 timed steps each.  Bench/test infrastructure only.

    python tests/cpu_baseline_scaling.py [P ...]
"""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

pkg = importlib.import_module(bench.PKG)
ps = [int(x) for x in sys.argv[1:]] or sorted({1, 2, 4, 8, os.cpu_count() or 1})
for P in ps:
    r = bench.cpu_arm(pkg, "3d-square", 0.05, 2, 1, P, budget_s=120.0)
    print(json.dumps({"threads": P, **r}), flush=True)
r = bench.cpu_arm(pkg, "3d-square", 0.05, 1, 1, 1, budget_s=120.0, serial=True)
print(json.dumps({"threads": "serial checker", **r}), flush=True)
