#!/usr/bin/env python
"""Experiment driver (not part of the product), kept as the record of how the slab sweep kernel was tuned in
round 1: it timed kernel variants (load batch, resident CTAs, prefetch, cp.async window fill, window capacity,
entry scheduling) selected through NSB_SLAB_VARIANT / NSB_SLAB_SCHED / NSB_SLAB_WINDOW on one resident problem,
one process per configuration.  Those switches were compiled out once the configuration was frozen
(csrc/slab.cuh: kSlabBatch, kSlabMinBlocks); results: profiles/r1_slab_kernels.md.  With the current library every
configuration times the same (adopted) kernel.  usage: sweep_variants.py [h]"""Experiment driver (not part of the product): times kernel variants of the slab sweep on one
resident problem.  Each configuration runs in its own process because the variant / layout
switches are read once per process.  usage: sweep_variants.py [h]"""Experiment driver (not part of the product), kept as the record of how the slab sweep kernel was tuned in
round 1: it timed kernel variants (load batch, resident CTAs, prefetch, cp.async window fill, window capacity,
entry scheduling) selected through NSB_SLAB_VARIANT / NSB_SLAB_SCHED / NSB_SLAB_WINDOW on one resident problem,
one process per configuration.  Those switches were compiled out once the configuration was frozen
(csrc/slab.cuh: kSlabBatch, kSlabMinBlocks); results: profiles/r1_slab_kernels.md.  With the current library every
configuration times the same (adopted) kernel.  usage: sweep_variants.py [h]"""
import importlib, json, os, subprocess, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def child(h):
    pkg = importlib.import_module("navierstokes-capoferri_cecchettini_untila_b200")
    prob = pkg.Problem.generate("3d-cylinder", h).build(inlet=(0, 0.45, 0.41, 0), expand_a00=False)
    dev = pkg.Device(3, 0).load_problem(prob, node_pattern=True)
    dev.set_params(0.01, prob.mean_velocity(0.0) * 0.4 / 20)
    dev.set_solver(1e-6, 28, 10000, 0.5)
    dev.assemble(0.01)
    if os.environ.get("NSB_SWEEP_ONLY"):
        out = {"sweep_ms": dev.bench_kernel(4, 30)}
    else:
        it, _, _ = dev.solve_time_step()
        out = {"its": it, "sweep_ms": dev.bench_kernel(4, 30), "spmv_ms": dev.bench_kernel(5, 30),
               "prec_ms": dev.bench_kernel(2, 10)}
    print("RESULT " + json.dumps(out), flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "child":
        child(float(sys.argv[2]))
        sys.exit(0)
    h = sys.argv[1] if len(sys.argv) > 1 else "0.011"
    combos = []
    for variant in os.environ.get("NSB_VARIANTS", "0").split(","):
        combos.append({"NSB_SLAB_VARIANT": variant})
    for extra in combos:
        env = dict(os.environ, **extra)
        t0 = time.time()
        p = subprocess.run([sys.executable, __file__, "child", h], env=env, capture_output=True, text=True, timeout=300)
        res = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
        print(extra, res[0] if res else ("FAILED rc=%d %s" % (p.returncode, p.stderr[-300:])), "wall %.1fs" % (time.time() - t0),
              flush=True)
