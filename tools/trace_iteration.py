#!/usr/bin/env python
"""Per-kernel timeline of the last outer GMRES iteration of a (multi-GPU) run, from the %globaltimer stamps the
device library records after every kernel when NSB_TRACE is set (csrc/nsb_capi.cu) -- the substitute for an ncu
launch list where ncu cannot be used (several ranks).  The stamps cost ~2 us per kernel: read SHARES, not absolutes.

    NSB_TRACE=gpurun_out/trace_ python -m torch.distributed.run --nproc-per-node N ... tools/trace_iteration.py [h]
    python tools/trace_iteration.py --summarise gpurun_out/trace_0.csv
"""
import collections
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def summarise(path):
    rows = sorted(((int(r["end_ns"]), r["kernel"]) for r in csv.DictReader(open(path))))
    # the last outer iteration: from the last block product (fs_slab_apply_kernel<...0>) to the end of the trace,
    # minus the tail after the solve (forces, final halo)
    ap = [i for i, (_, k) in enumerate(rows) if k in ("(fs_slab_apply_kernel<3, 0>)", "(fs_slab_apply_kernel<2, 0>)")]
    if len(ap) < 3:
        print("trace too short")
        return
    # the stamps inside the captured preconditioner graph keep the times of its LAST replay: take the last
    # iteration of the last solve, from its block product to the scaling of the new basis vector
    i0 = ap[-1]
    i1 = next(i for i in range(i0 + 1, len(rows)) if rows[i][1] == "scale_kernel" and rows[i][0] - rows[i0][0] > 300000)
    span = rows[i1][0] - rows[i0][0]
    agg = collections.OrderedDict()
    prev = rows[i0][0]
    for t, k in rows[i0 + 1:i1 + 1]:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += (t - prev) * 1e-3
        prev = t
    print(f"{path}: one outer iteration = {span * 1e-3:.1f} us, {i1 - i0} kernels")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {us:9.1f} us  {100 * us * 1e3 / span:5.1f} %  x{n:<3d} {k}")


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--summarise":
        for p in sys.argv[2:]:
            summarise(p)
        return
    import bench
    import importlib
    pkg = importlib.import_module(bench.PKG)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")
    sys.argv = [sys.argv[0]] + (["--h", sys.argv[1]] if len(sys.argv) > 1 else [])
    a = bench.parse()
    run = bench.NativeRun(pkg, a, a.mesh, a.h, dist, rank, world, int(os.environ.get("LOCAL_RANK", "0")))
    for _ in range(2):
        run.step(False)
    print(f"rank {rank}: {run.iters[-1]} iterations, solve {run.tsol[-1]:.1f} ms", flush=True)
    run.barrier()
    run.close()   # nsb_destroy writes the trace
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
