#!/usr/bin/env python
"""Headline benchmark: time per time step of the 3D Re=20 cylinder case
(BASELINE.json metric), with assembly DoF/s, SpMV GB/s and the HBM roofline of
the dominant kernel.

    python bench.py --gpus N --steps K --warmup W            # this repo (sm_100a path)
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference

One "step" = NavierStokes::assemble + solve_time_step + compute_forces
(reference src/NavierStokes.cpp:483-486) on the `3d-cylinder` mesh
(mesh/domain3D2.geo geometry, tests/3D/test_01 parameters).  Prints ONE JSON
line on rank 0.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "navierstokes-capoferri_cecchettini_untila_b200"

U_M, H_CH, DT, RE = 0.45, 0.41, 0.01, 20  # tests/3D/test_01/src/test_01.cpp:15-16, 57-58


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--mesh", default="3d-cylinder")
    ap.add_argument("--h", type=float, default=float(os.environ.get("NSB_BENCH_H", "0.011")),
                    help="target edge length of the mesh (0.011 ~ 10M DoFs)")
    ap.add_argument("--cpu-h", type=float, default=0.05, help="mesh of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-canonical-spmv", action="store_true",
                    help="skip y = A x on the materialised canonical (reference) block CSR (12 B x 9 nnz(F_s) + ... "
                         "= 11 GB at the default size)")
    ap.add_argument("--alpha", type=float, default=0.5)
    ap.add_argument("--schur", type=str, default="1,0,0,0,0", help="mode,nu,theta,omega,cycles of the Schur solver")
    ap.add_argument("--sweeps", type=str, default="0,0,0,0",
                    help="kF,ratioF,kS,ratioS of the inner sweeps (0 = automatic)")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clock / throttle sampling during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                parts = [x.strip() for x in line.split(",")]
                if len(parts) < 6:
                    continue
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                for n, v in zip(names, parts[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons)}
        return out


def spmv_canonical_bytes(info):
    """Algorithmic bytes of y = A x on the canonical (reference) block CSR
    (DESIGN.md §5): 12 B per stored non-zero, 8 B row offset per row, y written
    and x read once."""
    n_u, n_p = info["n_u"], info["n_p"]
    nnz = info["nnz_a00"] + info["nnz_a01"] + info["nnz_a10"]
    return 12 * nnz + 8 * (2 * n_u + n_p + 3) + 16 * (n_u + n_p)


def fs_slab_bytes(info, dim=3):
    """F_s in slab storage: 10 B per stored non-zero (8 B value + 16-bit window index;
    the ~3 % ELL padding is NOT counted), 4 B per window node, 4 B chunk-position word per row,
    8 B slice offset per 32 virtual rows (~1.3 per row)."""
    nnz = info["nnz_a00"] // (dim * dim)
    n_nodes = info["n_u"] // dim
    return 10 * nnz + 4 * info["slab_window_total"] + 4 * n_nodes + 8 * 8 * info["slab_count"]


def spmv_bytes(info, dim=3):
    """y = A x on the storage the solver uses: F_s in slab form, A01 in node-block slab form (8 B per
    value + 2 B index per node-level entry + pressure windows + 2 B permutation per node), A10 as
    CSR; x read once, y written once.  ELL padding is not counted."""
    n_u, n_p = info["n_u"], info["n_p"]
    g = 8 * info["nnz_a01"] + 2 * (info["nnz_a01"] // dim) + 4 * info["gslab_window_total"] + 2 * (n_u // dim) \
        + 8 * 8 * info["slab_count"]
    return fs_slab_bytes(info, dim) + g + 12 * info["nnz_a10"] + 8 * (n_p + 1) + 16 * (n_u + n_p)


def sweep_bytes(info, dim=3):
    """One Chebyshev-Jacobi sweep on F (three-term form): F_s stream + z read once (window staging;
    re-reads hit L2) + zold, Dinv.*b read, znew written, Dinv read per node."""
    n_u = info["n_u"]
    return fs_slab_bytes(info, dim) + 8 * n_u * 4 + 8 * (n_u // dim)


def sweep_s_bytes(info):
    n_p = info["n_p"]
    return 12 * info["nnz_s"] + 8 * (n_p + 1) + 8 * n_p * 7


def assembly_bytes(info, dim=3):
    """Compulsory traffic of the assembly: every stored value
    written once, dof ids + slots + vertex coordinates per cell, velocity read, rhs written."""
    nn, nv = (10, 4) if dim == 3 else (6, 3)
    per_cell = 4 * (nn + nv) + 4 * nv + 2 * (nn * nn + 2 * nn * nv)
    nnz = info["nnz_a00"] // (dim * dim) + info["nnz_a01"] + info["nnz_a10"]
    return 8 * nnz + info["n_cells"] * per_cell + 8 * info["n_u"] + 8 * (info["n_u"] + info["n_p"])


# --------------------------------------------------------------------------
def cpu_reference_run(pkg, mesh, h, steps, warmup, threads):
    """The CPU restatement of the reference algorithm (oracle) on a bounded
    sample mesh of the same geometry; returns (ms/step, n_dofs, iters)."""
    from oracle.ns_oracle import Oracle
    prob = pkg.Problem.generate(mesh, h).build(inlet=(pkg.INLET_PARABOLIC, U_M, H_CH, 0))
    dim = prob.sizes()["dim"]
    orc = Oracle(dim, prob.array("xyz"), prob.array("cells"), prob.array("bfaces"), prob.array("bids"))
    orc.set_inlet(0, U_M, H_CH, 0)
    orc.set_params(DT, 1e-3)
    orc.set_re_number(RE)
    orc.set_threads(threads)
    t, times, iters = 0.0, [], []
    for s in range(warmup + steps):
        t += DT
        t0 = time.perf_counter()
        orc.assemble(t)
        rc, it, _, _ = orc.solve_time_step()
        orc.compute_forces(t)
        if s >= warmup:
            times.append(time.perf_counter() - t0)
            iters.append(it)
    return 1e3 * float(np.mean(times)), orc.N, float(np.mean(iters))


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    pkg = importlib.import_module(PKG)
    hbm_peak, peak_src = load_peaks()
    kF, rF, kS, rS = a.sweeps.split(",")
    config = {"workload": f"{a.mesh} Re=20 (mesh/domain3D2.geo geometry, tests/3D/test_01 parameters), h={a.h}",
              "mesh": a.mesh, "h": a.h, "deltat": DT, "Re": RE, "quadrature": "dealii95 (14-pt)",
              "gmres": "left-preconditioned GMRES(28), rtol 1e-6 (reference stopping rule)",
              "preconditioner": f"aSIMPLE alpha={a.alpha}, Chebyshev-Jacobi sweeps F:{kF} S:{kS} (0 = automatic)",
              "l2_policy": "inputs larger than L2 (matrix >> 126 MB)", "parallelism": f"dd{a.gpus}"}

    if a.impl == "reference":
        if rank != 0:
            return 0
        threads = os.cpu_count() or 1
        ms, n_s, it = cpu_reference_run(pkg, a.mesh, a.cpu_h, max(1, min(a.steps, 2)), min(a.warmup, 1), threads)
        prob = pkg.Problem.generate(a.mesh, a.h)
        prob.build(inlet=(0, U_M, H_CH, 0), expand_a00=False)
        n_full = prob.sizes()["n_u"] + prob.sizes()["n_p"]
        val = ms * n_full / n_s
        sample = (f"oracle (CPU restatement of the reference algorithm: naive (q,i,j) assembly, GMRES(28)+aSIMPLE+"
                  f"ILU(0)+inner GMRES) on {a.mesh} h={a.cpu_h} ({n_s} DoFs, {it:.0f} GMRES its/step, "
                  f"{ms:.0f} ms/step measured), scaled linearly in DoFs to the {n_full}-DoF workload; "
                  f"assembly uses {threads} threads, the solve is serial like one Trilinos rank")
        line = {"impl": "reference", "metric": "time_per_step", "value": val, "unit": "ms/step", "n_gpus": a.gpus,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": val, "higher_is_better": False,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": "ms/step", "cores": threads, "kind": "port", "sample": sample},
                "e2e": {"value": val, "unit": "ms/step", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ---------------- native arm ----------------
    dist = None
    if world > 1:  # one process per GPU (torchrun); gloo carries the NCCL id, barriers and the max over ranks
        import torch
        import torch.distributed as dist
        dist.init_process_group("gloo")
    t_setup = time.perf_counter()
    prob = pkg.Problem.generate(a.mesh, a.h)
    prob.build(inlet=(pkg.INLET_PARABOLIC, U_M, H_CH, 0), expand_a00=False)
    sz = prob.sizes()
    dim = sz["dim"]
    loc = None
    if world > 1:
        prob.partition(world)  # same deterministic RCB partition on every rank
        loc = pkg.LocalProblem(prob, world, rank)
        ids = [pkg.Device.make_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        dev = pkg.Device(dim, local_rank).load_local_problem(prob, loc, ids[0])
    else:
        dev = pkg.Device(dim, local_rank).load_problem(prob, node_pattern=True)
    nu = prob.mean_velocity(0.0) * 0.4 / RE  # set_re_number, reference :332-341
    dev.set_params(DT, nu)
    dev.set_solver(1e-6, 28, 10000, a.alpha)
    dev.set_inner(int(kF), float(rF) if int(kF) > 0 else 0.0, int(kS), float(rS) if int(kS) > 0 else 0.0)
    sm = a.schur.split(",")
    dev.set_schur_solver(int(sm[0]), int(sm[1]), float(sm[2]), float(sm[3]), int(sm[4]))
    info = dev.info()
    N = sz["n_u"] + sz["n_p"]      # global unknowns
    N_loc = dev.N                  # this rank's vector (owned + ghost velocity, replicated pressure)
    t_setup = time.perf_counter() - t_setup

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    import ctypes as C
    L = dev.L
    if loc is None:
        bc_dofs = np.array(prob.array("bc.dofs"))
        bc_vals = prob.array("bc.values")
    else:
        bn = loc.array("bc_nodes")
        bc_dofs = (dim * bn[:, None] + np.arange(dim, dtype=np.uint32)[None, :]).astype(np.uint32).ravel()
        bc_vals = loc.array("bc_values")
    pin_bc = L.nsb_alloc_pinned(8 * max(bc_dofs.size, 1))
    pin_sol = L.nsb_alloc_pinned(8 * N_loc)
    bc_host = np.frombuffer((C.c_char * (8 * bc_dofs.size)).from_address(pin_bc), dtype=np.float64)
    sol_host = np.frombuffer((C.c_char * (8 * N_loc)).from_address(pin_sol), dtype=np.float64)
    bc_host[:] = bc_vals

    state = {"t": 0.0}
    iters, tasm, tprec, tsol = [], [], [], []

    def step(e2e):
        state["t"] += DT
        if e2e:  # the facade's per-step host traffic: Dirichlet values in, solution + forces out
            dev.set_dirichlet(bc_dofs, bc_host)
        dev.assemble(state["t"])
        it, tp, ts = dev.solve_time_step()
        f = dev.compute_forces(prob.mean_velocity(state["t"]))
        if e2e:
            dev.solution(out=sol_host)
        tm = dev.timers()
        iters.append(it)
        tasm.append(tm[0])
        tprec.append(tm[1])
        tsol.append(tm[2])
        return f

    for _ in range(a.warmup):
        step(False)
    del iters[:], tasm[:], tprec[:], tsol[:]
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = dev.launch_count()
    barrier()
    t0 = time.perf_counter()
    dev.timer_start()          # CUDA events on the stream all the kernels of this context run on
    for _ in range(a.steps):
        forces = step(False)   # every API call ends with a stream synchronise
    ms_events = dev.timer_stop() / a.steps
    barrier()
    ms_wall = max_over_ranks(1e3 * (time.perf_counter() - t0) / a.steps)
    ms_dev = max_over_ranks(ms_events)
    launches = dev.launch_count() - l0
    dev_stats = dict(iters=float(np.mean(iters)), asm=float(np.mean(tasm)), prec=float(np.mean(tprec)),
                     sol=float(np.mean(tsol)))
    # end to end through the C ABI with host buffers
    barrier()
    dev.timer_start()
    for _ in range(a.steps):
        forces = step(True)
    ms_e2e = max_over_ranks(dev.timer_stop() / a.steps)
    barrier()
    # kernel micro-benchmarks on the resident system (CUDA events on the ctx stream)
    reps = 20
    ms_spmv = dev.bench_kernel(5, reps)
    ms_sweep = dev.bench_kernel(4, reps)
    ms_sweep_s = dev.bench_kernel(6, reps)
    ms_asm = dev.bench_kernel(1, 5)
    ms_prec = dev.bench_kernel(2, 5)
    ms_schur = dev.bench_kernel(3, 5)
    ms_spmv_can = dev.bench_kernel(0, 10) if (not a.no_canonical_spmv and world == 1) else None
    clocks = sampler.stop()

    gb = 1e-9
    spmv_gbs = spmv_bytes(info, dim) * gb / (ms_spmv * 1e-3)
    sweep_gbs = sweep_bytes(info, dim) * gb / (ms_sweep * 1e-3)
    sweep_s_gbs = sweep_s_bytes(info) * gb / (ms_sweep_s * 1e-3)
    asm_gbs = assembly_bytes(info, dim) * gb / (ms_asm * 1e-3)
    # share of one outer GMRES iteration: (kF-1) F sweeps, the fine-level S sweeps, one block product
    info2 = dev.info()
    kF_eff, kS_eff = info2["sweeps_F"], info2["sweeps_S"] + (1 if info2["schur_mode"] == 1 else 0)
    shares = {"fs_slab_sweep_kernel (Jacobi-type sweep on F, slab storage)": ((kF_eff - 1) * ms_sweep, sweep_gbs,
                                                                               sweep_bytes(info, dim), ms_sweep),
              "cheb_sweep_kernel (Jacobi-type sweep on S)": (max(kS_eff - 1, 1) * ms_sweep_s, sweep_s_gbs,
                                                             sweep_s_bytes(info), ms_sweep_s),
              "fs_slab_apply_kernel + spmv_kernel (block product y = A x)": (ms_spmv, spmv_gbs, spmv_bytes(info, dim),
                                                                       ms_spmv)}
    top = max(shares, key=lambda k: shares[k][0])
    traffic = None
    try:  # measured DRAM bytes per launch of that kernel at this mesh size, when a capture exists (profiles/)
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        traffic = tj.get(top.split(" ")[0], {}).get(str(a.h)) if world == 1 else None
    except Exception:
        pass
    roof = {"bound": "hbm", "kernel": top, "achieved": shares[top][1], "peak": hbm_peak, "unit": "GB/s",
            "frac": shares[top][1] / hbm_peak, "traffic": traffic, "peak_source": peak_src,
            "ms_per_launch": shares[top][3], "algorithmic_bytes_per_launch": shares[top][2],
            "ms_per_gmres_iteration_by_kernel": {k: v[0] for k, v in shares.items()}}
    line = {"metric": "time_per_step", "value": ms_dev, "unit": "ms/step", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_dev, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "wall_ms_per_step": ms_wall, "n_dofs": N, "n_cells": info["n_cells"], "gmres_iters_per_step": dev_stats["iters"],
            "phase_ms": {"assemble": dev_stats["asm"], "prec_init": dev_stats["prec"], "solve": dev_stats["sol"]},
            "assembly_dofs_per_s": N / (ms_asm * 1e-3), "assembly_gbs": asm_gbs, "assembly_ms": ms_asm,
            "spmv_gbs": spmv_gbs, "spmv_frac_of_hbm": spmv_gbs / hbm_peak, "spmv_ms": ms_spmv,
            "sweep_F_gbs": sweep_gbs, "sweep_F_ms": ms_sweep, "sweep_S_gbs": sweep_s_gbs, "sweep_S_ms": ms_sweep_s,
            "schur_ms": ms_schur, "sweeps_F": info2["sweeps_F"], "schur_levels": info2["schur_levels"],
            "spmv_canonical_gbs": (spmv_canonical_bytes(info) * gb / (ms_spmv_can * 1e-3)) if ms_spmv_can else None,
            "spmv_canonical_ms": ms_spmv_can,
            "spmv_canonical_frac_of_hbm": (spmv_canonical_bytes(info) * gb / (ms_spmv_can * 1e-3) / hbm_peak) if ms_spmv_can else None,
            "prec_apply_ms": ms_prec, "cd": float(forces[2]), "cl": float(forces[3]),
            "setup_s": t_setup, "device_bytes": info["device_bytes"],
            "roofline": roof, "clocks": clocks,
            "e2e": {"value": ms_e2e, "unit": "ms/step", "h2d_bytes_per_step": int(8 * bc_dofs.size + 4 * bc_dofs.size),
                    "d2h_bytes_per_step": int(8 * N_loc + 32)},
            "gpu_launches": int(launches)}
    if world > 1:
        line["n_dofs_local"] = N_loc
        line["partition"] = "recursive coordinate bisection of the cells; velocity rows distributed, pressure replicated"
    if rank != 0:
        L.nsb_free_pinned(pin_bc)
        L.nsb_free_pinned(pin_sol)
        barrier()
        dev.close()
        dist.destroy_process_group()
        return 0
    if not a.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        ms, n_s, it = cpu_reference_run(pkg, a.mesh, a.cpu_h, 1, 1, threads)
        line["cpu_baseline"] = {
            "value": ms * N / n_s, "unit": "ms/step", "cores": threads, "kind": "port",
            "sample": (f"oracle on {a.mesh} h={a.cpu_h} ({n_s} DoFs): {ms:.0f} ms/step measured ({it:.0f} GMRES its), "
                       f"scaled linearly in DoFs to {N}; assembly on {threads} threads, solve serial")}
    L.nsb_free_pinned(pin_bc)
    L.nsb_free_pinned(pin_sol)
    print(json.dumps(line), flush=True)
    if dist is not None:
        barrier()
        dev.close()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
