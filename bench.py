#!/usr/bin/env python
"""Headline benchmark: time per time step of the 3D Re=20 cylinder case
(BASELINE.json metric), with assembly DoF/s, SpMV GB/s and the HBM roofline of
the dominant kernel.

    python bench.py --gpus N --steps K --warmup W            # this repo (sm_100a path)
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference, all host cores

The CPU arm cannot run the 9.67 M-DoF headline workload in minutes (about 10 minutes per step and >20 GB for
the canonical block matrix and its ILU factors), so it runs BASELINE.json's config C3 (3d-square, h = 0.05,
tests/3D/test_01 parameters) for the W + K steps it is asked for and prints THAT config, the steps it really
timed and the measured time; nothing is extrapolated into `value`.  The native arm times the same C3 mesh as
well (`c3` in its line), so one measured same-config pair exists.

One "step" = NavierStokes::assemble + solve_time_step + compute_forces
(reference src/NavierStokes.cpp:483-486) on the `3d-cylinder` mesh
(mesh/domain3D2.geo geometry, tests/3D/test_01 parameters).  Prints ONE JSON
line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
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
    ap.add_argument("--cpu-mesh", default="3d-square", help="mesh of the CPU arm / the paired C3 record")
    ap.add_argument("--cpu-h", type=float, default=0.05, help="edge length of that mesh (0.05 = BASELINE config C3)")
    ap.add_argument("--cpu-budget-s", type=float, default=300.0, help="wall-clock cap of the CPU arm's timed loop")
    ap.add_argument("--no-c3", action="store_true", help="skip the paired C3 record of the native arm")
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
def cpu_arm(pkg, mesh, h, steps, warmup, threads, budget_s=1e9, serial=False):
    """The CPU restatement of the reference algorithm (oracle/: naive (q,i,j) assembly, GMRES(28) + aSIMPLE with
    ILU(0)-preconditioned inner GMRES to 1e-2) on `mesh` at edge length `h`, tests/3D/test_01 parameters.
    serial=False: the way `mpirun -n threads` runs it (oracle/ns_baseline.cpp: one subdomain per thread,
    rank-local ILU(0), threaded SpMV / dots); serial=True: the single-thread checker itself.
    Times `steps` steps after `warmup` (stops early when the timed loop exceeds budget_s) and returns the
    measured figures of what really ran."""
    from oracle.ns_oracle import Oracle
    prob = pkg.Problem.generate(mesh, h).build(inlet=(pkg.INLET_PARABOLIC, U_M, H_CH, 0))
    dim = prob.sizes()["dim"]
    orc = Oracle(dim, prob.array("xyz"), prob.array("cells"), prob.array("bfaces"), prob.array("bids"))
    orc.set_inlet(0, U_M, H_CH, 0)
    orc.set_params(DT, 1e-3)
    orc.set_re_number(RE)
    if serial:
        orc.set_threads(1)
        asm, solve = orc.assemble, orc.solve_time_step
    else:
        part = np.array(prob.partition(threads)) if threads > 1 else np.zeros(orc.n_cells, np.int32)
        orc.baseline_partition(threads, part)
        asm, solve = orc.baseline_assemble, orc.baseline_solve_time_step
    t, times, iters, phases, f = 0.0, [], [], [], None
    t_loop = None
    for s in range(warmup + steps):
        if s == warmup:
            t_loop = time.perf_counter()
        t += DT
        t0 = time.perf_counter()
        asm(t)
        t1 = time.perf_counter()
        rc, it, tp, ts = solve()
        f = orc.compute_forces(t)
        t2 = time.perf_counter()
        if rc != 0:
            raise RuntimeError("CPU arm: GMRES did not converge")
        if s >= warmup:
            times.append(t2 - t0)
            iters.append(it)
            phases.append((t1 - t0, tp, ts))
            if time.perf_counter() - t_loop > budget_s:
                break
    ph = np.mean(np.array(phases), axis=0)
    return {"ms_per_step": 1e3 * float(np.mean(times)), "steps_timed": len(times), "warmup_run": warmup,
            "n_dofs": int(orc.N), "n_cells": int(orc.n_cells), "gmres_iters_per_step": float(np.mean(iters)),
            "phase_ms": {"assemble": 1e3 * float(ph[0]), "prec_init": 1e3 * float(ph[1]), "solve": 1e3 * float(ph[2])},
            "cd": float(f[2]), "cl": float(f[3]), "threads": 1 if serial else threads}


CPU_ARM_TEXT = ("CPU restatement of the reference algorithm (oracle/: naive (q,i,j) assembly loop, GMRES(28) + aSIMPLE, "
                "per-step SpGEMM for S, ILU(0)-preconditioned inner GMRES to 1e-2)")


def cpu_config(a):
    return {"workload": f"C3: {a.cpu_mesh} Re=20 (mesh/domain3D.geo geometry, tests/3D/test_01 parameters), h={a.cpu_h} "
                        f"-- the largest BASELINE.json config the CPU arm completes in minutes; the native arm's "
                        f"headline ({a.mesh} h={a.h}) is not CPU-runnable in this budget",
            "mesh": a.cpu_mesh, "h": a.cpu_h, "deltat": DT, "Re": RE, "quadrature": "dealii95 (14-pt)",
            "gmres": "left-preconditioned GMRES(28), rtol 1e-6 (reference stopping rule)",
            "preconditioner": "aSIMPLE alpha=0.5, ILU(0)-preconditioned inner GMRES to 1e-2 (reference :934-995)",
            "l2_policy": "n/a (CPU)", "parallelism": "one subdomain per host core, rank-local ILU(0)"}


class NativeRun:
    """One resident problem on this rank's GPU, stepped through the C ABI (include/nsb.h)."""

    def __init__(self, pkg, a, mesh, h, dist, rank, world, local_rank):
        self.pkg, self.a, self.dist, self.rank, self.world = pkg, a, dist, rank, world
        t_setup = time.perf_counter()
        prob = pkg.Problem.generate(mesh, h)
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
        kF, rF, kS, rS = a.sweeps.split(",")
        dev.set_inner(int(kF), float(rF) if int(kF) > 0 else 0.0, int(kS), float(rS) if int(kS) > 0 else 0.0)
        sm = a.schur.split(",")
        dev.set_schur_solver(int(sm[0]), int(sm[1]), float(sm[2]), float(sm[3]), int(sm[4]))
        self.prob, self.loc, self.dev, self.dim, self.sz = prob, loc, dev, dim, sz
        self.N = sz["n_u"] + sz["n_p"]   # global unknowns
        self.N_loc = dev.N               # this rank's vector (owned + ghost velocity, replicated pressure)
        self.info = dev.info()
        self.t_setup = time.perf_counter() - t_setup
        L = dev.L
        if loc is None:
            self.bc_dofs = np.array(prob.array("bc.dofs"))
            bc_vals = prob.array("bc.values")
        else:
            bn = loc.array("bc_nodes")
            self.bc_dofs = (dim * bn[:, None] + np.arange(dim, dtype=np.uint32)[None, :]).astype(np.uint32).ravel()
            bc_vals = loc.array("bc_values")
        self.pin_bc = L.nsb_alloc_pinned(8 * max(self.bc_dofs.size, 1))
        self.pin_sol = L.nsb_alloc_pinned(8 * self.N_loc)
        self.bc_host = np.frombuffer((C.c_char * (8 * self.bc_dofs.size)).from_address(self.pin_bc), dtype=np.float64)
        self.sol_host = np.frombuffer((C.c_char * (8 * self.N_loc)).from_address(self.pin_sol), dtype=np.float64)
        self.bc_host[:] = bc_vals
        self.t = 0.0
        self.iters, self.tasm, self.tprec, self.tsol = [], [], [], []
        self.forces = None

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def max_over_ranks(self, v):
        if self.dist is None:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, v):
        if self.dist is None:
            return np.asarray(v, np.float64)
        import torch
        t = torch.tensor(np.asarray(v, np.float64))
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.numpy()

    def step(self, e2e):
        dev = self.dev
        self.t += DT
        if e2e:  # the facade's per-step host traffic: Dirichlet values in, solution + forces out
            dev.set_dirichlet(self.bc_dofs, self.bc_host)
        dev.assemble(self.t)
        it, tp, ts = dev.solve_time_step()
        self.forces = dev.compute_forces(self.prob.mean_velocity(self.t))
        if e2e:
            dev.solution(out=self.sol_host)
        tm = dev.timers()
        self.iters.append(it)
        self.tasm.append(tm[0])
        self.tprec.append(tm[1])
        self.tsol.append(tm[2])

    def timed(self, steps, e2e):
        """K steps between two CUDA events on the context's stream; max over ranks.  Returns
        (ms/step by events, wall ms/step, launches, stats)."""
        dev = self.dev
        del self.iters[:], self.tasm[:], self.tprec[:], self.tsol[:]
        l0 = dev.launch_count()
        self.barrier()
        t0 = time.perf_counter()
        dev.timer_start()
        for _ in range(steps):
            self.step(e2e)  # every API call ends with a stream synchronise
        ms_events = dev.timer_stop() / steps
        self.barrier()
        ms_wall = self.max_over_ranks(1e3 * (time.perf_counter() - t0) / steps)
        stats = dict(iters=float(np.mean(self.iters)), asm=float(np.mean(self.tasm)), prec=float(np.mean(self.tprec)),
                     sol=float(np.mean(self.tsol)))
        return self.max_over_ranks(ms_events), ms_wall, dev.launch_count() - l0, stats

    def measure(self, steps, warmup):
        """warm-up, then `value` and `e2e` over THE SAME time steps: the state after the warm-up is saved and
        restored between the two timed loops."""
        for _ in range(warmup):
            self.step(False)
        snap, t_snap = self.dev.solution().copy(), self.t
        ms_dev, ms_wall, launches, stats = self.timed(steps, False)
        forces = self.forces.copy()
        checks = self.checksums()
        self.dev.set_solution(snap)
        self.t = t_snap
        ms_e2e, _, _, stats_e2e = self.timed(steps, True)
        return dict(ms=ms_dev, wall=ms_wall, launches=launches, stats=stats, e2e=ms_e2e, forces=forces,
                    e2e_iters=stats_e2e["iters"], checks=checks,
                    h2d=int(8 * self.bc_dofs.size + 4 * self.bc_dofs.size), d2h=int(8 * self.N_loc + 32))

    def checksums(self):
        """Rank-count-invariant digests of the state after the timed steps: l2 norms of the velocity (owned
        dofs summed over ranks) and of the (replicated) pressure."""
        x = self.dev.solution()
        n_u = self.dev.n_u
        n_uloc = getattr(self.dev, "n_uloc", n_u)
        u2 = float(self.sum_over_ranks([float(np.dot(x[:n_u], x[:n_u]))])[0])
        p = x[n_uloc:]
        return {"velocity_l2": u2 ** 0.5, "pressure_l2": float(np.dot(p, p)) ** 0.5}

    def close(self):
        L = self.dev.L
        L.nsb_free_pinned(self.pin_bc)
        L.nsb_free_pinned(self.pin_sol)
        self.dev.close()


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
        r = cpu_arm(pkg, a.cpu_mesh, a.cpu_h, a.steps, a.warmup, threads, budget_s=a.cpu_budget_s)
        sample = (f"{CPU_ARM_TEXT} as `mpirun -n {threads}` runs it (one subdomain per core, rank-local ILU(0), "
                  f"threaded SpMV/dots) on {a.cpu_mesh} h={a.cpu_h}: {r['n_dofs']} DoFs, {r['n_cells']} cells, "
                  f"{r['steps_timed']} steps timed after {r['warmup_run']} warm-up steps, "
                  f"{r['gmres_iters_per_step']:.1f} GMRES its/step; measured, nothing extrapolated")
        line = {"impl": "reference", "metric": "time_per_step", "value": r["ms_per_step"], "unit": "ms/step",
                "n_gpus": a.gpus, "steps": r["steps_timed"], "warmup": r["warmup_run"], "steps_requested": a.steps,
                "ms_per_step": r["ms_per_step"], "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": cpu_config(a), "n_dofs": r["n_dofs"],
                "n_cells": r["n_cells"], "gmres_iters_per_step": r["gmres_iters_per_step"], "phase_ms": r["phase_ms"],
                "cd": r["cd"], "cl": r["cl"],
                "cpu_baseline": {"value": r["ms_per_step"], "unit": "ms/step", "cores": threads, "kind": "port",
                                 "sample": sample},
                "e2e": {"value": r["ms_per_step"], "unit": "ms/step", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "paired_with": "the native arm's `c3` record (same mesh, same parameters, same number of steps)",
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return 0

    # ---------------- native arm ----------------
    dist = None
    if world > 1:  # one process per GPU (torchrun); gloo carries the NCCL id, barriers and the max over ranks
        import torch.distributed as dist
        dist.init_process_group("gloo")
    run = NativeRun(pkg, a, a.mesh, a.h, dist, rank, world, local_rank)
    dev, info, dim, N, N_loc = run.dev, run.info, run.dim, run.N, run.N_loc
    sampler = ClockSampler(local_rank)
    sampler.start()
    m = run.measure(a.steps, a.warmup)
    ms_dev, ms_wall, ms_e2e, launches, dev_stats, forces = m["ms"], m["wall"], m["e2e"], m["launches"], m["stats"], m["forces"]
    run.barrier()
    # kernel micro-benchmarks on the resident system (CUDA events on the ctx stream)
    reps = 20
    ms_spmv = dev.bench_kernel(5, reps)
    ms_sweep = dev.bench_kernel(4, reps)
    ms_sweep_s = dev.bench_kernel(6, reps)
    ms_asm = dev.bench_kernel(1, 5)
    ms_prec = dev.bench_kernel(2, 5)
    ms_schur = dev.bench_kernel(3, 5)
    ms_spmv_can = dev.bench_kernel(0, 10) if (not a.no_canonical_spmv and world == 1) else None
    parts = {"F_solve": dev.bench_kernel(9, 10), "schur_solve": dev.bench_kernel(10, 10),
             "block_product": ms_spmv, "gram_schmidt_k14": dev.bench_kernel(13, 10),
             "B_vec0": dev.bench_kernel(8, 10), "Bt_dst1": dev.bench_kernel(7, 10)}
    if world > 1:
        parts["velocity_halo_exchange"] = dev.bench_kernel(11, 20)
        parts["pressure_allgather"] = dev.bench_kernel(12, 20)
    clocks = sampler.stop()

    gb = 1e-9
    spmv_gbs = spmv_bytes(info, dim) * gb / (ms_spmv * 1e-3)
    sweep_gbs = sweep_bytes(info, dim) * gb / (ms_sweep * 1e-3)
    sweep_s_gbs = sweep_s_bytes(info) * gb / (ms_sweep_s * 1e-3)
    asm_gbs = assembly_bytes(info, dim) * gb / (ms_asm * 1e-3)
    # share of one outer GMRES iteration: the fine-level F passes, the fine-level S sweeps, one block product
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
        import glob
        tj = json.load(open(sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))[-1]))  # newest round
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
            "gram_schmidt_reorth_passes": info2["reorth_passes"], "inner_F_polynomial": dev.inner_params(),
            "exchanges": ("peer memory (IPC-mapped NVLink stores)" if info2["exchange_mode"] & 1 else "NCCL") if world > 1 else "none",
            "schur_fine_level_distributed": bool(info2["exchange_mode"] & 2),
            "spmv_canonical_gbs": (spmv_canonical_bytes(info) * gb / (ms_spmv_can * 1e-3)) if ms_spmv_can else None,
            "spmv_canonical_ms": ms_spmv_can,
            "spmv_canonical_frac_of_hbm": (spmv_canonical_bytes(info) * gb / (ms_spmv_can * 1e-3) / hbm_peak) if ms_spmv_can else None,
            "prec_apply_ms": ms_prec, "iteration_parts_ms": parts, "cd": float(forces[2]), "cl": float(forces[3]),
            "rank_count_invariants": dict(m["checks"], cd=float(forces[2]), cl=float(forces[3]),
                                          note="state after the timed steps; equal across --gpus N up to the GMRES tolerance"),
            "setup_s": run.t_setup, "device_bytes": info["device_bytes"],
            "roofline": roof, "clocks": clocks,
            "e2e": {"value": ms_e2e, "unit": "ms/step", "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": m["d2h"],
                    "same_steps_as_value": True, "gmres_iters_per_step": m["e2e_iters"]},
            "gpu_launches": int(launches)}
    if world > 1:
        line["n_dofs_local"] = N_loc
        line["partition"] = "recursive coordinate bisection of the cells; velocity rows distributed"
    run.barrier()
    run.close()
    if rank != 0:
        dist.destroy_process_group()
        return 0
    if world == 1 and not a.no_c3:
        # BASELINE.json config C3 on the same GPU: the mesh the CPU arm runs (paired same-config record)
        c3 = NativeRun(pkg, a, a.cpu_mesh, a.cpu_h, None, 0, 1, local_rank)
        mc = c3.measure(a.steps, a.warmup)
        i3 = c3.info
        mat_mb = (10 * (i3["nnz_a00"] // (dim * dim)) + 10 * i3["nnz_a01"] + 12 * i3["nnz_a10"] + 12 * i3["nnz_s"]) / 1e6
        line["c3"] = {"workload": cpu_config(a)["workload"], "mesh": a.cpu_mesh, "h": a.cpu_h, "n_dofs": c3.N,
                      "n_cells": i3["n_cells"], "steps": a.steps, "warmup": a.warmup, "ms_per_step": mc["ms"],
                      "wall_ms_per_step": mc["wall"], "gmres_iters_per_step": mc["stats"]["iters"],
                      "phase_ms": {"assemble": mc["stats"]["asm"], "prec_init": mc["stats"]["prec"],
                                   "solve": mc["stats"]["sol"]},
                      "e2e": {"value": mc["e2e"], "unit": "ms/step", "h2d_bytes_per_step": mc["h2d"],
                              "d2h_bytes_per_step": mc["d2h"], "same_steps_as_value": True},
                      "cd": float(mc["forces"][2]), "cl": float(mc["forces"][3]), "gpu_launches": int(mc["launches"]),
                      "caveat": f"L2-resident: the matrices hold {mat_mb:.0f} MB < 126 MB of L2; launch- and "
                                f"latency-bound, no roofline is quoted for it",
                      "paired_with": "bench.py --impl reference (same mesh, parameters and step count on the host cores)"}
        c3.close()
    if not a.no_cpu_baseline and world == 1:
        # bounded CPU sample on the same box: all cores (partitioned) and one thread, both measured on C3
        threads = os.cpu_count() or 1
        par = cpu_arm(pkg, a.cpu_mesh, a.cpu_h, 3, 1, threads, budget_s=60.0)
        ser = cpu_arm(pkg, a.cpu_mesh, a.cpu_h, 1, 1, 1, budget_s=60.0, serial=True)
        line["cpu_baseline"] = {
            "value": par["ms_per_step"], "unit": "ms/step", "cores": threads, "kind": "port",
            "sample": (f"{CPU_ARM_TEXT} on C3 ({a.cpu_mesh} h={a.cpu_h}, {par['n_dofs']} DoFs), NOT on the headline mesh: "
                       f"{par['steps_timed']} steps after 1 warm-up step as `mpirun -n {threads}` runs it (one subdomain per "
                       f"core, rank-local ILU(0)): {par['ms_per_step']:.0f} ms/step, {par['gmres_iters_per_step']:.1f} GMRES "
                       f"its; single thread: {ser['ms_per_step']:.0f} ms/step ({ser['steps_timed']} step). Compare with `c3`, "
                       f"not with `value`"),
            "single_thread_ms_per_step": ser["ms_per_step"], "all_cores_phase_ms": par["phase_ms"],
            "single_thread_phase_ms": ser["phase_ms"], "config": {"mesh": a.cpu_mesh, "h": a.cpu_h, "n_dofs": par["n_dofs"]},
            "native_same_config_ms_per_step": line.get("c3", {}).get("ms_per_step"),
            # labelled extrapolation, never used as a value: linear in DoFs ignores the growth of the GMRES and
            # ILU costs with the mesh, so it is a LOWER bound of the CPU time on the headline mesh
            "extrapolated_to_headline": {"extrapolated": True, "ms_per_step_lower_bound": par["ms_per_step"] * N / par["n_dofs"],
                                         "how": "all-cores figure x (headline DoFs / C3 DoFs)"}}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
