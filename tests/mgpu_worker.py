"""Multi-GPU worker (one process per GPU, launched by torchrun): steps the
3D cylinder case on WORLD_SIZE GPUs through the distributed C ABI and checks
solution / Cd / Cl against the single-rank CPU oracle at tight tolerance.
Used by tests/test_gpu_multi.py; not a pytest module itself."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import make_case  # noqa: E402

pkg = importlib.import_module("navierstokes-capoferri_cecchettini_untila_b200")
from oracle import ns_oracle  # noqa: E402


def main():
    key = sys.argv[1] if len(sys.argv) > 1 else "3d-cylinder"
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    prob, orc, dim, nu, um = make_case(pkg, ns_oracle, key)
    prob.partition(world)
    loc = pkg.LocalProblem(prob, world, rank)
    ids = [pkg.Device.make_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    dev = pkg.Device(dim, local_rank).load_local_problem(prob, loc, ids[0])
    dev.set_params(0.01, nu)
    dev.set_solver(gmres_rtol=1e-12, restart=60)
    orc.set_solver(1e-12, 30, 10000, 1e-10)
    x = np.zeros(orc.N)
    x[: orc.n_u] = 0.1 * np.sin(np.arange(orc.n_u))
    orc.set_solution(x)
    dev.set_solution(loc.to_local(x, dim))
    t = 0.0
    for step in range(2):
        t += 0.01
        orc.assemble(t)
        dev.assemble(t)
        if step == 0:
            # owned rows of the rhs equal the oracle's (entries parity under decomposition)
            out = loc.owned_to_global(dev.rhs(), dim, np.zeros(orc.N))
            tt = torch.from_numpy(out[: orc.n_u].copy())
            dist.all_reduce(tt)
            ref = orc.rhs()
            assert np.abs(tt.numpy() - ref[: orc.n_u]).max() <= 1e-10 * np.abs(ref).max()
        rc, it_o, _, _ = orc.solve_time_step()
        it_d, _, _ = dev.solve_time_step()
        f_o = orc.compute_forces(t)
        f_d = dev.compute_forces(prob.mean_velocity(t))
        out = loc.owned_to_global(dev.solution(), dim, np.zeros(orc.N))
        tt = torch.from_numpy(out[: orc.n_u].copy())
        dist.all_reduce(tt)
        xg = np.concatenate([tt.numpy(), out[orc.n_u:]])
        xo = orc.solution()
        err = np.linalg.norm(xg - xo) / np.linalg.norm(xo)
        assert rc == 0 and err < 1e-8, (step, err)
        assert abs(f_d[2] - f_o[2]) < 1e-6 * max(1.0, abs(f_o[2])) and abs(f_d[3] - f_o[3]) < 1e-6 * max(1.0, abs(f_o[3]))
        if rank == 0:
            print(f"step {step}: world {world} its gpu/oracle {it_d}/{it_o} rel err {err:.2e} Cd {f_d[2]:.8f}", flush=True)
    print(f"rank {rank} ok", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
