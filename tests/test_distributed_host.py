"""Host-side domain decomposition (include/nsb_host.h: nsh_partition /
nsh_localize) checked on CPU: structural invariants, and a world_size-2 gloo run
in which every rank multiplies its local rows after a halo exchange and the
result equals the single-rank product (the N > 1 data path without a GPU)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import ROOT, make_case


@pytest.mark.parametrize("key,parts", [("2d-cylinder", 2), ("3d-cylinder", 4), ("3d-square", 8)])
def test_localize_invariants(pkg, oracle_mod, key, parts):
    prob, orc, dim, nu, um = make_case(pkg, oracle_mod, key)
    prob.partition(parts)
    s0 = prob.sizes()
    locs = [pkg.LocalProblem(prob, parts, r) for r in range(parts)]
    sz = [l.sizes() for l in locs]
    assert sum(s["n_own"] for s in sz) == s0["n_nodes"] and sum(s["n_p_own"] for s in sz) == s0["n_p"]
    assert sum(s["n_bc_nodes"] for s in sz) * dim == s0["n_bc"]
    assert sum(s["n_force_faces"] for s in sz) == s0["n_force_faces"]
    perm = locs[0].array("node_perm")
    assert np.array_equal(np.sort(perm), np.arange(s0["n_nodes"]))
    for r, l in enumerate(locs):
        assert np.array_equal(l.array("node_perm"), perm)
        nb, sp_, rp_ = l.array("neighbors"), l.array("send_ptr"), l.array("recv_ptr")
        assert rp_[-1] == sz[r]["n_ghost"] and r not in nb
        for k, q in enumerate(nb):  # what r sends to q is what q expects from r
            lq = locs[q]
            kq = list(lq.array("neighbors")).index(r)
            n_send = sp_[k + 1] - sp_[k]
            n_recv_q = lq.array("recv_ptr")[kq + 1] - lq.array("recv_ptr")[kq]
            assert n_send == n_recv_q
            sent = sz[r]["node_offset"] + l.array("send_idx")[sp_[k]:sp_[k + 1]]
            got = lq.array("ghost_dist")[lq.array("recv_ptr")[kq]:lq.array("recv_ptr")[kq + 1]]
            assert np.array_equal(sent, got)
        # every column of an owned row is local, rows are sorted
        rp, ci = l.array("fs.rowptr"), l.array("fs.colind")
        assert ci.max() < sz[r]["n_own"] + sz[r]["n_ghost"]
        assert all(np.all(np.diff(ci[rp[i]:rp[i + 1]].astype(np.int64)) > 0) for i in range(0, rp.size - 1, 97))
        # local cells cover every cell the partition gives this rank
        mine = np.flatnonzero(prob.array("part.cell") == r)
        assert np.isin(mine, l.array("cells")).all()


def test_local_vectors_roundtrip(pkg, oracle_mod):
    prob, orc, dim, nu, um = make_case(pkg, oracle_mod, "3d-cylinder")
    parts = 3
    prob.partition(parts)
    x = np.random.default_rng(0).standard_normal(orc.N)
    out = np.zeros_like(x)
    for r in range(parts):
        loc = pkg.LocalProblem(prob, parts, r)
        out = loc.owned_to_global(loc.to_local(x, dim), dim, out)
    assert np.array_equal(out, x)


WORKER = r'''
import importlib, os, sys
import numpy as np, scipy.sparse as sp, torch, torch.distributed as dist
sys.path.insert(0, os.environ["NSB_ROOT"]); sys.path.insert(0, os.path.join(os.environ["NSB_ROOT"], "tests"))
from conftest import make_case
pkg = importlib.import_module("navierstokes-capoferri_cecchettini_untila_b200")
from oracle import ns_oracle
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
prob, orc, dim, nu, um = make_case(pkg, ns_oracle, "3d-cylinder")
prob.partition(world)
loc = pkg.LocalProblem(prob, world, rank)
s = loc.sizes()
# single-rank truth from the oracle's assembled blocks
xs = np.zeros(orc.N); xs[:orc.n_u] = 0.2 * np.sin(np.arange(orc.n_u))
orc.set_solution(xs); orc.assemble(0.01)
B = orc.scipy_blocks()
v = np.cos(0.37 * np.arange(orc.N))
y_ref = orc.vmult(v)
# local rows of F (node level, from the canonical A00), A01, A10 in local numbering
cn = loc._maps(); n_own, n_loc = s["n_own"], s["n_own"] + s["n_ghost"]
udofs = (dim * cn[:, None] + np.arange(dim)[None, :]).ravel()
pinv = np.argsort(loc.array("p_perm").astype(np.int64))            # distributed p id -> canonical
F_loc = B["a00"][udofs[:dim * n_own]][:, udofs]
A01_loc = B["a01"][udofs[:dim * n_own]][:, pinv]
own_p = pinv[s["p_offset"]:s["p_offset"] + s["n_p_own"]]
A10_loc = B["a10"][own_p][:, udofs]
# local pattern from nsh_localize equals the pattern of the extracted rows
rp, ci = loc.array("fs.rowptr"), loc.array("fs.colind")
Fn = F_loc[0::dim][:, 0::dim].tocsr(); Fn.sort_indices()
assert np.array_equal(Fn.indptr, rp) and np.array_equal(Fn.indices, ci), "fs pattern"
A10c = A10_loc.tocsr(); A10c.sort_indices()
assert np.array_equal(A10c.indptr, loc.array("a10.rowptr")) and np.array_equal(A10c.indices, loc.array("a10.colind"))
# distributed product: owned velocity entries only, ghosts by halo exchange over gloo
x = loc.to_local(v, dim)
x[dim * n_own: dim * n_loc] = np.nan                              # ghosts must come from the exchange
nb, sp_, si, rp_ = (loc.array(k) for k in ("neighbors", "send_ptr", "send_idx", "recv_ptr"))
reqs, bufs = [], []
for k, q in enumerate(nb):
    idx = si[sp_[k]:sp_[k + 1]].astype(np.int64)
    send = torch.from_numpy(np.ascontiguousarray(x[:dim * n_own].reshape(-1, dim)[idx]))
    recv = torch.empty((int(rp_[k + 1] - rp_[k]), dim), dtype=torch.float64)
    reqs += [dist.isend(send, int(q)), dist.irecv(recv, int(q))]
    bufs.append((k, recv))
for r in reqs: r.wait()
for k, recv in bufs:
    x[dim * (n_own + rp_[k]): dim * (n_own + rp_[k + 1])] = recv.numpy().ravel()
assert not np.isnan(x).any()
y_u = F_loc @ x[:dim * n_loc] + A01_loc @ x[dim * n_loc:]
y_p_own = A10_loc @ x[:dim * n_loc]
# replicate the pressure rows (allgatherv) and reduce a dot product over owned entries
pieces = [None] * world
dist.all_gather_object(pieces, (s["p_offset"], y_p_own))
y_p = np.empty(s["n_p"])
for off, arr in pieces: y_p[off:off + arr.size] = arr
y_loc = np.concatenate([y_u, np.zeros(dim * s["n_ghost"]), y_p])
out = np.zeros(orc.N)
out = loc.owned_to_global(y_loc, dim, out)
t = torch.from_numpy(out[:orc.n_u].copy()); dist.all_reduce(t)   # owned velocity rows are disjoint
y = np.concatenate([t.numpy(), out[orc.n_u:]])
assert np.abs(y - y_ref).max() < 1e-12 * np.abs(y_ref).max(), np.abs(y - y_ref).max()
d = torch.tensor([float(y_u @ y_u) + (float(y_p @ y_p) if rank == 0 else 0.0)], dtype=torch.float64); dist.all_reduce(d)
assert abs(d.item() - float(y_ref @ y_ref)) < 1e-10 * float(y_ref @ y_ref)
print("rank", rank, "ok", flush=True)
dist.destroy_process_group()
'''


def test_world_size_2_gloo_halo_spmv(pkg, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, NSB_ROOT=ROOT, OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == 2
