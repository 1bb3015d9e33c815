#!/usr/bin/env python
"""CPU study (scipy, no GPU): outer GMRES(28) iteration counts of the aSIMPLE preconditioner as a function of
the inner F solve -- exact, Chebyshev-Jacobi polynomial of degree k (what prec_apply runs), and a two-level
cycle whose coarse space is the P1 subspace of the P2 velocity space (p-coarsening).  Matrices come from the
CPU oracle (test infrastructure, which is why this study lives under tests/; it is an experiment, not a test).
STUDY=schur python tests/prec_study.py: the Schur V-cycle variants instead.

    python tests/prec_study.py [h=0.05] [extra diffusion factor=1]

deltat and nu are scaled with h so that the diffusion number and the CFL number match the 9.7 M-DoF workload
on a mesh the CPU can factorise.
"""
import importlib
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("navierstokes-capoferri_cecchettini_untila_b200")
from oracle.ns_oracle import Oracle  # noqa: E402

h = float(sys.argv[1]) if len(sys.argv) > 1 else 0.05
# Same nu*dt/h^2 (diffusion number) and |u|*dt/h (CFL) as the 9.7 M-DoF workload (h = 0.011, dt = 0.01, nu = 0.004):
# dt and nu both scale with h, so F = M/dt + nu K + C(u) keeps the relative weights of its three parts.
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0  # extra factor on the diffusion number
dt = 0.01 * h / 0.011
nu = 0.004 * h / 0.011 * scale
dim = 3
um = 0.45
prob = pkg.Problem.generate("3d-cylinder", h).build(inlet=(pkg.INLET_PARABOLIC, um, 0.41, 0))
orc = Oracle(dim, prob.array("xyz"), prob.array("cells"), prob.array("bfaces"), prob.array("bids"))
orc.set_inlet(0, um, 0.41, 0)
orc.set_params(dt, nu)
orc.set_threads(16)
n_u, n_p, N = orc.n_u, orc.n_p, orc.N
print(f"h={h} dt={dt} N={N} n_u={n_u} n_p={n_p} nu={nu}")
# a developed-looking velocity: two coarse time steps of the oracle itself
orc.set_solver(1e-6, 30, 10000, 1e-2)
for s in range(2):
    orc.assemble((s + 1) * dt)
    rc, it, _, _ = orc.solve_time_step()
    print("oracle step", s, "its", it)
orc.assemble(3 * dt)
B = orc.scipy_blocks()
A00, A01, A10 = B["a00"].tocsr(), B["a01"].tocsr(), B["a10"].tocsr()
rhs = orc.rhs()
x0 = orc.solution().copy()
A = sp.bmat([[A00, A01], [A10, None]]).tocsr()
n_nodes = n_u // dim
Fs = A00[0::dim, :][:, 0::dim].tocsr()
D = Fs.diagonal()
Dinv = 1.0 / D
Di_full = np.repeat(Dinv, dim)
S = (A10 @ sp.diags(Di_full) @ A01).tocsc()
S_lu = spla.splu(S)
alpha = 0.5

# gamma as auto_inner measures it needs the mass diagonal; estimate from M = F at nu=0,u=0 is not available here:
# use the spectrum instead
Dh = sp.diags(np.sqrt(Dinv))
ev_max = spla.eigs(sp.diags(Dinv) @ Fs, k=1, which="LM", return_eigenvectors=False, tol=1e-3)[0].real
print(f"lambda_max(D^-1 F_s) = {ev_max:.3f}")

# ---- P2 -> P1 interpolation on nodes ------------------------------------------------------
cn = prob.array("cell_nodes").reshape(-1, 10).astype(np.int64)
cp = prob.array("cell_pverts").reshape(-1, 4).astype(np.int64)
nxyz = prob.array("node_xyz").reshape(-1, dim)
rows, cols, vals = [], [], []
# vertex nodes
vn = cn[:, :4].ravel()
vp = cp.ravel()
rows.append(vn)
cols.append(vp)
vals.append(np.ones(vn.size))
pairs = [(i, j) for i in range(4) for j in range(i + 1, 4)]
vx = nxyz[cn[:, :4]]  # cells x 4 x 3
for e in range(4, 10):
    ex = nxyz[cn[:, e]]
    found = np.zeros(cn.shape[0], bool)
    for (i, j) in pairs:
        mid = 0.5 * (vx[:, i] + vx[:, j])
        m = np.linalg.norm(mid - ex, axis=1) < 1e-9
        m &= ~found
        found |= m
        for q in (i, j):
            rows.append(cn[m, e])
            cols.append(cp[m, q])
            vals.append(np.full(m.sum(), 0.5))
    assert found.all()
r_, c_, v_ = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
_, first = np.unique(r_ * n_p + c_, return_index=True)  # the same (node, vertex) pair comes from every cell around it
P = sp.coo_matrix((v_[first], (r_[first], c_[first])), shape=(n_nodes, n_p)).tocsr()
assert np.allclose(P.sum(axis=1), 1.0)
Fc = (P.T @ Fs @ P).tocsc()
Fc_lu = spla.splu(Fc)
print(f"P1 level: {n_p} rows, nnz/row {Fc.nnz / n_p:.1f}; fine nnz/row {Fs.nnz / n_nodes:.1f}")
Fs_lu = spla.splu(Fs.tocsc())


def to_nodes(v):
    return v.reshape(n_nodes, dim)


def cheb(b, k, lmax, ratio, z0=None, imag=0.0):
    """k Chebyshev-Jacobi sweeps on Fs z = b (b: nodes x dim); zero guess if z0 is None (first sweep free).
    Ellipse form: centre theta, real half-axis a, imaginary half-axis `imag`; only c2 = a^2 - imag^2 (the squared
    focal distance, negative for an upright ellipse) enters the recurrence t_{k+1} = 1 / (2 theta - c2 t_k)."""
    lmin = lmax / ratio
    theta, a = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
    c2 = a * a - imag * imag
    t = 1.0 / theta
    if z0 is None:
        d = Dinv[:, None] * b * t
        z = d.copy()
    else:
        d = Dinv[:, None] * (b - Fs @ z0) * t
        z = z0 + d
    for _ in range(1, k):
        tn = 1.0 / (2 * theta - c2 * t)
        d = c2 * tn * t * d + 2 * tn * Dinv[:, None] * (b - Fs @ z)
        z = z + d
        t = tn
    return z


def gmres_poly_roots(k, seed=1):
    """harmonic Ritz values of k Arnoldi steps on K = D^-1 F_s (the roots of the GMRES residual polynomial),
    in modified Leja order with conjugate pairs kept together"""
    rng = np.random.default_rng(seed)
    v = rng.standard_normal(n_nodes)
    V = [v / np.linalg.norm(v)]
    H = np.zeros((k + 1, k))
    for j in range(k):
        w = Dinv * (Fs @ V[j])
        for _ in range(2):
            for i in range(j + 1):
                hij = V[i] @ w
                H[i, j] += hij
                w -= hij * V[i]
        H[j + 1, j] = np.linalg.norm(w)
        V.append(w / H[j + 1, j])
    Hk = H[:k, :k]
    ek = np.zeros(k)
    ek[-1] = 1.0
    f = np.linalg.solve(Hk.T, ek)
    th = np.linalg.eigvals(Hk + H[k, k - 1] ** 2 * np.outer(f, ek))
    th = [t for t in th if t.imag >= -1e-14]  # one of each conjugate pair
    out = []
    rem = list(th)
    cur = max(rem, key=lambda t: abs(t))
    while rem:
        rem.remove(cur)
        out.append(cur)
        if not rem:
            break
        def score(t):
            p = 0.0
            for o in out:
                p += np.log(abs(t - o)) + (np.log(abs(t - np.conj(o))) if abs(o.imag) > 1e-14 else 0.0)
            return p
        cur = max(rem, key=score)
    return out


def gpoly_apply(b, roots):
    """z = p(K) D^-1 b with residual polynomial prod (1 - lambda/theta_i); same three-term sweep form as the
    Chebyshev kernel: znew = z + c1 (z - zold) + c2 Dinv (b - F z)"""
    bd = Dinv[:, None] * b
    z = np.zeros_like(b)
    zold = None
    first = True
    steps = []
    for t in roots:
        if abs(t.imag) < 1e-14:
            steps.append((0.0, 1.0 / t.real))
        else:
            a, m2 = t.real, abs(t) ** 2
            steps.append((0.0, 1.0 / a))
            steps.append((-(t.imag ** 2) / m2, a / m2))
    for c1, c2 in steps:
        if first:
            zn = c2 * bd  # z = 0: no product
            first = False
        else:
            zn = z + (c1 * (z - zold) if c1 != 0.0 else 0.0) + c2 * (bd - Dinv[:, None] * (Fs @ z))
            passes[0] += 1
        zold, z = z, zn
    return z


passes = [0]


def make_F(kind, **kw):
    lmax = 1.05 * ev_max

    def exact(b):
        return Fs_lu.solve(b)

    def poly(b):
        passes[0] += kw["k"] - 1
        return cheb(b, kw["k"], lmax, kw["ratio"], imag=kw.get("imag", 0.0))

    def twolevel(b):
        pre, post, r = kw["pre"], kw["post"], kw["ratio"]
        z = cheb(b, pre, lmax, r)
        passes[0] += pre - 1
        res = b - Fs @ z
        passes[0] += 1
        ec = Fc_lu.solve(P.T @ res)
        z = z + kw.get("omega", 1.0) * (P @ ec)
        if post:
            z = cheb(b, post, lmax, r, z0=z)
            passes[0] += post
        return z

    def gpoly(b):
        return gpoly_apply(b, kw["roots"])

    return {"exact": exact, "poly": poly, "two": twolevel, "gpoly": gpoly}[kind]


def asimple(Fsolve, Ssolve=None):
    Ssolve = Ssolve or S_lu.solve

    def apply(src):
        v0 = Fsolve(to_nodes(src[:n_u])).ravel()
        v1 = src[n_u:] - A10 @ v0
        d1 = -Ssolve(v1) / alpha
        d0 = v0 - Di_full * (A01 @ d1)
        return np.concatenate([d0, d1])
    return apply


def gmres_left(prec, tol_rel=1e-6, m=28, maxit=400):
    x = x0.copy()
    tol = tol_rel * np.linalg.norm(rhs)
    its = 0
    while True:
        r = prec(rhs - A @ x)
        beta = np.linalg.norm(r)
        if beta <= tol or its >= maxit:
            return its
        V = [r / beta]
        H = np.zeros((m + 1, m))
        g = np.zeros(m + 1)
        g[0] = beta
        cs, sn = np.zeros(m), np.zeros(m)
        k = 0
        for j in range(m):
            its += 1
            w = prec(A @ V[j])
            for _ in range(2):
                for i in range(j + 1):
                    hij = V[i] @ w
                    H[i, j] += hij
                    w -= hij * V[i]
            H[j + 1, j] = np.linalg.norm(w)
            V.append(w / H[j + 1, j])
            for i in range(j):
                t = cs[i] * H[i, j] + sn[i] * H[i + 1, j]
                H[i + 1, j] = -sn[i] * H[i, j] + cs[i] * H[i + 1, j]
                H[i, j] = t
            rr = np.hypot(H[j, j], H[j + 1, j])
            cs[j], sn[j] = H[j, j] / rr, H[j + 1, j] / rr
            H[j, j] = rr
            g[j + 1] = -sn[j] * g[j]
            g[j] *= cs[j]
            k = j + 1
            if abs(g[k]) <= tol or its >= maxit:
                break
        y = np.linalg.solve(np.triu(H[:k, :k]), g[:k])
        for i in range(k):
            x += y[i] * V[i]
        if abs(g[k]) <= tol or its >= maxit:
            return its


if os.environ.get("STUDY") == "schur":
    # the Schur V-cycle (tests/amg_emul.py emulates csrc/amg.cuh) with the F polynomial of degree 8
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from amg_emul import VCycle
    Fsolve = make_F("poly", k=8, ratio=(8 / 1.1) ** 2, imag=0.4)
    print("exact S:", gmres_left(asimple(Fsolve)))
    for kw in (dict(), dict(signed=True, rel=True, theta=0.25, theta_decay=1.0), dict(signed=True, rel=True, theta=0.35, theta_decay=1.0),
               dict(signed=True, rel=True, theta=0.5, theta_decay=1.0), dict(signed=True, rel=True, theta=0.5, theta_decay=0.5)):
        vc = VCycle(S.tocsr(), **kw)
        cx = sum(m.nnz for m in vc.M) / vc.M[0].nnz
        print(kw, "levels", vc.sizes(), f"operator complexity {cx:.2f}", "outer its", gmres_left(asimple(Fsolve, vc.solve)), flush=True)
    sys.exit(0)
cases = [("exact", {})]
if os.environ.get("STUDY") == "accurate":
    for k, ratio in ((16, 100.0), (24, 100.0), (24, 200.0), (40, 200.0)):
        cases.append(("poly", {"k": k, "ratio": ratio, "imag": 0.4}))
    cases.append(("gpoly", {"k": 20, "roots": None}))
for k in (4, 6, 8, 10):
    for imag in (0.0, 0.4):
        cases.append(("poly", {"k": k, "ratio": max(6.0, (k / 1.1) ** 2), "imag": imag}))
for pre, post, ratio in ((1, 2, 10.0), (2, 2, 10.0), (3, 3, 12.0)):
    cases.append(("two", {"pre": pre, "post": post, "ratio": ratio}))
for k in (3, 4, 5, 6, 8, 10):
    cases.append(("gpoly", {"k": k, "roots": gmres_poly_roots(k)}))
for kind, kw in cases:
    if kind == "gpoly" and kw.get("roots") is None:
        kw["roots"] = gmres_poly_roots(kw["k"])
    passes[0] = 0
    t0 = time.time()
    its = gmres_left(asimple(make_F(kind, **kw)))
    kw = {a: b for a, b in kw.items() if a != "roots"}
    print(f"{kind:6s} {kw}: outer its {its:4d}, fine F passes/application {passes[0] / max(1, its + its // 28 + 1):.1f} "
          f"-> F passes per step {passes[0]}  ({time.time() - t0:.1f}s)")
