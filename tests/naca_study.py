#!/usr/bin/env python
"""CPU study (scipy, no GPU; an experiment, not a test -- it lives under tests/ because it uses the oracle): the
NACA 2408 / 10 degrees case of tests/2D/test_naca/run_test.sh at h = 0.03, where round 2's first GPU run did not
converge.  Emulates the device preconditioner -- aSIMPLE with the Chebyshev-Jacobi polynomial on F (interval or
ellipse form) and the aggregation V-cycle on S (tests/amg_emul.py) -- inside GMRES(28) and prints outer iteration
counts next to an exact Schur solve.

    python tests/naca_study.py [time steps taken by the oracle before the system is frozen = 2]
"""
import importlib
import os
import sys

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
pkg = importlib.import_module("navierstokes-capoferri_cecchettini_untila_b200")
import oracle.ns_oracle as om  # noqa: E402
from amg_emul import VCycle  # noqa: E402
from conftest import make_configured_case  # noqa: E402

cfg = dict(mesh="airfoil:2408:0.4:10", h=0.03, uniform=True, um=1.0, re=None, dt=0.01, sin=False)
prob, orc, dim, nu = make_configured_case(pkg, om, **cfg)
orc.set_solver(1e-12, 30, 10000, 1e-10)
t = 0
for step in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    t += cfg["dt"]
    orc.assemble(t)
    print("oracle step", step, "its", orc.solve_time_step()[1])
t += cfg["dt"]
orc.assemble(t)
B = orc.scipy_blocks()
A00, A01, A10 = B["a00"].tocsr(), B["a01"].tocsr(), B["a10"].tocsr()
rhs, x0 = orc.rhs(), orc.solution().copy()
n_u, n_p = orc.n_u, orc.n_p
A = sp.bmat([[A00, A01], [A10, None]]).tocsr()
Fs = A00[0::dim, :][:, 0::dim].tocsr()
Dinv = 1 / Fs.diagonal()
Di_full = np.repeat(Dinv, dim)
S = (A10 @ sp.diags(Di_full) @ A01).tocsr()
S_lu = spla.splu(S.tocsc())
K = sp.diags(Dinv) @ Fs
lmax = 1.05 * spla.eigs(K, k=1, which="LR", return_eigenvectors=False, tol=1e-4)[0].real
ev = spla.eigs(K, k=4, which="LI", return_eigenvectors=False, tol=1e-4)
print(f"lambda_max(D^-1 F) = {lmax / 1.05:.3f}, eigenvalues with the largest imaginary part: {np.round(ev, 3)}")


def cheb(b, k, theta, c2):
    tk = 1 / theta
    d = Dinv[:, None] * b * tk
    z = d.copy()
    for _ in range(1, k):
        tn = 1 / (2 * theta - c2 * tk)
        d = c2 * tn * tk * d + 2 * tn * Dinv[:, None] * (b - Fs @ z)
        z = z + d
        tk = tn
    return z


def asimple(k, ratio, imag, ssolve=None):
    lmin = lmax / ratio
    theta, a = (lmax + lmin) / 2, (lmax - lmin) / 2
    c2 = a * a - imag * imag

    def apply(src):
        v0 = cheb(src[:n_u].reshape(-1, dim), k, theta, c2).ravel()
        v1 = src[n_u:] - A10 @ v0
        d1 = -(ssolve(v1) if ssolve else S_lu.solve(v1)) / 0.5
        return np.concatenate([v0 - Di_full * (A01 @ d1), d1])
    return apply


def gmres_left(prec, tol_rel=1e-6, m=28, maxit=1500):
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
                tt = cs[i] * H[i, j] + sn[i] * H[i + 1, j]
                H[i + 1, j] = -sn[i] * H[i, j] + cs[i] * H[i + 1, j]
                H[i, j] = tt
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


print("-- F polynomial (exact Schur solve): degree, imaginary half-axis -> outer iterations (1500 = no convergence)")
for k, ratio in ((3, 7.4), (4, 13.2), (11, 100.0)):
    print(f"   degree {k}:", {imag: gmres_left(asimple(k, ratio, imag)) for imag in (0.0, 1.0, 1.5, 2.0)}, flush=True)
print("-- Schur V-cycle (F polynomial of degree 11, imaginary half-axis 1.85): strength measure -> outer iterations")
for kw in (dict(), dict(theta=0.35), dict(signed=True, rel=True, theta=0.25, theta_decay=1.0),
           dict(signed=True, rel=True, theta=0.35, theta_decay=1.0), dict(signed=True, rel=True, theta=0.5, theta_decay=1.0)):
    vc = VCycle(S, **kw)
    E = np.eye(n_p) - np.column_stack([vc.solve(S @ e) for e in np.eye(n_p)])
    lam = 1 - np.linalg.eigvals(E).real
    print(f"   {kw or 'round-1 default (abs, 0.08, halved per level)'}: levels {vc.sizes()}, spectrum of V S in "
          f"[{lam.min():.4f}, {lam.max():.2f}], outer its {gmres_left(asimple(11, 100.0, 1.85, vc.solve))}", flush=True)
print("   exact Schur solve: outer its", gmres_left(asimple(11, 100.0, 1.85)))
