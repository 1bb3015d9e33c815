"""Python emulation (scipy) of the Schur V-cycle of csrc/amg.cuh + nsb_capi.cu (amg_build / amg_vcycle), for CPU
studies of the preconditioner (tests/prec_study.py, tests/naca_study.py).  Experiment code, not product."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def coarsen(M, theta, max_agg, signed=False, rel=False):
    M = M.tocsr()
    n = M.shape[0]
    diag = np.abs(M.diagonal())
    agg = np.full(n, -1, np.int64)
    na = 0
    rp, ci, va = M.indptr, M.indices, M.data
    rowmax = np.zeros(n)
    if rel:  # classical (Ruge-Stueben) strength: relative to the strongest coupling of the row
        for i in range(n):
            for k in range(rp[i], rp[i + 1]):
                if ci[k] != i:
                    rowmax[i] = max(rowmax[i], -va[k] if signed else abs(va[k]))
    for i in range(n):
        if agg[i] >= 0:
            continue
        nb = []
        best_w, best = -1.0, -1
        for k in range(rp[i], rp[i + 1]):
            j = ci[k]
            if j == i:
                continue
            w = -va[k] if signed else abs(va[k])
            if w < (theta * rowmax[i] if rel else theta * np.sqrt(diag[i] * diag[j])) or w <= 0:
                continue
            if agg[j] < 0:
                nb.append((w, j))
            elif w > best_w:
                best_w, best = w, j
        if not nb and best >= 0:
            agg[i] = agg[best]
            continue
        nb.sort(key=lambda t: -t[0])
        agg[i] = na
        for t in range(min(len(nb), max_agg - 1)):
            agg[nb[t][1]] = na
        na += 1
    P = sp.csr_matrix((np.ones(n), (np.arange(n), agg)), shape=(n, na))
    return P


def lmax_of(M, dinv):
    K = sp.diags(dinv) @ M
    if M.shape[0] < 4:
        return float(np.max(np.abs(np.linalg.eigvals(K.toarray()))))
    return float(abs(spla.eigs(K, k=1, which="LM", return_eigenvectors=False, tol=1e-3)[0]))


class VCycle:
    def __init__(self, S, theta=0.08, max_agg=8, nu=1, omega=1.5, smooth_ratio=4.0, coarse_sweeps=16, coarse_ratio=60.0,
                 min_rows=64, signed=False, rel=False, theta_decay=0.5):
        self.nu, self.omega, self.sr, self.cs, self.cr = nu, omega, smooth_ratio, coarse_sweeps, coarse_ratio
        self.M, self.P, self.dinv, self.lmax = [S.tocsr()], [], [], []
        while self.M[-1].shape[0] > min_rows and len(self.M) < 16:
            M = self.M[-1]
            P = coarsen(M, theta * theta_decay ** (len(self.M) - 1), max_agg, signed, rel)
            if P.shape[1] >= 0.9 * M.shape[0]:
                break
            self.P.append(P)
            self.M.append((P.T @ M @ P).tocsr())
        for M in self.M:
            d = 1.0 / M.diagonal()
            self.dinv.append(d)
            self.lmax.append(1.05 * lmax_of(M, d))

    def cheb(self, l, b, k, ratio, z0=None):
        M, dinv, lmax = self.M[l], self.dinv[l], self.lmax[l]
        lmin = lmax / ratio
        theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
        sigma = theta / delta
        if z0 is None:
            d = dinv * b / theta
            z = d.copy()
        else:
            d = dinv * (b - M @ z0) / theta
            z = z0 + d
        rho = 1.0 / sigma
        for _ in range(1, k):
            rn = 1.0 / (2 * sigma - rho)
            d = rn * rho * d + 2 * rn / delta * dinv * (b - M @ z)
            z = z + d
            rho = rn
        return z

    def cycle(self, l, b):
        if l + 1 == len(self.M):
            return self.cheb(l, b, self.cs, self.cr)
        z = self.cheb(l, b, self.nu, self.sr)
        r = b - self.M[l] @ z
        ec = self.cycle(l + 1, self.P[l].T @ r)
        z = z + self.omega * (self.P[l] @ ec)
        return self.cheb(l, b, self.nu, self.sr, z0=z)

    def solve(self, b):
        return self.cycle(0, b)

    def sizes(self):
        return [m.shape[0] for m in self.M]
