"""ctypes binding of include/nsb.h (libnsb.so, the sm_100a hot path)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None

A00, A01, A10, S = 0, 1, 2, 3
QUAD_DEALII93, QUAD_DEALII95 = 0, 1
PREC_ASIMPLE, PREC_IDENTITY, PREC_AYOSIDA = 0, 1, 2

# every symbol include/nsb.h declares (tests check the .so exports them all)
SYMBOLS = [
    "nsb_last_error", "nsb_device_count", "nsb_create", "nsb_destroy", "nsb_set_mesh", "nsb_set_dofs",
    "nsb_set_pattern", "nsb_set_node_pattern", "nsb_set_quadrature", "nsb_finalize_setup", "nsb_set_params",
    "nsb_set_bc_diag_mode", "nsb_set_solver", "nsb_set_inner", "nsb_set_solution", "nsb_get_solution",
    "nsb_set_dirichlet", "nsb_scale_dirichlet", "nsb_set_force_faces", "nsb_assemble", "nsb_solve_time_step",
    "nsb_compute_forces", "nsb_get_matrix_values", "nsb_get_pattern", "nsb_nnz", "nsb_get_rhs", "nsb_vmult",
    "nsb_bench_kernel", "nsb_launch_count", "nsb_timers", "nsb_info", "nsb_alloc_pinned", "nsb_free_pinned",
    "nsb_comm_unique_id", "nsb_comm_init", "nsb_set_local_dofs", "nsb_set_halo", "nsb_set_schur_solver", "nsb_gather_velocity",
    "nsb_slab_host_check", "nsb_gslab_host_check", "nsb_get_lumped_mass_inv",
    "nsb_timer_start", "nsb_timer_stop", "nsb_cheb_coeffs_host_check", "nsb_skew_radius_host_check", "nsb_inner_params", "nsb_set_schur_strength", "nsb_amg_coarsen_host_check", "nsb_fe_tables_host_check",
]


class DeviceError(RuntimeError):
    pass


def device_lib():
    global _lib
    if _lib is None:
        path = os.environ.get("NSB_LIBNSB") or os.path.join(_HERE, "libnsb.so")  # override: A/B experiments only
        if not os.path.exists(path):
            raise DeviceError("libnsb.so is not built (run `make cuda`); there is no CPU fallback")
        L = C.CDLL(path, mode=C.RTLD_GLOBAL)
        p = C.c_void_p
        f64p, u32p, i64p = (C.POINTER(t) for t in (C.c_double, C.c_uint32, C.c_int64))
        L.nsb_last_error.argtypes = [p]
        L.nsb_last_error.restype = C.c_char_p
        L.nsb_create.argtypes = [C.c_int, C.c_int, C.POINTER(p)]
        L.nsb_destroy.argtypes = [p]
        L.nsb_set_mesh.argtypes = [p, C.c_int64, f64p, C.c_int64, u32p]
        L.nsb_set_dofs.argtypes = [p, C.c_uint32, C.c_uint32, u32p]
        L.nsb_set_pattern.argtypes = [p, C.c_int, C.c_int64, i64p, u32p]
        L.nsb_set_node_pattern.argtypes = [p, C.c_int64, i64p, u32p]
        L.nsb_set_quadrature.argtypes = [p, C.c_int]
        L.nsb_finalize_setup.argtypes = [p]
        L.nsb_set_params.argtypes = [p, C.c_double, C.c_double]
        L.nsb_set_bc_diag_mode.argtypes = [p, C.c_int]
        L.nsb_set_solver.argtypes = [p, C.c_double, C.c_int, C.c_int, C.c_double, C.c_int]
        L.nsb_set_inner.argtypes = [p, C.c_int, C.c_double, C.c_int, C.c_double]
        L.nsb_set_schur_solver.argtypes = [p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int]
        L.nsb_set_solution.argtypes = [p, f64p]
        L.nsb_get_solution.argtypes = [p, f64p]
        L.nsb_set_dirichlet.argtypes = [p, C.c_int64, u32p, f64p]
        L.nsb_scale_dirichlet.argtypes = [p, C.c_double]
        L.nsb_set_force_faces.argtypes = [p, C.c_int64, u32p, f64p, f64p]
        L.nsb_assemble.argtypes = [p, C.c_double]
        L.nsb_solve_time_step.argtypes = [p, C.POINTER(C.c_int), f64p, f64p]
        L.nsb_compute_forces.argtypes = [p, C.c_double, f64p]
        L.nsb_get_matrix_values.argtypes = [p, C.c_int, f64p]
        L.nsb_get_pattern.argtypes = [p, C.c_int, i64p, u32p]
        L.nsb_nnz.argtypes = [p, C.c_int]
        L.nsb_nnz.restype = C.c_int64
        L.nsb_get_rhs.argtypes = [p, f64p]
        L.nsb_get_lumped_mass_inv.argtypes = [p, f64p]
        L.nsb_vmult.argtypes = [p, f64p, f64p]
        L.nsb_bench_kernel.argtypes = [p, C.c_int, C.c_int, f64p]
        L.nsb_timer_start.argtypes = [p]
        L.nsb_timer_stop.argtypes = [p, f64p]
        L.nsb_launch_count.argtypes = [p]
        L.nsb_launch_count.restype = C.c_int64
        L.nsb_timers.argtypes = [p, f64p]
        L.nsb_info.argtypes = [p, i64p]
        L.nsb_alloc_pinned.argtypes = [C.c_int64]
        L.nsb_alloc_pinned.restype = C.c_void_p
        L.nsb_free_pinned.argtypes = [C.c_void_p]
        L.nsb_comm_unique_id.argtypes = [C.c_char_p]
        L.nsb_comm_init.argtypes = [p, C.c_int, C.c_int, C.c_char_p]
        i32p = C.POINTER(C.c_int32)
        L.nsb_set_local_dofs.argtypes = [p, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, u32p, u32p, u32p]
        L.nsb_set_halo.argtypes = [p, C.c_int, i32p, i64p, u32p, i64p]
        L.nsb_gather_velocity.argtypes = [p, u32p, f64p]
        L.nsb_slab_host_check.argtypes = [C.c_int, C.c_int64, C.c_int64, i64p, u32p, f64p, C.c_uint32, f64p, f64p, i64p]
        L.nsb_gslab_host_check.argtypes = [C.c_int, C.c_int64, C.c_int64, i64p, u32p, C.c_uint32, i64p, u32p, f64p, f64p,
                                           f64p, i64p]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def slab_host_check(dim, rowptr, colind, val, x, n_cols=None, window_cap=1408):
    """Host evaluation of y = (F_s (x) I_dim) x through the slab layout of csrc/slab.cuh
    (no device needed).  Returns (y, stats dict)."""
    rowptr = np.ascontiguousarray(rowptr, np.int64)
    colind = np.ascontiguousarray(colind, np.uint32)
    val = np.ascontiguousarray(val, np.float64)
    x = np.ascontiguousarray(x, np.float64)
    n_rows = rowptr.size - 1
    n_cols = n_rows if n_cols is None else n_cols
    y = np.zeros(dim * n_rows)
    st = np.zeros(6, np.int64)
    rc = device_lib().nsb_slab_host_check(dim, n_rows, n_cols, _p(rowptr, C.c_int64), _p(colind, C.c_uint32),
                                          _p(val, C.c_double), window_cap, _p(x, C.c_double), _p(y, C.c_double),
                                          _p(st, C.c_int64))
    if rc != 0:
        raise DeviceError(f"nsb_slab_host_check failed ({rc})")
    return y, dict(zip(["slabs", "nnz", "padded", "max_window", "window_total", "bank_wavefronts_permille"], [int(v) for v in st]))


def cheb_coeffs(k, lmax, ratio, imag=0.0):
    """Coefficients (1/theta, c1[1:k], c2[1:k]) of the F polynomial for an ellipse (host only)."""
    L = device_lib()
    L.nsb_cheb_coeffs_host_check.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double),
                                             C.POINTER(C.c_double), C.POINTER(C.c_double)]
    it = C.c_double()
    c1, c2 = np.zeros(k), np.zeros(k)
    rc = L.nsb_cheb_coeffs_host_check(k, lmax, ratio, imag, C.byref(it), _p(c1, C.c_double), _p(c2, C.c_double))
    if rc != 0:
        raise DeviceError(f"nsb_cheb_coeffs_host_check failed ({rc})")
    return it.value, c1, c2


def skew_radius(H):
    """Largest singular value of the skew part of a small dense matrix and its right singular vector (host only)."""
    H = np.ascontiguousarray(H, np.float64)
    m = H.shape[0]
    L = device_lib()
    L.nsb_skew_radius_host_check.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    sig = C.c_double()
    y = np.zeros(m)
    rc = L.nsb_skew_radius_host_check(m, _p(H, C.c_double), C.byref(sig), _p(y, C.c_double))
    if rc != 0:
        raise DeviceError(f"nsb_skew_radius_host_check failed ({rc})")
    return sig.value, y


def fe_tables(dim, rule=QUAD_DEALII95):
    """Reference-cell contraction tables of the assembly kernel (host only): dict mhat [nn,nn], khat [dim,dim,nn,nn],
    chat [nn,dim,nn,nn] (chat[n,d,a,b] = int phi_a phi_n d_d phi_b), dhat [nn,nv,dim]."""
    nn, nv = (6, 3) if dim == 2 else (10, 4)
    L = device_lib()
    dp = C.POINTER(C.c_double)
    L.nsb_fe_tables_host_check.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp]
    m, k, ch, dh = np.zeros(nn * nn), np.zeros(dim * dim * nn * nn), np.zeros(nn * dim * nn * nn), np.zeros(nn * nv * dim)
    rc = L.nsb_fe_tables_host_check(dim, rule, _p(m, C.c_double), _p(k, C.c_double), _p(ch, C.c_double), _p(dh, C.c_double))
    if rc != 0:
        raise DeviceError(f"nsb_fe_tables_host_check failed ({rc})")
    return {"mhat": m.reshape(nn, nn), "khat": k.reshape(dim, dim, nn, nn), "chat": ch.reshape(nn, dim, nn, nn),
            "dhat": dh.reshape(nn, nv, dim)}


def amg_coarsen(rowptr, colind, val, theta, max_agg=8, measure=0, owner=None):
    """Aggregates of csrc/amg.cuh: coarsen on a CSR matrix (host only).  Returns (agg, n_coarse, coarse_nnz)."""
    rowptr = np.ascontiguousarray(rowptr, np.int64)
    colind = np.ascontiguousarray(colind, np.uint32)
    val = np.ascontiguousarray(val, np.float64)
    n = rowptr.size - 1
    L = device_lib()
    L.nsb_amg_coarsen_host_check.argtypes = [C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_uint32), C.POINTER(C.c_double),
                                             C.c_double, C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_uint32),
                                             C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    agg = np.zeros(n, np.uint32)
    nc, nnzc = C.c_int64(), C.c_int64()
    ow = None if owner is None else np.ascontiguousarray(owner, np.int32)
    rc = L.nsb_amg_coarsen_host_check(n, _p(rowptr, C.c_int64), _p(colind, C.c_uint32), _p(val, C.c_double), theta, max_agg,
                                      measure, None if ow is None else _p(ow, C.c_int32), _p(agg, C.c_uint32),
                                      C.byref(nc), C.byref(nnzc))
    if rc != 0:
        raise DeviceError(f"nsb_amg_coarsen_host_check failed ({rc})")
    return agg, nc.value, nnzc.value


def gslab_host_check(dim, node_rowptr, node_colind, rowptr01, colind01, val01, xp, window_cap=1408):
    """Host evaluation of y = A01 xp through the slab layout of A01 (no device needed)."""
    node_rowptr = np.ascontiguousarray(node_rowptr, np.int64)
    node_colind = np.ascontiguousarray(node_colind, np.uint32)
    rowptr01 = np.ascontiguousarray(rowptr01, np.int64)
    colind01 = np.ascontiguousarray(colind01, np.uint32)
    val01 = np.ascontiguousarray(val01, np.float64)
    xp = np.ascontiguousarray(xp, np.float64)
    n = node_rowptr.size - 1
    y = np.zeros(dim * n)
    st = np.zeros(3, np.int64)
    rc = device_lib().nsb_gslab_host_check(dim, n, n, _p(node_rowptr, C.c_int64), _p(node_colind, C.c_uint32),
                                           window_cap, _p(rowptr01, C.c_int64), _p(colind01, C.c_uint32),
                                           _p(val01, C.c_double), _p(xp, C.c_double), _p(y, C.c_double),
                                           _p(st, C.c_int64))
    if rc != 0:
        raise DeviceError(f"nsb_gslab_host_check failed ({rc})")
    return y, dict(zip(["nnz", "padded", "max_window"], [int(v) for v in st]))


class Device:
    """One ``nsb_ctx``: the device-resident system of one NavierStokes problem."""

    def __init__(self, dim: int, device_id: int = 0):
        self.L = device_lib()
        self.dim = dim
        h = C.c_void_p()
        rc = self.L.nsb_create(dim, device_id, C.byref(h))
        if rc != 0:
            raise DeviceError(f"nsb_create failed ({rc}): no usable CUDA device; there is no CPU fallback")
        self.h = h
        self.N = 0

    def _chk(self, rc):
        if rc != 0:
            raise DeviceError(f"nsb error {rc}: {self.L.nsb_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.L.nsb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- setup ---------------------------------------------------------
    def load_problem(self, prob, quad_rule=QUAD_DEALII95, node_pattern=False):
        """Uploads mesh, dofs, patterns, Dirichlet list and obstacle faces of a
        host :class:`Problem` (already built)."""
        s = prob.sizes()
        xyz, cells = prob.array("xyz"), prob.array("cells")
        self._chk(self.L.nsb_set_mesh(self.h, s["n_verts"], _p(xyz, C.c_double), s["n_cells"], _p(cells, C.c_uint32)))
        cd = prob.array("cell_dofs")
        self._chk(self.L.nsb_set_dofs(self.h, s["n_u"], s["n_p"], _p(cd, C.c_uint32)))
        for blk, name in ((A00, "a00"), (A01, "a01"), (A10, "a10"), (S, "s")):
            if blk == A00 and node_pattern:
                rp, ci = prob.array("nodes.rowptr"), prob.array("nodes.colind")
                self._chk(self.L.nsb_set_node_pattern(self.h, rp.size - 1, _p(rp, C.c_int64), _p(ci, C.c_uint32)))
                continue
            rp, ci = prob.array(name + ".rowptr"), prob.array(name + ".colind")
            self._chk(self.L.nsb_set_pattern(self.h, blk, rp.size - 1, _p(rp, C.c_int64), _p(ci, C.c_uint32)))
        self._chk(self.L.nsb_set_quadrature(self.h, quad_rule))
        self.set_dirichlet(prob.array("bc.dofs"), prob.array("bc.values"))
        fc, fn, fm = prob.array("ff.cell"), prob.array("ff.normal"), prob.array("ff.measure")
        self._chk(self.L.nsb_set_force_faces(self.h, fc.size, _p(fc, C.c_uint32), _p(fn, C.c_double),
                                             _p(fm, C.c_double)))
        self._chk(self.L.nsb_finalize_setup(self.h))
        self.N = s["n_u"] + s["n_p"]
        self.n_u, self.n_p = s["n_u"], s["n_p"]
        return self

    @staticmethod
    def make_unique_id() -> bytes:
        """128-byte NCCL id; create on rank 0 and broadcast to the other ranks."""
        buf = C.create_string_buffer(128)
        if device_lib().nsb_comm_unique_id(buf) != 0:
            raise DeviceError("nsb_comm_unique_id failed (libnccl.so.2 not loadable?)")
        return buf.raw

    def load_local_problem(self, prob, loc, unique_id: bytes, quad_rule=QUAD_DEALII95):
        """Distributed setup of rank ``loc.rank`` of ``loc.n_parts`` from a host
        :class:`LocalProblem` (one process per GPU)."""
        s = loc.sizes()
        sz = prob.sizes()
        xyz = prob.array("xyz")
        cv, cn, cp = loc.array("cell_verts"), loc.array("cell_nodes"), loc.array("cell_pverts")
        self._chk(self.L.nsb_set_mesh(self.h, sz["n_verts"], _p(xyz, C.c_double), s["n_local_cells"],
                                      _p(cv, C.c_uint32)))
        po = loc.array("p_offset")
        self._chk(self.L.nsb_set_local_dofs(self.h, loc.rank, loc.n_parts, s["n_own"], s["n_ghost"], s["n_p"],
                                            _p(po, C.c_uint32), _p(cn, C.c_uint32), _p(cp, C.c_uint32)))
        nb, sp_, si, rp_ = (loc.array(k) for k in ("neighbors", "send_ptr", "send_idx", "recv_ptr"))
        self._chk(self.L.nsb_set_halo(self.h, nb.size, _p(nb, C.c_int32), _p(sp_, C.c_int64), _p(si, C.c_uint32),
                                      _p(rp_, C.c_int64)))
        self._chk(self.L.nsb_comm_init(self.h, loc.rank, loc.n_parts, unique_id))
        rp, ci = loc.array("fs.rowptr"), loc.array("fs.colind")
        self._chk(self.L.nsb_set_node_pattern(self.h, rp.size - 1, _p(rp, C.c_int64), _p(ci, C.c_uint32)))
        for blk, name in ((A01, "a01"), (A10, "a10"), (S, "s")):
            rp, ci = loc.array(name + ".rowptr"), loc.array(name + ".colind")
            self._chk(self.L.nsb_set_pattern(self.h, blk, rp.size - 1, _p(rp, C.c_int64), _p(ci, C.c_uint32)))
        self._chk(self.L.nsb_set_quadrature(self.h, quad_rule))
        dim = self.dim
        bn = loc.array("bc_nodes")
        dofs = (dim * bn[:, None] + np.arange(dim, dtype=np.uint32)[None, :]).astype(np.uint32).ravel()
        self.set_dirichlet(dofs, loc.array("bc_values"))
        fc, fn, fm = loc.array("ff.cell"), loc.array("ff.normal"), loc.array("ff.measure")
        self._chk(self.L.nsb_set_force_faces(self.h, fc.size, _p(fc, C.c_uint32), _p(fn, C.c_double),
                                             _p(fm, C.c_double)))
        self._chk(self.L.nsb_finalize_setup(self.h))
        self.n_u = dim * s["n_own"]
        self.n_uloc = dim * (s["n_own"] + s["n_ghost"])
        self.n_p = s["n_p"]
        self.N = self.n_uloc + self.n_p
        return self

    def set_params(self, deltat, nu):
        self._chk(self.L.nsb_set_params(self.h, deltat, nu))

    def set_bc_diag_mode(self, mode):
        self._chk(self.L.nsb_set_bc_diag_mode(self.h, mode))

    def set_solver(self, gmres_rtol=1e-6, restart=28, max_it=10000, alpha=0.5, preconditioner=PREC_ASIMPLE):
        self._chk(self.L.nsb_set_solver(self.h, gmres_rtol, restart, max_it, alpha, preconditioner))

    def set_inner(self, sweeps_F, eig_ratio_F, sweeps_S, eig_ratio_S):
        self._chk(self.L.nsb_set_inner(self.h, sweeps_F, eig_ratio_F, sweeps_S, eig_ratio_S))

    def gather_velocity(self, node_offsets):
        """All ranks' owned velocity dofs in the distributed numbering (collective)."""
        no = np.ascontiguousarray(node_offsets, np.uint32)
        out = np.empty(self.dim * int(no[-1]), np.float64)
        self._chk(self.L.nsb_gather_velocity(self.h, _p(no, C.c_uint32), _p(out, C.c_double)))
        return out

    def set_schur_solver(self, mode=1, smoother_sweeps=0, theta=0.0, omega=0.0, cycles=0):
        self._chk(self.L.nsb_set_schur_solver(self.h, mode, smoother_sweeps, theta, omega, cycles))

    def set_schur_strength(self, measure=0, theta=0.35, decay=1.0):
        self.L.nsb_set_schur_strength.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
        self._chk(self.L.nsb_set_schur_strength(self.h, measure, theta, decay))

    def set_dirichlet(self, dofs, values):
        dofs = np.ascontiguousarray(dofs, np.uint32)
        values = np.ascontiguousarray(values, np.float64)
        self._chk(self.L.nsb_set_dirichlet(self.h, dofs.size, _p(dofs, C.c_uint32), _p(values, C.c_double)))

    def scale_dirichlet(self, factor):
        self._chk(self.L.nsb_scale_dirichlet(self.h, factor))

    # ---- state -----------------------------------------------------------
    def set_solution(self, x):
        x = np.ascontiguousarray(x, np.float64)
        assert x.size == self.N
        self._chk(self.L.nsb_set_solution(self.h, _p(x, C.c_double)))

    def solution(self, out=None):
        out = np.empty(self.N, np.float64) if out is None else out
        self._chk(self.L.nsb_get_solution(self.h, _p(out, C.c_double)))
        return out

    # ---- hot path -----------------------------------------------------------
    def assemble(self, time):
        self._chk(self.L.nsb_assemble(self.h, time))

    def solve_time_step(self):
        it, tp, ts = C.c_int(), C.c_double(), C.c_double()
        self._chk(self.L.nsb_solve_time_step(self.h, C.byref(it), C.byref(tp), C.byref(ts)))
        return it.value, tp.value, ts.value

    def compute_forces(self, u_mean):
        out = np.empty(4, np.float64)
        self._chk(self.L.nsb_compute_forces(self.h, u_mean, _p(out, C.c_double)))
        return out

    # ---- taps -----------------------------------------------------------------
    def nnz(self, block):
        return int(self.L.nsb_nnz(self.h, block))

    def values(self, block):
        out = np.empty(self.nnz(block), np.float64)
        self._chk(self.L.nsb_get_matrix_values(self.h, block, _p(out, C.c_double)))
        return out

    def pattern(self, block):
        n_rows = self.n_p if block in (A10, S) else self.n_u
        rp, ci = np.empty(n_rows + 1, np.int64), np.empty(self.nnz(block), np.uint32)
        self._chk(self.L.nsb_get_pattern(self.h, block, _p(rp, C.c_int64), _p(ci, C.c_uint32)))
        return rp, ci

    def rhs(self):
        out = np.empty(self.N, np.float64)
        self._chk(self.L.nsb_get_rhs(self.h, _p(out, C.c_double)))
        return out

    def timer_start(self):
        self._chk(self.L.nsb_timer_start(self.h))

    def timer_stop(self):
        """Milliseconds between timer_start and now, measured with CUDA events on the context's stream."""
        ms = C.c_double()
        self._chk(self.L.nsb_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def lumped_mass_inv(self, n_u):
        """deltat_lumped_mass_inv of the reference, velocity block (n_u values)."""
        out = np.empty(n_u, np.float64)
        self._chk(self.L.nsb_get_lumped_mass_inv(self.h, _p(out, C.c_double)))
        return out

    def vmult(self, x):
        x = np.ascontiguousarray(x, np.float64)
        y = np.empty(self.N, np.float64)
        self._chk(self.L.nsb_vmult(self.h, _p(x, C.c_double), _p(y, C.c_double)))
        return y

    def bench_kernel(self, which, reps):
        ms = C.c_double()
        self._chk(self.L.nsb_bench_kernel(self.h, which, reps, C.byref(ms)))
        return ms.value

    def launch_count(self):
        return int(self.L.nsb_launch_count(self.h))

    def timers(self):
        out = np.empty(4, np.float64)
        self._chk(self.L.nsb_timers(self.h, _p(out, C.c_double)))
        return out

    def inner_params(self):
        out = np.zeros(4)
        self.L.nsb_inner_params.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        self._chk(self.L.nsb_inner_params(self.h, _p(out, C.c_double)))
        return dict(zip(["degree_F", "ratio_F", "lambda_max_F", "imag_F"], [float(v) for v in out]))

    def info(self):
        out = (C.c_int64 * 20)()
        self._chk(self.L.nsb_info(self.h, out))
        keys = ["n_u", "n_p", "n_cells", "nnz_a00", "nnz_a01", "nnz_a10", "nnz_s", "n_q", "device_bytes", "sweeps_F",
                "sweeps_S", "schur_mode", "schur_levels", "slab_entries", "slab_window_total", "slab_count",
                "gslab_entries", "gslab_window_total", "reorth_passes", "exchange_mode"]
        return dict(zip(keys, [int(v) for v in out]))
