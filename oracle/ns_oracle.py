"""ctypes wrapper of the CPU oracle (oracle/libns_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Never by the product.
Parity unpinned -- see oracle/ns_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
QUAD_DEALII93, QUAD_DEALII95 = 0, 1
INLET_PARABOLIC, INLET_UNIFORM = 0, 1

_lib = None


def build():
    """Compiles the oracle (g++) if the shared object is missing or stale."""
    so = os.path.join(_HERE, "libns_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("ns_oracle.cpp", "ns_baseline.cpp", "ns_oracle.h", "ns_oracle_internal.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", os.path.dirname(_HERE), "oracle"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        p = C.c_void_p
        f64p, u32p, i32p, i64p = (C.POINTER(t) for t in (C.c_double, C.c_uint32, C.c_int32, C.c_int64))
        L.nso_create.restype = p
        L.nso_create.argtypes = [C.c_int, C.c_int64, f64p, C.c_int64, u32p, C.c_int64, u32p, i32p, C.c_int]
        L.nso_destroy.argtypes = [p]
        L.nso_sizes.argtypes = [p, i64p]
        L.nso_get_cell_dofs.argtypes = [p, u32p]
        L.nso_get_pattern.argtypes = [p, C.c_int, i64p, u32p]
        L.nso_get_values.argtypes = [p, C.c_int, f64p]
        L.nso_get_rhs.argtypes = [p, f64p]
        L.nso_get_lumped.argtypes = [p, f64p]
        L.nso_get_bc.argtypes = [p, u32p, f64p]
        L.nso_set_params.argtypes = [p, C.c_double, C.c_double]
        L.nso_set_bc_diag_mode.argtypes = [p, C.c_int]
        L.nso_set_inlet.argtypes = [p, C.c_int, C.c_double, C.c_double, C.c_int]
        L.nso_mean_velocity.argtypes = [p, C.c_double]
        L.nso_mean_velocity.restype = C.c_double
        L.nso_set_re_number.argtypes = [p, C.c_int]
        L.nso_set_re_number.restype = C.c_double
        L.nso_set_solution.argtypes = [p, f64p]
        L.nso_get_solution.argtypes = [p, f64p]
        L.nso_set_solver.argtypes = [p, C.c_double, C.c_int, C.c_int, C.c_double]
        L.nso_assemble.argtypes = [p, C.c_double]
        L.nso_solve_time_step.argtypes = [p, C.POINTER(C.c_int), f64p, f64p]
        L.nso_solve_time_step.restype = C.c_int
        L.nso_compute_forces.argtypes = [p, C.c_double, f64p]
        L.nso_vmult.argtypes = [p, f64p, f64p]
        L.nso_set_threads.argtypes = [p, C.c_int]
        L.nso_baseline_partition.argtypes = [p, C.c_int, i32p]
        L.nso_baseline_assemble.argtypes = [p, C.c_double]
        L.nso_baseline_solve_time_step.argtypes = [p, C.POINTER(C.c_int), f64p, f64p]
        L.nso_baseline_solve_time_step.restype = C.c_int
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class Oracle:
    BLOCKS = {"a00": 0, "a01": 1, "a10": 2, "s": 3}

    def __init__(self, dim, xyz, cells, bfaces, bids, quad_rule=QUAD_DEALII95):
        self.L = lib()
        self.dim = dim
        xyz = np.ascontiguousarray(xyz, np.float64)
        cells = np.ascontiguousarray(cells, np.uint32)
        bfaces = np.ascontiguousarray(bfaces, np.uint32)
        bids = np.ascontiguousarray(bids, np.int32)
        self.n_cells = cells.size // (dim + 1)
        self.h = C.c_void_p(self.L.nso_create(dim, xyz.size // dim, _p(xyz, C.c_double), self.n_cells,
                                              _p(cells, C.c_uint32), bids.size, _p(bfaces, C.c_uint32),
                                              _p(bids, C.c_int32), quad_rule))
        s = self.sizes()
        self.n_u, self.n_p, self.dpc = s["n_u"], s["n_p"], s["dpc"]
        self.N = self.n_u + self.n_p

    def __del__(self):
        try:
            self.L.nso_destroy(self.h)
        except Exception:
            pass

    def sizes(self):
        out = (C.c_int64 * 10)()
        self.L.nso_sizes(self.h, out)
        keys = ["n_u", "n_p", "nnz_a00", "nnz_a01", "nnz_a10", "nnz_s", "dpc", "n_q", "n_q_face", "n_bc"]
        return dict(zip(keys, [int(x) for x in out]))

    def cell_dofs(self):
        out = np.empty(self.n_cells * self.dpc, np.uint32)
        self.L.nso_get_cell_dofs(self.h, _p(out, C.c_uint32))
        return out

    def _rows(self, block):
        return self.n_p if block in ("a10", "s") else self.n_u

    def pattern(self, block):
        s = self.sizes()
        rowptr = np.empty(self._rows(block) + 1, np.int64)
        colind = np.empty(s["nnz_" + block], np.uint32)
        self.L.nso_get_pattern(self.h, self.BLOCKS[block], _p(rowptr, C.c_int64), _p(colind, C.c_uint32))
        return rowptr, colind

    def values(self, block):
        out = np.empty(self.sizes()["nnz_" + block], np.float64)
        self.L.nso_get_values(self.h, self.BLOCKS[block], _p(out, C.c_double))
        return out

    def rhs(self):
        out = np.empty(self.N, np.float64)
        self.L.nso_get_rhs(self.h, _p(out, C.c_double))
        return out

    def lumped(self):
        out = np.empty(self.N, np.float64)
        self.L.nso_get_lumped(self.h, _p(out, C.c_double))
        return out

    def bc(self):
        n = self.sizes()["n_bc"]
        d, v = np.empty(n, np.uint32), np.empty(n, np.float64)
        self.L.nso_get_bc(self.h, _p(d, C.c_uint32), _p(v, C.c_double))
        return d, v

    def set_params(self, deltat, nu):
        self.L.nso_set_params(self.h, deltat, nu)

    def set_bc_diag_mode(self, mode):
        self.L.nso_set_bc_diag_mode(self.h, mode)

    def set_inlet(self, kind, U_m, H=0.41, time_sin=0):
        self.L.nso_set_inlet(self.h, kind, U_m, H, time_sin)

    def mean_velocity(self, t=0.0):
        return self.L.nso_mean_velocity(self.h, t)

    def set_re_number(self, Re):
        return self.L.nso_set_re_number(self.h, Re)

    def set_solution(self, x):
        x = np.ascontiguousarray(x, np.float64)
        assert x.size == self.N
        self.L.nso_set_solution(self.h, _p(x, C.c_double))

    def solution(self):
        out = np.empty(self.N, np.float64)
        self.L.nso_get_solution(self.h, _p(out, C.c_double))
        return out

    def set_solver(self, outer_rtol=1e-6, n_tmp_vectors=30, max_it=10000, inner_rtol=1e-2):
        self.L.nso_set_solver(self.h, outer_rtol, n_tmp_vectors, max_it, inner_rtol)

    def set_threads(self, n):
        self.L.nso_set_threads(self.h, n)

    def assemble(self, time):
        self.L.nso_assemble(self.h, time)

    def solve_time_step(self):
        it, tp, ts = C.c_int(), C.c_double(), C.c_double()
        rc = self.L.nso_solve_time_step(self.h, C.byref(it), C.byref(tp), C.byref(ts))
        return rc, it.value, tp.value, ts.value

    def compute_forces(self, time=0.0):
        out = np.empty(4, np.float64)
        self.L.nso_compute_forces(self.h, time, _p(out, C.c_double))
        return out

    def vmult(self, x):
        x = np.ascontiguousarray(x, np.float64)
        y = np.empty(self.N, np.float64)
        self.L.nso_vmult(self.h, _p(x, C.c_double), _p(y, C.c_double))
        return y

    # ---- CPU baseline mode (bench.py only): the algorithm as `mpirun -n P` runs it ----
    def baseline_partition(self, n_parts, cell_part):
        cp = np.ascontiguousarray(cell_part, np.int32)
        assert cp.size == self.n_cells
        if self.L.nso_baseline_partition(self.h, n_parts, _p(cp, C.c_int32)) != 0:
            raise ValueError("bad partition")

    def baseline_assemble(self, time):
        self.L.nso_baseline_assemble(self.h, time)

    def baseline_solve_time_step(self):
        it, tp, ts = C.c_int(), C.c_double(), C.c_double()
        rc = self.L.nso_baseline_solve_time_step(self.h, C.byref(it), C.byref(tp), C.byref(ts))
        return rc, it.value, tp.value, ts.value

    def scipy_blocks(self):
        import scipy.sparse as sp
        out = {}
        for b in ("a00", "a01", "a10", "s"):
            rp, ci = self.pattern(b)
            ncols = self.n_p if b in ("a01", "s") else self.n_u
            out[b] = sp.csr_matrix((self.values(b), ci.astype(np.int64), rp), shape=(self._rows(b), ncols))
        return out
