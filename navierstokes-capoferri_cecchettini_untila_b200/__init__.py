"""ctypes bindings over the two C ABIs of this repository.

* ``include/nsb_host.h`` -> :class:`Problem` (mesh, Taylor-Hood numbering,
  block sparsity pattern, Dirichlet and obstacle-face lists; pure C++).
* ``include/nsb.h``      -> :class:`Device` (the sm_100a hot path: assembly,
  block-preconditioned GMRES, forces).

The product is the C++/CUDA behind those headers (host facade:
``host/NavierStokes.hpp``); this module exists so that the parity tests and
``bench.py`` can drive the same entry points a C++ caller uses.  There is no
CPU fallback: :class:`Device` raises when ``libnsb.so`` is missing or no CUDA
device is present.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)

_c_f64p = C.POINTER(C.c_double)
_c_u32p = C.POINTER(C.c_uint32)
_c_i32p = C.POINTER(C.c_int32)
_c_i64p = C.POINTER(C.c_int64)


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def _load(name):
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        raise RuntimeError(
            f"{name} is not built: run `make` (or `python -c 'import __graft_entry__ as g; g.build()'`) in {ROOT}"
        )
    return C.CDLL(path, mode=C.RTLD_GLOBAL)


# --------------------------------------------------------------------------
# host library
# --------------------------------------------------------------------------
_host = None


def host_lib():
    global _host
    if _host is None:
        L = _load("libnsb_host.so")
        L.nsh_last_error.restype = C.c_char_p
        L.nsh_problem_generate.argtypes = [C.c_char_p, C.c_double, C.POINTER(C.c_void_p)]
        L.nsh_problem_read.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
        L.nsh_problem_generate_airfoil.argtypes = [C.c_char_p, C.c_int] + [C.c_double] * 7 + [C.POINTER(C.c_void_p)]
        L.nsh_problem_from_arrays.argtypes = [C.c_int, C.c_int64, _c_f64p, C.c_int64, _c_u32p, C.c_int64, _c_u32p,
                                              _c_i32p, C.POINTER(C.c_void_p)]
        L.nsh_problem_write_msh.argtypes = [C.c_void_p, C.c_char_p]
        L.nsh_problem_free.argtypes = [C.c_void_p]
        L.nsh_build_space.argtypes = [C.c_void_p, C.c_int]
        L.nsh_set_inlet.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int]
        L.nsh_build_boundary.argtypes = [C.c_void_p]
        L.nsh_mean_velocity.argtypes = [C.c_void_p, C.c_double]
        L.nsh_mean_velocity.restype = C.c_double
        L.nsh_inlet_time_factor.argtypes = [C.c_void_p, C.c_double]
        L.nsh_inlet_time_factor.restype = C.c_double
        L.nsh_sizes.argtypes = [C.c_void_p, _c_i64p]
        L.nsh_array.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p), _c_i64p, C.POINTER(C.c_int)]
        L.nsh_partition.argtypes = [C.c_void_p, C.c_int]
        L.nsh_localize.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.nsh_local_free.argtypes = [C.c_void_p]
        L.nsh_local_sizes.argtypes = [C.c_void_p, _c_i64p]
        L.nsh_local_array.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p), _c_i64p, C.POINTER(C.c_int)]
        _host = L
    return _host


class HostError(RuntimeError):
    pass


_DT = {("f", 8): np.float64, ("u", 4): np.uint32, ("i", 4): np.int32, ("i", 8): np.int64}
_KIND = {"xyz": "f", "node_xyz": "f", "bc.values": "f", "ff.normal": "f", "ff.measure": "f", "bids": "i",
         "part.cell": "i"}

INLET_PARABOLIC, INLET_UNIFORM = 0, 1


class Problem:
    """Mesh + P2/P1 space + boundary lists (``nsh_problem``)."""

    def __init__(self, handle):
        self._h = handle
        self._L = host_lib()

    @staticmethod
    def _chk(rc):
        if rc != 0:
            raise HostError(host_lib().nsh_last_error().decode())

    @classmethod
    def generate(cls, name: str, h: float) -> "Problem":
        out = C.c_void_p()
        cls._chk(host_lib().nsh_problem_generate(name.encode(), h, C.byref(out)))
        return cls(out)

    @classmethod
    def generate_airfoil(cls, h: float, dat_path: str = "", naca4: int = 2408, chord: float = 0.4, aoa_deg: float = 0.0,
                         box=(2.2, 1.0), centre=(0.4, 0.5)) -> "Problem":
        """Airfoil mesh the way the reference prepares it (mesh/test.py + run_test.sh): contour from a .dat file
        (mesh/naca.dat layout) or the NACA 4-digit family, chord scaling, angle of attack (degrees, nose down for
        positive angles as in test.py), defaults = test.py's box and centre."""
        out = C.c_void_p()
        cls._chk(host_lib().nsh_problem_generate_airfoil(dat_path.encode(), naca4, chord, aoa_deg, box[0], box[1],
                                                         centre[0], centre[1], h, C.byref(out)))
        return cls(out)

    @classmethod
    def read_msh(cls, path: str, dim: int) -> "Problem":
        out = C.c_void_p()
        cls._chk(host_lib().nsh_problem_read(path.encode(), dim, C.byref(out)))
        return cls(out)

    @classmethod
    def from_arrays(cls, dim, xyz, cells, bfaces, bids) -> "Problem":
        xyz = np.ascontiguousarray(xyz, np.float64)
        cells = np.ascontiguousarray(cells, np.uint32)
        bfaces = np.ascontiguousarray(bfaces, np.uint32)
        bids = np.ascontiguousarray(bids, np.int32)
        out = C.c_void_p()
        cls._chk(host_lib().nsh_problem_from_arrays(dim, xyz.size // dim, _ptr(xyz, C.c_double),
                                                    cells.size // (dim + 1), _ptr(cells, C.c_uint32), bids.size,
                                                    _ptr(bfaces, C.c_uint32), _ptr(bids, C.c_int32), C.byref(out)))
        return cls(out)

    def write_msh(self, path):
        self._chk(self._L.nsh_problem_write_msh(self._h, path.encode()))

    def build(self, inlet=(INLET_PARABOLIC, 0.3, 0.41, 0), expand_a00=True) -> "Problem":
        self._chk(self._L.nsh_build_space(self._h, 1 if expand_a00 else 0))
        self._chk(self._L.nsh_set_inlet(self._h, int(inlet[0]), float(inlet[1]), float(inlet[2]), int(inlet[3])))
        self._chk(self._L.nsh_build_boundary(self._h))
        return self

    def partition(self, n_parts):
        self._chk(self._L.nsh_partition(self._h, n_parts))
        return self.array("part.cell")

    def mean_velocity(self, t=0.0):
        return self._L.nsh_mean_velocity(self._h, t)

    def inlet_time_factor(self, t):
        return self._L.nsh_inlet_time_factor(self._h, t)

    def sizes(self):
        out = (C.c_int64 * 10)()
        self._chk(self._L.nsh_sizes(self._h, out))
        keys = ["dim", "n_verts", "n_cells", "n_bfaces", "n_nodes", "n_u", "n_p", "dofs_per_cell", "n_bc",
                "n_force_faces"]
        return dict(zip(keys, [int(x) for x in out]))

    def array(self, name) -> np.ndarray:
        """Borrowed view of a named array (valid while this object lives)."""
        data, count, eb = C.c_void_p(), C.c_int64(), C.c_int()
        self._chk(self._L.nsh_array(self._h, name.encode(), C.byref(data), C.byref(count), C.byref(eb)))
        kind = _KIND.get(name, "i" if name.endswith("rowptr") else "u")
        dt = _DT[(kind, eb.value)]
        if count.value == 0:
            return np.zeros(0, dt)
        buf = (C.c_char * (count.value * eb.value)).from_address(data.value)
        a = np.frombuffer(buf, dtype=dt)
        a.flags.writeable = False
        return a

    def __del__(self):
        try:
            if self._h:
                self._L.nsh_problem_free(self._h)
                self._h = None
        except Exception:
            pass


class LocalProblem:
    """Rank-local view of a partitioned :class:`Problem` (``nsh_local``)."""
    _F64 = {"bc_values", "ff.normal", "ff.measure"}
    _I32 = {"neighbors"}
    _I64 = {"send_ptr", "recv_ptr"}

    def __init__(self, prob: Problem, n_parts: int, rank: int):
        self._L = host_lib()
        self._prob = prob  # keep the parent alive
        out = C.c_void_p()
        Problem._chk(self._L.nsh_localize(prob._h, n_parts, rank, C.byref(out)))
        self._h = out
        self.rank, self.n_parts = rank, n_parts

    def sizes(self):
        out = (C.c_int64 * 11)()
        Problem._chk(self._L.nsh_local_sizes(self._h, out))
        keys = ["n_own", "n_ghost", "n_p", "n_p_own", "p_offset", "n_local_cells", "n_neighbors", "n_bc_nodes",
                "n_force_faces", "n_nodes_global", "node_offset"]
        return dict(zip(keys, [int(x) for x in out]))

    def array(self, name) -> np.ndarray:
        data, count, eb = C.c_void_p(), C.c_int64(), C.c_int()
        Problem._chk(self._L.nsh_local_array(self._h, name.encode(), C.byref(data), C.byref(count), C.byref(eb)))
        if name in self._F64:
            dt = np.float64
        elif name in self._I32:
            dt = np.int32
        elif name in self._I64 or name.endswith("rowptr"):
            dt = np.int64
        else:
            dt = np.uint32
        if count.value == 0:
            return np.zeros(0, dt)
        buf = (C.c_char * (count.value * eb.value)).from_address(data.value)
        a = np.frombuffer(buf, dtype=dt)
        a.flags.writeable = False
        return a

    # ---- canonical (single-rank) numbering <-> this rank's local vector [u owned | u ghost | p] ----
    def _maps(self):
        if not hasattr(self, "_canon_nodes"):
            s = self.sizes()
            inv = np.empty(s["n_nodes_global"], np.int64)
            inv[self.array("node_perm").astype(np.int64)] = np.arange(s["n_nodes_global"])
            dist = np.concatenate([np.arange(s["node_offset"], s["node_offset"] + s["n_own"], dtype=np.int64),
                                   self.array("ghost_dist").astype(np.int64)])
            self._canon_nodes = inv[dist]  # canonical node of every local node
            self._n_own = s["n_own"]
        return self._canon_nodes

    def to_local(self, x_global, dim):
        """Local vector of a canonical global vector (velocity incl. ghosts, pressure replicated)."""
        cn = self._maps()
        n_u = dim * self.sizes()["n_nodes_global"]
        u = x_global[:n_u].reshape(-1, dim)[cn].ravel()
        p = np.empty(self.sizes()["n_p"])
        p[self.array("p_perm").astype(np.int64)] = x_global[n_u:]
        return np.concatenate([u, p])

    def owned_to_global(self, x_local, dim, out):
        """Writes this rank's owned velocity dofs (and the replicated pressure) of a local
        vector into the canonical global vector ``out``."""
        cn = self._maps()
        n_u = dim * self.sizes()["n_nodes_global"]
        n_loc = cn.size
        ug = out[:n_u].reshape(-1, dim)
        ug[cn[: self._n_own]] = x_local[: dim * n_loc].reshape(-1, dim)[: self._n_own]
        out[n_u:] = x_local[dim * n_loc:][self.array("p_perm").astype(np.int64)]
        return out

    def __del__(self):
        try:
            if self._h:
                self._L.nsh_local_free(self._h)
                self._h = None
        except Exception:
            pass


from .device import Device, DeviceError, device_lib  # noqa: E402,F401
