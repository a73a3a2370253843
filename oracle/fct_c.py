"""Oracle: ctypes front end of the C/OpenMP twin (oracle/fct_c.c).  TEST INFRASTRUCTURE (see oracle/__init__.py).

`CDriftProblem(n)` is the drift-control advection problem of BASELINE config 2/5 (advection_solidbody_FCT_PDECO_alltime.py)
on the n x n "right" RectangleMesh: mesh, DoF numbering and pattern, mass matrix, per-step drift
operator and the FCT state loop, all in C with OpenMP so that bench.py's CPU baseline can run the full-size workload on
every host core.  Pinned against the numpy oracle in tests/test_oracle_c.py."""
import ctypes as C
import os

import numpy as np

from . import build_c

_lib = None


def lib(native=False):
    """native=True (bench.py's CPU arm): try a -march=native rebuild on THIS host first (oracle/_fct_c_native.so); the
    portable object built in the build container is the fallback"""
    global _lib
    if _lib is None:
        path = None
        if native:
            try:
                path = build_c.build(native=True)
            except Exception:  # noqa: BLE001
                path = None
        if path is None:
            path = build_c.build()
        L = C.CDLL(path)
        L.fctc_threads.restype = C.c_int
        L.fctc_tpos.restype = C.c_int
        L.fctc_step.restype = C.c_int
        L.fctc_norm_sq_M.restype = C.c_double
        L.fctc_set_threads.restype = None
        L._path = path
        _lib = L
    return _lib


def use_all_host_threads():
    """set the OpenMP thread count to the cores this process may run on, whatever OMP_NUM_THREADS says (torchrun exports
    OMP_NUM_THREADS=1 to its ranks); returns the count"""
    try:
        t = len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        t = os.cpu_count() or 1
    lib().fctc_set_threads(C.c_int32(t))
    return lib().fctc_threads()


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class CDriftProblem:
    def __init__(self, n, a1=0.0, a2=1.0):
        L = lib()
        self.n = int(n)
        N = self.n + 1
        self.nodes = N * N
        self.ncells = 2 * self.n * self.n
        self.nnz = self.nodes + 2 * (2 * N * (N - 1) + (N - 1) ** 2)
        i32, f64 = np.int32, np.float64
        self.vertex_to_dof = np.empty(self.nodes, i32)
        self.cells = np.empty((self.ncells, 3), i32)
        self.dof_xy = np.empty((self.nodes, 2), f64)
        self.rowptr = np.empty(self.nodes + 1, i32)
        self.colidx = np.empty(self.nnz, i32)
        L.fctc_mesh(C.c_int32(self.n), C.c_double(a1), C.c_double(a2), _p(self.vertex_to_dof), _p(self.cells), _p(self.dof_xy),
                    _p(self.rowptr), _p(self.colidx))
        assert int(self.rowptr[-1]) == self.nnz
        self.tpos = np.empty(self.nnz, i32)
        self.diagpos = np.empty(self.nodes, i32)
        bad = L.fctc_tpos(C.c_int32(self.nodes), _p(self.rowptr), _p(self.colidx), _p(self.tpos), _p(self.diagpos))
        if bad:
            raise ValueError("pattern is not structurally symmetric / misses a diagonal")
        self.inc_ptr = np.empty(self.nodes + 1, i32)
        self.inc_idx = np.empty(6 * self.ncells, i32)
        L.fctc_incidence(C.c_int32(self.nodes), C.c_int64(self.ncells), _p(self.cells), _p(self.inc_ptr), _p(self.inc_idx))
        self.M = np.empty(self.nnz, f64)
        self._assemble(0, None, 0.0, 0.0, 1.0, self.M)
        self.ML = np.empty(self.nodes, f64)
        self.Md = np.empty(self.nodes, f64)
        L.fctc_row_lump(C.c_int32(self.nodes), _p(self.rowptr), _p(self.diagpos), _p(self.M), _p(self.ML), _p(self.Md))
        self._A = np.empty(self.nnz, f64)
        self._L = np.empty(self.nnz, f64)
        self._D = np.empty(self.nnz, f64)
        self._vec = np.empty(10 * self.nodes, f64)

    def threads(self):
        return lib().fctc_threads()

    def norm_sq_M(self, x, t=None):
        """(x - t)^T M (x - t)"""
        x = np.ascontiguousarray(x, dtype=np.float64)
        t = None if t is None else np.ascontiguousarray(t, dtype=np.float64)
        return float(lib().fctc_norm_sq_M(C.c_int32(self.nodes), _p(self.rowptr), _p(self.colidx), _p(self.M), _p(x),
                                          _p(t) if t is not None else None))

    def norm_sq_Q(self, phi, num_steps, dt, target=None):
        """helpers.py:330-360 (trapezoid in time) of phi - target"""
        phi = np.asarray(phi).reshape(num_steps + 1, self.nodes)
        tg = None if target is None else np.asarray(target).reshape(num_steps + 1, self.nodes)
        s = 0.0
        for k in range(num_steps + 1):
            w = 0.5 if k in (0, num_steps) else 1.0
            s += w * self.norm_sq_M(phi[k], None if tg is None else tg[k])
        return s * dt

    def _assemble(self, kind, c, bx, by, scale, out):
        lib().fctc_assemble(C.c_int32(kind), C.c_int32(self.nodes), _p(self.rowptr), _p(self.colidx), _p(self.inc_ptr),
                            _p(self.inc_idx), _p(self.dof_xy), _p(c) if c is not None else None, C.c_double(bx), C.c_double(by),
                            C.c_double(scale), _p(out))

    def drift_operator(self, c, bx=1.0, by=1.0, scale=1.0):
        """scale * [(b.grad c) u v + (b.grad v) c u] on the pattern (advection_solidbody_FCT_PDECO_alltime.py:222-226)"""
        out = np.empty(self.nnz, np.float64)
        self._assemble(1, np.ascontiguousarray(c, dtype=np.float64), bx, by, scale, out)
        return out

    def step(self, A, rhs, un, dt, rtol=1e-14, maxit=200):
        """one FCT step, FCT_alg_ref sign convention (helpers.py:1715-1872); returns (u_np1, jacobi sweeps)"""
        out = np.empty(self.nodes, np.float64)
        its = lib().fctc_step(C.c_int32(self.nodes), _p(self.rowptr), _p(self.colidx), _p(self.tpos), _p(self.diagpos), _p(A),
                              _p(rhs) if rhs is not None else None, _p(un), C.c_double(dt), _p(self.M), _p(self.ML),
                              _p(self.Md), _p(self._L), _p(self._D), _p(self._vec), C.c_double(rtol), C.c_int32(maxit), _p(out))
        return out, its

    def state(self, c_traj, u0, num_steps, dt, bx=1.0, by=1.0):
        """advection_solidbody_FCT_PDECO_alltime.py:210-228: u[i] = FCT_alg(A_u(c[i]), 0, u[i-1]) = FCT_alg_ref(-A_u, ...)"""
        c_traj = np.ascontiguousarray(c_traj, dtype=np.float64).reshape(num_steps + 1, self.nodes)
        u = np.zeros((num_steps + 1, self.nodes))
        u[0] = u0
        sweeps = 0
        for i in range(1, num_steps + 1):
            self._assemble(1, c_traj[i], bx, by, -1.0, self._A)
            u[i], its = self.step(self._A, None, np.ascontiguousarray(u[i - 1]), dt)
            sweeps += its
        return u, sweeps
