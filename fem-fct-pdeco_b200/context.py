"""Python handle on a libfctpdeco context: fixed CSR pattern + static matrices + device buffers.

Host logic only (numpy); all arithmetic happens in the CUDA library.
"""
import ctypes as C
import time

import numpy as np

from . import _lib
from ._lib import check, lib
from .pattern import HostPattern


def _hp(a):
    """host pointer of a C-contiguous numpy array"""
    return a.ctypes.data_as(C.c_void_p)


class DeviceArray:
    """fp64 device buffer owned by a context (thin wrapper over fct_malloc / fct_free)."""

    def __init__(self, ctx, size, dtype=np.float64):
        self.ctx = ctx
        self.size = int(size)
        self.dtype = np.dtype(dtype)
        ptr = C.c_void_p()
        check(lib.fct_malloc(ctx.handle, C.byref(ptr), self.size * self.dtype.itemsize))
        self.ptr = ptr.value
        self._owner = True

    @classmethod
    def view(cls, ctx, ptr, size, dtype=np.float64):
        self = cls.__new__(cls)
        self.ctx, self.ptr, self.size, self.dtype, self._owner = ctx, int(ptr), int(size), np.dtype(dtype), False
        return self

    def slice(self, start, size):
        """view of elements [start, start+size)"""
        assert 0 <= start and start + size <= self.size
        return DeviceArray.view(self.ctx, self.ptr + start * self.dtype.itemsize, size, self.dtype)

    def upload(self, host):
        host = np.ascontiguousarray(host, dtype=self.dtype).ravel()
        if host.size != self.size:
            raise ValueError(f"upload: expected {self.size} elements, got {host.size}")
        t0 = self.ctx._xfer_begin()
        check(lib.fct_h2d(self.ctx.handle, self.ptr, _hp(host), host.nbytes))
        check(lib.fct_ctx_sync(self.ctx.handle))      # the host array may be a temporary
        self.ctx._xfer_end(t0, host.nbytes, 0)
        return self

    def download(self, out=None):
        if out is None:
            out = np.empty(self.size, dtype=self.dtype)
        assert out.flags.c_contiguous and out.size == self.size and out.dtype == self.dtype
        t0 = self.ctx._xfer_begin()
        check(lib.fct_d2h(self.ctx.handle, _hp(out), self.ptr, out.nbytes))
        self.ctx._xfer_end(t0, 0, out.nbytes)
        return out

    def free(self):
        if self._owner and self.ptr and self.ctx.handle:
            lib.fct_free(self.ctx.handle, self.ptr)
        self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class FctContext:
    """One fixed P1 CSR pattern on one GPU.

    rowptr/colidx: host int32 arrays (columns ascending, diagonal present, structurally symmetric).
    """

    def __init__(self, rowptr, colidx, device=0, row_begin=0, row_end=None):
        if _lib.device_count() == 0:
            raise _lib.FctError("no CUDA device visible: fem-fct-pdeco_b200 has no CPU fallback")
        self.pattern = HostPattern(rowptr, colidx)
        self.rowptr, self.colidx = self.pattern.rowptr, self.pattern.colidx
        self.n, self.nnz = self.pattern.n, self.pattern.nnz
        self.row_begin = int(row_begin)
        self.row_end = self.n if row_end is None else int(row_end)
        h = C.c_void_p()
        check(lib.fct_ctx_create(C.byref(h), int(device), self.n, _hp(self.rowptr), _hp(self.colidx),
                                 self.row_begin, self.row_end))
        self.handle = h
        self.device = int(device)
        self.has_mesh = False
        self.has_mass = False

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "handle", None):
            lib.fct_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(lib.fct_ctx_sync(self.handle))

    # -- host<->device traffic of the numpy-facing API (DeviceArray.upload / download) ------------
    def _xfer_begin(self):
        """with accounting on (xfer_reset), a copy is timed on its own: wait for the queued kernels first"""
        if not getattr(self, "_xfer_on", False):
            return None
        check(lib.fct_ctx_sync(self.handle))
        return time.perf_counter()

    def _xfer_end(self, t0, h2d, d2h):
        if t0 is None:
            return
        x = self._xfer
        x["seconds"] += time.perf_counter() - t0
        x["h2d_bytes"] += h2d
        x["d2h_bytes"] += d2h

    def xfer_reset(self, on=True):
        """start (or stop) accounting; returns the counters collected so far"""
        old = getattr(self, "_xfer", None)
        self._xfer = {"seconds": 0.0, "h2d_bytes": 0, "d2h_bytes": 0}
        self._xfer_on = bool(on)
        return old

    def set_stream(self, cuda_stream):
        check(lib.fct_ctx_set_stream(self.handle, C.c_void_p(cuda_stream)))

    def launch_count(self):
        c = C.c_int64()
        check(lib.fct_launch_count(self.handle, C.byref(c)))
        return c.value

    def exchange_count(self):
        c = C.c_int64()
        check(lib.fct_exchange_count(self.handle, C.byref(c)))
        return c.value

    # -- buffers ----------------------------------------------------------------------------
    def empty(self, size):
        return DeviceArray(self, size)

    def array(self, host):
        host = np.ascontiguousarray(host, dtype=np.float64)
        return DeviceArray(self, host.size).upload(host)

    def pinned(self, size):
        """page-locked host fp64 array (numpy view; keep the context alive while using it)"""
        ptr = C.c_void_p()
        check(lib.fct_host_alloc(self.handle, C.byref(ptr), int(size) * 8))
        buf = (C.c_double * int(size)).from_address(ptr.value)
        arr = np.frombuffer(buf, dtype=np.float64)
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(ptr)
        return arr

    def static(self):
        """(M, ML, Mdiag, K) device views"""
        ptrs = [C.c_void_p() for _ in range(4)]
        check(lib.fct_ctx_static_dev(self.handle, *[C.byref(p) for p in ptrs]))
        return (DeviceArray.view(self, ptrs[0].value, self.nnz), DeviceArray.view(self, ptrs[1].value, self.n),
                DeviceArray.view(self, ptrs[2].value, self.n), DeviceArray.view(self, ptrs[3].value, self.nnz))

    # -- host-side pattern logic (fem-fct-pdeco_b200/pattern.py) ----------------------------------
    @property
    def rows(self):
        return self.pattern.rows

    def embed(self, mat):
        return self.pattern.embed(mat)

    def to_scipy(self, vals):
        return self.pattern.to_scipy(vals)

    # -- static data -------------------------------------------------------------------------
    def set_mesh(self, cells, dof_xy):
        cells = np.ascontiguousarray(cells, dtype=np.int32).reshape(-1, 3)
        dof_xy = np.ascontiguousarray(dof_xy, dtype=np.float64).reshape(-1, 2)
        if dof_xy.shape[0] != self.n:
            raise ValueError("dof_xy must have one row per DoF")
        check(lib.fct_ctx_set_mesh(self.handle, cells.shape[0], _hp(cells), _hp(dof_xy)))
        self.has_mesh = True

    def assemble_static(self):
        check(lib.fct_assemble_static(self.handle))
        self.has_mass = True

    def set_mass(self, M_vals_host, ML_host=None):
        tmp = self.array(M_vals_host)
        check(lib.fct_ctx_set_mass(self.handle, tmp.ptr))
        if ML_host is not None:
            t2 = self.array(ML_host)
            check(lib.fct_ctx_set_lumped(self.handle, t2.ptr))
            self.sync()
            t2.free()
        self.sync()
        tmp.free()
        self.has_mass = True

    def set_rect(self, n_cells, g0=0):
        """local row i is DoF g0 + i of the RectangleMesh(n_cells x n_cells) numbering: enables the overlapped-tile kernels"""
        check(lib.fct_ctx_set_rect(self.handle, int(n_cells), int(g0)))

    def tiles_active(self):
        a = C.c_int32()
        check(lib.fct_tiles_active(self.handle, C.byref(a)))
        return bool(a.value)

    def set_solver(self, rtol=1e-14, max_sweeps=200):
        check(lib.fct_ctx_set_solver(self.handle, float(rtol), int(max_sweeps)))

    # -- kernels on device arrays ------------------------------------------------------------
    def spmv(self, A, x, y, alpha=1.0, beta=0.0, z=None):
        check(lib.fct_spmv(self.handle, A.ptr, x.ptr, alpha, beta, z.ptr if z is not None else None, y.ptr))

    def chebsi(self, M, Md, b, y, iters=20, lmin=0.5, lmax=2.0):
        check(lib.fct_chebsi(self.handle, M.ptr, Md.ptr, b.ptr, y.ptr, int(iters), float(lmin), float(lmax)))

    def artificial_diffusion(self, mat, D):
        check(lib.fct_artificial_diffusion(self.handle, mat.ptr, D.ptr))

    def row_lump(self, mat, out):
        check(lib.fct_row_lump(self.handle, mat.ptr, out.ptr))

    def step(self, A, u_n, dt, u_out, sign=1.0, S=None, rhs=None, want_info=True):
        info = _lib.StepInfo() if want_info else None
        check(lib.fct_step(self.handle, A.ptr, float(sign), S.ptr if S is not None else None,
                           rhs.ptr if rhs is not None else None, u_n.ptr, float(dt), u_out.ptr,
                           C.byref(info) if want_info else None))
        return info

    def step_host(self, A_vals, u_n, dt, sign=1.0, S_vals=None, rhs=None):
        """FCT step on host arrays (copies inside); returns (u_out, info)."""
        A_vals = np.ascontiguousarray(A_vals, dtype=np.float64)
        u_n = np.ascontiguousarray(u_n, dtype=np.float64)
        if A_vals.size != self.nnz or u_n.size != self.n:
            raise ValueError("step_host: operand sizes do not match the pattern")
        S_vals = None if S_vals is None else np.ascontiguousarray(S_vals, dtype=np.float64)
        rhs = None if rhs is None else np.ascontiguousarray(rhs, dtype=np.float64)
        out = np.empty(self.n)
        info = _lib.StepInfo()
        check(lib.fct_step_host(self.handle, _hp(A_vals), float(sign), _hp(S_vals) if S_vals is not None else None,
                                _hp(rhs) if rhs is not None else None, _hp(u_n), float(dt), _hp(out), C.byref(info)))
        return out, info

    def solve(self, kind, mat, b, x, rtol=1e-14, maxit=10000):
        its = C.c_int32()
        res = C.c_double()
        check(lib.fct_solve(self.handle, int(kind), mat.ptr, b.ptr, x.ptr, float(rtol), int(maxit), C.byref(its),
                            C.byref(res)))
        return its.value, res.value

    def vals_axpby(self, a, X, b, Y, out):
        check(lib.fct_vals_axpby(self.handle, float(a), X.ptr, float(b), Y.ptr if Y is not None else None, out.ptr))

    def axpby(self, a, x, b, y, out, length=None):
        length = out.size if length is None else length
        check(lib.fct_axpby(self.handle, int(length), float(a), x.ptr, float(b), y.ptr if y is not None else None, out.ptr))

    def clip_axpy(self, x, s, d, lo, hi, out):
        check(lib.fct_clip_axpy(self.handle, out.size, x.ptr, float(s), d.ptr, float(lo), float(hi), out.ptr))

    def dot_M(self, M, x, y):
        out = C.c_double()
        check(lib.fct_dot_M(self.handle, M.ptr, x.ptr, y.ptr, C.byref(out)))
        return out.value

    def norm_sq_Q(self, M, phi, num_steps, dt, target=None):
        out = C.c_double()
        check(lib.fct_norm_sq_Q(self.handle, M.ptr, phi.ptr, target.ptr if target is not None else None,
                                int(num_steps), float(dt), C.byref(out)))
        return out.value

    def assemble_matrix(self, kind, out, c0=None, c1=None, c2=None, s0=0.0, s1=0.0, scale=1.0, accumulate=False):
        p = [a.ptr if a is not None else None for a in (c0, c1, c2)]
        check(lib.fct_assemble_matrix(self.handle, int(kind), p[0], p[1], p[2], float(s0), float(s1), float(scale),
                                      1 if accumulate else 0, out.ptr))

    def assemble_vector(self, kind, out, c0=None, c1=None, c2=None, c3=None, s0=0.0, s1=0.0, scale=1.0,
                        accumulate=False):
        p = [a.ptr if a is not None else None for a in (c0, c1, c2, c3)]
        check(lib.fct_assemble_vector(self.handle, int(kind), p[0], p[1], p[2], p[3], float(s0), float(s1),
                                      float(scale), 1 if accumulate else 0, out.ptr))

    # -- drift-control advection PDECO loops ------------------------------------------------------
    def advdrift_state(self, c_traj, u_traj, num_steps, dt, bx=1.0, by=1.0, eps=0.0):
        sw = C.c_int32()
        check(lib.fct_advdrift_state(self.handle, c_traj.ptr, u_traj.ptr, int(num_steps), float(dt), float(bx),
                                     float(by), float(eps), C.byref(sw)))
        return sw.value

    def advdrift_adjoint(self, c_traj, u_traj, uhat_traj, p_traj, num_steps, dt, bx=1.0, by=1.0, eps=0.0):
        sw = C.c_int32()
        check(lib.fct_advdrift_adjoint(self.handle, c_traj.ptr, u_traj.ptr, uhat_traj.ptr, p_traj.ptr, int(num_steps),
                                       float(dt), float(bx), float(by), float(eps), C.byref(sw)))
        return sw.value

    def advdrift_gradient(self, c_traj, u_traj, p_traj, d_traj, num_steps, beta, bx=1.0, by=1.0):
        check(lib.fct_advdrift_gradient(self.handle, c_traj.ptr, u_traj.ptr, p_traj.ptr, d_traj.ptr, int(num_steps),
                                        float(beta), float(bx), float(by)))

    def advdrift_state_host(self, c_traj_host, u_traj_host, num_steps, dt, bx=1.0, by=1.0, eps=0.0):
        """host (numpy) trajectories; u_traj_host[0:n] holds the IC, slices 1.. are written in place"""
        assert c_traj_host.flags.c_contiguous and u_traj_host.flags.c_contiguous
        assert c_traj_host.dtype == np.float64 and u_traj_host.dtype == np.float64
        assert c_traj_host.size == (num_steps + 1) * self.n == u_traj_host.size
        sw = C.c_int32()
        check(lib.fct_advdrift_state_host(self.handle, _hp(c_traj_host), _hp(u_traj_host), int(num_steps), float(dt),
                                          float(bx), float(by), float(eps), C.byref(sw)))
        return sw.value

    def template_count(self):
        c = C.c_int32()
        check(lib.fct_template_count(self.handle, C.byref(c)))
        return c.value

    def geom_template_count(self):
        c = C.c_int32()
        check(lib.fct_geom_template_count(self.handle, C.byref(c)))
        return c.value

    def bench_jacobi_sweeps(self, A, u_n, dt, reps=20):
        ms = C.c_float()
        check(lib.fct_bench_jacobi_sweeps(self.handle, A.ptr, u_n.ptr, float(dt), int(reps), C.byref(ms)))
        return ms.value

    def bench_dominant_kernel(self, c, u_n, dt):
        """CUDA-event time of the kernel with the largest share of an FCT step (the low-order Jacobi solve: fused K-sweep
        tile launches on structured numberings, else one launch per sweep) on the drift operator of control `c`, with the
        bytes it has to move per launch and the SURVEY App. E accounting bytes of the same work; bench.py's
        `roofline`"""
        n, nnz = self.n, self.nnz
        A = self.empty(nnz)
        self.assemble_matrix(_lib.FORM_DRIFT, A, c0=c, s0=1.0, s1=1.0, scale=-1.0)
        appE = 12 * nnz + 4 * n + 3 * 8 * n
        try:
            if self.tiles_active():
                K = 4
                ms = min(self.bench_jacobi_fused(A, u_n, dt, sweeps=K, reps=6) for _ in range(3)) * K
                return {"name": "k_tile", "ms_per_launch": ms, "appE_bytes": K * appE, "sweeps_per_launch": K,
                        "what": f"{K} fused Jacobi sweeps of the low-order solve on overlapped (diagonal, position) tiles "
                                "(4 launches per FCT step); shared-memory-bound, not HBM-bound: it moves a quarter of the "
                                "per-sweep kernels' bytes",
                        "bytes_per_launch": 8 * nnz + 3 * 8 * n,
                        "bytes_model": "8 B/nnz row-scaled L values + b, x, x_new: 3 x 8 B/row, once per 4 sweeps (frame "
                                       "re-reads are L2 hits)"}
            ms = self.bench_jacobi_sweeps(A, u_n, dt, reps=20)
        finally:
            A.free()
        if self.template_count():
            return {"name": "k_jacobi_sweep_tpl", "ms_per_launch": ms, "appE_bytes": appE, "sweeps_per_launch": 1,
                    "what": "one Jacobi sweep of the low-order solve (~14 launches per FCT step)",
                    "bytes_per_launch": 8 * nnz + 2 * n + 4 * n + 3 * 8 * n,
                    "bytes_model": "8 B/nnz row-scaled L values + 2 B/row template code + 4 B/row rowptr + b, x (gathered), "
                                   "x_new: 3 x 8 B/row"}
        return {"name": "k_jacobi_sweep", "ms_per_launch": ms, "appE_bytes": appE, "bytes_per_launch": appE,
                "sweeps_per_launch": 1, "what": "one Jacobi sweep of the low-order solve (CSR kernel)",
                "bytes_model": "12 B/nnz values + column indices, 4 B/row rowptr, 3 x 8 B/row vectors"}

    def debug_jacobi_fixed(self, A, u_n, dt, sweeps, fused, x_out):
        check(lib.fct_debug_jacobi_fixed(self.handle, A.ptr, u_n.ptr, float(dt), int(sweeps), int(fused), x_out.ptr))

    def bench_jacobi_fused(self, A, u_n, dt, sweeps=4, reps=5):
        """ms per sweep of the overlapped-tile kernel (`sweeps` = 2..4 Jacobi sweeps per launch)"""
        ms = C.c_float()
        check(lib.fct_bench_jacobi_fused(self.handle, A.ptr, u_n.ptr, float(dt), int(sweeps), int(reps), C.byref(ms)))
        return ms.value

    # -- timing ------------------------------------------------------------------------------------
    def event(self):
        e = C.c_void_p()
        check(lib.fct_event_create(self.handle, C.byref(e)))
        return e

    def record(self, ev):
        check(lib.fct_event_record(self.handle, ev))

    def elapsed_ms(self, e0, e1):
        ms = C.c_float()
        check(lib.fct_event_elapsed_ms(self.handle, e0, e1, C.byref(ms)))
        return ms.value
