"""ctypes binding of libfctpdeco.so (the C ABI declared in include/fctpdeco.h).

There is no CPU fallback: if the shared library is missing the import of this module raises, and if no CUDA
device is present ``FctContext`` creation raises.  Build with ``python fem-fct-pdeco_b200/build.py`` or
``__graft_entry__.build()``.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfctpdeco.so")


class FctError(RuntimeError):
    """Error reported by libfctpdeco (message from fct_last_error())."""


class StepInfo(C.Structure):
    _fields_ = [("solver_sweeps", C.c_int32), ("converged", C.c_int32), ("last_delta", C.c_double),
                ("x_norm", C.c_double), ("min_rowsum_low", C.c_double)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA library has not been built "
            "(run `python fem-fct-pdeco_b200/build.py`).  This package has no CPU fallback.")
    return C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)


lib = _load()

_p = C.c_void_p
_i32 = C.c_int32
_i64 = C.c_int64
_f64 = C.c_double
_pi32 = C.POINTER(C.c_int32)
_pi64 = C.POINTER(C.c_int64)
_pf64 = C.POINTER(C.c_double)

# name -> argtypes  (every entry point of include/fctpdeco.h; tests/test_abi.py checks the two stay in sync)
SIGNATURES = {
    "fct_last_error": [],
    "fct_version": [],
    "fct_device_count": [],
    "fct_mesh_rect_sizes": [_i32, _pi64, _pi64, _pi64],
    "fct_mesh_rect_build": [_i32, _f64, _f64, _p, _p, _p, _p, _p],
    "fct_ctx_create": [C.POINTER(_p), C.c_int, _i32, _p, _p, _i32, _i32],
    "fct_ctx_destroy": [_p],
    "fct_ctx_set_stream": [_p, _p],
    "fct_ctx_sync": [_p],
    "fct_ctx_sizes": [_p, _pi32, _pi64, _pi32, _pi32],
    "fct_ctx_pattern_dev": [_p, C.POINTER(_p), C.POINTER(_p), C.POINTER(_p)],
    "fct_ctx_set_mesh": [_p, _i64, _p, _p],
    "fct_ctx_set_mass": [_p, _p],
    "fct_ctx_set_lumped": [_p, _p],
    "fct_assemble_static": [_p],
    "fct_ctx_static_dev": [_p, C.POINTER(_p), C.POINTER(_p), C.POINTER(_p), C.POINTER(_p)],
    "fct_ctx_set_solver": [_p, _f64, _i32],
    "fct_ctx_set_rect": [_p, _i32, _i64],
    "fct_tiles_active": [_p, _pi32],
    "fct_debug_tile_list": [_i32, _i64, _i32, _i32, _i32, _p, _i32, _pi32, _pi32],
    "fct_malloc": [_p, C.POINTER(_p), _i64],
    "fct_free": [_p, _p],
    "fct_h2d": [_p, _p, _p, _i64],
    "fct_d2h": [_p, _p, _p, _i64],
    "fct_host_alloc": [_p, C.POINTER(_p), _i64],
    "fct_host_free": [_p, _p],
    "fct_spmv": [_p, _p, _p, _f64, _f64, _p, _p],
    "fct_chebsi": [_p, _p, _p, _p, _p, _i32, _f64, _f64],
    "fct_artificial_diffusion": [_p, _p, _p],
    "fct_row_lump": [_p, _p, _p],
    "fct_step": [_p, _p, _f64, _p, _p, _p, _f64, _p, C.POINTER(StepInfo)],
    "fct_step_host": [_p, _p, _f64, _p, _p, _p, _f64, _p, C.POINTER(StepInfo)],
    "fct_solve": [_p, _i32, _p, _p, _p, _f64, _i32, _pi32, _pf64],
    "fct_vals_axpby": [_p, _f64, _p, _f64, _p, _p],
    "fct_dot_M": [_p, _p, _p, _p, _pf64],
    "fct_norm_sq_Q": [_p, _p, _p, _p, _i32, _f64, _pf64],
    "fct_clip_axpy": [_p, _i64, _p, _f64, _p, _f64, _f64, _p],
    "fct_axpby": [_p, _i64, _f64, _p, _f64, _p, _p],
    "fct_assemble_matrix": [_p, _i32, _p, _p, _p, _f64, _f64, _f64, _i32, _p],
    "fct_assemble_vector": [_p, _i32, _p, _p, _p, _p, _f64, _f64, _f64, _i32, _p],
    "fct_forward_nonlinear": [_p, _p, _f64, _p, _i32, _f64, _f64, _p, _pi32],
    "fct_adjoint_nonlinear": [_p, _p, _p, _i32, _f64, _f64, _p, _pi32],
    "fct_forward_schnak": [_p, _p, _f64, _p, _p, _i32, _f64, _p, _p, _f64, _pi32],
    "fct_adjoint_schnak": [_p, _p, _p, _p, _p, _i32, _f64, _p, _p, _pi32],
    "fct_forward_chtxs": [_p, _p, _f64, _p, _p, _i32, _f64, _p, _f64, _pi32],
    "fct_adjoint_chtxs": [_p, _p, _p, _p, _p, _p, _p, _p, _i32, _f64, _p, _f64, _pi32],
    "fct_advdrift_state": [_p, _p, _p, _i32, _f64, _f64, _f64, _f64, _pi32],
    "fct_advdrift_adjoint": [_p, _p, _p, _p, _p, _i32, _f64, _f64, _f64, _f64, _pi32],
    "fct_advdrift_gradient": [_p, _p, _p, _p, _p, _i32, _f64, _f64, _f64],
    "fct_advdrift_state_host": [_p, _p, _p, _i32, _f64, _f64, _f64, _f64, _pi32],
    "fct_nccl_unique_id": [_p],
    "fct_ctx_init_comm": [_p, _p, _i32, _i32, _i32, _i32, _i32, _i32],
    "fct_halo_exchange": [_p, _p],
    "fct_ctx_set_rings": [_p, _i32, _pi32, _pi32],
    "fct_p2p_create": [_p, _i32, _i32, _i32, _p],
    "fct_p2p_connect": [_p, _p],
    "fct_p2p_error": [_p, _pi32],
    "fct_launch_count": [_p, _pi64],
    "fct_exchange_count": [_p, _pi64],
    "fct_guard_check": [_pi64, _pi64],
    "fct_profiler_range": [_i32],
    "fct_bench_jacobi_sweeps": [_p, _p, _p, _f64, _i32, C.POINTER(C.c_float)],
    "fct_debug_jacobi_fixed": [_p, _p, _p, _f64, _i32, _i32, _p],
    "fct_bench_jacobi_fused": [_p, _p, _p, _f64, _i32, _i32, C.POINTER(C.c_float)],
    "fct_template_count": [_p, _pi32],
    "fct_geom_template_count": [_p, _pi32],
    "fct_event_create": [_p, C.POINTER(_p)],
    "fct_event_record": [_p, _p],
    "fct_event_elapsed_ms": [_p, _p, _p, C.POINTER(C.c_float)],
    "fct_event_destroy": [_p, _p],
}

for _name, _args in SIGNATURES.items():
    _fn = getattr(lib, _name)          # AttributeError here = header/library mismatch: fail loudly
    _fn.argtypes = _args
    _fn.restype = C.c_char_p if _name == "fct_last_error" else C.c_int

# form / load kinds (include/fctpdeco.h)
FORM_MASS, FORM_STIFFNESS, FORM_DRIFT, FORM_WIND_P1, FORM_WIND_P1_T = 0, 1, 2, 3, 4
FORM_WMASS1, FORM_WMASS2, FORM_WMASS3, FORM_CHTX, FORM_CHTX_EXP, FORM_CHTX_ADJ = 5, 6, 7, 8, 9, 10
FORM_WIND_POLY3, FORM_WIND_POLY3_T, FORM_DRIFT_MASS, FORM_DRIFT_CONV, FORM_DIVW_MASS = 11, 12, 13, 14, 15
LOAD_P1_1, LOAD_P1_2, LOAD_P1_3, LOAD_P1_4, LOAD_CONST, LOAD_DRIFT_GRAD, LOAD_CHTX_ADJ, LOAD_POLY3 = 0, 1, 2, 3, 4, 5, 6, 7
SOLVER_JACOBI, SOLVER_PCG, SOLVER_BICGSTAB, SOLVER_CHEB_PCG = 0, 1, 2, 3


def check(rc):
    if rc != 0:
        msg = lib.fct_last_error()
        raise FctError(msg.decode() if msg else f"libfctpdeco call failed with code {rc}")


def device_count():
    return int(lib.fct_device_count())
