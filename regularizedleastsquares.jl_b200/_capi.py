"""ctypes binding of librls_b200.so (include/rls_b200.h).

This is the Python twin of the `ccall` layer in julia/RLSB200.jl: the same symbols,
the same POD structs.  There is no fallback — if the shared library is missing the
import fails loudly, and if no sm_100 device is present `rls_ctx_create` fails.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "librls_b200.so")

RLS_OK = 0
RLS_F32, RLS_C32 = 0, 1
RLS_FISTA, RLS_POGM, RLS_OPTISTA, RLS_CGNR, RLS_ADMM, RLS_SPLITBREGMAN = 0, 1, 2, 3, 4, 5
RLS_REG_NONE, RLS_REG_L1, RLS_REG_L2, RLS_REG_L21, RLS_REG_TV, RLS_REG_NUCLEAR, RLS_REG_LLR = 0, 1, 2, 3, 4, 5, 6
RLS_LLR_RANDSHIFT, RLS_LLR_OVERLAPPING = 1, 2
RLS_PROJ_REAL, RLS_PROJ_POSITIVE = 1, 2
RLS_NORMAL_TWOPASS, RLS_NORMAL_ONEPASS, RLS_NORMAL_GRAM, RLS_NORMAL_AUTO, RLS_NORMAL_MATRIXFREE = 0, 1, 2, 3, 4
RLS_TRAFO_IDENTITY, RLS_TRAFO_GRADIENT = 0, 1
RLS_VARY_RHO_NONE, RLS_VARY_RHO_BALANCE, RLS_VARY_RHO_PNP = 0, 1, 2
RLS_DIST_UNIFORM01, RLS_DIST_IH4 = 0, 1
RLS_LAYOUT_COLMAJOR, RLS_LAYOUT_ROWMAJOR, RLS_LAYOUT_AUTO = 0, 1, 2
RLS_MAX_TV_DIMS = 4


class RlsError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"librls_b200 status {status}: {msg}")
        self.status = status


class RegDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("lambda_is_f64", C.c_int32), ("lambda_", C.c_double), ("slices", C.c_int64),
        ("tv_ndims", C.c_int32), ("tv_ndirs", C.c_int32), ("tv_shape", C.c_int64 * RLS_MAX_TV_DIMS),
        ("tv_dims", C.c_int32 * RLS_MAX_TV_DIMS), ("tv_iterations", C.c_int32), ("trafo", C.c_int32),
        ("rho", C.c_float), ("_pad", C.c_int32),
    ]


class SolverDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("iterations", C.c_int32), ("restart", C.c_int32), ("proj_mask", C.c_int32),
        ("rho", C.c_float), ("theta", C.c_float), ("sigma_fac", C.c_float), ("rel_tol", C.c_float),
        ("abs_tol", C.c_float), ("tol_inner", C.c_float), ("iterations_cg", C.c_int32), ("vary_rho", C.c_int32),
        ("n_reg", C.c_int32), ("iterations_inner", C.c_int32), ("reg", RegDesc * 4),
    ]


class SolverScalars(C.Structure):
    _fields_ = [
        ("iteration", C.c_int32), ("done", C.c_int32),
        ("rho", C.c_float), ("theta", C.c_float), ("theta_old", C.c_float), ("theta_n", C.c_float),
        ("alpha", C.c_float), ("beta", C.c_float), ("gamma", C.c_float), ("gamma_old", C.c_float), ("sigma", C.c_float),
        ("norm_x0", C.c_float), ("rel_res_norm", C.c_float), ("res_norm", C.c_float),
        ("cg_alpha", C.c_float * 2), ("cg_beta", C.c_float * 2), ("cg_zeta", C.c_float * 2),
        ("admm_rk", C.c_float * 4), ("admm_sk", C.c_float * 4), ("admm_eps_pri", C.c_float * 4),
        ("admm_eps_dua", C.c_float * 4), ("admm_delta", C.c_float * 4), ("admm_rho", C.c_float * 4),
        ("admm_sigma_abs", C.c_float), ("cg_iterations_last", C.c_int32), ("cg_iterations_total", C.c_int32),
        ("outer_iteration", C.c_int32),
    ]


_P = C.c_void_p
_I32, _I64, _U64, _F32, _F64 = C.c_int32, C.c_int64, C.c_uint64, C.c_float, C.c_double
_PI32, _PI64, _PF32, _PF64 = C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_float), C.POINTER(C.c_double)
_PP = C.POINTER(C.c_void_p)
# rls_apply_fn: int32 (*)(void* user, const void* x_dev, void* res_dev, void* cuda_stream)
APPLY_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p)

# name -> argtypes ; every function returns int32 unless listed in _SPECIAL
SIGNATURES = {
    "rls_device_count": [_PI32],
    "rls_ctx_create": [_I32, _PP],
    "rls_ctx_destroy": [_P],
    "rls_ctx_sync": [_P],
    "rls_ctx_device_info": [_P, _PI32, _PI32, _PI32, _PI64, _PI64],
    "rls_timer_start": [_P],
    "rls_timer_stop": [_P, _PF32],
    "rls_ctx_launch_count": [_P, _PI64],
    "rls_ctx_flush_l2": [_P],
    "rls_comm_unique_id": [_P],
    "rls_ctx_comm_init": [_P, _I32, _I32, _P],
    "rls_ctx_comm_info": [_P, _PI32, _PI32],
    "rls_vec_allreduce": [_P],
    "rls_ctx_allreduce_f64": [_P, _PF64, _I32],
    "rls_ctx_peer_export": [_P, _I64, _P],
    "rls_ctx_peer_import": [_P, _P, _I32],
    "rls_group_create": [_I32, _PI32, _PP],
    "rls_group_destroy": [_P],
    "rls_group_size": [_P, _PI32],
    "rls_group_ctx": [_P, _I32, _PP],
    "rls_group_mat_create": [_P, _I32, _I64, _I64, _P, _I64, _PP],
    "rls_group_mat_fill_philox": [_P, _U64, _I32, _F32],
    "rls_group_mat_part": [_P, _I32, _PP, _PI64, _PI64],
    "rls_group_mat_destroy": [_P],
    "rls_group_solver_create": [_P, _I32, C.POINTER(SolverDesc), _PP],
    "rls_group_solver_destroy": [_P],
    "rls_group_solver_solve_host": [_P, _P, _I64, _P, _I64, _PI32, C.POINTER(SolverScalars)],
    "rls_vec_create": [_P, _I32, _I64, _PP],
    "rls_vec_destroy": [_P],
    "rls_vec_len": [_P, _PI64, _PI32],
    "rls_vec_upload": [_P, _P, _I64],
    "rls_vec_download": [_P, _P, _I64],
    "rls_vec_copy": [_P, _P],
    "rls_vec_fill": [_P, _F32, _F32],
    "rls_vec_fill_philox": [_P, _U64, _U64, _I32, _F32, _I64],
    "rls_vec_device_ptr": [_P, _PP],
    "rls_vec_nrm2": [_P, _PF64],
    "rls_vec_asum": [_P, _PF64],
    "rls_vec_dot": [_P, _P, _PF64],
    "rls_mat_create": [_P, _I32, _I64, _I64, _P, _I64, _PP],
    "rls_mat_create_layout": [_P, _I32, _I64, _I64, _P, _I64, _I32, _PP],
    "rls_mat_layout": [_P, _PI32],
    "rls_mat_wrap_device": [_P, _I32, _I64, _I64, _P, _I64, _PP],
    "rls_mat_destroy": [_P],
    "rls_mat_relayout": [_P, _I32, _PP],
    "rls_mat_shape": [_P, _PI64, _PI64, _PI32],
    "rls_mat_upload": [_P, _P, _I64],
    "rls_mat_download": [_P, _P, _I64],
    "rls_mat_fill_philox": [_P, _U64, _I32, _F32, _I64, _I64],
    "rls_mat_frob2": [_P, _PF64],
    "rls_gemv_n": [_P, _P, _P],
    "rls_gemv_c": [_P, _P, _P],
    "rls_normal_create": [_P, _I32, _PP],
    "rls_normal_from_gram": [_P, _PP],
    "rls_normal_destroy": [_P],
    "rls_normal_form": [_P, _PI32],
    "rls_normal_describe": [_P, C.c_char_p, _I32],
    "rls_normal_apply": [_P, _P, _P],
    "rls_normal_apply_batch": [_P, _I32, _PP, _PP],
    "rls_normal_batch_debug": [_P, _I32, _P, _I64],
    "rls_power_iterations": [_P, _P, _F64, _I32, _PF64],
    "rls_linop_sampling_create": [_P, _I32, _I64, _I64, _PI64, _PP],
    "rls_linop_fft_create": [_P, _I32, _PI64, _I32, _I32, _PP],
    "rls_linop_compose": [_P, _P, _PP],
    "rls_linop_destroy": [_P],
    "rls_linop_shape": [_P, _PI64, _PI64, _PI32],
    "rls_linop_mul": [_P, _P, _P],
    "rls_linop_mul_adjoint": [_P, _P, _P],
    "rls_normal_from_linop": [_P, _PP],
    "rls_normal_from_callback": [_P, _I32, _I64, APPLY_FN, _P, _PP],
    "rls_prox_l1": [_P, _F32],
    "rls_prox_l2": [_P, _F32],
    "rls_prox_l21": [_P, _F32, _I64],
    "rls_prox_tv": [_P, _F32, _I32, _PI64, _I32, _PI32, _I32],
    "rls_prox_positive": [_P],
    "rls_prox_real": [_P],
    "rls_prox_nuclear": [_P, _F32, _I64, _I64],
    "rls_prox_llr": [_P, _F32, _I32, _PI64, _PI64, _PI64, _I32],
    "rls_grad_rows": [_I32, _PI64, _I32, _PI32, _PI64],
    "rls_grad_apply": [_P, _P, _I32, _PI64, _I32, _PI32],
    "rls_grad_apply_t": [_P, _P, _I32, _PI64, _I32, _PI32],
    "rls_solver_create": [_P, _P, C.POINTER(SolverDesc), _PP],
    "rls_solver_destroy": [_P],
    "rls_solver_set_reg": [_P, _I32, C.POINTER(RegDesc)],
    "rls_solver_init": [_P, _P, _P],
    "rls_solver_iterate": [_P, _PI32, C.POINTER(SolverScalars)],
    "rls_solver_run": [_P, _PI32, C.POINTER(SolverScalars)],
    "rls_solver_solve": [_P, _P, _P, _PI32, C.POINTER(SolverScalars)],
    "rls_solver_solve_host": [_P, _P, _I64, _P, _I64, _PI32, C.POINTER(SolverScalars)],
    "rls_solver_scalars_get": [_P, C.POINTER(SolverScalars)],
    "rls_solver_vec": [_P, C.c_char_p, _PP],
    "rls_solver_solve_batch_host": [_P, _P, _I64, _I32, _P, _I64, _PI32],
    "rls_kaczmarz_create": [_P, _I32, _PP],
    "rls_kaczmarz_destroy": [_P],
    "rls_kaczmarz_block_rows": [_P, _PI32],
    "rls_kaczmarz_rownorm2": [_P, _P, _I64],
    "rls_kaczmarz_set_rows": [_P, _P, _P, _I64],
    "rls_kaczmarz_init": [_P, _P, _P, _F32],
    "rls_kaczmarz_sweep": [_P],
    "rls_kaczmarz_vec": [_P, C.c_char_p, _PP],
    "rls_kaczmarz_debug": [_P, _I32, _P, _I64],
    "rls_kaczmarz_check": [_P],
    "rls_kaczmarz_describe": [_P, C.c_char_p, _I32],
}
_SPECIAL = {"rls_abi_version": ([], _I32), "rls_last_error": ([], C.c_char_p)}

_lib = None


def load():
    """dlopen the library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, args in SIGNATURES.items():
        f = getattr(lib, name)
        f.argtypes = args
        f.restype = _I32
    for name, (args, res) in _SPECIAL.items():
        f = getattr(lib, name)
        f.argtypes = args
        f.restype = res
    _lib = lib
    return lib


def last_error():
    return load().rls_last_error().decode("utf-8", "replace")


def check(status):
    if status != RLS_OK:
        raise RlsError(status, last_error())


def call(name, *args):
    check(getattr(load(), name)(*args))
