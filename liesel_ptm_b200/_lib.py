"""ctypes binding of the C ABI in ``include/ptm_b200.h``.

The shared library is built in-tree by ``__graft_entry__.build()`` (nvcc,
sm_100a).  There is NO fallback: if the library is missing or a CUDA call fails
the functions here raise -- the product path never routes through a CPU
implementation.
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PTM_B200_LIB", os.path.join(_HERE, "libptm_b200.so"))

PTM_F32, PTM_F64 = 0, 1
PTM_LOC, PTM_SCALE = 0, 1
PTM_TRAFO_PTM, PTM_TRAFO_ONION, PTM_TRAFO_IDENTITY = 0, 1, 2

STATUS = {
    0: "PTM_OK", 1: "PTM_ERR_NULL", 2: "PTM_ERR_SHAPE", 3: "PTM_ERR_DTYPE",
    4: "PTM_ERR_RANGE", 5: "PTM_ERR_CUDA", 6: "PTM_ERR_WORKSPACE",
}

# every symbol include/ptm_b200.h declares
SYMBOLS = [
    "ptm_problem_create", "ptm_problem_destroy", "ptm_num_params", "ptm_param_layout",
    "ptm_workspace_bytes", "ptm_problem_shard_view", "ptm_problem_shard",
    "ptm_logp_grad_block_f64", "ptm_logp_grad_block_f32",
    "ptm_logp_f64", "ptm_logp_f32", "ptm_logp_grad_f64", "ptm_logp_grad_f32",
    "ptm_basis_f64", "ptm_basis_f32", "ptm_transform_f64", "ptm_transform_f32",
    "ptm_xtwx_f64", "ptm_xtwx_f32", "ptm_loc_precision_f64", "ptm_loc_precision_f32",
    "ptm_quadform_f64", "ptm_quadform_f32", "ptm_intercept_stats_f64", "ptm_intercept_stats_f32",
    "ptm_loc_precision_all_f64", "ptm_loc_precision_all_f32",
    "ptm_hessian_workspace_bytes", "ptm_hessian_block_f64", "ptm_hessian_block_f32",
    "ptm_pointwise_workspace_bytes", "ptm_pointwise_f64", "ptm_pointwise_f32",
    "ptm_problem_shared_designs", "ptm_predict_workspace_bytes", "ptm_predict_f64", "ptm_predict_f32",
    "ptm_predict_quantile_f64", "ptm_predict_quantile_f32", "ptm_predict_sample_f64", "ptm_predict_sample_f32", "ptm_inverse_f64", "ptm_inverse_f32",
    "ptm_last_launch_count", "ptm_last_main_route", "ptm_last_xtwx_route", "ptm_profile_enable", "ptm_last_kernel_ms", "ptm_last_error",
    "ptm_version",
]


class PtmProblemDesc(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32),
        ("n", C.c_int64),
        ("y", C.c_void_p),
        ("n_terms", C.c_int32),
        ("x", C.POINTER(C.c_void_p)),
        ("nbases", C.POINTER(C.c_int32)),
        ("ncoef", C.POINTER(C.c_int32)),
        ("predictor", C.POINTER(C.c_int32)),
        ("knots", C.POINTER(C.c_void_p)),
        ("Z", C.POINTER(C.c_void_p)),
        ("K", C.POINTER(C.c_void_p)),
        ("rank", C.POINTER(C.c_int32)),
        ("nparam", C.c_int32),
        ("trafo_knots", C.c_void_p),
        ("trafo_Z", C.c_void_p),
        ("trafo_lambda", C.c_double),
        ("trafo_grid", C.c_void_p),
        ("noncentered", C.POINTER(C.c_int32)),
        ("trafo_noncentered", C.c_int32),
        ("trafo_variant", C.c_int32),
    ]


class PtmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        self.code = code
        super().__init__(f"{STATUS.get(code, code)}: {msg}")


_lib = None


def load() -> C.CDLL:
    """Load the CUDA library; fail loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a).  liesel_ptm_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
    lib.ptm_problem_create.argtypes = [C.POINTER(PtmProblemDesc), C.POINTER(vp)]
    lib.ptm_problem_create.restype = C.c_int
    lib.ptm_problem_destroy.argtypes = [vp]
    lib.ptm_problem_destroy.restype = C.c_int
    lib.ptm_num_params.argtypes = [vp]
    lib.ptm_num_params.restype = i64
    lib.ptm_param_layout.argtypes = [vp] + [C.POINTER(i64)] * 6
    lib.ptm_param_layout.restype = C.c_int
    lib.ptm_workspace_bytes.argtypes = [vp, i64]
    lib.ptm_workspace_bytes.restype = sz
    lib.ptm_problem_shard_view.argtypes = [vp, i64, i64, i32, C.POINTER(vp)]
    lib.ptm_problem_shard_view.restype = C.c_int
    lib.ptm_problem_shard.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i32)]
    lib.ptm_problem_shard.restype = C.c_int
    for sfx in ("f64", "f32"):
        f = getattr(lib, f"ptm_logp_{sfx}")
        f.argtypes = [vp, vp, i64, vp, vp, sz, vp]
        f.restype = C.c_int
        f = getattr(lib, f"ptm_logp_grad_{sfx}")
        f.argtypes = [vp, vp, i64, vp, vp, vp, sz, vp]
        f.restype = C.c_int
        f = getattr(lib, f"ptm_logp_grad_block_{sfx}")
        f.argtypes = [vp, vp, i64, vp, vp, sz, vp]
        f.restype = C.c_int
        f = getattr(lib, f"ptm_basis_{sfx}")
        f.argtypes = [vp, i32, vp, vp, vp]
        f.restype = C.c_int
        f = getattr(lib, f"ptm_transform_{sfx}")
        f.argtypes = [vp, vp, i64, vp, i64, vp, vp, vp, vp]
        f.restype = C.c_int
        f = getattr(lib, f"ptm_xtwx_{sfx}")
        f.argtypes = [vp, i32, vp, i64, vp, vp, sz, vp]
        f.restype = C.c_int
        f = getattr(lib, f"ptm_loc_precision_all_{sfx}")
        f.argtypes = [vp, vp, i64, vp, vp, vp, sz, vp]
        f.restype = C.c_int
        f = getattr(lib, f"ptm_intercept_stats_{sfx}")
        f.argtypes = [vp, vp, i64, vp, vp, sz, vp]
        f.restype = C.c_int
        f = getattr(lib, f"ptm_quadform_{sfx}")
        f.argtypes = [vp, vp, i64, vp, vp]
        f.restype = C.c_int
        f = getattr(lib, f"ptm_loc_precision_{sfx}")
        f.argtypes = [vp, i32, vp, i64, vp, vp, vp, sz, vp]
        f.restype = C.c_int
    lib.ptm_hessian_workspace_bytes.argtypes = [vp, i64, i32]
    lib.ptm_hessian_workspace_bytes.restype = sz
    for sfx in ("f64", "f32"):
        f = getattr(lib, f"ptm_hessian_block_{sfx}")
        f.argtypes = [vp, i32, i32, vp, i64, vp, vp, vp, sz, vp]
        f.restype = C.c_int
    lib.ptm_pointwise_workspace_bytes.argtypes = [vp, i64]
    lib.ptm_pointwise_workspace_bytes.restype = sz
    for sfx in ("f64", "f32"):
        f = getattr(lib, f"ptm_pointwise_{sfx}")
        f.argtypes = [vp, vp, i64, vp, vp, vp, sz, vp]
        f.restype = C.c_int
    lib.ptm_problem_shared_designs.argtypes = [vp]
    lib.ptm_problem_shared_designs.restype = C.c_int
    lib.ptm_predict_workspace_bytes.argtypes = [vp, i64]
    lib.ptm_predict_workspace_bytes.restype = sz
    for sfx in ("f64", "f32"):
        f = getattr(lib, f"ptm_predict_{sfx}")
        f.argtypes = [vp, vp, i64, vp, vp, i64, vp, vp, vp, vp, sz, vp]
        f.restype = C.c_int
        f = getattr(lib, f"ptm_predict_quantile_{sfx}")
        f.argtypes = [vp, vp, i64, vp, vp, i64, vp, vp, sz, vp]
        f.restype = C.c_int
        f = getattr(lib, f"ptm_predict_sample_{sfx}")
        f.argtypes = [vp, vp, i64, vp, vp, i64, vp, vp, sz, vp]
        f.restype = C.c_int
        f = getattr(lib, f"ptm_inverse_{sfx}")
        f.argtypes = [vp, vp, i64, vp, i64, vp, vp]
        f.restype = C.c_int
    lib.ptm_last_launch_count.restype = C.c_int
    lib.ptm_last_xtwx_route.restype = C.c_int
    lib.ptm_last_main_route.restype = C.c_int
    lib.ptm_profile_enable.argtypes = [C.c_int]
    lib.ptm_profile_enable.restype = C.c_int
    lib.ptm_last_kernel_ms.argtypes = [C.POINTER(C.c_float)] * 3
    lib.ptm_last_kernel_ms.restype = C.c_int
    lib.ptm_last_error.restype = C.c_char_p
    lib.ptm_version.restype = C.c_char_p
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise PtmError(rc, load().ptm_last_error().decode())
