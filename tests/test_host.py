"""CPU tests of the host-side mirror (set-up code, parameter layout, API errors) and of
the C-ABI library as a shared object (symbols, argument validation that needs no GPU)."""

import ctypes as C
import os
import re
import sys

import numpy as np
import pytest

import liesel_ptm_b200 as lp
from liesel_ptm_b200 import _lib
from oracle import ptm_oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_knots_match_oracle():
    x = np.random.default_rng(0).uniform(-2, 2, 100)
    assert np.array_equal(lp.equidistant_knots(x, 20), o.equidistant_knots(x, 20))
    a, b = lp.PTMKnots(-4.0, 4.0, 20), o.PTMKnots(-4.0, 4.0, 20)
    assert np.array_equal(a.knots, b.knots) and a.step == b.step
    assert np.array_equal(lp.approx_grid(a.knots), o.PTMSpline(b.knots).bspline.grid)
    with pytest.raises(ValueError):
        lp.equidistant_knots(x, 2)


def test_banded_basis_equals_dense():
    x = np.random.default_rng(1).uniform(-2, 2, 500)
    knots = lp.equidistant_knots(x, 15)
    x[0], x[1] = knots[3], knots[15]  # both ends of the interior range
    j, v = lp.banded_basis(x, knots)
    D = o.basis_matrix(x, knots)
    dense = np.zeros((500, 16))
    for m in range(4):
        dense[np.arange(500), j - 3 + m] = v[:, m]
    assert np.array_equal(dense[:, :15], D)
    assert np.all(dense[:, 15] == 0)
    assert np.array_equal(j, np.searchsorted(knots, x, side="right") - 1)


def test_ps_properties_and_errors():
    x = np.random.default_rng(2).uniform(-2, 2, 300)
    b = lp.ps(x, nbases=20, xname="x")
    assert b.Z.shape == (20, 19) and b.penalty.shape == (19, 19)
    D = o.basis_matrix(x, b.knots)
    assert np.abs((D @ b.Z).sum(axis=0)).max() < 1e-9  # sum-to-zero term constraint
    assert np.linalg.matrix_rank(b.penalty) == 18
    t = lp.term.f_ig(b)
    assert t.name == "f(x)" and t.penalty_rank == 18 and t.nbases == 19
    with pytest.raises(ValueError):
        lp.ps(x, nbases=20)  # unnamed
    with pytest.raises(ValueError):
        lp.ps(np.append(x, 9.0), nbases=20, xname="x", knots=b.knots)  # outside the knots


def test_trafo_reparam():
    Z = lp.trafo_reparam(20)
    assert Z.shape == (20, 19)
    pen = lp.Penalty.pspline(20, 1)
    pen = pen / np.linalg.norm(pen, ord=np.inf)
    assert np.allclose(Z.T @ pen @ Z, np.eye(19), atol=1e-8)
    assert np.abs(Z.sum(axis=0)).max() < 1e-8  # delta sums to zero


def _model(n=50, to_float32=False):
    rng = np.random.default_rng(3)
    y = rng.normal(size=n)
    m = lp.LocScalePTM.new_ptm(y, nparam=20, to_float32=to_float32, device="cuda:0")
    m.loc += lp.term.f_ig(lp.ps(rng.uniform(-2, 2, n), nbases=12, xname="x1"))
    m.scale += lp.term.f_ig(lp.ps(rng.uniform(-2, 2, n), nbases=10, xname="x2"), fname="g")
    return m


def test_param_layout_and_pack():
    m = _model()
    L, P = m.param_layout()
    assert L == {"beta0": 0, "gamma0": 1, "beta": [2, 13], "delta_z": 22, "tau": [41, 42], "tau_delta": 43}
    assert P == 44
    beta = [np.full((3, 11), 1.0), np.full((3, 9), 2.0)]
    par = m.pack(beta, np.full((3, 19), 3.0), tau=np.array([[4.0, 5.0]] * 3), tau_delta=6.0, beta0=7.0,
                 gamma0=8.0)
    assert par.shape == (3, 44)
    assert np.all(par[:, 0] == 7) and np.all(par[:, 1] == 8)
    assert np.all(par[:, 2:13] == 1) and np.all(par[:, 13:22] == 2) and np.all(par[:, 22:41] == 3)
    assert np.all(par[:, 41] == 4) and np.all(par[:, 42] == 5) and np.all(par[:, 43] == 6)


def test_model_api_errors():
    m = _model()
    with pytest.raises(TypeError):
        m.loc += 3.0
    with pytest.raises(RuntimeError):
        m.loc += list(m.loc.terms.values())[0]  # duplicate name
    with pytest.raises(ValueError):
        m.loc = "something else"
    import torch

    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU path"):
            m.build()  # the product path fails loudly without a GPU


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libptm_b200.so")
    with pytest.raises(ImportError, match="no CPU fallback"):
        _lib.load()


def _header_functions():
    txt = open(os.path.join(ROOT, "include", "ptm_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ptm_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _header_functions()
    assert len(names) >= 20
    for nm in names:
        assert hasattr(lib, nm), nm
    assert sorted(_lib.SYMBOLS) == names
    assert b"sm_100a" in lib.ptm_version()


def test_capi_argument_validation_without_gpu():
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.ptm_problem_create(None, C.byref(h)) == 1  # PTM_ERR_NULL
    d = _lib.PtmProblemDesc()
    d.dtype = 7
    assert lib.ptm_problem_create(C.byref(d), C.byref(h)) == 3  # PTM_ERR_DTYPE
    # a covariate outside its interior knots is rejected before any CUDA call,
    # like the reference's ValueError (bspline/approx.py:197-206)
    y = np.zeros(4)
    x = np.array([0.0, 0.5, 1.0, 9.0])
    knots = lp.equidistant_knots(np.array([-2.0, 2.0]), 8)
    Z, K = np.eye(8), np.eye(8)
    tk = lp.PTMKnots(-4.0, 4.0, 20).knots
    tz = lp.trafo_reparam(20)
    d = _lib.PtmProblemDesc()
    d.dtype, d.n, d.y, d.n_terms = _lib.PTM_F64, 4, y.ctypes.data, 1
    xs = (C.c_void_p * 1)(x.ctypes.data)
    ks = (C.c_void_p * 1)(knots.ctypes.data)
    zs = (C.c_void_p * 1)(Z.ctypes.data)
    Ks = (C.c_void_p * 1)(K.ctypes.data)
    one = (C.c_int32 * 1)
    d.x, d.knots, d.Z, d.K = xs, ks, zs, Ks
    d.nbases, d.ncoef, d.predictor, d.rank = one(8), one(8), one(0), one(8)
    d.nparam, d.trafo_knots, d.trafo_Z, d.trafo_lambda = 20, tk.ctypes.data, tz.ctypes.data, 0.1
    assert lib.ptm_problem_create(C.byref(d), C.byref(h)) == 4  # PTM_ERR_RANGE
    assert b"not inside the range of interior knots" in lib.ptm_last_error()
    assert lib.ptm_num_params(None) == -1
    assert lib.ptm_workspace_bytes(None, 4) == 0
    assert lib.ptm_pointwise_workspace_bytes(None, 4) == 0
    # every per-call entry point rejects a null problem / null arrays with PTM_ERR_NULL
    # before touching the device
    for sfx in ("f64", "f32"):
        assert getattr(lib, f"ptm_logp_{sfx}")(None, None, 1, None, None, 0, None) == 1
        assert getattr(lib, f"ptm_logp_grad_{sfx}")(None, None, 1, None, None, None, 0, None) == 1
        assert getattr(lib, f"ptm_pointwise_{sfx}")(None, None, 1, None, None, None, 0, None) == 1
        assert getattr(lib, f"ptm_loc_precision_all_{sfx}")(None, None, 1, None, None, None, 0, None) == 1
        assert getattr(lib, f"ptm_intercept_stats_{sfx}")(None, None, 1, None, None, 0, None) == 1
        assert getattr(lib, f"ptm_xtwx_{sfx}")(None, 0, None, 1, None, None, 0, None) == 1
    assert lib.ptm_last_xtwx_route() in (0, 1, 2)


def test_bench_workload_regions():
    import bench_workload as bw

    wl = bw.build("c3_n1M_5loc_5scale_nb20_1024chains", chains=2, device="cpu", n_override=20000)
    assert wl.params.shape == (2, 222) and wl.bytes_alg == 88
    terms = [o.term_from_reparam(wl.X[i], t.basis.knots, t.basis.Z, t.penalty, t.penalty_rank,
                                 "loc" if k == 0 else "scale") for i, (t, k) in enumerate(wl.model.all_terms())]
    om = o.PTMModel(wl.y, terms, nparam=20)
    om.Zd = lp.trafo_reparam(20)
    from tests.golden_io import state_from_params

    st = state_from_params(om, wl.params, wl.model.param_layout()[0])
    frac = np.bincount(om.spline.region_code(om.loglik_terms(st, 0)["r"]), minlength=5) / wl.n
    assert frac[2] > 0.95 and frac.min() > 1e-4  # mostly core, all five regions exercised


def test_generated_dispatch_header_is_current():
    """liesel_ptm_b200/csrc/ptm_dispatch.cuh is the output of tools/gen_dispatch.py."""
    import subprocess
    import sys as _sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([_sys.executable, os.path.join(root, "tools", "gen_dispatch.py")], capture_output=True,
                         text=True, check=True).stdout
    with open(os.path.join(root, "liesel_ptm_b200", "csrc", "ptm_dispatch.cuh")) as f:
        assert f.read() == out



# ---- the JAX side of the boundary: committed, type-checked, not executable here ----
def test_xla_ffi_translation_unit_type_checks_and_matches_fused_py(tmp_path):
    """csrc/ptm_xla_ffi.cc compiles against a stand-in of the XLA FFI header (jax is absent
    from this image) and links against libptm_b200.so; the handler symbols it exports are the
    ones integration/fused.py registers; fused.py byte-compiles."""
    import importlib.util
    import py_compile
    import shutil
    import subprocess

    if shutil.which("g++") is None:
        pytest.skip("g++ not available: the XLA FFI translation unit is NOT type-checked")
    _lib.load()
    src = os.path.join(ROOT, "liesel_ptm_b200", "csrc", "ptm_xla_ffi.cc")
    out = str(tmp_path / "libptm_xla_ffi_stub.so")
    cuda_inc = "/usr/local/cuda/include"
    cmd = ["g++", "-std=c++17", "-Wall", "-Werror", "-shared", "-fPIC", "-DPTM_FFI_STUB",
           "-I" + os.path.join(ROOT, "tests", "ffi_stub"), "-I" + cuda_inc, "-I" + os.path.join(ROOT, "include"),
           src, "-L" + os.path.join(ROOT, "liesel_ptm_b200"), "-lptm_b200", "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    syms = subprocess.run(["nm", "-D", "--defined-only", out], capture_output=True, text=True).stdout
    exported = sorted(ln.split()[-1] for ln in syms.splitlines() if " T Ptm" in ln)
    fused_path = os.path.join(ROOT, "integration", "fused.py")
    py_compile.compile(fused_path, doraise=True)
    spec = importlib.util.spec_from_file_location("ptm_fused", fused_path)
    fused = importlib.util.module_from_spec(spec)
    sys.modules["ptm_fused"] = fused  # dataclasses look the module up
    spec.loader.exec_module(fused)  # top level imports only numpy / ctypes
    assert sorted(fused.all_target_symbols()) == exported
    # every entry point the handlers call is declared in the header
    names = _header_functions()
    for name in fused.FFI_TARGETS:
        assert f"{name}_f64" in names and f"{name}_f32" in names, name
    # without the stub macro and without jax the TU compiles to nothing (guarded)
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I" + cuda_inc,
                        "-I" + os.path.join(ROOT, "include"), src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
