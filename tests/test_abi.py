"""The C-ABI library loads without a GPU and exports every symbol include/ruvnb.h declares (no compute)."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "ruvnb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ruvnb_[A-Za-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    from ruviiinb_b200 import _lib
    names = _declared()
    assert len(names) >= 30
    assert sorted(_lib.SYMBOLS) == names, set(names) ^ set(_lib.SYMBOLS)
    for n in names:
        assert getattr(lib, n) is not None


def test_abi_version_and_defaults(lib):
    from ruviiinb_b200 import _lib
    assert lib.ruvnb_abi_version() == 3
    o = _lib.default_opts()
    # defaults of ruvIII.nb (R/ruvIIInb.R:27-28)
    assert (o.k, o.robust, o.lambda_a, o.lambda_b, o.step_fac, o.inner_maxit, o.outer_maxit) == (2, 0, 0.01, 16.0, 0.5, 50, 25)
    assert (o.inner_tol, o.outer_tol, o.psi_tol) == (1e-4, 1e-4, 1e-8)


def test_no_cpu_fallback(lib):
    """Without a CUDA device ruvnb_create fails loudly; with one it succeeds."""
    import torch
    from ruviiinb_b200 import _lib
    if torch.cuda.is_available():
        _lib.Context(0).close()
        return
    h = C.c_void_p()
    assert lib.ruvnb_create(C.byref(h), 0) < 0 and not h.value
    try:
        _lib.Context(0)
    except _lib.RuvnbError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("Context() must raise without a GPU")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "ruviiinb_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt, f
