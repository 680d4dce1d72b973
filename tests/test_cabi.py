"""The C-ABI shared library loads and exports every symbol include/mupopp_b200.h declares (no compute, no GPU)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from mupopp_b200 import _lib, build
    build.build()
    return _lib.load()


def _declared():
    txt = open(os.path.join(ROOT, "include", "mupopp_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mp_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(lib):
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), n


def test_bindings_cover_header(lib):
    from mupopp_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()


def test_fails_loudly_without_gpu(lib):
    import ctypes as C
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = lib.mp_ctx_create(0, None, C.byref(h))
    assert rc != 0
    assert b"no CPU fallback" in lib.mp_last_error()
    from mupopp_b200 import _lib
    from mupopp_b200.device import Context
    with pytest.raises(_lib.MpError):
        Context(0)


def test_struct_sizes_match_header():
    """ctypes mirrors of the header structs (layout is what crosses the boundary)."""
    import ctypes as C
    from mupopp_b200 import _lib
    assert C.sizeof(_lib.MpMesh) == 56
    assert C.sizeof(_lib.MpTransportParams) == 8 * 15 + 16
    assert C.sizeof(_lib.MpCsr) == 96
    assert C.sizeof(_lib.MpSolveInfo) == 40
    assert C.sizeof(_lib.MpFunctional) == 184
