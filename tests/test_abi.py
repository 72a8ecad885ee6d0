"""CPU tests of the drop-in boundary: the shared library builds, loads, and exports every symbol that
include/moda_b200.h declares (and nothing the Python binding expects is missing).  No compute calls."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    from moda_b200 import build
    return build.build()


def _declared():
    src = open(os.path.join(ROOT, "include", "moda_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(moda_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(libpath):
    L = ctypes.CDLL(libpath)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), "declared in include/moda_b200.h but not exported: %s" % n


def test_binding_matches_header(libpath):
    from moda_b200 import _lib
    declared = set(_declared())
    for n in _lib.SIGNATURES:
        assert n in declared, "bound in moda_b200/_lib.py but not declared in the header: %s" % n
    for n in declared - {"moda_version", "moda_last_error"}:
        assert n in _lib.SIGNATURES, "declared in the header but not bound: %s" % n


def test_version_and_error_strings(libpath):
    L = ctypes.CDLL(libpath)
    L.moda_version.restype = ctypes.c_char_p
    L.moda_last_error.restype = ctypes.c_char_p
    assert b"sm_100a" in L.moda_version()
    assert isinstance(L.moda_last_error(), bytes)


def test_argument_checks_fail_loudly_without_gpu(libpath):
    """Argument validation happens before any launch, so it is testable on a CPU-only box."""
    from moda_b200 import _lib
    L = _lib.lib()
    rc = L.moda_skin_warp_fwd(None, None, None, None, None, None, None, None, 4, 8, 200, 0, 0, 0, 0, None)
    assert rc < 0 and b"B=200" in L.moda_last_error()


def test_product_refuses_cpu_tensors(libpath):
    import torch
    from moda_b200.nerf import Embedding, NeRF
    from moda_b200 import geom_utils as G
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Embedding(3, 10)(torch.zeros(2, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        NeRF()(torch.zeros(2, 90))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        G.skinning(torch.zeros(25, 10), torch.zeros(2, 4, 3), None, skin_aux=torch.zeros(2))


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under moda_b200/ may reference it."""
    pkg = os.path.join(ROOT, "moda_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
