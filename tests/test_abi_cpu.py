"""CPU suite: the C-ABI shared library loads and exports every symbol include/*.h declares, the
pure-host entry points behave, and the product path refuses to run without CUDA (no fallback)."""
import ctypes
import glob
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for hdr in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = open(hdr).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names.update(re.findall(r"\b(b200[a-z]*_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    from b200qp import _lib
    L = _lib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 9
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/ but not exported by {_lib.LIB_PATH}"
    for name in _lib.EXPORTS:
        assert name in declared


def test_version_and_workspace_are_host_only():
    from b200qp import _lib
    L = _lib.lib()
    assert b"sm_100a" in L.b200qp_version()
    pr = _lib.Problem(128, 30, 60, 0, _lib.F64, 20, 3, 0, 1e-12, 900, 30, 1800, 60, 0, 0)
    nbytes = L.b200qp_workspace_bytes(ctypes.byref(pr))
    assert nbytes > 128 * 8 * (30 * 30 + 60 * 30)  # at least Q^-1 and [A;G]Q^-1 per problem
    pr2 = _lib.Problem(256, 30, 60, 0, _lib.F64, 20, 3, 0, 1e-12, 900, 30, 1800, 60, 0, 0)
    assert L.b200qp_workspace_bytes(ctypes.byref(pr2)) > nbytes
    bad = _lib.Problem(0, 30, 60, 0, _lib.F64, 20, 3, 0, 1e-12, 900, 30, 1800, 60, 0, 0)
    assert L.b200qp_workspace_bytes(ctypes.byref(bad)) == 0
    bad = _lib.Problem(4, 30, 60, 0, 7, 20, 3, 0, 1e-12, 900, 30, 1800, 60, 0, 0)
    assert L.b200qp_workspace_bytes(ctypes.byref(bad)) == 0


def test_null_pointers_are_rejected_before_any_launch():
    from b200qp import _lib
    L = _lib.lib()
    pr = _lib.Problem(4, 5, 6, 0, _lib.F64, 20, 3, 0, 1e-12, 25, 5, 30, 6, 0, 0)
    null = ctypes.c_void_p(0)
    rc = L.b200qp_forward(ctypes.byref(pr), *([null] * 13))
    assert rc == -1
    rc = L.b200qp_backward(ctypes.byref(pr), *([null] * 13))
    assert rc == -1
    with pytest.raises(RuntimeError, match="invalid argument"):
        _lib.check(rc, "b200qp_backward")


def test_no_cpu_fallback():
    from b200qp.qp import QPFunction
    from oracle import qp_oracle as O
    Q, p, G, h, A, b = O.random_qp(2, 4, 3, 0, seed=0)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        QPFunction(check_Q_spd=False)(Q, p, G, h, A, b)


def test_shape_errors_match_reference_strings():
    from b200qp.util import expandParam, extract_nBatch, get_sizes
    with pytest.raises(RuntimeError, match="Unexpected number of dimensions."):
        expandParam(torch.zeros(2, 2, 2, 2), 2, 3)
    x, e = expandParam(torch.zeros(3, 3), 5, 3)
    assert e and x.shape == (5, 3, 3)
    assert extract_nBatch(torch.zeros(3, 3), torch.zeros(7, 3), *[torch.zeros(1)] * 4) == 7
    assert get_sizes(torch.zeros(4, 6, 3), torch.zeros(4, 0, 3)) == (6, 3, 0, 4)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the package may import it."""
    pkg = os.path.join(ROOT, "diff-qp-mpc_b200")
    for path in glob.glob(os.path.join(pkg, "**", "*.py"), recursive=True):
        src = open(path).read()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), path
