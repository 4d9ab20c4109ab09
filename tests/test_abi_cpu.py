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


def test_my_envs_host_mirror_selects_the_fused_models():
    """deqmpc/my_envs call surface without a GPU: the env exposes what Tracking_MPC reads, and both
    `env.dynamics` (module) and `env.dynamics_derivatives` (bound method, deqmpc/policies.py:576-577) map to the
    same fused model + parameter vector; a reference-style module whose `.package` is an extension MODULE named
    like the reference's packages is recognised by that name."""
    import types
    import torch
    from b200qp import _lib, my_envs
    from b200qp.envs import dyn_spec
    kw = dict(dtype=torch.float64, device="cpu")
    env = my_envs.CartpoleEnv(nx=4, dt=0.05, kwargs=kw)
    assert (env.nx, env.nu, env.nq, env.T, env.u_bounds) == (4, 1, 2, 200, 100.0)
    assert float(env.Rlqr[0]) == 1e-8 and env.action_space.high[0] == 100.0
    s1, s2 = dyn_spec(env.dynamics), dyn_spec(env.dynamics_derivatives)
    assert s1 == s2 and s1[0] == _lib.ENV_CARTPOLE1L and s1[1][:4] == [0.05, 11.0, 1.0, 2.0] and s1[2:] == (4, 1)
    env2 = my_envs.CartpoleEnv(nx=6, dt=0.03, kwargs=kw)
    assert dyn_spec(env2.dynamics)[0] == _lib.ENV_CARTPOLE2L and dyn_spec(env2.dynamics)[2:] == (6, 1)
    pend = my_envs.PendulumEnv(nx=2, dt=0.05, kwargs=kw)
    assert dyn_spec(pend.dynamics)[0] == _lib.ENV_PENDULUM1L
    ref_like = types.SimpleNamespace(package=types.ModuleType("cartpole1l_v2"), dt=0.05)
    assert dyn_spec(ref_like)[1][1:4] == [0.7, 0.1, 0.05]
    import pytest
    with pytest.raises(NotImplementedError):
        my_envs.package_spec("cartpole3l", 0.05)
    with pytest.raises(RuntimeError, match="CUDA"):
        env.dynamics(torch.zeros(2, 4, dtype=torch.float64), torch.zeros(2, 1, dtype=torch.float64))


def test_new_entry_points_validate_before_any_launch():
    """b200qp_forward_cb_step / _phase_cb (residual callbacks), b200qp_solve_host_* slots, b200data_sample_windows: argument
    validation returns EINVAL (-1) without touching the device."""
    from b200qp import _lib
    L = _lib.lib()
    pr = _lib.Problem(4, 5, 6, 0, _lib.F64, 20, 3, 0, 1e-12, 25, 5, 30, 6, 0, 0)
    null = ctypes.c_void_p(0)
    one = ctypes.c_void_p(16)   # never dereferenced: validation fails first
    assert L.b200qp_forward_cb_step(ctypes.byref(pr), 0, null, one, null) == -1          # no x_out
    assert L.b200qp_forward_cb_step(ctypes.byref(pr), 20, one, one, null) == -1          # iteration out of range
    assert L.b200qp_forward_phase_cb(ctypes.byref(pr), -1001, *([one] * 12), null, null, null) == -1   # BEGIN is not a cb phase
    assert L.b200qp_solve_host_wait(_lib.HOST_SLOTS) == -1 and L.b200qp_solve_host_wait(-1) == -1
    assert L.b200qp_solve_host_wait(_lib.HOST_SLOTS - 1) == 0                            # idle slot: nothing to wait for
    assert L.b200data_sample_windows(null, one, one, 10, 2, 1, one, 8, 4, 3, 0, one, one, one, one, one, null) == -1
    assert L.b200data_sample_windows(one, one, one, 10, 2, 1, one, 8, 4, 3, 5, one, one, one, one, one, null) == -1   # bad mode
    assert _lib.FLAG_FACTORED_GRAD == 4 and _lib.HOST_SLOTS == 4
