"""CPU tests of the my_envs oracle: the re-derived rigid-body models (oracle/myenvs_oracle.py
PortPackage) against golden vectors of the reference's own CasADi-generated code (oracle/gen_golden_myenvs.py),
and -- where oracle/_ref was built -- against that code directly on fresh samples."""
import os

import numpy as np
import pytest

from oracle import myenvs_oracle as MO

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODELS = list(MO.NQ)


def _rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("name", MODELS)
def test_port_matches_reference_golden(name):
    g = np.load(os.path.join(GOLD, f"dyn_myenvs_{name}.npz"))
    port = MO.Dynamics(MO.PortPackage(name), 2 * MO.NQ[name], float(g["dt"]))
    xn, (A, B) = port.dynamics_derivatives(g["x"], g["u"])
    assert _rel(xn, g["xn"]) < 1e-13 and _rel(A, g["A"]) < 1e-12 and _rel(B, g["B"]) < 1e-12


@pytest.mark.parametrize("name", MODELS)
def test_port_matches_compiled_reference(name):
    if not os.path.exists(os.path.join(os.path.dirname(MO.__file__), "_ref", f"lib{name}_ref.so")):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    rs = np.random.RandomState(7)
    n, N = MO.NQ[name], 100
    q, qd, tau = rs.uniform(-6, 6, (N, n)), rs.uniform(-5, 5, (N, n)), rs.uniform(-20, 20, (N, n))
    h = rs.uniform(0.01, 0.06, (N, 1))          # the generated code takes the step per row
    R, P = MO.RefPackage(name), MO.PortPackage(name)
    for a, b in zip(R.dynamics(q, qd, tau, h), P.dynamics(q, qd, tau, h)):
        assert _rel(b, a) < 1e-13
    for a, b in zip(R.derivatives(q, qd, tau, h), P.derivatives(q, qd, tau, h)):
        assert np.abs(b - a).max() < 1e-12 * max(1.0, np.abs(a).max())


def test_jacobian_blocks_are_out_by_in():
    """dynamics.py:100-108 ends with transposes: A[b, i, j] = d xnext_i / d x_j (finite-difference check)."""
    name = "cartpole1l"
    d = MO.Dynamics(MO.PortPackage(name), 4, 0.05)
    rs = np.random.RandomState(1)
    x, u = rs.uniform(-1, 1, (3, 4)), rs.uniform(-5, 5, (3, 1))
    _, (A, B) = d.dynamics_derivatives(x, u)
    eps = 1e-6
    for j in range(4):
        e = np.zeros(4); e[j] = eps
        fd = (d.forward(x + e, u) - d.forward(x - e, u)) / (2 * eps)
        assert np.abs(fd - A[:, :, j]).max() < 1e-7
    fd = (d.forward(x, u + eps) - d.forward(x, u - eps)) / (2 * eps)
    assert np.abs(fd - B[:, :, 0]).max() < 1e-7
