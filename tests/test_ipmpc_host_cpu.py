"""CPU suite: the host-side logic of b200qp/qp_wrapper.py (block-structured QP assembly, the residual callback on LinDx
dynamics, the cost) against matrices assembled by the real reference (oracle/gen_golden_ipmpc_matrices.py).  Pure torch
index arithmetic: bit-exact.  (The solve itself needs the CUDA library and is tested in tests/test_mpc_parity_gpu.py.)"""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_dense_qp_assembly_matches_reference():
    from b200qp import qp_wrapper as ip_mpc
    g = {k: torch.tensor(v) for k, v in np.load(os.path.join(GOLDEN, "ipmpc_matrices.npz")).items()}
    Tm1, B, nx, nt = g["F"].shape
    nu, T = nt - nx, Tm1 + 1
    ctrl = ip_mpc.MPC(nx, nu, T, u_lower=g["ul"], u_upper=g["uu"], n_batch=B, solver_type="dense")
    Q, q = ctrl.compute_Qq_dense(g["C"], g["c"])
    A, b = ctrl.compute_Ab_dense(g["F"], g["f"], g["x0"])
    G, h = ctrl.compute_Gh_dense(g["x0"])
    for name, ours in (("Q", Q), ("q", q), ("A", A), ("b", b), ("G", G), ("h", h)):
        assert torch.equal(ours, g[name]), name
    res = ctrl.dyn_res(g["z"], ip_mpc.LinDx(g["F"], g["f"]), g["x0"])
    assert torch.equal(res, g["dyn_res"])
    assert torch.allclose(res, torch.bmm(A, g["z"].unsqueeze(-1)).squeeze(-1) - b, atol=1e-12)
    cost = ctrl.compute_cost(g["z"].reshape(B, T, nt), ip_mpc.QuadCost(g["C"], g["c"]))
    assert torch.allclose(cost, g["cost"], rtol=1e-14, atol=0)
    # roll-out on linear dynamics (qp_wrapper.py:604-617)
    u = g["z"].reshape(B, T, nt)[:, :, nx:].transpose(0, 1)
    xs = ctrl.rollout(g["x0"], u, ip_mpc.LinDx(g["F"], g["f"]))
    assert xs.shape == (T, B, nx) and torch.equal(xs[0], g["x0"])
    x1 = torch.bmm(g["F"][0], torch.cat([g["x0"], u[0]], -1).unsqueeze(-1)).squeeze(-1) + g["f"][0]
    assert torch.allclose(xs[1], x1)


def test_unsupported_modes_raise():
    import pytest
    from b200qp import qp_wrapper as ip_mpc
    with pytest.raises(NotImplementedError):
        ip_mpc.MPC(2, 1, 5, grad_method=ip_mpc.GradMethods.AUTO_DIFF)
    with pytest.raises(NotImplementedError):
        ip_mpc.MPC(2, 1, 5, slew_rate_penalty=0.1)
