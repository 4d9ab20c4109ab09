"""CPU suite: oracle/mpc_oracle.py against the golden vectors of the real reference's
qpth.AL_mpc.MPC (cold call + warm-started calls, forward outputs, solver state and the implicit
backward)."""
import pytest
import torch

from oracle import mpc_oracle as MO
from tests.mpc_cases import CASES, load, oracle_dyn, rel


@pytest.mark.parametrize("case", list(CASES))
def test_mpc_oracle_matches_reference_golden(case):
    g = load(case)
    dyn = oracle_dyn(case)
    B, T = g["u_init"].shape[:2]
    nx, nu = dyn.nx, dyn.nu
    ub = float(g["ub"]) * torch.ones(nu, dtype=torch.float64)
    st = MO.ALState(B, T * nx + 2 * T * nu)
    u, x = g["u_init"].clone(), None
    for k in range(int(g["n_calls"])):
        C = g["Cd"]
        c = -(C * g[f"xref{k}"])
        if x is None:
            x = MO.rollout(g["x0"], u, dyn)
        xs, us, ctx = MO.al_solve(x.double(), u.double(), g["x0"], C, c, dyn, -ub, ub, st)
        dC, dc = MO.al_backward(ctx, torch.ones(B, T, nx + nu, dtype=torch.float64))
        assert rel(xs.float(), g[f"out_x{k}"]) < 1e-6 and rel(us.float(), g[f"out_u{k}"]) < 1e-6
        assert rel(st.lam, g[f"out_lam{k}"]) < 1e-8 and rel(st.rho, g[f"out_rho{k}"]) == 0
        assert rel(dC, g[f"out_dC{k}"]) < 1e-8 and rel(dc, g[f"out_dc{k}"]) < 1e-8
        x, u = xs.float(), us.float()


def test_block_tridiagonal_structure_of_the_dense_hessian():
    """The kernels factor H block-tridiagonally; check on the oracle's DENSE H that everything
    outside the block tridiagonal band is exactly zero and the sub-diagonal block only has nx rows."""
    g = load("pend_B16_T5")
    dyn = oracle_dyn("pend_B16_T5")
    B, T = g["u_init"].shape[:2]
    nx, nu, nt = dyn.nx, dyn.nu, dyn.nx + dyn.nu
    xu = torch.cat((MO.rollout(g["x0"], g["u_init"], dyn), g["u_init"]), 2)
    lam = torch.randn(B, T * nx + 2 * T * nu, dtype=torch.float64)
    rho = torch.full((B, 1), 3.0, dtype=torch.float64)
    ub = 0.1 * torch.ones(nu, dtype=torch.float64)
    _, H = MO.merit_grad_hess(xu, g["Cd"], g["Cd"] * 0, g["x0"], lam, rho, dyn, -ub, ub)
    for t in range(T):
        for s in range(T):
            blk = H[:, t * nt:(t + 1) * nt, s * nt:(s + 1) * nt]
            if abs(t - s) > 1:
                assert blk.abs().max() == 0
            if t == s + 1:
                assert blk[:, nx:, :].abs().max() == 0
