"""Generate tests/golden/dyn_rex.npz and tests/golden/mpc_rex_B4_T8.npz from the REAL reference
(deqmpc/rex_quadrotor.py RexQuadrotor_dynamics(_jac) and qpth.AL_mpc.MPC on it).  Build container
only (needs /root/reference).  TEST INFRASTRUCTURE, NOT PRODUCT."""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "_stubs"))
for p in ("/root/reference", "/root/reference/deqmpc"):
    sys.path.append(p)
warnings.filterwarnings("ignore")


def main():
    import rex_quadrotor as RQ
    from qpth import AL_mpc as al_mpc, al_utils
    rs = np.random.RandomState(0)
    dyn, dynj = RQ.RexQuadrotor_dynamics(), RQ.RexQuadrotor_dynamics_jac()
    N = 64
    x = np.concatenate([rs.uniform(-2, 2, (N, 3)), rs.uniform(-0.4, 0.4, (N, 3)), rs.uniform(-0.5, 0.5, (N, 3)),
                        rs.uniform(-0.25, 0.25, (N, 3))], 1)
    u = rs.uniform(11.5, 18.3, (N, 4))
    xt, ut = torch.tensor(x), torch.tensor(u)
    xn = dyn(xt, ut)
    xr, ur = xt.clone().requires_grad_(True), ut.clone().requires_grad_(True)
    out, (A, B) = dynj(xr, ur)
    assert torch.equal(out.detach(), xn.detach())
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "dyn_rex.npz"), x=x, u=u, xn=xn.detach().numpy(),
                        A=A.detach().numpy(), B=B.detach().numpy())
    print("rex dynamics: |xn|", float(xn.norm()), "|A|", float(A.norm()), "|B|", float(B.norm()))

    # AL-MPC on the rex quadrotor: cold call + backward
    Bsz, T, nx, nu = 4, 8, 12, 4
    x0 = torch.tensor(x[:Bsz])
    u_init = torch.tensor(rs.uniform(14.0, 16.0, (Bsz, T, nu)))
    Cd = torch.tensor(np.tile(np.array([10.0] * 3 + [0.01] * 3 + [1.0] * 3 + [0.01] * 3 + [1e-4] * 4), (Bsz, T, 1)))
    xref = torch.tensor(np.concatenate([0.1 * rs.randn(Bsz, T, nx), 14.9 + 0.1 * rs.randn(Bsz, T, nu)], 2))
    ul, uu = 11.5 * torch.ones(nu, dtype=torch.float64), 18.3 * torch.ones(nu, dtype=torch.float64)
    ctrl = al_mpc.MPC(nx, nu, T, u_lower=ul, u_upper=uu, exit_unconverged=False, eps=1e-5, n_batch=Bsz, backprop=False,
                      verbose=0, u_init=u_init.clone(), solver_type="dense", dtype=torch.float64)
    ctrl.reinitialize(x0, None)
    ctrl.u_init = u_init.clone()
    Cfull = torch.diag_embed(Cd).clone().requires_grad_(True)
    c = (-(Cd * xref)).clone().requires_grad_(True)
    xs, us = ctrl(x0, al_utils.QuadCost(Cfull, c), dyn, dynj)
    (xs.sum() + us.sum()).backward()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "mpc_rex_B4_T8.npz"), x0=x0.numpy(), u_init=u_init.numpy(),
                        Cd=Cd.numpy(), xref=xref.numpy(), out_x=xs.detach().numpy(), out_u=us.detach().numpy(),
                        out_lam=ctrl.lamda_prev.numpy(), out_rho=ctrl.rho_prev.numpy(),
                        out_dC=Cfull.grad.diagonal(dim1=-2, dim2=-1).numpy(), out_dc=c.grad.numpy())
    print("rex MPC: |x|", float(xs.norm()), "|u|", float(us.norm()), "|lam|", float(ctrl.lamda_prev.norm()))

    # BASELINE configs[3] horizon (T = 40: the per-problem state no longer fits in shared memory, the kernel
    # works out of its global scratch slab), bench-like inputs
    torch.manual_seed(0)
    Bsz, T = 4, 40
    x0 = torch.cat((torch.rand(Bsz, 3, dtype=torch.float64) * 2 - 1, torch.rand(Bsz, 3, dtype=torch.float64) * 0.4 - 0.2,
                    torch.rand(Bsz, 6, dtype=torch.float64) * 0.4 - 0.2), 1)
    u_init = 14.9 + 0.1 * torch.randn(Bsz, T, 4, dtype=torch.float64)
    Cd = torch.tensor([10.] * 3 + [0.01] * 3 + [1.] * 3 + [0.01] * 3 + [1e-4] * 4, dtype=torch.float64).repeat(Bsz, T, 1)
    xref = torch.zeros(Bsz, T, nx + nu, dtype=torch.float64)
    ctrl = al_mpc.MPC(nx, nu, T, u_lower=ul, u_upper=uu, exit_unconverged=False, eps=1e-5, n_batch=Bsz, backprop=False,
                      verbose=0, u_init=u_init.clone(), solver_type="dense", dtype=torch.float64)
    ctrl.reinitialize(x0, None)
    ctrl.u_init = u_init.clone()
    Cfull = torch.diag_embed(Cd).clone().requires_grad_(True)
    c = (-(Cd * xref)).clone().requires_grad_(True)
    xs, us = ctrl(x0, al_utils.QuadCost(Cfull, c), dyn, dynj)
    (xs.sum() + us.sum()).backward()
    assert torch.isfinite(xs).all() and torch.isfinite(us).all()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "mpc_rex_B4_T40.npz"), x0=x0.numpy(), u_init=u_init.numpy(),
                        Cd=Cd.numpy(), xref=xref.numpy(), out_x=xs.detach().numpy(), out_u=us.detach().numpy(),
                        out_lam=ctrl.lamda_prev.numpy(), out_rho=ctrl.rho_prev.numpy(),
                        out_dC=Cfull.grad.diagonal(dim1=-2, dim2=-1).numpy(), out_dc=c.grad.numpy())
    print("rex MPC T=40: |x|", float(xs.norm()), "|u|", float(us.norm()), "|lam|", float(ctrl.lamda_prev.norm()))


if __name__ == "__main__":
    main()
