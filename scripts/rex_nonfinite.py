"""Which problems of a bench-like rex-quadrotor batch end non-finite?  Saves their inputs for a reference run."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "diff-qp-mpc_b200"))
import numpy as np, torch
from b200qp import envs
from b200qp.AL_mpc import MPC
from b200qp.al_utils import QuadCost
dev = torch.device("cuda:0")
B, T, nx, nu = 1024, 40, 12, 4
one = lambda v, n: v * torch.ones(n, dtype=torch.float64, device=dev)
for seed in range(1, 8):
    torch.manual_seed(seed)
    x0 = torch.cat((torch.rand(B, 3, dtype=torch.float64) * 2 - 1, torch.rand(B, 3, dtype=torch.float64) * 0.4 - 0.2,
                    torch.rand(B, 6, dtype=torch.float64) * 0.4 - 0.2), 1).to(dev)
    u0 = (14.9 + 0.1 * torch.randn(B, T, 4, dtype=torch.float64)).to(dev)
    Cd = torch.tensor([10.] * 3 + [0.01] * 3 + [1.] * 3 + [0.01] * 3 + [1e-4] * 4, dtype=torch.float64, device=dev).repeat(B, T, 1)
    ctrl = MPC(nx, nu, T, u_lower=one(11.5, 4), u_upper=one(18.3, 4), n_batch=B, u_init=u0, eps=1e-5, dtype=torch.float64)
    ctrl.reinitialize(x0, None); ctrl.u_init = u0
    x, u = ctrl(x0, QuadCost(torch.diag_embed(Cd), torch.zeros(B, T, nx + nu, dtype=torch.float64, device=dev)),
                envs.RexQuadrotor_dynamics(), envs.RexQuadrotor_dynamics_jac())
    bad = (~torch.isfinite(x).reshape(B, -1).all(1)) | (~torch.isfinite(u).reshape(B, -1).all(1))
    idx = bad.nonzero().flatten().cpu()
    print("seed", seed, "non-finite problems:", int(bad.sum()), "of", B, "first:", idx[:8].tolist(),
          "status:", ctrl.status[idx[:8].to(dev)].tolist() if len(idx) else [], flush=True)
    if len(idx):
        good = (~bad).nonzero().flatten().cpu()[:2]
        sel = torch.cat((idx[:6], good))
        np.savez(os.path.join(ROOT, "gpurun_out", "rex_bad.npz"), sel=sel.numpy(), x0=x0[sel.to(dev)].cpu().numpy(),
                 u0=u0[sel.to(dev)].cpu().numpy(), x=x[sel.to(dev)].cpu().numpy(), u=u[sel.to(dev)].cpu().numpy(),
                 lam=ctrl.lamda_prev[sel.to(dev)].cpu().numpy())
        break
