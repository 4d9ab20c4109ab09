"""Tiny driver for profiling: the my_envs cart-pole AL-MPC solve (BASELINE configs[2] shape: T = 20, B = 4096)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "diff-qp-mpc_b200"))
import torch
from b200qp import my_envs
from b200qp.AL_mpc import MPC
from b200qp.al_utils import QuadCost
dev = torch.device("cuda:0")
B, T, nx = 4096, 20, 4
g = torch.Generator(device="cpu").manual_seed(0)
r = lambda *sh: torch.rand(*sh, generator=g, dtype=torch.float64)
one = lambda v, n: v * torch.ones(n, dtype=torch.float64, device=dev)
d = my_envs.CartpoleDynamics(nx=nx, dt=0.05, kwargs=dict(dtype=torch.float64, device=dev))
x0 = torch.cat((r(B, nx // 2) * 0.6 - 0.3, r(B, nx // 2) * 0.2 - 0.1), 1).to(dev)
u0 = (0.1 * (r(B, T, 1) - 0.5)).to(dev)
Cd = torch.tensor([1.] * nx + [1e-8], dtype=torch.float64, device=dev).repeat(B, T, 1)
ctrl = MPC(nx, 1, T, u_lower=one(-100., 1), u_upper=one(100., 1), n_batch=B, u_init=u0, eps=1e-5, dtype=torch.float64)
for _ in range(2):
    ctrl.reinitialize(x0, None); ctrl.u_init = u0
    with torch.no_grad():
        x, u = ctrl(x0, QuadCost(torch.diag_embed(Cd), torch.zeros(B, T, nx + 1, dtype=torch.float64, device=dev)), d, d.dynamics_derivatives)
torch.cuda.synchronize()
print("finite", bool(torch.isfinite(x).all()), float(x.double().norm()))
