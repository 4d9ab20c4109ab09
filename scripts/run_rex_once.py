"""Tiny driver for profiling: the rex-quadrotor AL-MPC solve (BASELINE configs[3] shape, T = 40) a few times."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "diff-qp-mpc_b200"))
import torch
from b200qp import envs
from b200qp.AL_mpc import MPC
from b200qp.al_utils import QuadCost
dev = torch.device("cuda:0")
B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, 40
g = torch.Generator(device="cpu").manual_seed(0)
r = lambda *sh: torch.rand(*sh, generator=g, dtype=torch.float64)
one = lambda v, n: v * torch.ones(n, dtype=torch.float64, device=dev)
x0 = torch.cat((r(B, 3) * 2 - 1, r(B, 3) * 0.2 - 0.1, r(B, 6) * 0.2 - 0.1), 1).to(dev)
u0 = (14.9 + 0.1 * (r(B, T, 4) - 0.5)).to(dev)
Cd = torch.tensor([10.] * 3 + [0.01] * 3 + [1.] * 3 + [0.01] * 3 + [1e-4] * 4, dtype=torch.float64, device=dev).repeat(B, T, 1)
ctrl = MPC(12, 4, T, u_lower=one(11.5, 4), u_upper=one(18.3, 4), n_batch=B, u_init=u0, eps=1e-5, dtype=torch.float64)
dx, dxj = envs.RexQuadrotor_dynamics(), envs.RexQuadrotor_dynamics_jac()
for _ in range(2):
    ctrl.reinitialize(x0, None); ctrl.u_init = u0
    with torch.no_grad():
        x, u = ctrl(x0, QuadCost(torch.diag_embed(Cd), torch.zeros(B, T, 16, dtype=torch.float64, device=dev)), dx, dxj)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ctrl.reinitialize(x0, None); ctrl.u_init = u0
e0.record()
with torch.no_grad():
    x, u = ctrl(x0, QuadCost(torch.diag_embed(Cd), torch.zeros(B, T, 16, dtype=torch.float64, device=dev)), dx, dxj)
e1.record()
torch.cuda.synchronize()
print("B", B, "forward ms", e0.elapsed_time(e1), "rollouts/s", B / (e0.elapsed_time(e1) * 1e-3))
print("finite", bool(torch.isfinite(x).all()), float(x.double().norm()))
