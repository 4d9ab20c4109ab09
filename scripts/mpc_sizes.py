"""Time the fused AL-MPC solve (forward kernel only and forward+backward through the drop-in) at
the BASELINE config shapes.  Run on the GPU box."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "diff-qp-mpc_b200"))
import torch
from b200qp import envs
from b200qp.AL_mpc import MPC
from b200qp.al_utils import QuadCost

dev = torch.device("cuda:0")


def run(name, dx, dxj, nx, nu, T, B, x0, u0, qd, ub_lo, ub_hi, reps=5):
    Cd = torch.tensor(qd, dtype=torch.float64, device=dev).repeat(B, T, 1)
    ctrl = MPC(nx, nu, T, u_lower=ub_lo, u_upper=ub_hi, n_batch=B, u_init=u0, eps=1e-5, dtype=torch.float64)
    Cfull = torch.diag_embed(Cd).requires_grad_(True)
    c = torch.zeros(B, T, nx + nu, dtype=torch.float64, device=dev, requires_grad=True)

    def step():
        ctrl.reinitialize(x0, None)
        ctrl.u_init = u0
        Cfull.grad = None; c.grad = None
        x, u = ctrl(x0, QuadCost(Cfull, c), dx, dxj)
        (x.sum() + u.sum()).backward()
        return x

    step(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        x = step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:34s} B={B:5d} T={T:3d}  {ms:9.3f} ms/call  {B / ms * 1e3:12.0f} rollouts/s  finite={bool(torch.isfinite(x).all())}", flush=True)


torch.manual_seed(0)
B, T = 1024, 5
x0 = torch.stack((torch.rand(B, dtype=torch.float64) * 6.28 - 3.14, torch.rand(B, dtype=torch.float64) * 2 - 1), 1).to(dev)
one = lambda v, n: v * torch.ones(n, dtype=torch.float64, device=dev)
run("pendulum (cfg[1])", envs.PendulumDynamics(), envs.PendulumDynamics_jac(), 2, 1, T, B, x0,
    torch.randn(B, T, 1, dtype=torch.float64, device=dev), [10., 1., 0.01], one(-3., 1), one(3., 1))
B, T = 4096, 20
th = torch.rand(B, dtype=torch.float64) * 0.6 - 0.3
x0 = torch.stack((torch.rand(B, dtype=torch.float64) - 0.5, torch.zeros(B, dtype=torch.float64), torch.cos(th), torch.sin(th),
                  torch.zeros(B, dtype=torch.float64)), 1).to(dev)
run("cartpole env_dx (cfg[2] shape)", envs.CartpoleDx(), envs.CartpoleDx_jac(), 5, 1, T, B, x0,
    0.1 * torch.randn(B, T, 1, dtype=torch.float64, device=dev), [0.1, 0.1, 1., 1., 0.1, 0.001], one(-100., 1), one(100., 1))
for B in (1024, 4096, 8192):
    T = 40
    x0 = torch.cat((torch.rand(B, 3, dtype=torch.float64) * 2 - 1, torch.rand(B, 3, dtype=torch.float64) * 0.4 - 0.2,
                    torch.rand(B, 6, dtype=torch.float64) * 0.4 - 0.2), 1).to(dev)
    run("rex quadrotor (cfg[3] shape)", envs.RexQuadrotor_dynamics(), envs.RexQuadrotor_dynamics_jac(), 12, 4, T, B, x0,
        14.9 + 0.1 * torch.randn(B, T, 4, dtype=torch.float64, device=dev), [10.] * 3 + [0.01] * 3 + [1.] * 3 + [0.01] * 3 + [1e-4] * 4,
        one(11.5, 4), one(18.3, 4), reps=2)
