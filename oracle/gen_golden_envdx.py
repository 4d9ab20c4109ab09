"""Generate tests/golden/dyn_envdx_{pendulum,cartpole}.npz from the REAL reference modules
qpth/env_dx/pendulum.py:PendulumDx and qpth/env_dx/cartpole.py:CartpoleDx (imported from /root/reference with the
matplotlib stub of oracle/_stubs): next state, and the Jacobians the reference itself would obtain -- autograd
through its own forward (the pattern of deqmpc/envs.py:74-82; env_dx ships no dx_jac).
Build container only.  TEST INFRASTRUCTURE, NOT PRODUCT."""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "_stubs"))
sys.path.insert(1, "/root/reference")
warnings.filterwarnings("ignore")


def jac(mod, x, u):
    xr, ur = x.clone().requires_grad_(True), u.clone().requires_grad_(True)
    out = mod(xr, ur)
    nx = out.shape[1]
    A = torch.stack([torch.autograd.grad(out[:, i].sum(), xr, retain_graph=True)[0] for i in range(nx)], 1)
    B = torch.stack([torch.autograd.grad(out[:, i].sum(), ur, retain_graph=True)[0] for i in range(nx)], 1)
    return out.detach(), A, B


def main():
    # the reference builds its parameters with torch.Tensor(...) = float32 (default dtype left alone on purpose):
    # with float64 states they are promoted, i.e. the model constants are the float32-rounded values
    from qpth.env_dx.pendulum import PendulumDx
    from qpth.env_dx.cartpole import CartpoleDx
    gold = os.path.join(ROOT, "tests", "golden")
    rs = np.random.RandomState(7)
    N = 200
    for name, mod, nx in (("pendulum", PendulumDx(), 3), ("cartpole", CartpoleDx(), 5)):
        x = rs.randn(N, nx)
        # controls on both sides of the reference's clamp (pendulum +-2, cart-pole +-100 as coded)
        u = rs.randn(N, 1) * (1.5 if name == "pendulum" else 60.0)
        xn, A, B = jac(mod, torch.tensor(x), torch.tensor(u))
        np.savez_compressed(os.path.join(gold, f"dyn_envdx_{name}.npz"), x=x, u=u, xn=xn.numpy(), A=A.numpy(), B=B.numpy())
        print(f"{name}: |xn| {float(xn.norm()):.9f} |A| {float(A.norm()):.9f} |B| {float(B.norm()):.9f} "
              f"clamped {int((B.abs().sum((1, 2)) == 0).sum())}/{N}")


if __name__ == "__main__":
    main()
