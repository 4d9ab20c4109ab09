"""Generate tests/golden/dyn_envsv1_{cartpole1l,cartpole2l}.npz from the REAL reference modules
deqmpc/envs_v1.py:OneLinkCartpoleDynamics / TwoLinkCartpoleDynamics (closed-form accelerations under classical RK4): next
state, and the Jacobians autograd gives through the reference's own forward (the pattern of deqmpc/envs.py:74-82).
Build container only.  TEST INFRASTRUCTURE, NOT PRODUCT."""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "_stubs"))
for p in ("/root/reference", "/root/reference/deqmpc"):
    sys.path.append(p)
warnings.filterwarnings("ignore")

from gen_golden_envdx import jac  # noqa: E402


def main():
    import envs_v1 as V1
    gold = os.path.join(ROOT, "tests", "golden")
    rs = np.random.RandomState(11)
    N = 200
    for name, mod, nx, umax in (("cartpole1l", V1.OneLinkCartpoleDynamics(), 4, 50.0), ("cartpole2l", V1.TwoLinkCartpoleDynamics(), 6, 5.0)):
        nq = nx // 2
        x = np.concatenate([rs.uniform(-np.pi, np.pi, (N, nq)), rs.uniform(-3, 3, (N, nq))], 1)
        u = rs.uniform(-umax, umax, (N, 1))
        xn, A, B = jac(mod, torch.tensor(x), torch.tensor(u))
        np.savez_compressed(os.path.join(gold, f"dyn_envsv1_{name}.npz"), x=x, u=u, xn=xn.numpy(), A=A.numpy(), B=B.numpy())
        print(f"{name}: |xn| {float(xn.norm()):.9f} |A| {float(A.norm()):.9f} |B| {float(B.norm()):.9f}")


if __name__ == "__main__":
    main()
