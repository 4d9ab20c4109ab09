"""Generate tests/golden/tracking_pend_B8_T5.npz from the REAL reference's deqmpc Tracking_MPC
(policies.py:560-690) on its PendulumEnv (jit-scripted dynamics).  Build container only."""
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "_stubs"))
for p in ("/root/reference", "/root/reference/deqmpc"):
    sys.path.append(p)
warnings.filterwarnings("ignore")


def main():
    import envs as RE
    import policies as RP
    env = RE.PendulumEnv(stabilization=False)
    env.nq = 1
    B, T = 8, 5
    args = types.SimpleNamespace(T=T, bsz=B, Q=env.Qlqr.double(), R=env.Rlqr.double(), dtype="double", solver_type="al",
                                 qp_iter=2, eps=1e-2, warm_start=True, device=torch.device("cpu"))
    torch.manual_seed(0)
    mpc = RP.Tracking_MPC(args, env)
    rs = np.random.RandomState(5)
    x0 = torch.tensor(np.stack([rs.uniform(-2, 2, B), rs.uniform(-1, 1, B)], 1))
    save = dict(x0=x0.numpy())
    mpc.reinitialize(x0, None)
    for k in range(3):
        x_ref = torch.tensor(0.3 * rs.randn(B, T, 2)).requires_grad_(True)
        u_ref = torch.tensor(0.3 * rs.randn(B, T, 1)).requires_grad_(True)
        xs, us = mpc(x0, None, x_ref, u_ref)
        (xs.sum() + 2 * us.sum()).backward()
        save.update({f"x_ref{k}": x_ref.detach().numpy(), f"u_ref{k}": u_ref.detach().numpy(), f"out_x{k}": xs.detach().numpy(),
                     f"out_u{k}": us.detach().numpy(), f"g_xref{k}": x_ref.grad.numpy(), f"g_uref{k}": u_ref.grad.numpy()})
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "tracking_pend_B8_T5.npz"), **save)
    print("Tracking_MPC golden written; |x2|", float(xs.norm()), "|g_xref2|", float(x_ref.grad.norm()))


if __name__ == "__main__":
    main()
