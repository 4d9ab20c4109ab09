"""Generate tests/golden/ipmpc_matrices.npz: the dense QP matrices the REAL reference's qpth.qp_wrapper.MPC assembles
(compute_Qq_dense / compute_Ab_dense / compute_Gh_dense, qp_wrapper.py:639-680) and its dyn_res / compute_cost on LinDx
dynamics, for random inputs -- the host-side logic of b200qp/qp_wrapper.py, checkable without a GPU.  Build container only."""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "_stubs"))
sys.path.insert(1, "/root/reference")
warnings.filterwarnings("ignore")


def main():
    from qpth import qp_wrapper as ip_mpc
    rs = np.random.RandomState(3)
    nx, nu, T, B = 4, 1, 6, 3
    nt = nx + nu
    ul, uu = -2.0 * torch.ones(nu, dtype=torch.float64), 1.5 * torch.ones(nu, dtype=torch.float64)
    ctrl = ip_mpc.MPC(nx, nu, T, u_lower=ul, u_upper=uu, n_batch=B, solver_type="dense")
    F = torch.tensor(rs.randn(T - 1, B, nx, nt))
    f = torch.tensor(rs.randn(T - 1, B, nx))
    x0 = torch.tensor(rs.randn(B, nx))
    C = torch.tensor(rs.randn(T, B, nt, nt))
    c = torch.tensor(rs.randn(T, B, nt))
    Q, q = ctrl.compute_Qq_dense(C, c)
    A, b = ctrl.compute_Ab_dense(F, f, x0)
    G, h = ctrl.compute_Gh_dense(x0)
    z = torch.tensor(rs.randn(B, T * nt))
    res = ctrl.dyn_res(z, ip_mpc.LinDx(F, f), x0)
    cost = ctrl.compute_cost(z.reshape(B, T, nt), ip_mpc.QuadCost(C, c))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ipmpc_matrices.npz"), F=F.numpy(), f=f.numpy(), x0=x0.numpy(),
                        C=C.numpy(), c=c.numpy(), ul=ul.numpy(), uu=uu.numpy(), Q=Q.numpy(), q=q.numpy(), A=A.numpy(), b=b.numpy(),
                        G=G.numpy(), h=h.numpy(), z=z.numpy(), dyn_res=res.numpy(), cost=cost.numpy())
    # the linear residual IS A z - b: what DenseQPFunction's canonical-callback probe relies on
    assert torch.allclose(res, torch.bmm(A, z.unsqueeze(-1)).squeeze(-1) - b, atol=1e-12)
    print("ipmpc matrices:", tuple(Q.shape), tuple(A.shape), tuple(G.shape))


if __name__ == "__main__":
    main()
