"""One forward+backward at a given size (ncu target): qp_one.py NZ NB"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "diff-qp-mpc_b200"))
import torch
from b200qp.qp import QPFunction
nz, nb = int(sys.argv[1]), int(sys.argv[2])
m = 2 * nz
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
Lm = torch.rand(nb, nz, nz, generator=g, device=dev, dtype=torch.float64)
Q = torch.bmm(Lm, Lm.transpose(1, 2)) + 1e-3 * torch.eye(nz, device=dev, dtype=torch.float64)
G = torch.randn(nb, m, nz, generator=g, device=dev, dtype=torch.float64)
z0 = torch.randn(nb, nz, generator=g, device=dev, dtype=torch.float64)
s0 = torch.rand(nb, m, generator=g, device=dev, dtype=torch.float64)
p = torch.randn(nb, nz, generator=g, device=dev, dtype=torch.float64)
h = torch.bmm(G, z0.unsqueeze(2)).squeeze(2) + s0
A = torch.zeros(nb, 0, nz, device=dev, dtype=torch.float64); b = torch.zeros(nb, 0, device=dev, dtype=torch.float64)
for t in (Q, p, G, h):
    t.requires_grad_(True)
fn = QPFunction(verbose=-1, check_Q_spd=False, maxIter=int(sys.argv[3]) if len(sys.argv) > 3 else 20)
z = fn(Q, p, G, h, A, b)
z.backward(torch.ones_like(z))
torch.cuda.synchronize()
print("n_iter", fn.info["n_iter"], float(z.sum()))
