"""BASELINE configs[4]: random-QP sweep nz in {30,100,200} (m = 2 nz), forward+backward solves/s."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "diff-qp-mpc_b200"))
import torch
from b200qp.qp import QPFunction

dev = torch.device("cuda:0")


def gen(nb, nz, m, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    L = torch.rand(nb, nz, nz, generator=g, device=dev, dtype=torch.float64)
    Q = torch.bmm(L, L.transpose(1, 2)) + 1e-3 * torch.eye(nz, device=dev, dtype=torch.float64)
    G = torch.randn(nb, m, nz, generator=g, device=dev, dtype=torch.float64)
    z0 = torch.randn(nb, nz, generator=g, device=dev, dtype=torch.float64)
    s0 = torch.rand(nb, m, generator=g, device=dev, dtype=torch.float64)
    p = torch.randn(nb, nz, generator=g, device=dev, dtype=torch.float64)
    h = torch.bmm(G, z0.unsqueeze(2)).squeeze(2) + s0
    return Q, p, G, h, torch.zeros(nb, 0, nz, device=dev, dtype=torch.float64), torch.zeros(nb, 0, device=dev, dtype=torch.float64)


for nz, nb in ((30, 1000), (30, 10000), (30, 100000), (100, 1000), (100, 10000), (200, 1000)):
    m = 2 * nz
    try:
        Q, p, G, h, A, b = gen(nb, nz, m)
        for t in (Q, p, G, h):
            t.requires_grad_(True)
        fn = QPFunction(verbose=-1, check_Q_spd=False)
        ones = torch.ones(nb, nz, device=dev, dtype=torch.float64)

        def step():
            for t in (Q, p, G, h):
                t.grad = None
            z = fn(Q, p, G, h, A, b)
            z.backward(ones)
            return z

        for _ in range(3):
            z = step()
        torch.cuda.synchronize()
        reps = 5 if nz == 30 else 2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            z = step()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        rx = (torch.bmm(Q, z.unsqueeze(2)).squeeze(2) + p).detach()
        print(f"nz={nz:4d} m={m:4d} nb={nb:7d}: {ms:10.2f} ms/step {nb / ms * 1e3:12.0f} solves/s  n_iter={fn.info['n_iter']} finite={bool(torch.isfinite(z).all())}", flush=True)
    except Exception as ex:
        print(f"nz={nz} nb={nb}: FAILED {ex!r}", flush=True)
