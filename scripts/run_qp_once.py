"""Tiny driver for profiling: forward (+ backward) of one random QP batch through the public API, a few times.
usage: python scripts/run_qp_once.py NB [REPS]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "diff-qp-mpc_b200"))
import torch  # noqa: E402
import bench  # noqa: E402
from b200qp.qp import QPFunction  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
Q, p, G, h, A, b = bench.gen_batch(nb, dev, seed=1000)
for t in (Q, p, G, h):
    t.requires_grad_(True)
fn = QPFunction(verbose=-1, check_Q_spd=False)
for _ in range(reps):
    z = fn(Q, p, G, h, A, b)
    z.backward(torch.ones_like(z))
torch.cuda.synchronize()
print("n_iter", fn.info["n_iter"], "launches", fn.info["launches"], "nan_onset", fn.info.get("nan_onset"), "z", float(z.sum()))
