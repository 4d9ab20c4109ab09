"""Generate tests/golden/dense_*.npz from the REAL reference's qpth.qp.DenseQPFunction (full-KKT LU
variant, qp.py:187-271 + batch_LU.py) with the canonical callbacks, and assert that the oracle's
restatement (qp_oracle.dense_forward/backward) reproduces it.  Build container only."""
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

from oracle import qp_oracle as O  # noqa: E402

# name -> (nb, nz, nineq, neq, seed)
CASES = {
    "dense_nb16_nz10_m12_p4": (16, 10, 12, 4, 11),
    "dense_nb32_nz30_m60_p0": (32, 30, 60, 0, 12),
    "dense_nb8_nz15_m10_p10": (8, 15, 10, 10, 13),
    # 64 < nineq <= 128: the size class where a second workspace layout used to be derived in backward
    "dense_nb8_nz30_m96_p0": (8, 30, 96, 0, 14),
    "dense_nb4_nz40_m100_p6": (4, 40, 100, 6, 15),
}


def make_inputs(case):
    nb, nz, m, p, seed = CASES[case]
    Q, pp, G, h, A, b = O.random_qp(nb, nz, m, p, seed=seed, well_conditioned=True)
    return dict(Q=Q, p=pp, G=G, h=h, A=A, b=b)


def run_reference(inp):
    from qpth.qp import DenseQPFunction
    t = {k: v.clone().requires_grad_(True) for k, v in inp.items()}
    Q, p, G, h, A, b = (t[k] for k in "QpGhAb")
    dyn_res = lambda x: torch.bmm(A, x.unsqueeze(-1)).squeeze(-1) - b
    z = DenseQPFunction(verbose=-1)(Q, p, G, h, A, b, dyn_res)
    z.backward(torch.ones_like(z))
    neq = A.shape[1]
    return dict(zhat=z.detach(), dQ=Q.grad, dp=p.grad, dG=G.grad, dh=h.grad,
                dA=A.grad if neq > 0 else torch.zeros_like(A), db=b.grad if neq > 0 else torch.zeros_like(b))


def main():
    for case in CASES:
        inp = make_inputs(case)
        ref = run_reference(inp)
        fwd = O.dense_forward(*(inp[k].clone() for k in "QpGhAb"))
        gr = O.dense_backward(fwd, torch.ones_like(fwd["zhat"]))
        ora = dict(zhat=fwd["zhat"], **gr)
        worst = 0.0
        for k in ref:
            a, b = ref[k], ora[k]
            if a.numel() == 0:
                continue
            err = ((a - b).norm() / max(b.norm().item(), 1e-300)).item()
            worst = max(worst, err)
            assert err < 1e-9, (case, k, err)
        save = {k: v.numpy() for k, v in ref.items()}
        save.update(lams=fwd["lams"].numpy(), slacks=fwd["slacks"].numpy(), nus=fwd["nus"].numpy(), n_iter=np.int64(fwd["n_iter"]))
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"{case}.npz"), **save)
        print(f"{case}: n_iter={fwd['n_iter']} oracle vs reference worst rel err {worst:.2e}", flush=True)


if __name__ == "__main__":
    main()
