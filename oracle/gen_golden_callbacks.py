"""Generate tests/golden/cb_*.npz: the REAL reference's QPFunction (8-argument form) and DenseQPFunction run with
NON-linear residual callbacks -- this fork evaluates cost_grad(x) / dyn_res(x) at the top of every PDIPM iteration
(qpth/solvers/pdipm/batch.py:93-102, batch_LU.py:88-97), and its MPC callers pass the non-linear dynamics residual
(qpth/qp_wrapper.py:303-316) -- and assert that the oracle restatement reproduces them.  Build container only.

The callbacks are defined in tests/qp_cases.py (`nonlinear_callbacks`) so that the GPU tests evaluate the same
functions on the device."""
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
from tests.qp_cases import CB_CASES, nonlinear_callbacks  # noqa: E402


def make_inputs(case):
    nb, nz, m, p, seed, _ = CB_CASES[case]
    Q, pp, G, h, A, b = O.random_qp(nb, nz, m, p, seed=seed, well_conditioned=True)
    return dict(Q=Q, p=pp, G=G, h=h, A=A, b=b)


def run_reference(case, inp):
    from qpth.qp import DenseQPFunction, QPFunction
    dense = CB_CASES[case][5]
    t = {k: v.clone().requires_grad_(True) for k, v in inp.items()}
    Q, p, G, h, A, b = (t[k] for k in "QpGhAb")
    cost_grad, dyn_res = nonlinear_callbacks(Q.detach(), p.detach(), A.detach(), b.detach())
    if dense:
        z = DenseQPFunction(verbose=-1)(Q, p, G, h, A, b, dyn_res)  # its backward returns 7 gradients: 8 arguments cannot be differentiated
    else:
        z = QPFunction(verbose=-1, check_Q_spd=False)(Q, p, G, h, A, b, dyn_res, cost_grad)
    z.backward(torch.ones_like(z))
    return dict(zhat=z.detach(), dQ=Q.grad, dp=p.grad, dG=G.grad, dh=h.grad, dA=A.grad, db=b.grad)


def main():
    for case in CB_CASES:
        inp = make_inputs(case)
        ref = run_reference(case, inp)
        cost_grad, dyn_res = nonlinear_callbacks(inp["Q"], inp["p"], inp["A"], inp["b"])
        if CB_CASES[case][5]:
            fwd = O.dense_forward(*(inp[k].clone() for k in "QpGhAb"), dyn_res=dyn_res)
            gr = O.dense_backward(fwd, torch.ones_like(fwd["zhat"]))
        else:
            fwd = O.qp_forward(*(inp[k].clone() for k in "QpGhAb"), cost_grad=cost_grad, dyn_res=dyn_res)
            gr = O.qp_backward(fwd, *(inp[k] for k in "QpGhAb"), torch.ones_like(fwd["zhat"]))
        ora = dict(zhat=fwd["zhat"], **gr)
        worst = 0.0
        for k in ref:
            a, b = ref[k], ora[k]
            if a is None or a.numel() == 0:
                continue
            err = ((a - b).norm() / max(b.norm().item(), 1e-300)).item()
            worst = max(worst, err)
            assert err < 1e-9, (case, k, err)
        # the callbacks must matter: the plain QP has a different solution
        plain = O.qp_forward(*(inp[k].clone() for k in "QpGhAb"))["zhat"]
        shift = ((plain - ref["zhat"]).norm() / ref["zhat"].norm()).item()
        assert shift > 1e-3, (case, shift)
        save = {k: v.numpy() for k, v in ref.items() if v is not None}
        save.update(lams=fwd["lams"].numpy(), slacks=fwd["slacks"].numpy(), nus=fwd["nus"].numpy(), n_iter=np.int64(fwd["n_iter"]))
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"{case}.npz"), **save)
        print(f"{case}: n_iter={fwd['n_iter']} oracle vs reference worst rel err {worst:.2e}; callbacks move zhat by {shift:.2e}", flush=True)


if __name__ == "__main__":
    main()
