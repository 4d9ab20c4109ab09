"""CPU suite (-m "not gpu"): the oracle against the committed golden vectors of the real reference,
and the CPU model of the GPU elimination (oracle/algo_model.py) against the same vectors at the
1e-6 parity gate.  No CUDA anywhere in this file."""
import pytest
import torch

from tests.qp_cases import CASES, ITER_EXACT, checksum, compare_with_golden, load_golden, make_inputs

SMALL = [c for c in CASES if CASES[c][5] == "float64" and CASES[c][0] * CASES[c][2] <= 128 * 60]


def _oracle_out(inp):
    from oracle.gen_golden import run_oracle
    return run_oracle(inp)


@pytest.mark.parametrize("case", SMALL)
def test_oracle_matches_reference_golden(case):
    """The oracle issues the reference's torch.linalg calls in the reference's order; on the same
    torch build the golden vectors are reproduced bit for bit, on another build to ~1e-9."""
    inp = make_inputs(case)
    g = load_golden(case)
    assert abs(checksum(inp) - float(g["input_checksum"])) <= 1e-9 * abs(float(g["input_checksum"])), "generator drift"
    out = _oracle_out(inp)
    compare_with_golden(case, out, rtol=1e-8)
    if case in ITER_EXACT:
        assert out["n_iter"] == int(g["n_iter"])


@pytest.mark.parametrize("case", ["cfg1_nb128_nz30_m60", "eq_nb32_nz20_m16_p6", "kktshape_nb2_nz5_m4_p3",
                                  "shared_ph_nb8_nz10_m10"])
def test_gpu_elimination_model_matches_golden(case):
    """oracle/algo_model.py = the kernels' algebra (Cholesky, explicit Q^-1, Schur on the
    inequality block) on the CPU: same iteration count, 1e-6 parity."""
    from oracle import algo_model as M
    from oracle import qp_oracle as O
    inp = make_inputs(case)
    g = load_golden(case)
    Q, p, G, h, A, b = (inp[k] for k in "QpGhAb")
    nb = O._nbatch(Q, p, G, h, A, b)
    Qe, pe, Ge, he = (O._expand(x, nb, d)[0].contiguous() for x, d in ((Q, 3), (p, 2), (G, 3), (h, 2)))
    if A.shape[-2] > 0:
        Ae, be = O._expand(A, nb, 3)[0].contiguous(), O._expand(b, nb, 2)[0].contiguous()
    else:
        Ae, be = torch.zeros(nb, 0, Qe.shape[-1], dtype=Qe.dtype), torch.zeros(nb, 0, dtype=Qe.dtype)
    fwd = M.pdipm_model(Qe, pe, Ge, he, Ae, be)
    assert fwd["n_iter"] == int(g["n_iter"])
    from tests.qp_cases import gate
    gate(fwd["zhat"], g["zhat"], 1e-6, "zhat")
    gate(fwd["lams"], g["lams"], 1e-6, "lams")
    gate(fwd["slacks"], g["slacks"], 1e-6, "slacks")


def test_block_kkt_equals_full_kkt():
    """The reference's own self-consistency test (test.py:222-247): block-LU KKT solve == full-LU
    KKT solve, here on the oracle's restatement of both."""
    from oracle import qp_oracle as O
    torch.manual_seed(0)
    nb, n, m, p = 4, 8, 6, 3
    Q, pp, G, h, A, b = O.random_qp(nb, n, m, p, seed=17)
    d = torch.rand(nb, m, dtype=torch.float64) + 0.1
    rx, rs, rz, ry = (torch.randn(nb, k, dtype=torch.float64) for k in (n, m, m, p))
    kkt = O.BlockKKT(Q, G, A)
    kkt.refactor(d)
    dx, ds, dz, dy = kkt.solve(d, rx, rs, rz, ry)
    fx, fs, fz, fy = O.full_kkt_solve(Q, torch.diag_embed(d), G, A, rx, rs, rz, ry)
    for a_, b_ in ((dx, fx), (ds, fs), (dz, fz), (dy, fy)):
        assert torch.allclose(a_, b_, rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize("case", ["dense_nb16_nz10_m12_p4", "dense_nb8_nz15_m10_p10"])
def test_dense_oracle_matches_reference_golden(case):
    """qp_oracle.dense_forward/backward (DenseQPFunction restatement) against the reference's outputs."""
    import os
    import numpy as np
    from oracle import qp_oracle as O
    from oracle.gen_golden_dense import make_inputs as dense_inputs
    from tests.qp_cases import GOLDEN_DIR, gate
    g = dict(np.load(os.path.join(GOLDEN_DIR, f"{case}.npz")))
    inp = dense_inputs(case)
    fwd = O.dense_forward(*(inp[k].clone() for k in "QpGhAb"))
    gr = O.dense_backward(fwd, torch.ones_like(fwd["zhat"]))
    assert fwd["n_iter"] == int(g["n_iter"])
    gate(fwd["zhat"], g["zhat"], 1e-9, "zhat")
    for k in ("dQ", "dp", "dG", "dh", "dA", "db"):
        gate(gr[k], g[k], 1e-9, k)


def test_oracle_reproduces_callback_goldens():
    """The oracle with this fork's NON-linear residual callbacks against the goldens of the real reference
    (oracle/gen_golden_callbacks.py): forward solution and all gradients."""
    import os
    import numpy as np
    import torch
    from oracle import qp_oracle as O
    from tests.qp_cases import CB_CASES, GOLDEN_DIR, nonlinear_callbacks
    for case in ("cb_qp_nb16_nz12_m10_p4", "cb_dense_nb16_nz15_m10_p10"):
        nb, nz, m, p, seed, dense = CB_CASES[case]
        Q, pp, G, h, A, b = O.random_qp(nb, nz, m, p, seed=seed, well_conditioned=True)
        g = dict(np.load(os.path.join(GOLDEN_DIR, f"{case}.npz")))
        cost_grad, dyn_res = nonlinear_callbacks(Q, pp, A, b)
        ones = torch.ones(nb, nz, dtype=torch.float64)
        if dense:
            fwd = O.dense_forward(Q.clone(), pp.clone(), G.clone(), h.clone(), A.clone(), b.clone(), dyn_res=dyn_res)
            gr = O.dense_backward(fwd, ones)
        else:
            fwd = O.qp_forward(Q.clone(), pp.clone(), G.clone(), h.clone(), A.clone(), b.clone(), cost_grad=cost_grad, dyn_res=dyn_res)
            gr = O.qp_backward(fwd, Q, pp, G, h, A, b, ones)
        assert fwd["n_iter"] == int(g["n_iter"])
        assert torch.equal(fwd["zhat"], torch.tensor(g["zhat"]))
        for k in ("dQ", "dp", "dG", "dh", "dA", "db"):
            assert (gr[k] - torch.tensor(g[k])).abs().max() <= 1e-12 * max(1.0, float(np.abs(g[k]).max())), k
