"""Shared helpers for the QP parity tests: golden-case inputs and the parity gate.

Parity gate: fp64 results within 1e-6 relative WITH a norm floor,
    ||a - b|| <= rtol * (||b|| + median_batch ||b||)   per problem, and
    ||a - b|| <= rtol * ||b||                           on the whole-batch tensor;
bare per-problem relative error is ill-posed for problems whose gradient is numerically zero.
"""
import os

import numpy as np
import torch

from oracle import qp_oracle as O
from oracle.gen_golden import CASES, make_inputs, checksum  # noqa: F401  (no reference import at module level)

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Cases whose reference run stops on the stall counter / NaN logic (deterministic iteration count)
# as opposed to `best.resids.max() < eps`, where rounding noise at the 1e-12 threshold decides.
ITER_EXACT = {
    "cfg1_nb128_nz30_m60", "wellcond_nb64_nz30_m60", "eq_nb32_nz20_m16_p6", "kktshape_nb2_nz5_m4_p3",
    "shared_QG_nb16_nz12_m20_p3", "shared_ph_nb8_nz10_m10", "single_nb1_nz10_m1_p2", "mid_nb16_nz64_m64_p16",
}


def reference_stop(inp, eps=1e-12, lim=3, max_iter=20, noise=6e-13):
    """How the reference's loop ends on these inputs (the oracle is bit-identical to it on the CPU) and whether that
    iteration count is DETERMINISTIC, i.e. a property of the problem rather than of rounding noise.

    The loop returns on one of three batch-global tests (qpth/solvers/pdipm/batch.py:141): the stall counter, the
    divergence test, or `best.resids.max() < eps`.  With eps = 1e-12 the last one compares the residual floor of the
    slowest problem -- pure rounding noise of Qx + p + G'z, a few 1e-13 at these scales -- with the threshold.  When
    that maximum comes within `noise` (absolute) of eps at some iteration, two correct implementations (LU vs LDL^T,
    different summation order) can legitimately land on different sides of it; once a run misses the threshold
    nothing else stops it before the stall counter or maxIter (measured: 16 vs 20 iterations on a 256-problem
    batch whose reference run stops with 7.4e-13).  Such cases are reported as not deterministic and the tests
    compare the solutions only.  Returns (n_iter, deterministic, why)."""
    t = {k: v.double().clone() for k, v in inp.items()}
    nb = O._nbatch(*(t[k] for k in "QpGhAb"))
    Q, p = O._expand(t["Q"], nb, 3)[0], O._expand(t["p"], nb, 2)[0]
    G, h = O._expand(t["G"], nb, 3)[0], O._expand(t["h"], nb, 2)[0]
    A, b = O._expand(t["A"], nb, 3)[0], O._expand(t["b"], nb, 2)[0]
    trace = []
    kkt = O.BlockKKT(Q.contiguous(), G.contiguous(), A.contiguous())
    n_iter = O.pdipm_solve(Q.contiguous(), p.contiguous(), G.contiguous(), h.contiguous(), A.contiguous(), b.contiguous(),
                           kkt, eps, lim, max_iter, trace=trace)[4]
    best, worst = None, []
    for it, tr in enumerate(trace):
        r = tr["resids"]
        if best is not None:
            # the stall counter (batch.py:127-131) resets when ANY problem improves: an iteration whose only
            # improvements (or near-improvements) are residual-floor noise -- a change of less than 5 % -- is a coin flip
            rel = (r - best) / best
            solid = bool((rel < -0.05).any())
            noisy = bool((rel.abs() <= 0.05).any())
            if noisy and not solid:
                return n_iter, False, f"iteration {it}: the only candidate improvements are within 5 % of the best residual (noise)"
        best = r.clone() if best is None else torch.where(r < best, r, best)
        worst.append(best.max().item())
    near = [i for i, w in enumerate(worst) if abs(w - eps) < noise]
    if near:
        return n_iter, False, f"worst best-residual {worst[near[0]]:.2e} at iteration {near[0]} is within {noise:g} of eps"
    return n_iter, True, "stall counter / maxIter / a residual far from eps decides"


# Deterministic by the two rules above, yet a different count: recorded with the measured count and the reason.  The
# tests assert the recorded count exactly (a record, not a tolerance).
ITER_RECORDED = {
    "huge_nb4_nz200_m400": dict(ours=20, ref=19, why=(
        "problem 1 of 4 turns NaN in the reference at iteration 16 while still converging (mu = 1.6e-15, residual "
        "1.6e-11): a breakdown of the partial-pivot LU on T = R + diag(s/z) with s/z spanning 30 orders of "
        "magnitude, not a property of the iterate.  The LDL^T kernels factor the same T without breaking down, the "
        "problem improves once more at iteration 16 and the stall counter reaches 3 one iteration later.  The "
        "returned solutions agree to 1e-9.")),
}


def load_golden(case):
    return dict(np.load(os.path.join(GOLDEN_DIR, f"qp_{case}.npz")))


def gate(a, b, rtol, what):
    """a: ours, b: reference.  Both torch tensors on CPU, float64."""
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if b.numel() == 0:
        return 0.0
    assert torch.isfinite(a).all() == torch.isfinite(b).all(), f"{what}: finiteness differs"
    whole = (a - b).norm().item() / max(b.norm().item(), 1e-300)
    assert whole <= rtol, f"{what}: whole-tensor rel err {whole:.3e} > {rtol}"
    if a.dim() >= 2 and a.shape[0] > 1:
        af, bf = a.reshape(a.shape[0], -1), b.reshape(b.shape[0], -1)
        bn = bf.norm(dim=1)
        # norm floor: the batch median, and 1e-3 of the batch maximum for families where most
        # problems have a numerically zero answer (e.g. the multipliers of inactive constraints at nineq = 1)
        err = (af - bf).norm(dim=1) / (bn + torch.maximum(bn.median(), 1e-3 * bn.max()) + 1e-300)
        worst = err.max().item()
        assert worst <= rtol, f"{what}: per-problem rel err {worst:.3e} > {rtol} (problem {int(err.argmax())})"
    return whole


def compare_with_golden(case, out, rtol):
    """out: dict with zhat, lams, slacks, nus, dQ, dp, dG, dh, dA, db (CPU tensors)."""
    g = load_golden(case)
    worst = {}
    for k in ("zhat", "lams", "slacks", "nus", "dp", "dh", "db"):
        if k in out and out[k] is not None:
            worst[k] = gate(out[k], g[k], rtol, f"{case}:{k}")
    for k in ("dQ", "dG", "dA"):
        if out.get(k) is None:
            continue
        v = torch.as_tensor(out[k], dtype=torch.float64)
        head = torch.as_tensor(g[k + "_head"])
        if k + "_rownorm" in g:
            worst[k] = gate(v[: head.shape[0]], head, rtol, f"{case}:{k}[:8]")
            rn = v.reshape(v.shape[0], -1).norm(dim=1)
            gate(rn, g[k + "_rownorm"], rtol, f"{case}:{k} row norms")
        else:
            worst[k] = gate(v, head, rtol, f"{case}:{k}")
    return worst


# ---- this fork's residual callbacks (qpth/solvers/pdipm/batch.py:93-102) with NON-linear functions ----------------
# name -> (nb, nz, nineq, neq, seed, dense)
CB_CASES = {
    "cb_qp_nb16_nz12_m10_p4": (16, 12, 10, 4, 21, False),
    "cb_qp_nb32_nz30_m60_p6": (32, 30, 60, 6, 22, False),      # fast-path kernels (nineq <= 64)
    "cb_qp_nb8_nz40_m96_p8": (8, 40, 96, 8, 23, False),        # generic / blocked kernels (nineq > 64)
    "cb_dense_nb16_nz15_m10_p10": (16, 15, 10, 10, 24, True),  # the shape of qp_wrapper.MPC's pendulum QP (T = 5)
    "cb_dense_nb8_nz30_m24_p8": (8, 30, 24, 8, 25, True),
}


def nonlinear_callbacks(Q, p, A, b):
    """cost_grad(x) = Qx + p + 0.05 tanh(x)  (the gradient of a non-quadratic convex cost);
    dyn_res(x) = Ax - b + 0.1 sin(x[:, :neq])  (a non-linear equality residual, like the dynamics residual the
    reference's MPC callers pass).  Same expressions on whatever device / dtype the tensors live on."""
    neq = A.shape[-2]

    def cost_grad(x):
        return torch.bmm(Q, x.unsqueeze(-1)).squeeze(-1) + p + 0.05 * torch.tanh(x)

    def dyn_res(x):
        return torch.bmm(A, x.unsqueeze(-1)).squeeze(-1) - b + 0.1 * torch.sin(x[:, :neq])

    return cost_grad, dyn_res
