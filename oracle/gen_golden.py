"""Generate tests/golden/qp_*.npz from the REAL reference (run in the build container only).

    python oracle/gen_golden.py            # needs /root/reference

For every case it (1) runs the unmodified reference qpth.qp.QPFunction on CPU, (2) asserts that
oracle/qp_oracle.py reproduces it BIT-FOR-BIT (this is what pins the oracle), and (3) writes
the outputs as a small .npz.  Inputs are not stored: they are regenerated from the seed by
oracle.qp_oracle.random_qp (numpy RandomState is platform independent); an input checksum is
stored to catch generator drift.  TEST INFRASTRUCTURE, NOT PRODUCT.
"""
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

# name -> (nb, nz, nineq, neq, seed, dtype, well_conditioned, shared (names of un-batched params))
CASES = {
    "cfg1_nb128_nz30_m60": (128, 30, 60, 0, 0, "float64", False, ()),
    "wellcond_nb64_nz30_m60": (64, 30, 60, 0, 3, "float64", True, ()),
    "eq_nb32_nz20_m16_p6": (32, 20, 16, 6, 1, "float64", False, ()),
    "kktshape_nb2_nz5_m4_p3": (2, 5, 4, 3, 2, "float64", False, ()),
    "shared_QG_nb16_nz12_m20_p3": (16, 12, 20, 3, 4, "float64", False, ("Q", "G", "A")),
    "shared_ph_nb8_nz10_m10": (8, 10, 10, 0, 5, "float64", False, ("p", "h")),
    "single_nb1_nz10_m1_p2": (1, 10, 1, 2, 6, "float64", False, ()),
    "big_nb8_nz100_m200": (8, 100, 200, 0, 7, "float64", False, ()),
    "mid_nb16_nz64_m64_p16": (16, 64, 64, 16, 8, "float64", False, ()),
    "fp32_nb32_nz30_m60": (32, 30, 60, 0, 9, "float32", False, ()),
    # 64 < nineq <= 128 and equality-constrained large problems: the global-resident blocked kernels (qp_blocked.cuh)
    "m96_nb8_nz40_m96": (8, 40, 96, 0, 10, "float64", False, ()),
    "m80eq_nb8_nz48_m80_p8": (8, 48, 80, 8, 11, "float64", False, ()),
    "bigeq_nb4_nz70_m150_p12": (4, 70, 150, 12, 12, "float64", False, ()),
    # round 2: the largest size of BASELINE configs[4], fp32 above nineq = 128, the edges of the resident route
    # (nineq = 63 / 31: the last size with a spare row for the bordered right-hand side) and a family whose ratio
    # tests are fill-only (nineq = 1: the resident route must fall back to the exact one)
    "huge_nb4_nz200_m400": (4, 200, 400, 0, 14, "float64", False, ()),
    "fp32big_nb8_nz64_m160": (8, 64, 160, 0, 15, "float32", False, ()),
    "m63_nb64_nz32_m63": (64, 32, 63, 0, 16, "float64", False, ()),
    "m31_nb16_nz16_m31": (16, 16, 31, 0, 17, "float64", True, ()),
    "fillonly_nb24_nz10_m1": (24, 10, 1, 0, 21, "float64", False, ()),
}
FULL_GRAD_ROWS = 8  # dQ/dG/dA are stored in full for the first rows, as norms for the rest


def make_inputs(case):
    nb, nz, m, p, seed, dt, wc, shared = CASES[case]
    dtype = getattr(torch, dt)
    Q, pp, G, h, A, b = O.random_qp(nb, nz, m, p, seed=seed, dtype=torch.float64, well_conditioned=wc)
    d = dict(Q=Q, p=pp, G=G, h=h, A=A, b=b)
    for k in shared:
        d[k] = d[k][0].clone()
    if shared:
        # keep every problem feasible when one side of a constraint is shared across the batch
        rs = np.random.RandomState(seed + 1000)
        z0 = torch.tensor(rs.randn(nb, nz))
        s0 = torch.tensor(rs.rand(nb, m))
        Gb = d["G"] if d["G"].dim() == 3 else d["G"].unsqueeze(0).expand(nb, m, nz)
        Ab = d["A"] if d["A"].dim() == 3 else d["A"].unsqueeze(0).expand(nb, p, nz)
        if "h" in shared:
            d["h"] = torch.ones(m, dtype=torch.float64)  # z = 0 is strictly feasible for every G_i
            assert p == 0
        else:
            d["h"] = torch.bmm(Gb, z0.unsqueeze(2)).squeeze(2) + s0
            d["b"] = torch.bmm(Ab, z0.unsqueeze(2)).squeeze(2)
    return {k: v.to(dtype).contiguous() for k, v in d.items()}


def checksum(inp):
    return float(sum(v.double().sum().item() for v in inp.values()))


def run_reference(inp, eps=1e-12, maxIter=20, notImprovedLim=3):
    from qpth.qp import QPFunction
    import qpth.solvers.pdipm.batch as RB
    RB.factor_kkt_eye = None
    t = {k: v.clone().requires_grad_(True) for k, v in inp.items()}
    Q, p, G, h, A, b = (t[k] for k in "QpGhAb")
    nb = O._nbatch(Q, p, G, h, A, b)
    Qe, pe = O._expand(Q, nb, 3)[0], O._expand(p, nb, 2)[0]
    Ae, be = O._expand(A, nb, 3)[0], O._expand(b, nb, 2)[0]
    neq = A.shape[-2]
    cg = lambda x: torch.bmm(x.unsqueeze(1), Qe.transpose(1, 2)).squeeze(1) + pe
    dr = (lambda x: torch.bmm(x.unsqueeze(1), Ae.transpose(1, 2)).squeeze(1) - be) if neq > 0 else (lambda x: 0.0)
    iters = []
    orig = RB.factor_kkt

    def counting(S_LU, R, d):
        iters.append(1)
        return orig(S_LU, R, d)

    RB.factor_kkt = counting
    try:
        fn = QPFunction(eps=eps, verbose=-1, notImprovedLim=notImprovedLim, maxIter=maxIter, check_Q_spd=False)
        z = fn(Q, p, G, h, A, b, dr, cg)
        n_iter = len(iters) - 1  # one factor_kkt for the initial point, then one per loop body
        gy = torch.ones_like(z)
        z.backward(gy)
    finally:
        RB.factor_kkt = orig
    fn_ctx = z.grad_fn
    out = dict(zhat=z.detach(), lams=fn_ctx.lams, slacks=fn_ctx.slacks,
               nus=fn_ctx.nus if neq > 0 else torch.zeros(nb, 0, dtype=z.dtype),
               n_iter=n_iter, dQ=Q.grad, dp=p.grad, dG=G.grad, dh=h.grad,
               dA=A.grad if neq > 0 else torch.zeros_like(A), db=b.grad if neq > 0 else torch.zeros_like(b))
    return out


def run_oracle(inp, eps=1e-12, maxIter=20, notImprovedLim=3):
    Q, p, G, h, A, b = (inp[k].clone() for k in "QpGhAb")
    fwd = O.qp_forward(Q, p, G, h, A, b, eps, notImprovedLim, maxIter)
    g = O.qp_backward(fwd, Q, p, G, h, A, b, torch.ones_like(fwd["zhat"]))
    nb = fwd["zhat"].shape[0]
    neq = A.shape[-2]
    return dict(zhat=fwd["zhat"], lams=fwd["lams"], slacks=fwd["slacks"],
                nus=fwd["nus"] if neq > 0 else torch.zeros(nb, 0, dtype=Q.dtype), n_iter=fwd["n_iter"],
                dQ=g["dQ"], dp=g["dp"], dG=g["dG"], dh=g["dh"],
                dA=g["dA"] if neq > 0 else torch.zeros_like(A), db=g["db"] if neq > 0 else torch.zeros_like(b))


def main():
    outdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    import faulthandler
    faulthandler.dump_traceback_later(240, exit=True)
    only = sys.argv[1:]   # optional: generate just these cases
    for case in CASES:
        if only and case not in only:
            continue
        inp = make_inputs(case)
        ref = run_reference(inp)
        ora = run_oracle(inp)
        assert ref["n_iter"] == ora["n_iter"], (case, ref["n_iter"], ora["n_iter"])
        for k in ref:
            if k == "n_iter":
                continue
            a, b = ref[k], ora[k]
            same = torch.equal(torch.nan_to_num(a, nan=12345.0), torch.nan_to_num(b, nan=12345.0))
            assert same, f"oracle is not bit-identical to the reference: case={case} key={k} maxdiff={(a-b).abs().max()}"
        save = dict(n_iter=np.int64(ref["n_iter"]), input_checksum=np.float64(checksum(inp)))
        for k in ("zhat", "lams", "slacks", "nus", "dp", "dh", "db"):
            save[k] = ref[k].numpy()
        for k in ("dQ", "dG", "dA"):
            v = ref[k]
            if v.dim() == 3 and v.shape[0] > FULL_GRAD_ROWS:
                save[k + "_head"] = v[:FULL_GRAD_ROWS].numpy()
                save[k + "_rownorm"] = v.reshape(v.shape[0], -1).norm(dim=1).numpy()
                save[k + "_rowsum"] = v.reshape(v.shape[0], -1).sum(dim=1).numpy()
            else:
                save[k + "_head"] = v.numpy()
        np.savez_compressed(os.path.join(outdir, f"qp_{case}.npz"), **save)
        print(f"{case}: n_iter={ref['n_iter']} |zhat|={ref['zhat'].norm():.6f} bit-identical oracle OK", flush=True)


if __name__ == "__main__":
    main()
