"""GPU tests of the RESIDENT route (csrc/qp_resident.cuh: several PDIPM iterations per launch, the batch-global
get_step fill of qpth/solvers/pdipm/batch.py:211-214 speculated and repaired).  Through the C ABI, `-m gpu`.

* chunk invariance: the same kernels with 1, 3, 4 or 20 iterations per launch return bit-identical results --
  with one iteration per launch there is next to nothing to speculate wrongly, so this pins the repair logic;
* the resident route against the one-launch-per-iteration route and against the oracle / reference goldens
  (iteration count included);
* the fallback (`B200QP_ST_SPEC_FAIL` -> exact route) on a problem family whose ratio tests are fill-only.
"""
import pytest
import torch

from tests.qp_cases import compare_with_golden, gate, make_inputs, reference_stop

pytestmark = pytest.mark.gpu


@pytest.fixture()
def opts():
    from b200qp import _lib
    yield _lib.set_option
    for k, v in (("res", 1), ("res_ch", 4), ("res_spec", 1)):
        _lib.set_option(k, v)


def solve(inp, device, backward=True):
    from b200qp.qp import QPFunction
    t = {k: v.to(device).requires_grad_(backward) for k, v in inp.items()}
    fn = QPFunction(verbose=-1, check_Q_spd=False)
    z = fn(t["Q"], t["p"], t["G"], t["h"], t["A"], t["b"])
    out = dict(zhat=z.detach().cpu())
    if backward:
        z.backward(torch.ones_like(z))
        ctx = z.grad_fn
        out.update(lams=ctx.lams.cpu(), slacks=ctx.slacks.cpu(), nus=ctx.nus.cpu(), dQ=t["Q"].grad.cpu(), dp=t["p"].grad.cpu(),
                   dG=t["G"].grad.cpu(), dh=t["h"].grad.cpu(), dA=None, db=None)
    return out, dict(fn.info)


def rand_inputs(nb, nz, m, seed, wc=False):
    from oracle import qp_oracle as O
    Q, p, G, h, A, b = O.random_qp(nb, nz, m, 0, seed=seed, well_conditioned=wc)
    return dict(Q=Q, p=p, G=G, h=h, A=A, b=b)


@pytest.mark.parametrize("shape", [(512, 30, 60, 0, False), (300, 30, 60, 3, True), (200, 12, 20, 7, False), (64, 32, 63, 9, False)])
def test_chunk_invariance_bitwise(shape, cuda_device, opts):
    nb, nz, m, seed, wc = shape
    inp = rand_inputs(nb, nz, m, seed, wc)
    ref = None  # the compile-time-size specialisation does the same arithmetic in the same order as the generic kernel
    for spec in (1, 0):
        opts("res_spec", spec)
        for ch in (1, 3, 4, 20):
            opts("res_ch", ch)
            out, info = solve(inp, cuda_device)
            assert not info.get("exact_rerun"), "unexpected fallback"
            if ref is None:
                ref = (out, info)
                continue
            assert info["n_iter"] == ref[1]["n_iter"], (ch, info, ref[1])
            for k in ("zhat", "lams", "slacks", "dp", "dG"):
                a, b = out[k], ref[0][k]
                same = (a == b) | (torch.isnan(a) & torch.isnan(b))
                assert bool(same.all()), f"spec {spec} chunk {ch}: {k} differs from (spec 1, chunk 1) in {int((~same).sum())} entries"


# the last three shapes: odd nz (scalar loads in k_wres_prefactor), nz = 32 (no padding row in Q's tiles), nineq = 32 (four tile rows
# of G with a bordered fifth row of T)
@pytest.mark.parametrize("shape", [(256, 30, 60, 11, False), (128, 30, 60, 0, False), (64, 30, 60, 3, True), (96, 10, 10, 5, False), (50, 20, 31, 6, False),
                                   (40, 7, 9, 12, False), (33, 32, 40, 13, False), (48, 17, 32, 14, True)])
def test_resident_vs_per_iteration_route_and_oracle(shape, cuda_device, opts):
    from oracle import qp_oracle as O
    nb, nz, m, seed, wc = shape
    inp = rand_inputs(nb, nz, m, seed, wc)
    fwd = O.qp_forward(*(inp[k].clone() for k in "QpGhAb"))
    gr = O.qp_backward(fwd, *(inp[k] for k in "QpGhAb"), torch.ones_like(fwd["zhat"]))
    _, det, why = reference_stop(inp)
    opts("res", 0)
    ex, ex_info = solve(inp, cuda_device)
    print(f"reference {fwd['n_iter']} iterations, deterministic={det} ({why}); per-iteration route {ex_info['n_iter']}")
    if det:
        assert ex_info["n_iter"] == fwd["n_iter"]
    opts("res", 1)
    for spec in (1, 0):
        opts("res_spec", spec)
        out, info = solve(inp, cuda_device)
        print(f"  resident route spec={spec}: {info['n_iter']} iterations, NaN onset {info['nan_onset']}")
        if det:
            assert info["n_iter"] == fwd["n_iter"], (spec, info["n_iter"], fwd["n_iter"])
        for k in ("zhat", "lams", "slacks"):
            gate(out[k], fwd[k], 1e-6, f"{k} vs oracle")
            gate(out[k], ex[k], 1e-7, f"{k} vs per-iteration route")
        for k in ("dp", "dG", "dQ", "dh"):
            gate(out[k], gr[k], 1e-6, f"{k} vs oracle")


@pytest.mark.parametrize("case", ["cfg1_nb128_nz30_m60", "wellcond_nb64_nz30_m60", "shared_ph_nb8_nz10_m10"])
def test_resident_goldens_all_variants(case, cuda_device, opts):
    from tests.qp_cases import load_golden
    inp = make_inputs(case)
    g = load_golden(case)
    for spec, ch in ((1, 4), (0, 4), (1, 7), (0, 20), (1, 1)):
        opts("res_spec", spec); opts("res_ch", ch)
        from tests.test_qp_parity_gpu import run_ours
        out, info = run_ours(inp, cuda_device)
        compare_with_golden(case, out, rtol=1e-6)
        assert info["n_iter"] == int(g["n_iter"]), (spec, ch, info["n_iter"], int(g["n_iter"]))


def test_fill_only_problems_fall_back_to_exact_route(cuda_device, opts):
    """nineq = 1: the single ratio of a problem is either unfilled or the whole test is fill-only, which the
    resident route never speculates -> B200QP_ST_SPEC_FAIL -> the Python layer repeats the call on the exact route."""
    from oracle import qp_oracle as O
    inp = rand_inputs(24, 10, 1, 21)
    fwd = O.qp_forward(*(inp[k].clone() for k in "QpGhAb"))
    out, info = solve(inp, cuda_device)
    gate(out["zhat"], fwd["zhat"], 1e-6, "zhat")
    gate(out["lams"], fwd["lams"], 1e-6, "lams")
    if reference_stop(inp)[1]:
        assert info["n_iter"] == fwd["n_iter"]
    opts("res", 0)
    ex, ex_info = solve(inp, cuda_device)
    if info.get("exact_rerun"):
        assert torch.equal(out["zhat"], ex["zhat"]), "the rerun must be the exact route"
    print("fill-only family: exact_rerun =", info.get("exact_rerun", False))


def test_bench_batch_against_oracle(cuda_device, opts):
    """The bench configuration itself (BASELINE configs[1]: nb = 32768, nz = 30, nineq = 60, 20 iterations because of
    the batch-global termination) against the oracle on the same inputs: iteration count, solution, duals at 1e-6 for
    every problem; gradients
      (a) against the reference backward evaluated AT OUR forward outputs (isolates the backward kernel): 1e-6, every
          problem;
      (b) end to end against the reference's own forward + backward: 1e-6, except for problems whose gradient is
          ill-conditioned in the reference itself.  qp.py:139 forms d = clamp(lams) / clamp(slacks); a weakly active
          constraint (lam ~ slack ~ 1e-9, d = O(1e-2)) makes the reference's own gradient move by 1e-5 when its own forward
          outputs are perturbed by 1e-9 (1000x inside the forward gate) -- measured below, never assumed: the tolerance of
          a problem is max(1e-6, 10 x the change of the reference gradient under such a perturbation), and at most 0.01 %
          of the batch may need more than 1e-6."""
    from oracle import qp_oracle as O
    inp = rand_inputs(32768, 30, 60, 0)
    fwd = O.qp_forward(*(inp[k].clone() for k in "QpGhAb"))
    ones = torch.ones_like(fwd["zhat"])
    args = tuple(inp[k] for k in "QpGhAb")
    gr = O.qp_backward(fwd, *args, ones)
    out, info = solve(inp, cuda_device)
    det = reference_stop(inp)
    print("bench batch: n_iter", info["n_iter"], "oracle", fwd["n_iter"], det, "exact_rerun", info.get("exact_rerun", False))
    if det[1]:
        assert info["n_iter"] == fwd["n_iter"]
    for k in ("zhat", "lams", "slacks"):
        gate(out[k], fwd[k], 1e-6, k)
    # (a) the backward kernel against the reference backward at the same (our) forward outputs
    at_ours = dict(fwd)
    at_ours.update(zhat=out["zhat"], lams=out["lams"], slacks=out["slacks"])
    gr_ours = O.qp_backward(at_ours, *args, ones)
    for k in ("dp", "dh", "dG", "dQ"):
        gate(out[k], gr_ours[k], 1e-6, k + " (backward kernel)")
    # (b) end to end, with the measured conditioning of the reference's own gradient
    sens = {k: torch.zeros(gr[k].shape[0], dtype=torch.float64) for k in ("dp", "dh", "dG", "dQ")}
    for trial in range(3):
        gen = torch.Generator().manual_seed(100 + trial)
        pert = dict(fwd)
        for k in ("zhat", "lams", "slacks"):
            v = fwd[k]
            pert[k] = v + 1e-9 * v.norm(dim=1, keepdim=True) / v.shape[1] ** 0.5 * torch.randn(v.shape, generator=gen, dtype=v.dtype)
        g2 = O.qp_backward(pert, *args, ones)
        for k in sens:
            a_, b_ = g2[k].reshape(g2[k].shape[0], -1), gr[k].reshape(gr[k].shape[0], -1)
            bn = b_.norm(dim=1)
            sens[k] = torch.maximum(sens[k], (a_ - b_).norm(dim=1) / (bn + bn.median() + 1e-300))
    for k in ("dp", "dh", "dG", "dQ"):
        a_, b_ = out[k].reshape(out[k].shape[0], -1), gr[k].reshape(gr[k].shape[0], -1)
        bn = b_.norm(dim=1)
        err = (a_ - b_).norm(dim=1) / (bn + bn.median() + 1e-300)
        tol = torch.clamp(10.0 * sens[k], min=1e-6)
        over = int((err > 1e-6).sum())
        worst = int(err.argmax())
        print(f"  {k}: worst per-problem rel err {err.max().item():.3e} (problem {worst}, reference sensitivity {sens[k][worst].item():.3e}); "
              f"{over} of {err.numel()} problems above 1e-6")
        assert bool((err <= tol).all()), f"{k}: problem {int((err / tol).argmax())} exceeds its conditioned tolerance"
        assert over <= max(1, err.numel() // 10000), f"{k}: {over} problems above 1e-6"
        whole = (a_ - b_).norm().item() / b_.norm().item()
        assert whole <= 1e-6, f"{k}: whole-tensor rel err {whole:.3e}"
