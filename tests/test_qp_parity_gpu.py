"""GPU parity tests proper: the CUDA path (through the C ABI) against the committed golden vectors
of the real reference and against the oracle on seeded inputs.  Run with `-m gpu` on a B200."""
import numpy as np
import pytest
import torch

from tests.qp_cases import CASES, ITER_EXACT, compare_with_golden, gate, load_golden, make_inputs, checksum

pytestmark = pytest.mark.gpu

F64_CASES = [c for c in CASES if CASES[c][5] == "float64"]


def check_iterations(case, ours, ref, route="default", inp=None):
    """Iteration count: bit-exact wherever it is deterministic (tests/qp_cases.py:reference_stop decides that from
    the reference's own residual trace, never from our result); a non-deterministic case is reported with its
    reason and the two counts."""
    from tests.qp_cases import ITER_RECORDED, reference_stop
    n_ref, det, why = reference_stop(inp if inp is not None else make_inputs(case))
    assert n_ref == ref, "oracle and golden disagree on the reference's iteration count"
    print(f"[iterations] {case}/{route}: ours {ours} reference {ref} deterministic={det} ({why})")
    if case in ITER_RECORDED:
        rec = ITER_RECORDED[case]
        assert (ours, ref) == (rec["ours"], rec["ref"]), f"{case}/{route}: {ours} iterations, recorded {rec['ours']} (reference {ref})"
        return
    if det or case in ITER_EXACT:
        assert ours == ref, f"{case}/{route}: {ours} iterations, reference {ref}"


def run_ours(inp, device, **kw):
    from b200qp.qp import QPFunction
    t = {k: v.to(device).requires_grad_(True) for k, v in inp.items()}
    fn = QPFunction(verbose=-1, check_Q_spd=False, **kw)
    z = fn(t["Q"], t["p"], t["G"], t["h"], t["A"], t["b"])
    z.backward(torch.ones_like(z))
    ctx = z.grad_fn
    neq = inp["A"].shape[-2]
    out = dict(zhat=z.detach().cpu(), lams=ctx.lams.cpu(), slacks=ctx.slacks.cpu(), nus=ctx.nus.cpu(),
               dQ=t["Q"].grad.cpu(), dp=t["p"].grad.cpu(), dG=t["G"].grad.cpu(), dh=t["h"].grad.cpu(),
               dA=t["A"].grad.cpu() if neq > 0 else None, db=t["b"].grad.cpu() if neq > 0 else None)
    return out, fn.info


@pytest.mark.parametrize("case", F64_CASES)
def test_golden_fp64(case, cuda_device):
    inp = make_inputs(case)
    g = load_golden(case)
    assert abs(checksum(inp) - float(g["input_checksum"])) <= 1e-9 * abs(float(g["input_checksum"])), "generator drift"
    out, info = run_ours(inp, cuda_device)
    worst = compare_with_golden(case, out, rtol=1e-6)
    print(case, "n_iter", info["n_iter"], "ref", int(g["n_iter"]), {k: f"{v:.1e}" for k, v in worst.items()})
    check_iterations(case, info["n_iter"], int(g["n_iter"]))


def test_golden_fp32(cuda_device):
    """fp32: zhat within 1e-4 of the fp64 reference answer is not meaningful per problem for this
    generator (the reference's own fp32 run is 1.2e-4 off); gate against the
    reference's fp32 golden at the reference's own fp32-vs-fp64 error level."""
    case = "fp32_nb32_nz30_m60"
    inp = make_inputs(case)
    out, info = run_ours(inp, cuda_device)
    g = load_golden(case)
    z, zr = out["zhat"].double(), torch.as_tensor(g["zhat"]).double()
    rel = ((z - zr).norm() / zr.norm()).item()
    print("fp32 zhat rel", rel, "n_iter", info["n_iter"], int(g["n_iter"]))
    assert rel <= 1e-3
    lam_rel = ((out["lams"].double() - torch.as_tensor(g["lams"]).double()).norm() / torch.as_tensor(g["lams"]).double().norm()).item()
    assert lam_rel <= 2e-2


def test_oracle_seeded_neq0(cuda_device):
    """Fresh seeded batch (not a golden): CUDA path vs the oracle run here on the CPU."""
    from oracle import qp_oracle as O
    Q, p, G, h, A, b = O.random_qp(48, 24, 40, 0, seed=123)
    fwd = O.qp_forward(Q.clone(), p.clone(), G.clone(), h.clone(), A.clone(), b.clone())
    gr = O.qp_backward(fwd, Q, p, G, h, A, b, torch.ones_like(fwd["zhat"]))
    out, info = run_ours(dict(Q=Q, p=p, G=G, h=h, A=A, b=b), cuda_device)
    gate(out["zhat"], fwd["zhat"], 1e-6, "zhat")
    gate(out["lams"], fwd["lams"], 1e-6, "lams")
    gate(out["slacks"], fwd["slacks"], 1e-6, "slacks")
    gate(out["dp"], gr["dp"], 1e-6, "dp")
    gate(out["dG"], gr["dG"], 1e-6, "dG")
    gate(out["dQ"], gr["dQ"], 1e-6, "dQ")
    gate(out["dh"], gr["dh"], 1e-6, "dh")
    assert info["n_iter"] == fwd["n_iter"]


def test_oracle_seeded_eq(cuda_device):
    from oracle import qp_oracle as O
    Q, p, G, h, A, b = O.random_qp(40, 18, 22, 7, seed=321)
    fwd = O.qp_forward(Q.clone(), p.clone(), G.clone(), h.clone(), A.clone(), b.clone())
    gr = O.qp_backward(fwd, Q, p, G, h, A, b, torch.ones_like(fwd["zhat"]))
    out, info = run_ours(dict(Q=Q, p=p, G=G, h=h, A=A, b=b), cuda_device)
    for k in ("zhat", "lams", "slacks", "nus"):
        gate(out[k], fwd[k], 1e-6, k)
    for k in ("dQ", "dp", "dG", "dh", "dA", "db"):
        gate(out[k], gr[k], 1e-6, k)
    assert info["n_iter"] == fwd["n_iter"]


def test_kkt_residual_property_large_batch(cuda_device):
    """Size-independent property at a batch the oracle cannot finish quickly: the returned
    (zhat, lams, slacks) satisfy the KKT conditions of every problem."""
    from oracle import qp_oracle as O
    from b200qp.qp import QPFunction
    nb = 4096
    Q, p, G, h, A, b = (t.to(cuda_device) for t in O.random_qp(nb, 30, 60, 0, seed=5))
    fn = QPFunction(verbose=-1, check_Q_spd=False)
    z = fn(Q, p, G, h, A, b)
    ctx_lams = None
    zz = z
    # recompute through autograd-free path to fetch duals
    Qg = Q.clone().requires_grad_(True)
    z2 = fn(Qg, p, G, h, A, b)
    lams, slacks = z2.grad_fn.lams, z2.grad_fn.slacks
    assert torch.equal(z, z2.detach()), "forward is not run-to-run deterministic"
    rx = torch.bmm(Q, z.unsqueeze(2)).squeeze(2) + p + torch.bmm(lams.unsqueeze(1), G).squeeze(1)
    rz = torch.bmm(G, z.unsqueeze(2)).squeeze(2) + slacks - h
    scale = 1 + p.norm(dim=1)
    assert (rx.norm(dim=1) / scale).max().item() < 1e-7
    assert (rz.norm(dim=1) / (1 + h.norm(dim=1))).max().item() < 1e-7
    assert (lams * slacks).abs().max().item() < 1e-6
    assert lams.min().item() > -1e-9 and slacks.min().item() > -1e-9
    assert fn.info["n_iter"] <= 20


def test_kkt_solver_backends_agree(cuda_device):
    """The reference's own self-consistency tests (test.py:222-247): block pre-factor + factor +
    solve == one-shot factor_solve == iterative-refinement solve; here all three entry points of
    b200qp.solvers.pdipm.batch against the oracle's full-KKT LU solve."""
    from b200qp.solvers.pdipm import batch as pdipm_b
    from oracle import qp_oracle as O
    torch.manual_seed(0)
    nb, n, m, p = 9, 10, 7, 3
    Q, pp, G, h, A, b = O.random_qp(nb, n, m, p, seed=21)
    d = torch.rand(nb, m, dtype=torch.float64) + 0.05
    rx, rs, rz, ry = (torch.randn(nb, k, dtype=torch.float64) for k in (n, m, m, p))
    want = O.full_kkt_solve(Q, torch.diag_embed(d), G, A, rx, rs, rz, ry)
    dev = cuda_device
    Qc, Gc, Ac, dc = Q.to(dev), G.to(dev), A.to(dev), d.to(dev)
    r = [t.to(dev) for t in (rx, rs, rz, ry)]
    got1 = pdipm_b.factor_solve_kkt(Qc, torch.diag_embed(dc), Gc, Ac, *r)
    Q_LU, S_LU, R = pdipm_b.pre_factor_kkt(Qc, Gc, Ac)
    pdipm_b.factor_kkt(S_LU, R, dc)
    got2 = pdipm_b.solve_kkt(Q_LU, dc, Gc, Ac, S_LU, *r)
    got3 = pdipm_b.solve_kkt_ir(Qc, torch.diag_embed(dc), Gc, Ac, *r, niter=1)
    for got in (got1, got2):
        for a_, w_, name in zip(got, want, ("dx", "ds", "dz", "dy")):
            gate(a_.cpu(), w_, 1e-8, name)
    # the regularise-and-refine back-end converges to the system regularised by -1e-7 I in the constraint block
    # only (batch.py:245-272); the reference's own test gates it at rtol 1e-4 (test.py:237-247)
    for a_, w_, name in zip(got3, want, ("dx", "ds", "dz", "dy")):
        gate(a_.cpu(), w_, 1e-5, name + " (ir)")
    # factor_solve_kkt_reg / kkt_resid_reg against the full regularised KKT matrix solved densely on the CPU
    eps = 1e-7
    Qt = Q + eps * torch.eye(n, dtype=torch.float64)
    Dt = torch.diag_embed(d) + eps * torch.eye(m, dtype=torch.float64)
    K = torch.zeros(nb, n + 2 * m + p, n + 2 * m + p, dtype=torch.float64)
    K[:, :n, :n] = Qt
    K[:, :n, n + m:n + 2 * m] = G.transpose(1, 2); K[:, :n, n + 2 * m:] = A.transpose(1, 2)
    K[:, n:n + m, n:n + m] = Dt; K[:, n:n + m, n + m:n + 2 * m] = torch.eye(m, dtype=torch.float64)
    K[:, n + m:n + 2 * m, :n] = G; K[:, n + m:n + 2 * m, n:n + m] = torch.eye(m, dtype=torch.float64)
    K[:, n + m:n + 2 * m, n + m:n + 2 * m] = -eps * torch.eye(m, dtype=torch.float64)
    K[:, n + 2 * m:, :n] = A; K[:, n + 2 * m:, n + 2 * m:] = -eps * torch.eye(p, dtype=torch.float64)
    sol = torch.linalg.solve(K, -torch.cat((rx, rs, rz, ry), 1).unsqueeze(2)).squeeze(2)
    want_reg = (sol[:, :n], sol[:, n:n + m], sol[:, n + m:n + 2 * m], sol[:, n + 2 * m:])
    got4 = pdipm_b.factor_solve_kkt_reg(Qt.to(dev), Dt.to(dev), Gc, Ac, *r, eps)
    for a_, w_, name in zip(got4, want_reg, ("dx", "ds", "dz", "dy")):
        gate(a_.cpu(), w_, 1e-8, name + " (reg)")
    res = pdipm_b.kkt_resid_reg(Qt.to(dev), Dt.to(dev), Gc, Ac, eps, *got4, *r)
    assert max(float(v.abs().max()) for v in res) < 1e-9
    # solver-level forward: same best iterates as QPFunction
    x, y, z, s = pdipm_b.forward(Qc, pp.to(dev), Gc, h.to(dev), Ac, b.to(dev))
    fwd = O.qp_forward(Q.clone(), pp.clone(), G.clone(), h.clone(), A.clone(), b.clone())
    gate(x.cpu(), fwd["zhat"], 1e-6, "x")
    gate(z.cpu(), fwd["lams"], 1e-6, "z")
    gate(s.cpu(), fwd["slacks"], 1e-6, "s")
    gate(y.cpu(), fwd["nus"], 1e-6, "y")


@pytest.mark.parametrize("nb", [37, 4099])
def test_host_buffer_entry_point_matches_device_path(nb, cuda_device):
    """b200qp_solve_host (chunked H2D -> pre-factorisation and backward -> D2H pipelines around the
    whole-batch PDIPM loop) returns bit-identical results to the device-resident QPFunction path;
    nb = 4099 exercises the 8-chunk pipeline with ragged chunks."""
    import ctypes
    from b200qp import _lib
    from oracle import qp_oracle as O
    nz, m = 12, 18
    Q, p, G, h, A, b = O.random_qp(nb, nz, m, 0, seed=77)
    out, info = run_ours(dict(Q=Q, p=p, G=G, h=h, A=A, b=b), cuda_device)
    L = _lib.lib()
    prob = _lib.Problem(nb, nz, m, 0, _lib.F64, 20, 3, 0, 1e-12, nz * nz, nz, m * nz, m, 0, 0)
    host = {k: torch.empty(s, dtype=torch.float64) for k, s in dict(
        zhat=(nb, nz), lams=(nb, m), slacks=(nb, m), dQ=(nb, nz, nz), dp=(nb, nz), dG=(nb, m, nz), dh=(nb, m)).items()}
    st = torch.zeros(8, dtype=torch.float64)
    gz = torch.ones(nb, nz, dtype=torch.float64)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    N0 = ctypes.c_void_p(0)
    rc = L.b200qp_solve_host(ctypes.byref(prob), P(Q), P(p), P(G), P(h), N0, N0, P(gz), P(host["zhat"]), P(host["lams"]), N0,
                             P(host["slacks"]), P(host["dQ"]), P(host["dp"]), P(host["dG"]), P(host["dh"]), N0, N0, P(st))
    assert rc == 0
    assert int(st[0]) == info["n_iter"]
    for k in ("zhat", "lams", "slacks", "dQ", "dp", "dG", "dh"):
        assert torch.equal(host[k], out[k]), k


def test_pipelined_host_entry_points(cuda_device):
    """b200qp_solve_host_submit / _wait: three jobs in flight in three slots return what the
    synchronous call returns (different inputs per slot, pinned buffers)."""
    import ctypes
    from b200qp import _lib
    from oracle import qp_oracle as O
    nb, nz, m = 4100, 10, 14
    L = _lib.lib()
    prob = _lib.Problem(nb, nz, m, 0, _lib.F64, 20, 3, 0, 1e-12, nz * nz, nz, m * nz, m, 0, 0)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    N0 = ctypes.c_void_p(0)
    jobs = []
    for seed in (5, 6, 7):
        Q, p, G, h, A, b = O.random_qp(nb, nz, m, 0, seed=seed)
        inp = {k: v.pin_memory() for k, v in dict(Q=Q, p=p, G=G, h=h).items()}
        out = {k: torch.empty(s, dtype=torch.float64).pin_memory() for k, s in dict(
            zhat=(nb, nz), lams=(nb, m), slacks=(nb, m), dQ=(nb, nz, nz), dp=(nb, nz), dG=(nb, m, nz), dh=(nb, m)).items()}
        jobs.append((inp, out, torch.zeros(8, dtype=torch.float64).pin_memory()))
    gz = torch.ones(nb, nz, dtype=torch.float64).pin_memory()

    def call(fn, *pre):
        def run(inp, out, st):
            return fn(*pre, ctypes.byref(prob), P(inp["Q"]), P(inp["p"]), P(inp["G"]), P(inp["h"]), N0, N0, P(gz), P(out["zhat"]),
                      P(out["lams"]), N0, P(out["slacks"]), P(out["dQ"]), P(out["dp"]), P(out["dG"]), P(out["dh"]), N0, N0, P(st))
        return run

    for slot, job in enumerate(jobs):
        assert call(L.b200qp_solve_host_submit, slot)(*job) == 0
    assert L.b200qp_solve_host_wait(0) == 0 and L.b200qp_solve_host_wait(1) == 0 and L.b200qp_solve_host_wait(2) == 0
    assert L.b200qp_solve_host_wait(4) != 0  # B200QP_HOST_SLOTS = 4
    got = [{k: v.clone() for k, v in job[1].items()} for job in jobs]
    iters = [int(job[2][0]) for job in jobs]
    for i, job in enumerate(jobs):
        for v in job[1].values():
            v.zero_()
        assert call(L.b200qp_solve_host)(*job) == 0
        assert int(job[2][0]) == iters[i]
        for k in got[i]:
            assert torch.equal(got[i][k], job[1][k]), (i, k)
    # B200QP_FLAG_FACTORED_GRAD: dQ / dG are not written; the four factors reproduce them (qpth/qp.py:158-174)
    inp, out, st = jobs[0]
    for v in out.values():
        v.fill_(7.0)
    prob.flags = 4
    assert call(L.b200qp_solve_host)(inp, out, st) == 0
    prob.flags = 0
    assert bool((out["dQ"] == 7.0).all()) and bool((out["dG"] == 7.0).all())
    for k in ("zhat", "lams", "slacks", "dp", "dh"):
        assert torch.equal(out[k], got[0][k]), k
    dx, dlam, z, lam = out["dp"], -out["dh"], out["zhat"], out["lams"]
    dQ = 0.5 * (dx.unsqueeze(2) * z.unsqueeze(1) + z.unsqueeze(2) * dx.unsqueeze(1))
    dG = dlam.unsqueeze(2) * z.unsqueeze(1) + lam.unsqueeze(2) * dx.unsqueeze(1)
    assert (dQ - got[0]["dQ"]).abs().max() <= 1e-12 * got[0]["dQ"].abs().max()
    assert (dG - got[0]["dG"]).abs().max() <= 1e-12 * got[0]["dG"].abs().max()


@pytest.mark.parametrize("case", ["dense_nb16_nz10_m12_p4", "dense_nb32_nz30_m60_p0", "dense_nb8_nz15_m10_p10"])
def test_dense_qp_function_golden(case, cuda_device):
    """DenseQPFunction (full-KKT variant: regularised solve + refinement, qp.py:187-271) against the
    real reference's outputs and gradients (oracle/gen_golden_dense.py)."""
    import os
    from b200qp.qp import DenseQPFunction
    from oracle.gen_golden_dense import make_inputs as dense_inputs
    from tests.qp_cases import GOLDEN_DIR
    g = dict(np.load(os.path.join(GOLDEN_DIR, f"{case}.npz")))
    inp = dense_inputs(case)
    t = {k: v.to(cuda_device).requires_grad_(True) for k, v in inp.items()}
    neq = inp["A"].shape[1]
    dyn_res = lambda x: torch.bmm(t["A"], x.unsqueeze(-1)).squeeze(-1) - t["b"]
    fn = DenseQPFunction(verbose=-1)
    z = fn(t["Q"], t["p"], t["G"], t["h"], t["A"], t["b"], dyn_res)
    z.backward(torch.ones_like(z))
    ctx = z.grad_fn
    assert fn.info["n_iter"] == int(g["n_iter"]), (fn.info["n_iter"], int(g["n_iter"]))
    gate(z.detach().cpu(), g["zhat"], 1e-6, "zhat")
    gate(ctx.lams.cpu(), g["lams"], 1e-6, "lams")
    gate(ctx.slacks.cpu(), g["slacks"], 1e-6, "slacks")
    for k in ("dQ", "dp", "dG", "dh") + (("dA", "db") if neq > 0 else ()):
        gate(t[k[1:]].grad.cpu(), g[k], 1e-6, k)
    if neq > 0:
        gate(ctx.nus.cpu(), g["nus"], 1e-6, "nus")


def _exact_worker(rank, world, port, nb, q):
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p_ in (root, os.path.join(root, "diff-qp-mpc_b200")):
        if p_ not in sys.path:
            sys.path.insert(0, p_)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    ndev = torch.cuda.device_count()
    dev = torch.device("cuda", rank % ndev)
    torch.cuda.set_device(dev)
    # NCCL needs one device per rank; on a single-GPU box the two ranks share cuda:0 over gloo
    dist.init_process_group("nccl" if ndev >= world else "gloo", rank=rank, world_size=world)
    try:
        from b200qp.dist import shard_slice
        from b200qp.qp import QPFunction
        from oracle import qp_oracle as O
        Q, p, G, h, A, b = O.random_qp(nb, 30, 60, 0, seed=0)
        sl = shard_slice(nb, rank, world)
        t = [x[sl].to(dev).requires_grad_(True) for x in (Q, p, G, h)] + [A[sl].to(dev), b[sl].to(dev)]
        fn = QPFunction(verbose=-1, check_Q_spd=False, process_group=dist.group.WORLD)
        z = fn(*t)
        z.backward(torch.ones_like(z))
        # numpy, not tensors: torch shares CPU tensors through the sender's fd server, which dies with the worker
        q.put((rank, fn.info["n_iter"], z.detach().cpu().numpy(), t[2].grad.cpu().numpy()))
    except Exception as ex:  # pragma: no cover
        q.put((rank, -1, repr(ex), None))
    finally:
        dist.destroy_process_group()


def test_exact_global_batch_mode_two_ranks(cuda_device):
    """QPFunction(process_group=...): a batch sharded over two ranks is solved bit-for-bit like the
    unsharded batch on one device (the batch-global termination and step fill are all-reduced per
    iteration), whereas independent per-shard solves stop at different iterations."""
    import socket
    import torch.multiprocessing as mp
    from b200qp.qp import QPFunction
    from oracle import qp_oracle as O
    nb, world = 256, 2
    Q, p, G, h, A, b = O.random_qp(nb, 30, 60, 0, seed=0)
    t = [x.to(cuda_device).requires_grad_(True) for x in (Q, p, G, h)] + [A.to(cuda_device), b.to(cuda_device)]
    fn = QPFunction(verbose=-1, check_Q_spd=False)
    from b200qp import _lib
    _lib.set_option("res", 0)  # the phased (sharded) mode is the one-launch-per-iteration route: compare like with like
    try:
        z = fn(*t)
        z.backward(torch.ones_like(z))
    finally:
        _lib.set_option("res", 1)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_exact_worker, args=(r, world, port, nb, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = sorted([q.get(timeout=240) for _ in range(world)], key=lambda r: r[0])
    for pr in procs:
        pr.join(timeout=60)
    assert all(r[1] >= 0 for r in res), res
    assert all(r[1] == fn.info["n_iter"] for r in res), ([r[1] for r in res], fn.info["n_iter"])
    assert np.array_equal(np.concatenate([r[2] for r in res]), z.detach().cpu().numpy())
    assert np.array_equal(np.concatenate([r[3] for r in res]), t[2].grad.cpu().numpy())


@pytest.mark.parametrize("opt,case", [("force_generic", "cfg1_nb128_nz30_m60"), ("force_generic", "eq_nb32_nz20_m16_p6"),
                                      ("force_generic", "mid_nb16_nz64_m64_p16"), ("mid_fast", "m96_nb8_nz40_m96"),
                                      ("mid_fast", "m80eq_nb8_nz48_m80_p8"), ("res", "cfg1_nb128_nz30_m60"),
                                      ("res", "wellcond_nb64_nz30_m60"), ("res", "m63_nb64_nz32_m63")])
def test_alternative_kernel_routes_golden(opt, case, cuda_device):
    """The routes the default dispatch no longer takes stay parity-green: the generic shared-memory-resident
    kernels (force_generic: no fast path), the register-tile fast path for 64 < nineq <= 128 (mid_fast) and the
    one-launch-per-iteration fast kernels where the resident route is the default (res = 0).  The overrides are
    process-wide options of the library (b200qp_set_option), restored afterwards."""
    from b200qp import _lib
    _lib.set_option(opt, 0 if opt == "res" else 1)
    try:
        inp = make_inputs(case)
        g = load_golden(case)
        out, info = run_ours(inp, cuda_device)
    finally:
        _lib.set_option(opt, 1 if opt == "res" else 0)
    worst = compare_with_golden(case, out, rtol=1e-6)
    print(opt, case, "n_iter", info["n_iter"], "ref", int(g["n_iter"]), {k: f"{v:.1e}" for k, v in worst.items()})
    check_iterations(case, info["n_iter"], int(g["n_iter"]), route=opt)


@pytest.mark.parametrize("case", list(__import__("tests.qp_cases", fromlist=["CB_CASES"]).CB_CASES))
def test_nonlinear_callbacks_golden(case, cuda_device):
    """This fork's residual callbacks with NON-linear functions (qpth/solvers/pdipm/batch.py:93-102, batch_LU.py:88-97):
    QPFunction (8-argument form) and DenseQPFunction against goldens of the real reference run with the same callbacks
    (oracle/gen_golden_callbacks.py).  The callbacks are evaluated between the launches of b200qp_forward_cb_step /
    b200qp_forward_phase_cb, on the device."""
    from oracle import qp_oracle as O
    from tests.qp_cases import CB_CASES, GOLDEN_DIR, nonlinear_callbacks
    from b200qp.qp import DenseQPFunction, QPFunction
    import os
    nb, nz, m, p, seed, dense = CB_CASES[case]
    Q, pp, G, h, A, b = O.random_qp(nb, nz, m, p, seed=seed, well_conditioned=True)
    g = dict(np.load(os.path.join(GOLDEN_DIR, f"{case}.npz")))
    t = [v.to(cuda_device).requires_grad_(True) for v in (Q, pp, G, h, A, b)]
    cost_grad, dyn_res = nonlinear_callbacks(t[0].detach(), t[1].detach(), t[4].detach(), t[5].detach())
    calls = {"cg": 0, "ry": 0}

    def cg(x):
        calls["cg"] += 1
        return cost_grad(x)

    def ry(x):
        calls["ry"] += 1
        return dyn_res(x)

    if dense:
        fn = DenseQPFunction(verbose=-1)
        z = fn(*t, ry)
    else:
        fn = QPFunction(verbose=-1, check_Q_spd=False)
        z = fn(*t, ry, cg)
    z.backward(torch.ones_like(z))
    assert calls["ry"] >= int(g["n_iter"]) and (dense or calls["cg"] >= int(g["n_iter"]))
    print(case, "n_iter", fn.info["n_iter"], "reference", int(g["n_iter"]))
    gate(z.detach().cpu(), torch.tensor(g["zhat"]), 1e-6, "zhat")
    for k, ten in zip(("dQ", "dp", "dG", "dh", "dA", "db"), t):
        gate(ten.grad.cpu(), torch.tensor(g[k]), 1e-6, k)
    assert fn.info["n_iter"] == int(g["n_iter"])
    # a callback that IS the canonical form stays inside the fused kernels (no call per iteration)
    calls["ry"] = 0
    lin = lambda x: (calls.__setitem__("ry", calls["ry"] + 1), torch.bmm(t[4].detach(), x.unsqueeze(-1)).squeeze(-1) - t[5].detach())[1]
    z2 = (DenseQPFunction(verbose=-1) if dense else QPFunction(verbose=-1, check_Q_spd=False))(*[v.detach() for v in t], lin)
    assert calls["ry"] == 1 and torch.isfinite(z2).all()


def test_check_Q_spd_matches_reference_criterion(cuda_device):
    """qpth/qp.py:82-86: `check_Q_spd=True` raises 'Q is not SPD.' when some eigenvalue of Q has a non-positive real part.
    Symmetric indefinite Q: caught by the pivots of the pre-factorisation.  NON-symmetric Q whose leading pivots are all
    positive but which has a negative eigenvalue: caught by the reference's own criterion (eigvals); a non-symmetric Q
    with eigenvalues in the right half plane passes, as in the reference."""
    from oracle import qp_oracle as O
    from b200qp.qp import QPFunction
    Q, p, G, h, A, b = O.random_qp(4, 6, 5, 0, seed=3)
    t = [v.to(cuda_device) for v in (Q, p, G, h, A, b)]
    QPFunction(verbose=-1, check_Q_spd=True)(*t)                     # SPD: fine
    Qi = t[0].clone(); Qi[1] = -Qi[1]                                  # symmetric, negative definite
    with pytest.raises(RuntimeError, match="Q is not SPD"):
        QPFunction(verbose=-1, check_Q_spd=True)(Qi, *t[1:])
    # [[1, 4], [1, 1]] block: pivots 1, -3?  use [[1, -4], [-1, 1]]... pivots 1 and 1 - 4 = -3; instead a matrix with positive
    # LU pivots and a negative eigenvalue: [[1, 3], [3, 10]] is SPD, so take the non-symmetric [[1, -3], [3, -8.9]]:
    # pivots 1 and -8.9 + 9 = 0.1 > 0, eigenvalues approx -0.01 and -7.9
    Qn = torch.eye(6, dtype=torch.float64, device=cuda_device).repeat(4, 1, 1)
    Qn[:, 0, 0], Qn[:, 0, 1], Qn[:, 1, 0], Qn[:, 1, 1] = 1.0, -3.0, 3.0, -8.9
    assert bool((torch.linalg.eigvals(Qn[0]).real <= 0).any())
    with pytest.raises(RuntimeError, match="Q is not SPD"):
        QPFunction(verbose=-1, check_Q_spd=True)(Qn, *t[1:])
    Qr = torch.eye(6, dtype=torch.float64, device=cuda_device).repeat(4, 1, 1)
    Qr[:, 0, 1], Qr[:, 1, 0] = 0.5, -0.5                              # non-symmetric, eigenvalues 1 +- 0.5i
    z = QPFunction(verbose=-1, check_Q_spd=True)(Qr, *t[1:])
    assert torch.isfinite(z).all()
