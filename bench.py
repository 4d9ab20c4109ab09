#!/usr/bin/env python
"""bench.py -- headline metric of BASELINE.json on B200: batched QP solves/s (forward+backward,
nz=30, nineq=60, neq=0, fp64) through the QPFunction drop-in, one process per GPU.

    python bench.py --gpus 1 --steps 10 --warmup 3            # our CUDA path
    python bench.py --impl reference --steps 3 --warmup 1     # the reference's CPU algorithm
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one QPFunction forward + backward over one batch of synthetic random QPs (recipe of
the reference's prof-linear.py:64-75).  The batch is sharded across ranks with NO data-path
collective (problems are independent) => weak scaling, `value` = problems
all ranks solved / max-over-ranks device time.  Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "diff-qp-mpc_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

NZ, NINEQ, NEQ = 30, 60, 0
METRIC = "batched QP solves/sec (fwd+bwd, nz=30, nineq=60, neq=0, fp64)"
UNIT = "solves/s"


def algorithmic_flops(n, m, n_iter):
    """SURVEY.md section 8d (minimal-work, Cholesky based, neq=0)."""
    f_pre = n ** 3 / 3 + n * n * m + n * m * m
    f_solve = 4 * n * n + 4 * n * m + 2 * m * m
    f_init = m ** 3 / 3 + f_solve
    f_iter = m ** 3 / 3 + 10 * n * n + 12 * n * m + 4 * m * m
    f_bwd = m ** 3 / 3 + f_solve + 3 * n * n + 3 * n * m
    return dict(pre=f_pre, init=f_init, iter=f_iter, bwd=f_bwd, total=f_pre + f_init + n_iter * f_iter + f_bwd)


def algorithmic_bytes_per_solve(n, m, es=8):
    """read Q,p,G,h once in fwd and once in bwd; write zhat, lam, s, grads."""
    return es * (3 * (n * n + m * n + n + m) + n + 2 * m)


def iter_kernel_bytes(n, m, es=8):
    """Bytes one launch of the fused iteration kernel must move per problem in THIS design (the
    d-independent matrices are re-staged every launch because the reference's batch-global
    termination forces a grid-wide dependency per iteration): R, BQi, Qi, Q, G + iterate/direction."""
    ldn, ldm = n | 1, m | 1
    mats = m * ldm + m * ldn + n * ldn + n * n + m * n
    vecs = 2 * (2 * n + 4 * m) + (n + 2 * m)
    return es * (mats + vecs)


def gen_batch(nb, device, seed, dtype=torch.float64):
    g = torch.Generator(device=device).manual_seed(seed)
    L = torch.rand(nb, NZ, NZ, generator=g, device=device, dtype=dtype)
    Q = torch.bmm(L, L.transpose(1, 2)) + 1e-3 * torch.eye(NZ, device=device, dtype=dtype)
    G = torch.randn(nb, NINEQ, NZ, generator=g, device=device, dtype=dtype)
    z0 = torch.randn(nb, NZ, generator=g, device=device, dtype=dtype)
    s0 = torch.rand(nb, NINEQ, generator=g, device=device, dtype=dtype)
    p = torch.randn(nb, NZ, generator=g, device=device, dtype=dtype)
    h = torch.bmm(G, z0.unsqueeze(2)).squeeze(2) + s0
    A = torch.zeros(nb, 0, NZ, device=device, dtype=dtype)
    b = torch.zeros(nb, 0, device=device, dtype=dtype)
    return Q, p, G, h, A, b


def gen_batch_sized(nb, nz, m, device, seed):
    """The same recipe at another size (BASELINE configs[4])."""
    g = torch.Generator(device=device).manual_seed(seed)
    L = torch.rand(nb, nz, nz, generator=g, device=device, dtype=torch.float64)
    Q = torch.bmm(L, L.transpose(1, 2)) + 1e-3 * torch.eye(nz, device=device, dtype=torch.float64)
    G = torch.randn(nb, m, nz, generator=g, device=device, dtype=torch.float64)
    z0 = torch.randn(nb, nz, generator=g, device=device, dtype=torch.float64)
    s0 = torch.rand(nb, m, generator=g, device=device, dtype=torch.float64)
    p = torch.randn(nb, nz, generator=g, device=device, dtype=torch.float64)
    h = torch.bmm(G, z0.unsqueeze(2)).squeeze(2) + s0
    return Q, p, G, h


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, mx, pw = [], set(), None, []
        try:
            for line in open(self.path):
                c = [x.strip() for x in line.split(",")]
                if len(c) < 8:
                    continue
                try:
                    sm.append(float(c[1])); mx = float(c[2]); pw.append(float(c[3]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(pw) if pw else None)
        return out


def _fp64_peak():
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "fp64_peaks_r01.json"))).get("dfma_tflops_sustained", 34.0)
    except Exception:
        return 34.0


FP64_PEAK_TFLOPS = _fp64_peak()


def cpu_oracle_rate(nb, repeats, threads):
    """The reference's algorithm (oracle port, bit-identical to qpth on CPU) on the host cores."""
    from oracle import qp_oracle as O
    # torch's intra-op pool already spans the host cores; calling torch.set_num_threads here was found to corrupt
    # MKL's threaded LU of larger matrices later in the same process (oneMKL DLASWP parameter errors, then a hang)
    Q, p, G, h, A, b = gen_batch(nb, torch.device("cpu"), seed=0)
    times = []
    n_iter = None
    for r in range(repeats + 1):
        t0 = time.perf_counter()
        fwd = O.qp_forward(Q, p, G, h, A, b)
        O.qp_backward(fwd, Q, p, G, h, A, b, torch.ones_like(fwd["zhat"]))
        dt = time.perf_counter() - t0
        n_iter = fwd["n_iter"]
        if r > 0 or repeats == 0:
            times.append(dt)
    return nb / statistics.median(times), n_iter, statistics.median(times)


def mpc_problem(B, T, device, seed=0):
    """BASELINE configs[1] shape: pendulum (deqmpc/envs.py), T=5, bounds +-3, Q=diag(10,1,0.01)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    th = (torch.rand(B, generator=g, dtype=torch.float64) * 2 - 1) * 3.141592653589793
    thd = torch.rand(B, generator=g, dtype=torch.float64) * 2 - 1
    x0 = torch.stack((th, thd), 1).to(device)
    u0 = torch.randn(B, T, 1, generator=g, dtype=torch.float64).to(device)
    Cd = torch.tensor([10.0, 1.0, 0.01], dtype=torch.float64).repeat(B, T, 1).to(device)
    return x0, u0, Cd


def bench_mpc(dev, B=1024, T=5, reps=20):
    """MPC rollouts/s: al_mpc.MPC forward (AL solve) + T-step open-loop simulation of the nominal
    controls + implicit backward of loss = x.sum() + u.sum()  (SURVEY.md section 8d)."""
    from b200qp import envs
    from b200qp.AL_mpc import MPC
    from b200qp.al_utils import QuadCost
    x0, u0, Cd = mpc_problem(B, T, dev)
    dx, dxj = envs.PendulumDynamics(), envs.PendulumDynamics_jac()
    ub = 3.0 * torch.ones(1, dtype=torch.float64, device=dev)
    ctrl = MPC(2, 1, T, u_lower=-ub, u_upper=ub, n_batch=B, u_init=u0, eps=1e-5, dtype=torch.float64)
    Cfull = torch.diag_embed(Cd).requires_grad_(True)
    xref = torch.zeros(B, T, 3, dtype=torch.float64, device=dev, requires_grad=True)

    def step():
        ctrl.reinitialize(x0, None)
        ctrl.u_init = u0
        Cfull.grad = None
        xref.grad = None
        c = -(Cd * xref)
        x, u = ctrl(x0, QuadCost(Cfull, c), dx, dxj)
        sim = ctrl.rollout(x0, u.double(), dx)
        (x.sum() + u.sum()).backward()
        return x, sim

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    # the same step captured once in a CUDA graph and replayed (the launch-bound regime at this size: ~40 tiny
    # launches of torch glue + autograd around one fused solve; everything is enqueued on the capturing stream)
    graph_ms = None
    try:
        torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        gs = torch.cuda.Stream()
        gs.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(gs):
            step()
        torch.cuda.current_stream().wait_stream(gs)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=gs):
            gx, gsim = step()
        graph.replay()
        torch.cuda.synchronize()
        xe, sime = step()
        assert torch.equal(gx, xe) and torch.equal(gsim, sime), "graph replay differs from the eager step"
        e0.record()
        for _ in range(reps):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        graph_ms = e0.elapsed_time(e1) / reps
    except Exception as ex:  # pragma: no cover
        graph_ms = repr(ex)[:200]
    # kernel-only time of the fused AL solve (the rest of the step is torch glue + tiny kernels)
    c = -(Cd * xref).detach()
    Cdiag = Cd.clone()
    from b200qp.al_utils import ALSolve, ALState
    spec = envs.dyn_spec(dx)
    xi = ctrl.rollout(x0, u0, dx)
    ul, uu = (-ub).expand(B, T, 1).contiguous(), ub.expand(B, T, 1).contiguous()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kms = []
    for _ in range(5):
        st = ALState(torch.zeros(B, T * 2 + 2 * T, dtype=torch.float64, device=dev), torch.ones(B, 1, dtype=torch.float64, device=dev))
        torch.cuda.synchronize()
        k0.record()
        ALSolve.apply(Cdiag, c, xi, u0, x0, ul, uu, st, spec, 2)
        k1.record()
        torch.cuda.synchronize()
        kms.append(k0.elapsed_time(k1))
    return {"workload": f"BASELINE configs[1] shape: pendulum AL-MPC (deqmpc/envs.py dynamics) T={T} batch={B} fp64, "
                        "forward + open-loop simulation + implicit backward",
            "rollouts_per_s": B / (ms * 1e-3), "ms_per_call": ms, "al_solve_call_ms": min(kms),
            "cuda_graph_ms_per_call": graph_ms,
            "cuda_graph_rollouts_per_s": (B / (graph_ms * 1e-3)) if isinstance(graph_ms, float) else None,
            "launches_per_call": 1 + (T - 1) * 2 + 1}


def bench_mpc_shapes(dev):
    """Fused AL-MPC solve + backward at the other BASELINE config shapes (device-resident, fp64):
    configs[2] shape with the env_dx cartpole, configs[3] shape with the rex quadrotor."""
    from b200qp import envs
    from b200qp.AL_mpc import MPC
    from b200qp.al_utils import QuadCost
    out = {}
    one = lambda v, n: v * torch.ones(n, dtype=torch.float64, device=dev)
    g = torch.Generator(device="cpu").manual_seed(0)
    r = lambda *sh: torch.rand(*sh, generator=g, dtype=torch.float64)

    def run(name, dx, dxj, nx, nu, T, B, x0, u0, qd, lo, hi, reps, fdyn):
        Cd = torch.tensor(qd, dtype=torch.float64, device=dev).repeat(B, T, 1)
        ctrl = MPC(nx, nu, T, u_lower=lo, u_upper=hi, n_batch=B, u_init=u0, eps=1e-5, dtype=torch.float64)
        Cfull = torch.diag_embed(Cd).requires_grad_(True)
        c = torch.zeros(B, T, nx + nu, dtype=torch.float64, device=dev, requires_grad=True)

        def step():
            ctrl.reinitialize(x0, None)
            ctrl.u_init = u0
            Cfull.grad = None
            c.grad = None
            x, u = ctrl(x0, QuadCost(Cfull, c), dx, dxj)
            (x.sum() + u.sum()).backward()

        step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # three groups of `reps` steps, the fastest group counts: with 8 ranks on 16 host cores a descheduled Python process
        # otherwise shows up as a 2x outlier of one rank (and the aggregate takes the MAX over ranks)
        groups = []
        for _ in range(3):
            e0.record()
            for _ in range(reps):
                step()
            e1.record()
            torch.cuda.synchronize()
            groups.append(e0.elapsed_time(e1) / reps)
        ms = min(groups)
        # forward alone (the fused AL solve k_al_solve<Dyn> + a handful of tiny torch kernels) for the roofline
        fms = []
        for _ in range(max(2, reps // 2)):
            ctrl.reinitialize(x0, None)
            ctrl.u_init = u0
            torch.cuda.synchronize()
            e0.record()
            with torch.no_grad():
                ctrl(x0, QuadCost(Cfull.detach(), c.detach()), dx, dxj)
            e1.record()
            torch.cuda.synchronize()
            fms.append(e0.elapsed_time(e1))
        fwd_ms = min(fms)
        nt = nx + nu
        f_dyn = fdyn  # flops of one dynamics step (estimate, stated in DESIGN.md section 4)
        f_jac = 2.0 * nt * f_dyn  # forward-mode duals: nt directions, ~2 flops per primal flop and direction
        flops = 8.0 * T * ((7.0 / 3.0) * nt ** 3 + nx * nt ** 2 + 6 * nt ** 2 + f_jac + 21 * f_dyn)
        tf = flops * B / (fwd_ms * 1e-3) / 1e12
        out[name] = {"B": B, "T": T, "nx": nx, "nu": nu, "ms_per_call": ms, "rollouts_per_s": B / (ms * 1e-3),
                     "forward_ms": fwd_ms,
                     "roofline": {"kernel": "k_al_solve<Dyn,double> (whole AL solve, one warp per problem)", "bound": "fp64",
                                  "flops_per_solve": flops, "achieved": tf, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s",
                                  "frac": tf / FP64_PEAK_TFLOPS,
                                  "formula": "8 T [(7/3) nt^3 + nx nt^2 + 6 nt^2 + F_jac + 21 F_dyn] (SURVEY.md 8d), "
                                             f"F_dyn ~ {f_dyn:.0f}, F_jac ~ 2 nt F_dyn"}}

    B, T = 4096, 20
    th = r(B) * 0.6 - 0.3
    x0 = torch.stack((r(B) - 0.5, torch.zeros(B, dtype=torch.float64), torch.cos(th), torch.sin(th),
                      torch.zeros(B, dtype=torch.float64)), 1).to(dev)
    run("cartpole_env_dx_T20_B4096", envs.CartpoleDx(), envs.CartpoleDx_jac(), 5, 1, T, B, x0,
        (0.1 * (r(B, T, 1) - 0.5)).to(dev), [0.1, 0.1, 1., 1., 0.1, 0.001], one(-100., 1), one(100., 1), 5, 60.)
    # configs[2] with the dynamics deqmpc actually trains on: my_envs CartpoleEnv(nx=4, dt=0.05), u in +-100,
    # Qlqr = 1, Rlqr = 1e-8 (deqmpc/my_envs/cartpole.py:43-84, train.py:102); and the two-link variant (train.py:105)
    from b200qp import my_envs
    kw = dict(dtype=torch.float64, device=dev)
    for nm, nx, dt_, um, fd in (("cartpole1l_myenvs_T20_B4096", 4, 0.05, 100., 240.), ("cartpole2l_myenvs_T20_B4096", 6, 0.03, 250., 600.)):
        d = my_envs.CartpoleDynamics(nx=nx, dt=dt_, kwargs=kw)
        x0 = torch.cat((r(B, nx // 2) * 0.6 - 0.3, r(B, nx // 2) * 0.2 - 0.1), 1).to(dev)
        run(nm, d, d.dynamics_derivatives, nx, 1, T, B, x0, (0.1 * (r(B, T, 1) - 0.5)).to(dev), [1.] * nx + [1e-8],
            one(-um, 1), one(um, 1), 5, fd)
    B, T = 1024, 40
    x0 = torch.cat((r(B, 3) * 2 - 1, r(B, 3) * 0.2 - 0.1, r(B, 6) * 0.2 - 0.1), 1).to(dev)
    run("rex_quadrotor_T40_B1024", envs.RexQuadrotor_dynamics(), envs.RexQuadrotor_dynamics_jac(), 12, 4, T, B, x0,
        (14.9 + 0.1 * (r(B, T, 4) - 0.5)).to(dev), [10.] * 3 + [0.01] * 3 + [1.] * 3 + [0.01] * 3 + [1e-4] * 4,
        one(11.5, 4), one(18.3, 4), 2, 1600.)
    return out


def bench_policy(dev, world, B=256, T=5, reps=20):
    """deqmpc training step on the run.sh configuration (integrator, T=5, bsz=256, deq_iter=6, hdim=128): DEQMPCPolicy
    forward (6 x DEQLayer + Tracking_MPC AL solve), loss, backward through the implicit MPC adjoints, one flat-bucket
    all-reduce of the DEQLayer gradients over the ranks (NCCL), Adam step.  Eager and captured in one CUDA graph."""
    import types
    import numpy as np
    import torch.distributed as dist
    from b200qp import envs, policies

    class _Sp:
        def __init__(self, low, high):
            self.low, self.high = low, high

    class IntegratorEnv:  # deqmpc/envs.py:246-268
        def __init__(self):
            self.dynamics, self.dynamics_derivatives = envs.IntegratorDynamics(), envs.IntegratorDynamics_jac()
            self.nx, self.nu, self.nq, self.dt = 2, 1, 1, 0.1
            self.action_space = _Sp(-np.full(1, 2.0), np.full(1, 2.0))
            self.Qlqr, self.Rlqr = torch.Tensor([10.0, 1.00]), torch.Tensor([0.01])

    env = IntegratorEnv()
    args = types.SimpleNamespace(T=T, bsz=B, dtype="double", solver_type="al", nq=1, hdim=128, layer_type="mlp", deq_out_type=1,
                                 policy_out_type=1, kernel_width=3, pooling="mean", deq_iter=6, qp_iter=1, eps=1e-2, warm_start=True,
                                 device=dev, deq=True, en_qp_solve=True, Q=env.Qlqr, R=env.Rlqr)
    torch.manual_seed(0)
    group = dist.group.WORLD if world > 1 else None
    out = {"workload": f"deqmpc/run.sh shape: integrator, T={T}, bsz={B} per GPU, deq_iter=6, hdim=128, solver_type=al; "
                       "policy forward + loss + backward + gradient all-reduce + Adam", "n_gpus": world}
    kw = dict(device=dev)
    batch = (torch.rand(B, 2, **kw) * 2 - 1, 0.5 * torch.randn(B, T, 2, **kw), 0.5 * torch.randn(B, T, 1, **kw), torch.ones(B, T, **kw))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, use_graph in (("eager", False), ("cuda_graph", True)):
        policy = policies.DEQMPCPolicy(args, env)
        opt = torch.optim.Adam(policy.model.parameters(), lr=1e-3)
        step = policies.GraphedTrainStep(policy, opt, args, batch, process_group=group, use_graph=use_graph)
        for _ in range(3):
            step(*batch)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0.record()
        for _ in range(reps):
            step(*batch)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = ms.item()
        out[name] = {"ms_per_step": ms, "train_steps_per_s": 1e3 / ms, "policy_forwards_per_s": B * world * 1e3 / ms,
                     "mpc_solves_per_s": 6 * B * world * 1e3 / ms}
        # forward only (rollout / evaluation)
        with torch.no_grad():
            for _ in range(2):
                policy(*batch, qp_solve=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                policy(*batch, qp_solve=True)
            e1.record()
            torch.cuda.synchronize()
        if not use_graph:
            out["forward_only_eager_ms"] = e0.elapsed_time(e1) / reps
    out["gradient_allreduce"] = ("one flat-bucket NCCL all-reduce of %d DEQLayer parameters per step" % sum(p.numel() for p in policy.model.parameters())
                                 if world > 1 else "n/a at 1 GPU")
    return out


def bench_qp_sizes(dev):
    """BASELINE configs[4], the larger sizes of the sweep (device-resident, fp64, forward+backward): the
    nineq x nineq system no longer fits in shared memory and the blocked tensor-core kernels of
    csrc/qp_blocked.cuh take over."""
    from b200qp.qp import QPFunction
    out = {}
    for nz, nbs in ((100, 1184), (200, 592)):
        m = 2 * nz
        Q, p, G, h = gen_batch_sized(nbs, nz, m, dev, 5)
        A = torch.zeros(nbs, 0, nz, device=dev, dtype=torch.float64)
        b = torch.zeros(nbs, 0, device=dev, dtype=torch.float64)
        for t in (Q, p, G, h):
            t.requires_grad_(True)
        fn = QPFunction(verbose=-1, check_Q_spd=False)
        ones = torch.ones(nbs, nz, device=dev, dtype=torch.float64)

        def step():
            for t in (Q, p, G, h):
                t.grad = None
            z = fn(Q, p, G, h, A, b)
            z.backward(ones)
            return z

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps_ms = []
        for _ in range(3):
            e0.record()
            z = step()
            e1.record()
            torch.cuda.synchronize()
            reps_ms.append(e0.elapsed_time(e1))
        ms = statistics.median(reps_ms)
        n_it = fn.info["n_iter"]
        fl = algorithmic_flops(nz, m, n_it)["total"]
        out[f"nz{nz}_m{m}"] = {"nb": nbs, "ms_per_step": ms, "solves_per_s": nbs / (ms * 1e-3), "n_iter": n_it,
                               "tflops_algorithmic": fl * nbs / (ms * 1e-3) / 1e12, "finite": bool(torch.isfinite(z).all())}
    return out


def cpu_mpc_rate(B=256, T=5, threads=1):
    """The reference's AL-MPC algorithm (oracle port) on the host cores, bounded sample."""
    from oracle import mpc_oracle as MO
    x0, u0, Cd = mpc_problem(B, T, torch.device("cpu"))
    dyn = MO.Pendulum()
    ub = 3.0 * torch.ones(1, dtype=torch.float64)
    c = torch.zeros(B, T, 3, dtype=torch.float64)
    times = []
    for r in range(3):
        t0 = time.perf_counter()
        st = MO.ALState(B, T * 2 + 2 * T)
        xs, us, ctx = MO.al_solve(MO.rollout(x0, u0, dyn), u0, x0, Cd, c, dyn, -ub, ub, st)
        MO.rollout(x0, us, dyn)
        MO.al_backward(ctx, torch.ones(B, T, 3, dtype=torch.float64))
        times.append(time.perf_counter() - t0)
    sec = statistics.median(times[1:])
    return B / sec, sec


def run_reference(args, rank):
    if rank != 0:
        return
    threads = torch.get_num_threads()
    from oracle import qp_oracle as O
    nb = args.ref_nb
    Q, p, G, h, A, b = gen_batch(nb, torch.device("cpu"), seed=0)

    def step():
        fwd = O.qp_forward(Q, p, G, h, A, b)
        O.qp_backward(fwd, Q, p, G, h, A, b, torch.ones_like(fwd["zhat"]))
        return fwd["n_iter"]

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        n_iter = step()
    dt = time.perf_counter() - t0
    value = nb * args.steps / dt
    sample = f"nb={nb} random QPs per step (same generator), {n_iter} PDIPM iterations, torch CPU fp64"
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"random-QP batch (prof-linear.py:64-75 recipe, torch RNG) nz={NZ} nineq={NINEQ} neq=0 "
                               f"fp64, QPFunction forward+backward; BASELINE configs[4] sweep point, {args.nb} QPs per GPU",
                   "batch_per_step": nb,
                   "sample": f"each step is a bounded sample of that workload: the first {nb} problems of the same generator "
                             "(the reference's CPU time is linear in the batch: 20 iterations either way; the full batch was "
                             "timed once, profiles/r02/bench_r02_reference_arm_nb32768.json)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


_REAL_STDOUT = None


def protect_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner at
    NCCL_DEBUG=VERSION/WARN): from here on file descriptor 1 is stderr, and only emit() writes to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(obj):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def main():
    protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nb", type=int, default=32768, help="QPs per GPU per step")
    ap.add_argument("--ref-nb", type=int, default=1024, help="QPs per step of the CPU reference arm")
    ap.add_argument("--cpu-nb", type=int, default=1024, help="bounded sample for cpu_baseline")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--quick", action="store_true", help="headline leg only (no size sweep, no MPC leg, no nb=128 leg)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line (NCCL prints its version there)
        dist.init_process_group("nccl", device_id=dev)

    from b200qp import _lib
    from b200qp.qp import QPFunction
    L = _lib.lib()

    nb = args.nb
    Q, p, G, h, A, b = gen_batch(nb, dev, seed=1000 + rank)
    for t in (Q, p, G, h):
        t.requires_grad_(True)
    fn = QPFunction(verbose=-1, check_Q_spd=False)
    ones = torch.ones(nb, NZ, device=dev, dtype=torch.float64)

    def step():
        for t in (Q, p, G, h):
            t.grad = None
        z = fn(Q, p, G, h, A, b)
        z.backward(ones)
        return z

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    # ---- timed region: device-resident inputs --------------------------------------------
    L.b200qp_profile_enable(1)
    ms_buf = (ctypes.c_float * 256)()
    kind_buf = (ctypes.c_int * 256)()
    iter_ms, all_kernel_ms, n_iters = [], [], []
    kind_ms = {0: [], 1: [], 2: [], 4: [], 5: [], 6: []}
    res_ms, res_launches = [], []  # resident route: the chunk launches of one step (kind 5)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
        n = L.b200qp_profile_read(ms_buf, kind_buf, 256)
        ni = fn.info["n_iter"]
        n_iters.append(ni)
        k2 = [ms_buf[i] for i in range(n) if kind_buf[i] == 2]
        iter_ms.extend(k2[:ni])
        k5 = [ms_buf[i] for i in range(n) if kind_buf[i] == 5]
        if k5:
            res_ms.append(sum(k5))
            res_launches.append(len(k5))
        all_kernel_ms.append(sum(ms_buf[i] for i in range(n) if kind_buf[i] != 3))
        for kd in kind_ms:
            kind_ms[kd].append(sum(ms_buf[i] for i in range(n) if kind_buf[i] == kd))
    e1.record()
    barrier()
    clocks = sampler.stop()
    L.b200qp_profile_enable(0)
    ms_total = e0.elapsed_time(e1)
    tmax = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total = tmax.item()
    value = nb * world * args.steps / (ms_total * 1e-3)
    n_iter = int(statistics.median(n_iters))
    launches_per_step = fn.info["launches"] + 1

    # ---- roofline of the dominant kernel (fused PDIPM iteration) ---------------------------
    fl = algorithmic_flops(NZ, NINEQ, n_iter)
    avg_iter_ms = sum(iter_ms) / max(len(iter_ms), 1)
    peaks_fp64 = {}
    try:
        peaks_fp64 = json.load(open(os.path.join(ROOT, "profiles", "fp64_peaks_r01.json")))
    except Exception:
        pass
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        hbm_src = "MEASURED_PEAKS.json"
    except Exception:
        hbm_peak, hbm_src = 6650.0, "fallback"
    fp64_peak = peaks_fp64.get("dfma_tflops_sustained", 34.0)
    traffic = None
    try:  # DRAM bytes of one launch, from the committed `ncu --set full` capture, scaled to this batch
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01", "ncu_traffic.json")))
        traffic = tj["dram_bytes_per_problem_per_launch"] * nb
    except Exception:
        pass
    resident = len(res_ms) > 0
    if resident:
        # resident route: the iteration work of a step is spread over the k_res_chunk launches (several iterations of every
        # problem per launch + one repair launch).  Algorithmic flops of those launches together: the initial point and
        # n_iter iterations of EVERY problem, as the reference executes them (it keeps iterating NaN problems
        # until the batch-global test stops the loop).  The kernels stop a problem once its iterate has turned NaN (nothing
        # of it can be returned any more), so the executed flops are lower than the credited ones on batches with NaN
        # problems; the repair launch is not credited.
        step_flops = (fl["init"] + n_iter * fl["iter"]) * nb
        n_l = statistics.mean(res_launches)
        avg_iter_ms = statistics.mean(res_ms) / n_l
        flops_per_launch = step_flops / n_l
        # bytes those launches must move in this design: Q, G, Q^-1 staged once per launch, R once per iteration (L2 hits
        # after the first), the iterate history written once per iteration
        ch_bytes = 8 * (n_l * (NZ * NZ + NINEQ * NZ + NZ * (NZ | 1)) + (n_iter + 1) * (36 * 64 + NZ + 2 * NINEQ + 2)) * nb
        bytes_per_launch = ch_bytes / n_l
        kname = ("k_wres_chunk<NTI=8,NC=30,MC=60,WPC=1> (ten PDIPM iterations of one QP per launch, one warp per QP, the factor "
                 "in registers as DMMA accumulator tiles; csrc/qp_wres.cuh)")
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r02", "ncu_traffic.json")))
            traffic = tj["dram_bytes_per_problem_per_launch"] * nb
            traffic_src = "profiles/r02/ncu_traffic.json: dram__bytes_read+write of the first k_wres_chunk launch of a call (initial point + 10 iterations), per problem, x this batch"
        except Exception:
            traffic, traffic_src = None, "no ncu capture of k_wres_chunk committed yet"
    else:
        flops_per_launch = fl["iter"] * nb
        bytes_per_launch = iter_kernel_bytes(NZ, NINEQ) * nb
        kname = "k_fast_iter<double,MPAD=64,NT=128,DMMA factor> (one PDIPM iteration, one CTA per QP)"
        traffic_src = "profiles/r01/ncu_traffic.json: dram__bytes_read+write of one launch at nb=4096, per problem, x this batch"
    ach_tf = flops_per_launch / (avg_iter_ms * 1e-3) / 1e12
    ach_gbs = bytes_per_launch / (avg_iter_ms * 1e-3) / 1e9
    roofline = {
        "kernel": kname, "bound": "fp64",
        "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach_tf / fp64_peak,
        "peak_source": "profiles/fp64_peaks_r01.json (DFMA microbenchmark on this pool's B200; MEASURED_PEAKS.json has no FP64 figure)",
        "flops_per_launch": flops_per_launch, "avg_launch_ms": avg_iter_ms,
        "flops_basis": "algorithmic = what the reference executes (init + n_iter iterations of every problem); problems whose "
                       "iterate turned NaN are not iterated further by the kernels",
        "launches_timed": int(sum(res_launches)) if resident else len(iter_ms),
        "traffic": traffic,
        "traffic_source": traffic_src,
        "hbm": {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                "bytes_per_launch": bytes_per_launch, "peak_source": hbm_src},
        "whole_solve": {"flops_per_solve": fl["total"], "bytes_per_solve": algorithmic_bytes_per_solve(NZ, NINEQ),
                        "tflops": fl["total"] * value / 1e12, "frac_of_fp64_peak": fl["total"] * value / 1e12 / (fp64_peak * world),
                        "kernel_share_of_step": (sum(res_ms) if resident else sum(iter_ms)) / max(sum(all_kernel_ms), 1e-9),
                        "ms_per_step_by_kernel": {name: statistics.mean(kind_ms[kd]) for name, kd in
                                                  (("prefactor", 0), ("initial_point", 1), ("iterations", 2), ("backward", 4),
                                                   ("resident_chunks", 5), ("reduce_select", 6))}},
    }

    # ---- end to end through the C ABI with HOST buffers ------------------------------------
    e2e = None
    if not args.no_e2e:
        es = 8
        host = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in
                dict(Q=Q, p=p, G=G, h=h).items()}
        for k, v in dict(Q=Q, p=p, G=G, h=h).items():
            host[k].copy_(v.detach())
        gz = torch.ones(nb, NZ, dtype=torch.float64).pin_memory()
        shapes = dict(zhat=(nb, NZ), lams=(nb, NINEQ), slacks=(nb, NINEQ), dQ=(nb, NZ, NZ), dp=(nb, NZ),
                      dG=(nb, NINEQ, NZ), dh=(nb, NINEQ))
        NS = 3  # jobs in flight: one per pipeline stage (H2D, kernels, D2H)
        outs = [{k: torch.empty(sh, dtype=torch.float64).pin_memory() for k, sh in shapes.items()} for _ in range(NS)]
        sts = [torch.zeros(8, dtype=torch.float64).pin_memory() for _ in range(NS)]
        out = outs[0]
        prob = _lib.Problem(nb, NZ, NINEQ, 0, _lib.F64, 20, 3, 0, 1e-12, NZ * NZ, NZ, NINEQ * NZ, NINEQ, 0, 0)
        P = lambda t: ctypes.c_void_p(t.data_ptr())
        NULL = ctypes.c_void_p(0)

        def submit(slot, pr=prob):
            o = outs[slot]
            rc = L.b200qp_solve_host_submit(slot, ctypes.byref(pr), P(host["Q"]), P(host["p"]), P(host["G"]), P(host["h"]),
                                            NULL, NULL, P(gz), P(o["zhat"]), P(o["lams"]), NULL, P(o["slacks"]), P(o["dQ"]),
                                            P(o["dp"]), P(o["dG"]), P(o["dh"]), NULL, NULL, P(sts[slot]))
            _lib.check(rc, "b200qp_solve_host_submit")

        def wait(slot):
            _lib.check(L.b200qp_solve_host_wait(slot), "b200qp_solve_host_wait")

        def serve(pr):
            """serving loop: every step copies its inputs from pinned host memory and returns its results to host
            memory; steps k+1 and k+2 are submitted before step k is awaited (three arenas, three streams)"""
            for i in range(NS):
                submit(i, pr)
            for i in range(NS):
                wait(i)
            barrier()
            t0 = time.perf_counter()
            for i in range(args.steps):
                submit(i % NS, pr)
                if i >= NS - 1:
                    wait((i - (NS - 1)) % NS)
            for i in range(max(0, args.steps - (NS - 1)), args.steps):
                wait(i % NS)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            tm = torch.tensor([dt], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            return tm.item()

        # three timed loops of `steps` steps each; the median is reported and all three are listed (the PCIe path of a shared
        # host shows +-5 % run-to-run)
        e2e_runs = sorted(serve(prob) for _ in range(3))
        dt = e2e_runs[1]
        # the synchronous single-call form, for the record
        t1 = time.perf_counter()
        for i in range(max(2, args.steps // 3)):
            submit(0); wait(0)
        dt_sync = (time.perf_counter() - t1) / max(2, args.steps // 3)
        # opt-in factored gradients (B200QP_FLAG_FACTORED_GRAD): dQ, dG stay on the device as their four factors
        prob_f = _lib.Problem(nb, NZ, NINEQ, 0, _lib.F64, 20, 3, 4, 1e-12, NZ * NZ, NZ, NINEQ * NZ, NINEQ, 0, 0)
        dt_f = serve(prob_f)
        h2d = es * nb * (NZ * NZ + NZ + NINEQ * NZ + NINEQ + NZ)
        d2h = es * nb * (NZ + 2 * NINEQ + NZ * NZ + NZ + NINEQ * NZ + NINEQ) + 64
        d2h_f = es * nb * (NZ + 2 * NINEQ + NZ + NINEQ) + 64
        e2e = {"value": nb * world * args.steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * dt / args.steps,
               "api": "b200qp_solve_host_submit/_wait (C ABI, pinned host buffers; every step copies its inputs in "
                      "and its results out inside the timed region; three slots so that the copies of adjacent steps "
                      "overlap the kernels)",
               "pcie_gb_s_per_direction": 1e-9 * max(h2d, d2h) / (dt / args.steps),
               # all ranks together, both directions: on one box this saturates at ~130 GB/s (host side: PCIe root / DRAM),
               # which is what bounds the end-to-end number at 4 and 8 GPUs (46 KB cross the bus per solve with dense gradients)
               "host_aggregate_gb_s": 1e-9 * (h2d + d2h) * world / (dt / args.steps),
               "bus_bytes_per_solve": (h2d + d2h) / nb,
               "runs_ms_per_step": [1e3 * t / args.steps for t in e2e_runs], "reported": "median of the three runs",
               "single_call_ms_per_step": 1e3 * dt_sync, "single_call_value": nb * world / dt_sync,
               "factored_grad": {"value": nb * world * args.steps / dt_f, "unit": UNIT, "ms_per_step": 1e3 * dt_f / args.steps,
                                 "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_f,
                                 "note": "opt-in B200QP_FLAG_FACTORED_GRAD: dQ, dG returned as their factors (dp, dh, zhat, "
                                         "lams; qpth/qp.py:158-174), not part of the headline"}}
        for i in range(NS):  # leave dense results in the host buffers for the check below
            submit(i); wait(i)
        # keep the device path honest: same answer from both entry points
        zz = step().detach().cpu()
        assert torch.allclose(zz, out["zhat"], rtol=0, atol=0), "host-buffer path and device path disagree"
        del host, out, outs

    # ---- exact global-batch mode (N > 1): the sharded batch solved as ONE reference batch -- the batch-global termination
    # test and step fill are all-reduced once per iteration (96 bytes, NCCL MAX), one launch per iteration by construction
    exact = None
    if world > 1 and not args.quick:
        try:
            fx = QPFunction(verbose=-1, check_Q_spd=False, process_group=dist.group.WORLD)

            def xstep():
                for t in (Q, p, G, h):
                    t.grad = None
                fx(Q, p, G, h, A, b).backward(ones)

            for _ in range(2):
                xstep()
            barrier()
            x0_, x1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            x0_.record()
            for _ in range(3):
                xstep()
            x1_.record()
            barrier()
            xm = torch.tensor([x0_.elapsed_time(x1_) / 3], device=dev, dtype=torch.float64)
            dist.all_reduce(xm, op=dist.ReduceOp.MAX)
            exact = {"workload": f"the same {nb} QPs per GPU treated as one batch of {nb * world} (QPFunction(process_group=...)): "
                                 "bit-identical to the unsharded reference batch", "ms_per_step": xm.item(),
                     "value": nb * world / (xm.item() * 1e-3), "unit": UNIT, "n_iter": fx.info["n_iter"],
                     "collectives_per_step": fx.info["n_iter"] + 1}
        except Exception as ex:  # pragma: no cover
            exact = {"error": repr(ex)[:300]}

    # ---- BASELINE configs[0] shape (nb=128) for reference: latency-bound -----------------------
    small = None
    try:
        if args.quick:
            raise RuntimeError("skipped (--quick)")
        Qs, ps, Gs, hs, As, bs = gen_batch(128, dev, seed=0)
        for t in (Qs, ps, Gs, hs):
            t.requires_grad_(True)
        o128 = torch.ones(128, NZ, device=dev, dtype=torch.float64)
        fs = QPFunction(verbose=-1, check_Q_spd=False)

        def sstep():
            for t in (Qs, ps, Gs, hs):
                t.grad = None
            fs(Qs, ps, Gs, hs, As, bs).backward(o128)

        for _ in range(5):
            sstep()
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(20):
            sstep()
        a1.record()
        torch.cuda.synchronize()
        ms = a0.elapsed_time(a1) / 20
        small = {"workload": "BASELINE configs[0]: nb=128 nz=30 nineq=60 fwd+bwd", "ms_per_call": ms,
                 "solves_per_s": 128 / (ms * 1e-3), "n_iter": fs.info["n_iter"]}
    except Exception as ex:  # pragma: no cover
        small = {"error": repr(ex)}

    sizes = None
    if rank == 0 and not args.quick:
        try:
            sizes = bench_qp_sizes(dev)
        except Exception as ex:  # pragma: no cover
            sizes = {"error": repr(ex)}

    # MPC rollouts/s (BASELINE's second metric): every rank runs the same batch on its own GPU (weak scaling, problems are
    # independent, no collective); the aggregate uses the slowest rank's time
    mpc = None
    if not args.quick:
        try:
            barrier()
            mpc = bench_mpc(dev)
            shapes_ = bench_mpc_shapes(dev)
            if world > 1:
                names = sorted(shapes_)
                t = torch.tensor([mpc["ms_per_call"], mpc["cuda_graph_ms_per_call"] if isinstance(mpc["cuda_graph_ms_per_call"], float) else 0.0]
                                 + [shapes_[k]["ms_per_call"] for k in names], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                t = t.tolist()
                mpc["ms_per_call"] = t[0]
                if t[1] > 0:
                    mpc["cuda_graph_ms_per_call"] = t[1]
                for i, k in enumerate(names):
                    shapes_[k]["ms_per_call"] = t[2 + i]
            B0 = 1024
            mpc["n_gpus"] = world
            mpc["rollouts_per_s"] = B0 * world / (mpc["ms_per_call"] * 1e-3)
            if isinstance(mpc["cuda_graph_ms_per_call"], float):
                mpc["cuda_graph_rollouts_per_s"] = B0 * world / (mpc["cuda_graph_ms_per_call"] * 1e-3)
            for k, v in shapes_.items():
                v["rollouts_per_s"] = v["B"] * world / (v["ms_per_call"] * 1e-3)
                v["n_gpus"] = world
            mpc["other_shapes"] = shapes_
            if rank == 0 and world == 1 and not args.no_cpu:
                threads = torch.get_num_threads()
                rate, sec = cpu_mpc_rate(256, 5, threads)
                mpc["cpu_baseline"] = {"value": rate, "unit": "rollouts/s", "cores": threads, "kind": "port",
                                       "sample": f"oracle/mpc_oracle.py (dense restatement of qpth AL_mpc) B=256 T=5, median of 2 ({sec:.2f} s each)"}
        except Exception as ex:  # pragma: no cover
            mpc = {"error": repr(ex)}
            if world > 1:
                raise

    pol = None
    if not args.quick:
        try:
            barrier()
            pol = bench_policy(dev, world)
        except Exception as ex:  # pragma: no cover
            pol = {"error": repr(ex)[:300]}
            if world > 1:
                raise

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = torch.get_num_threads()
        rate, it_cpu, sec = cpu_oracle_rate(args.cpu_nb, 2, threads)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"oracle/qp_oracle.py (bit-identical restatement of qpth on torch CPU), nb={args.cpu_nb} of the same "
                         f"generator, fwd+bwd, {it_cpu} iterations, median of 2 after 1 warm-up ({sec:.2f} s each)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"random-QP batch (prof-linear.py:64-75 recipe, torch RNG) nz={NZ} nineq={NINEQ} neq=0 "
                                   f"fp64, QPFunction forward+backward; BASELINE configs[4] sweep point, {nb} QPs per GPU",
                       "batch_per_gpu": nb, "global_batch": nb * world, "pdipm_iterations": n_iter,
                       "eps": 1e-12, "maxIter": 20, "exact_rerun": bool(fn.info.get("exact_rerun", False)),
                       "nan_onset_iteration": fn.info.get("nan_onset"), "parallelism": f"batch-sharded x{world}, no data-path collective",
                       "l2": "inputs+workspace per step exceed the 126 MB L2 (no flush needed)"},
            "clocks": clocks, "e2e": e2e, "exact_global_batch": exact, "gpu_launches": launches_per_step * args.steps,
            "roofline": roofline, "cpu_baseline": cpu, "cfg1_nb128": small, "cfg4_sizes": sizes, "mpc": mpc, "policy": pol,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
