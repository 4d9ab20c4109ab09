"""GPU parity tests of the AL-MPC path: the fused kernels (through the C ABI, behind the MPC
drop-in) against the golden vectors of the real reference and against the oracle."""
import os

import numpy as np
import pytest
import torch

from tests.mpc_cases import CASES, GOLDEN_DIR as GOLDEN, load, oracle_dyn, rel

pytestmark = pytest.mark.gpu

# fp64 parity gate of BASELINE.json's north_star: 1e-6 relative on solution, state and gradients;
# the float32 outputs are compared at float32 resolution.
RTOL64, RTOL32 = 1e-6, 2e-6


def _ours(case, dev):
    from b200qp import envs
    from b200qp.AL_mpc import MPC
    from b200qp.al_utils import QuadCost
    g = load(case)
    env = CASES[case][0]
    dx, dxj = (envs.PendulumDynamics(), envs.PendulumDynamics_jac()) if env == "pendulum" else \
              (envs.IntegratorDynamics(), envs.IntegratorDynamics_jac())
    B, T = g["u_init"].shape[:2]
    nx, nu = dx.nx, dx.nu
    ub = float(g["ub"]) * torch.ones(nu, dtype=torch.float64, device=dev)
    x0 = g["x0"].to(dev)
    ctrl = MPC(nx, nu, T, u_lower=-ub, u_upper=ub, exit_unconverged=False, eps=1e-5, n_batch=B, backprop=False,
               verbose=0, u_init=g["u_init"].to(dev), solver_type="dense", dtype=torch.float64)
    ctrl.reinitialize(x0, None)
    ctrl.u_init = g["u_init"].to(dev)
    outs = []
    for k in range(int(g["n_calls"])):
        Cfull = torch.diag_embed(g["Cd"]).to(dev).requires_grad_(True)
        c = (-(g["Cd"] * g[f"xref{k}"])).to(dev).requires_grad_(True)
        x, u = ctrl(x0, QuadCost(Cfull, c), dx, dxj)
        assert x.dtype == torch.float32 and u.dtype == torch.float32
        (x.sum() + u.sum()).backward()
        outs.append(dict(x=x.detach().cpu(), u=u.detach().cpu(), lam=ctrl.lamda_prev.cpu(), rho=ctrl.rho_prev.cpu(),
                         dC=Cfull.grad.diagonal(dim1=-2, dim2=-1).cpu(), dc=c.grad.cpu()))
    return g, outs


@pytest.mark.parametrize("case", list(CASES))
def test_al_mpc_matches_reference_golden(case, cuda_device):
    g, outs = _ours(case, cuda_device)
    for k, o in enumerate(outs):
        errs = {key: rel(o[key], g[f"out_{key}{k}"]) for key in o}
        print(case, "call", k, {a: f"{b:.1e}" for a, b in errs.items()})
        assert errs["x"] <= RTOL32 and errs["u"] <= RTOL32, errs
        for key in ("lam", "rho", "dC", "dc"):
            assert errs[key] <= RTOL64, (key, errs)


@pytest.mark.parametrize("name", ["pendulum", "cartpole"])
def test_envdx_against_reference_modules(name, cuda_device):
    """PendulumDx / CartpoleDx kernels against goldens produced by the reference's OWN modules
    (qpth/env_dx/pendulum.py:49-84, cartpole.py:63-96; oracle/gen_golden_envdx.py imports them from the reference
    and differentiates their forward with autograd): next state and both Jacobians, controls on both sides of
    the clamp."""
    from b200qp import envs
    g = dict(np.load(os.path.join(GOLDEN, f"dyn_envdx_{name}.npz")))
    mod = envs.PendulumDx_jac() if name == "pendulum" else envs.CartpoleDx_jac()
    x, u = torch.tensor(g["x"]).to(cuda_device), torch.tensor(g["u"]).to(cuda_device)
    xn, (A, B) = mod(x, u)
    assert rel(xn.cpu(), torch.tensor(g["xn"])) < 1e-12
    assert rel(A.cpu(), torch.tensor(g["A"])) < 1e-12
    assert rel(B.cpu(), torch.tensor(g["B"])) < 1e-12
    clamped = torch.tensor(g["B"]).abs().sum((1, 2)) == 0
    assert clamped.any() and not clamped.all(), "the golden must exercise both sides of the control clamp"
    assert bool((B.cpu()[clamped] == 0).all())


@pytest.mark.parametrize("env", ["pendulum", "integrator", "pendulum_dx", "cartpole_dx"])
def test_dynamics_step_and_jacobian(env, cuda_device):
    """b200dyn_step / b200dyn_jac against a torch restatement + autograd Jacobians on the CPU."""
    from b200qp import envs
    from oracle import mpc_oracle as MO
    torch.manual_seed(0)
    N = 257
    if env == "pendulum":
        mod, nx, nu = envs.PendulumDynamics_jac(), 2, 1
        ref = MO.Pendulum().step
    elif env == "integrator":
        mod, nx, nu = envs.IntegratorDynamics_jac(), 2, 1
        ref = MO.Integrator().step
    elif env == "pendulum_dx":
        mod, nx, nu = envs.PendulumDx_jac(), 3, 1

        def ref(x, u):  # qpth/env_dx/pendulum.py:49-84
            g_, m, l, dt = 10., 1., 1., 0.05
            uc = torch.clamp(u[:, 0], -2., 2.)
            c, s, dth = x.unbind(1)
            th = torch.atan2(s, c)
            nd = dth + dt * (-3. * g_ / (2. * l) * (-s) + 3. * uc / (m * l ** 2))
            nt = th + nd * dt
            return torch.stack((torch.cos(nt), torch.sin(nt), nd), 1)
    else:
        mod, nx, nu = envs.CartpoleDx_jac(), 5, 1

        def ref(state, u):  # qpth/env_dx/cartpole.py:63-96
            gravity, masscart, masspole, length = torch.tensor((9.8, 1.0, 0.1, 0.5)).unbind()
            total_mass = masspole + masscart
            pml = masspole * length
            uc = torch.clamp(u[:, 0], -100., 100.)
            x, dx, c, s, dth = state.unbind(1)
            th = torch.atan2(s, c)
            cart_in = (uc + pml * dth ** 2 * s) / total_mass
            th_acc = (gravity * s - c * cart_in) / (length * (4. / 3. - masspole * c ** 2 / total_mass))
            xacc = cart_in - pml * th_acc * c / total_mass
            return torch.stack((x + 0.05 * dx, dx + 0.05 * xacc, torch.cos(th + 0.05 * dth), torch.sin(th + 0.05 * dth),
                                dth + 0.05 * th_acc), 1)
    x = torch.randn(N, nx, dtype=torch.float64)
    u = 1.5 * torch.randn(N, nu, dtype=torch.float64)
    xr, ur = x.clone().requires_grad_(True), u.clone().requires_grad_(True)
    out = ref(xr, ur)
    A = torch.stack([torch.autograd.grad(out[:, i].sum(), xr, retain_graph=True)[0] for i in range(nx)], 1)
    Bm = torch.stack([torch.autograd.grad(out[:, i].sum(), ur, retain_graph=True)[0] for i in range(nx)], 1)
    xn, (Ag, Bg) = mod(x.to(cuda_device), u.to(cuda_device))
    assert rel(xn.cpu(), out.detach()) < 1e-12
    assert rel(Ag.cpu(), A) < 1e-12 and rel(Bg.cpu(), Bm) < 1e-12
    # plain step entry point + autograd through it
    base = type(mod).__mro__[2]()
    xg, ug = x.to(cuda_device).requires_grad_(True), u.to(cuda_device).requires_grad_(True)
    y = base(xg, ug)
    assert rel(y.detach().cpu(), out.detach()) < 1e-12
    y.sum().backward()
    assert rel(xg.grad.cpu(), A.sum(1)) < 1e-12 and rel(ug.grad.cpu(), Bm.sum(1)) < 1e-12


def test_al_mpc_vs_oracle_seeded_large_horizon(cuda_device):
    """Fresh seeded problem (not a golden): pendulum, T=20, B=48, tight bounds; CUDA vs oracle."""
    from b200qp import envs
    from b200qp.AL_mpc import MPC
    from b200qp.al_utils import QuadCost
    from oracle import mpc_oracle as MO
    torch.manual_seed(7)
    B, T, nx, nu = 48, 20, 2, 1
    dyn = MO.Pendulum()
    x0 = torch.stack((torch.rand(B, dtype=torch.float64) * 6 - 3, torch.rand(B, dtype=torch.float64) * 2 - 1), 1)
    u0 = torch.randn(B, T, nu, dtype=torch.float64)
    Cd = torch.tensor([10., 1., 0.01], dtype=torch.float64).repeat(B, T, 1)
    c = 0.1 * torch.randn(B, T, nx + nu, dtype=torch.float64)
    ub = 2.0 * torch.ones(nu, dtype=torch.float64)
    st = MO.ALState(B, T * nx + 2 * T * nu)
    xs, us, ctx = MO.al_solve(MO.rollout(x0, u0, dyn), u0, x0, Cd, c, dyn, -ub, ub, st)
    dC, dc = MO.al_backward(ctx, torch.ones(B, T, nx + nu, dtype=torch.float64))
    dev = cuda_device
    ctrl = MPC(nx, nu, T, u_lower=-ub.to(dev), u_upper=ub.to(dev), n_batch=B, u_init=u0.to(dev), dtype=torch.float64)
    ctrl.reinitialize(x0.to(dev), None)
    ctrl.u_init = u0.to(dev)
    Cf = torch.diag_embed(Cd).to(dev).requires_grad_(True)
    cg = c.to(dev).requires_grad_(True)
    x, u = ctrl(x0.to(dev), QuadCost(Cf, cg), envs.PendulumDynamics(), envs.PendulumDynamics_jac())
    (x.sum() + u.sum()).backward()
    assert rel(x.detach().cpu(), xs.float()) < RTOL32 and rel(u.detach().cpu(), us.float()) < RTOL32
    assert rel(ctrl.lamda_prev.cpu(), st.lam) < RTOL64
    assert rel(Cf.grad.diagonal(dim1=-2, dim2=-1).cpu(), dC) < RTOL64 and rel(cg.grad.cpu(), dc) < RTOL64
    assert torch.equal(ctrl.status.cpu(), ctx["status"])


def _npz(name):
    import os
    import numpy as np
    from tests.mpc_cases import GOLDEN_DIR
    return {k: torch.as_tensor(v) for k, v in dict(np.load(os.path.join(GOLDEN_DIR, name))).items()}


def test_rex_quadrotor_dynamics_golden(cuda_device):
    """b200dyn_step / b200dyn_jac for the rex quadrotor (RK4, MRP attitude) against outputs of the
    reference's RexQuadrotor_dynamics(_jac) (oracle/gen_golden_rex.py)."""
    from b200qp import envs
    g = _npz("dyn_rex.npz")
    mod = envs.RexQuadrotor_dynamics_jac()
    xn, (A, B) = mod(g["x"].to(cuda_device), g["u"].to(cuda_device))
    assert rel(xn.cpu(), g["xn"]) < 1e-13
    assert rel(A.cpu(), g["A"]) < 1e-12 and rel(B.cpu(), g["B"]) < 1e-12
    y = envs.RexQuadrotor_dynamics()(g["x"].to(cuda_device), g["u"].to(cuda_device))
    assert rel(y.cpu(), g["xn"]) < 1e-13


@pytest.mark.parametrize("gold", ("mpc_rex_B4_T8.npz", "mpc_rex_B4_T40.npz"))
def test_al_mpc_rex_quadrotor_golden(gold, cuda_device):
    """AL-MPC on the rex quadrotor (nx=12, nu=4, B=4; T=8 and the configs[3] horizon T=40, where the kernel works
    out of its global scratch slab) against the real reference run."""
    from b200qp import envs
    from b200qp.AL_mpc import MPC
    from b200qp.al_utils import QuadCost
    g = _npz(gold)
    dev = cuda_device
    B, T, nu = g["u_init"].shape
    nx = g["x0"].shape[1]
    ul, uu = 11.5 * torch.ones(nu, dtype=torch.float64, device=dev), 18.3 * torch.ones(nu, dtype=torch.float64, device=dev)
    ctrl = MPC(nx, nu, T, u_lower=ul, u_upper=uu, exit_unconverged=False, eps=1e-5, n_batch=B, backprop=False, verbose=0,
               u_init=g["u_init"].to(dev), solver_type="dense", dtype=torch.float64)
    x0 = g["x0"].to(dev)
    ctrl.reinitialize(x0, None)
    ctrl.u_init = g["u_init"].to(dev)
    Cfull = torch.diag_embed(g["Cd"]).to(dev).requires_grad_(True)
    c = (-(g["Cd"] * g["xref"])).to(dev).requires_grad_(True)
    x, u = ctrl(x0, QuadCost(Cfull, c), envs.RexQuadrotor_dynamics(), envs.RexQuadrotor_dynamics_jac())
    (x.sum() + u.sum()).backward()
    errs = dict(x=rel(x.detach().cpu(), g["out_x"]), u=rel(u.detach().cpu(), g["out_u"]), lam=rel(ctrl.lamda_prev.cpu(), g["out_lam"]),
                rho=rel(ctrl.rho_prev.cpu(), g["out_rho"]), dC=rel(Cfull.grad.diagonal(dim1=-2, dim2=-1).cpu(), g["out_dC"]),
                dc=rel(c.grad.cpu(), g["out_dc"]))
    print("rex", {a: f"{b:.1e}" for a, b in errs.items()})
    assert errs["x"] <= RTOL32 and errs["u"] <= RTOL32, errs
    for key in ("lam", "rho", "dC", "dc"):
        assert errs[key] <= RTOL64, (key, errs)


def test_al_mpc_float32_solver_precision(cuda_device):
    """dtype=float32 (the reference's non-"double" Tracking_MPC setting): the fp32 kernels land
    within fp32 conditioning of the fp64 solve; MPC state stays float32."""
    from b200qp import envs
    from b200qp.AL_mpc import MPC
    from b200qp.al_utils import QuadCost
    torch.manual_seed(3)
    dev = cuda_device
    B, T, nx, nu = 32, 6, 2, 1
    x0 = torch.stack((torch.rand(B) * 2 - 1, torch.rand(B) - 0.5), 1).to(dev)
    u0 = (0.1 * torch.randn(B, T, nu)).to(dev)
    Cd = torch.tensor([10.0, 1.0, 0.01]).repeat(B, T, 1).to(dev)
    outs = {}
    for dt in (torch.float64, torch.float32):
        ub = (3.0 * torch.ones(nu)).to(dev)
        ctrl = MPC(nx, nu, T, u_lower=-ub, u_upper=ub, n_batch=B, u_init=u0, dtype=dt)
        ctrl.reinitialize(x0, None)
        ctrl.u_init = u0
        x, u = ctrl(x0, QuadCost(torch.diag_embed(Cd), torch.zeros(B, T, nx + nu, device=dev)), envs.PendulumDynamics(),
                    envs.PendulumDynamics_jac())
        assert ctrl.lamda_prev.dtype == dt and x.dtype == torch.float32
        outs[dt] = (x, u)
    assert rel(outs[torch.float32][0].cpu(), outs[torch.float64][0].cpu()) < 5e-3
    assert rel(outs[torch.float32][1].cpu(), outs[torch.float64][1].cpu()) < 5e-2


def test_unknown_dynamics_is_rejected(cuda_device):
    from b200qp.envs import dyn_spec

    class Mystery(torch.nn.Module):
        def forward(self, x, u):
            return x

    with pytest.raises(NotImplementedError, match="no fused kernel"):
        dyn_spec(Mystery())


def test_tracking_mpc_matches_reference_golden(cuda_device):
    """The production caller (deqmpc/policies.py Tracking_MPC, solver_type='al'): three consecutive
    calls with warm start hand-off and gradients w.r.t. the tracked references, against the real
    reference run on its jit-scripted PendulumEnv (oracle/gen_golden_tracking.py)."""
    import types
    import numpy as np
    from b200qp import envs
    from b200qp.policies import Tracking_MPC
    g = _npz("tracking_pend_B8_T5.npz")
    dev = cuda_device
    B, T = g["x_ref0"].shape[:2]
    space = types.SimpleNamespace(high=np.array([3.0]), low=np.array([-3.0]))
    env = types.SimpleNamespace(nx=2, nu=1, nq=1, dt=0.05, dynamics=envs.PendulumDynamics(),
                                dynamics_derivatives=envs.PendulumDynamics_jac(), action_space=space)
    # PendulumEnv keeps Qlqr / Rlqr in float32 (deqmpc/envs.py:100-101): 0.01 enters rounded through float32
    args = types.SimpleNamespace(T=T, bsz=B, Q=torch.Tensor([10.0, 1.0]).double(),
                                 R=torch.Tensor([0.01]).double(), dtype="double", solver_type="al", qp_iter=2,
                                 eps=1e-2, warm_start=True, device=dev)
    mpc = Tracking_MPC(args, env)
    x0 = g["x0"].to(dev)
    mpc.reinitialize(x0, None)
    for k in range(3):
        x_ref = g[f"x_ref{k}"].to(dev).requires_grad_(True)
        u_ref = g[f"u_ref{k}"].to(dev).requires_grad_(True)
        xs, us = mpc(x0, None, x_ref, u_ref)
        (xs.sum() + 2 * us.sum()).backward()
        errs = dict(x=rel(xs.detach().cpu(), g[f"out_x{k}"]), u=rel(us.detach().cpu(), g[f"out_u{k}"]),
                    gx=rel(x_ref.grad.cpu(), g[f"g_xref{k}"]), gu=rel(u_ref.grad.cpu(), g[f"g_uref{k}"]))
        print("tracking call", k, {a: f"{b:.1e}" for a, b in errs.items()})
        assert errs["x"] <= RTOL32 and errs["u"] <= RTOL32, errs
        assert errs["gx"] <= RTOL64 and errs["gu"] <= RTOL64, errs


MYENVS = ("pendulum1l", "cartpole1l", "cartpole1l_v2", "cartpole2l")


def _my_dynamics(name, dt, dev):
    from b200qp import my_envs
    kw = dict(dtype=torch.float64, device=dev)
    if name == "pendulum1l":
        return my_envs.PendulumDynamics(nx=2, dt=dt, kwargs=kw)
    return my_envs.CartpoleDynamics(nx=6 if name == "cartpole2l" else 4, dt=dt, kwargs=kw, version=2 if name.endswith("v2") else 1)


@pytest.mark.parametrize("name", MYENVS)
def test_myenvs_dynamics_golden(name, cuda_device):
    """deqmpc/my_envs Dynamics.forward / derivatives / dynamics_derivatives (CasADi-generated code in the
    reference) against outputs of that generated code driven through the reference's own dynamics.py
    (oracle/gen_golden_myenvs.py), and against the oracle port on a larger seeded sample."""
    import numpy as np
    from oracle import myenvs_oracle as MO
    g = _npz(f"dyn_myenvs_{name}.npz")
    d = _my_dynamics(name, float(g["dt"]), cuda_device)
    x, u = g["x"].to(cuda_device), g["u"].to(cuda_device)
    xn, (A, B) = d.dynamics_derivatives(x, u)
    assert rel(xn.cpu(), g["xn"]) < 1e-13 and rel(A.cpu(), g["A"]) < 1e-12 and rel(B.cpu(), g["B"]) < 1e-12
    assert rel(d(x, u).cpu(), g["xn"]) < 1e-13
    A2, B2 = d.derivatives(x, u)
    assert torch.equal(A2, A) and torch.equal(B2, B)
    # autograd through the step uses the same Jacobians
    xr, ur = x.clone().requires_grad_(True), u.clone().requires_grad_(True)
    d(xr, ur).sum().backward()
    assert rel(xr.grad.cpu(), g["A"].sum(1)) < 1e-12 and rel(ur.grad.cpu(), g["B"].sum(1)) < 1e-12
    # larger sample against the port
    rs = np.random.RandomState(11)
    nq, N = MO.NQ[name], 4096
    xs = np.concatenate([rs.uniform(-7, 7, (N, nq)), rs.uniform(-6, 6, (N, nq))], 1)
    us = rs.uniform(-50, 50, (N, 1))
    pxn, (pA, pB) = MO.Dynamics(MO.PortPackage(name), 2 * nq, float(g["dt"])).dynamics_derivatives(xs, us)
    xn, (A, B) = d.dynamics_derivatives(torch.tensor(xs, device=cuda_device), torch.tensor(us, device=cuda_device))
    assert rel(xn.cpu(), pxn) < 1e-13 and rel(A.cpu(), pA) < 1e-12 and rel(B.cpu(), pB) < 1e-12


@pytest.mark.parametrize("case", ("cartpole1l_B8_T10", "cartpole2l_B4_T8", "pendulum1l_B8_T6", "cartpole1l_B8_T20"))
def test_al_mpc_myenvs_golden(case, cuda_device):
    """AL-MPC on the my_envs dynamics (the production configuration: CartpoleEnv + Tracking_MPC's al_mpc.MPC)
    against the real reference run on its generated code: cold call, warm-started call, state and backward."""
    from b200qp.AL_mpc import MPC
    from b200qp.al_utils import QuadCost
    g = _npz(f"mpc_{case}.npz")
    name = case.split("_B")[0]
    dev = cuda_device
    d = _my_dynamics(name, float(g["dt"]), dev)
    B, T, nu = g["u_init"].shape
    nx = g["x0"].shape[1]
    ub = float(g["umax"]) * torch.ones(nu, dtype=torch.float64, device=dev)
    ctrl = MPC(nx, nu, T, u_lower=-ub, u_upper=ub, exit_unconverged=False, eps=1e-5, n_batch=B, backprop=False, verbose=0,
               u_init=g["u_init"].to(dev), solver_type="dense", dtype=torch.float64)
    x0 = g["x0"].to(dev)
    ctrl.reinitialize(x0, None)
    ctrl.u_init = g["u_init"].to(dev)
    for k in range(2):
        Cfull = torch.diag_embed(g["Cd"]).to(dev).requires_grad_(True)
        c = (-(g["Cd"] * g["xref"])).to(dev).requires_grad_(True)
        x, u = ctrl(x0, QuadCost(Cfull, c), d, d.dynamics_derivatives)
        (x.sum() + u.sum()).backward()
        errs = dict(x=rel(x.detach().cpu(), g[f"out_x{k}"]), u=rel(u.detach().cpu(), g[f"out_u{k}"]),
                    lam=rel(ctrl.lamda_prev.cpu(), g[f"out_lam{k}"]), rho=rel(ctrl.rho_prev.cpu(), g[f"out_rho{k}"]),
                    dC=rel(Cfull.grad.diagonal(dim1=-2, dim2=-1).cpu(), g[f"out_dC{k}"]), dc=rel(c.grad.cpu(), g[f"out_dc{k}"]))
        print(case, k, {a: f"{b:.1e}" for a, b in errs.items()})
        assert errs["x"] <= RTOL32 and errs["u"] <= RTOL32, errs
        for key in ("lam", "rho", "dC", "dc"):
            assert errs[key] <= RTOL64, (key, errs)


@pytest.mark.parametrize("case", ["ipmpc_pendulum1l_B8_T5_single", "ipmpc_pendulum1l_B8_T5_sqp3",
                                  "ipmpc_cartpole1l_B4_T10_single", "ipmpc_cartpole1l_B4_T10_sqp3",
                                  "ipmpc_cartpole1l_B4_T20_single"])
def test_ip_mpc_matches_reference_golden(case, cuda_device):
    """b200qp.qp_wrapper.MPC (the interior-point MPC: SQP loop, DenseQPFunction with the NON-linear dynamics residual as
    its dyn_res callback, line search) against goldens of the real qpth.qp_wrapper.MPC on the reference's my_envs
    dynamics (oracle/gen_golden_ipmpc.py): nominal states / controls and the gradients w.r.t. the cost."""
    from b200qp import qp_wrapper as ip_mpc
    g = dict(np.load(os.path.join(GOLDEN, f"{case}.npz")))
    name = case.split("_")[1]
    d = _my_dynamics(name, float(g["dt"]), cuda_device)
    T, B, nu = g["u_init"].shape
    nx = g["x0"].shape[1]
    qp_iter = 1 if case.endswith("single") else 3
    um = float(g["umax"])
    dev = cuda_device
    ctrl = ip_mpc.MPC(nx, nu, T, u_lower=-um * torch.ones(nu, dtype=torch.float64, device=dev),
                      u_upper=um * torch.ones(nu, dtype=torch.float64, device=dev), qp_iter=qp_iter, exit_unconverged=False,
                      eps=1e-5, n_batch=B, backprop=False, verbose=0, u_init=torch.tensor(g["u_init"]).to(dev),
                      grad_method=ip_mpc.GradMethods.ANALYTIC, solver_type="dense", single_qp_solve=(qp_iter == 1))
    C = torch.diag(torch.tensor(g["Cd"])).repeat(T, B, 1, 1).to(dev).requires_grad_(True)
    c = torch.tensor(g["c"]).to(dev).requires_grad_(True)
    xs, us = ctrl(torch.tensor(g["x0"]).to(dev), ip_mpc.QuadCost(C, c), d, d.dynamics_derivatives)
    (xs.sum() + 2 * us.sum()).backward()
    errs = {k: rel(a.detach().cpu(), torch.tensor(g[n])) for k, a, n in (("x", xs, "out_x"), ("u", us, "out_u"))}
    # Gradients.  The outputs are x + alpha (x_hat - x) with alpha the step of the LAST line search (qp_wrapper.py:399-401),
    # so every gradient is alpha_j times the adjoint of the last QP.  In SQP mode that line search runs at a converged
    # point where `cost_new < cost` is rounding noise, and alpha_j is an arbitrary power of linesearch_decay in the
    # reference as here (oracle/gen_golden_ipmpc.py records it): compare gradient / alpha, per problem.
    a_ref, a_ours = torch.tensor(g["alpha"]), ctrl.info["alpha"].cpu()
    if qp_iter == 1:
        assert torch.equal(a_ref, a_ours), "single-QP mode: the line search starts far from convergence, alpha must agree"
    for k, a, n in (("dC", C.grad, "dC"), ("dc", c.grad, "dc")):
        ours = a.detach().cpu().transpose(0, 1).reshape(B, -1) / a_ours[:, None]
        ref = torch.tensor(g[n]).transpose(0, 1).reshape(B, -1) / a_ref[:, None]
        errs[k] = float(((ours - ref).norm(dim=1) / (ref.norm(dim=1) + ref.norm(dim=1).median())).max())
    print(case, "QP iterations", ctrl.info["qp_iters"], {k: f"{v:.1e}" for k, v in errs.items()}, "alpha ours", a_ours.tolist(),
          "reference", a_ref.tolist())
    for k, v in errs.items():
        assert v <= RTOL64, (k, errs)


def test_tracking_mpc_ip_branch(cuda_device):
    """Tracking_MPC(solver_type='ip') (deqmpc/policies.py:622-639, 649-663) drives b200qp.qp_wrapper.MPC.  The reference's
    own 'ip' branch raises before it reaches the solver (qp_wrapper.py:487, `.view` on the transposed u_init), so there is
    no golden: the shim is checked against a direct qp_wrapper.MPC call with the cost the reference would build
    (p = -Q x_ref), which is itself golden-tested above."""
    import types
    from b200qp import my_envs, policies, qp_wrapper as ip_mpc
    dev = cuda_device
    kw = dict(dtype=torch.float64, device=dev)
    env = my_envs.PendulumEnv(nx=2, dt=0.05, kwargs=kw)
    B, T = 8, 5
    args = types.SimpleNamespace(T=T, bsz=B, Q=torch.tensor([1.0, 1.0], dtype=torch.float64), R=torch.tensor([1e-2], dtype=torch.float64), dtype="double", solver_type="ip",
                                 qp_iter=1, eps=1e-2, warm_start=True, device=dev)
    torch.manual_seed(0)
    mpc = policies.Tracking_MPC(args, env)
    u_init = mpc.u_init.clone()
    x0 = torch.rand(B, 2, **kw) - 0.5
    x_ref, u_ref = (0.3 * torch.randn(B, T, 2, **kw)).requires_grad_(True), (0.3 * torch.randn(B, T, 1, **kw)).requires_grad_(True)
    xs, us = mpc(x0, torch.cat([x_ref, u_ref], -1), x_ref, u_ref)
    assert xs.shape == (B, T, 2) and us.shape == (B, T, 1)
    (xs.sum() + us.sum()).backward()
    assert torch.isfinite(x_ref.grad).all() and x_ref.grad.abs().sum() > 0
    ctrl = ip_mpc.MPC(2, 1, T, u_lower=mpc.u_lower.double(), u_upper=mpc.u_upper.double(), qp_iter=1, exit_unconverged=False, eps=1e-5,
                      n_batch=B, backprop=False, verbose=0, u_init=u_init.transpose(0, 1).contiguous(),
                      grad_method=ip_mpc.GradMethods.ANALYTIC, solver_type="dense", single_qp_solve=True)
    Q = torch.diag(torch.tensor([1.0, 1.0, 1e-2], **kw)).repeat(T, B, 1, 1)
    p = -(torch.tensor([1.0, 1.0, 1e-2], **kw) * torch.cat([x_ref, u_ref], -1).detach()).transpose(0, 1)
    xd, ud = ctrl(x0, ip_mpc.QuadCost(Q, p), env.dynamics, env.dynamics_derivatives)
    assert torch.equal(xd.transpose(0, 1), xs.detach()) and torch.equal(ud.transpose(0, 1), us.detach())
    assert torch.equal(mpc.u_init, us.detach())


def _policy_setup(dev, B=8, T=5):
    import types
    from b200qp import envs, policies

    class _Spaces:
        def __init__(self, low, high):
            self.low, self.high = low, high

    class IntegratorEnv:  # deqmpc/envs.py:246-268
        def __init__(self):
            self.dynamics, self.dynamics_derivatives = envs.IntegratorDynamics(), envs.IntegratorDynamics_jac()
            self.nx, self.nu, self.nq, self.dt = 2, 1, 1, 0.1
            self.action_space = _Spaces(-np.full(1, 2.0), np.full(1, 2.0))
            self.Qlqr, self.Rlqr = torch.Tensor([10.0, 1.00]), torch.Tensor([0.01])

    env = IntegratorEnv()
    args = types.SimpleNamespace(T=T, bsz=B, dtype="double", solver_type="al", nq=1, hdim=128, layer_type="mlp", deq_out_type=1,
                                 policy_out_type=1, kernel_width=3, pooling="mean", deq_iter=6, qp_iter=1, eps=1e-2, warm_start=True,
                                 device=dev, deq=True, en_qp_solve=True, Q=env.Qlqr, R=env.Rlqr)
    return policies, env, args


def test_deqmpc_policy_matches_reference_golden(cuda_device):
    """DEQMPCPolicy (deqmpc/policies.py:426-529: six rounds of DEQLayer -> Tracking_MPC, the run.sh configuration on the
    integrator) with the reference's own weights against the reference's trajectories of every DEQ iteration, its loss
    (policies.py:800-808) and the gradients of all network parameters (oracle/gen_golden_policy.py).

    The chain is chaotic at float32 noise level IN THE REFERENCE (measured by the generator: 1e-7 on the input moves DEQ
    iteration 2 by 1e-2 -- the AL solve's discrete 20-way line search, fed back six times), so the golden and this test run
    the network in float64 (two dtype casts, nothing else changed); the native float32 policy is compared on iterations 0-1,
    before the amplification sets in."""
    policies, env, args = _policy_setup(cuda_device)
    g = dict(np.load(os.path.join(GOLDEN, "policy_integrator_B8_T5.npz")))
    sd = {k[2:]: torch.tensor(v) for k, v in g.items() if k.startswith("w_")}
    t = lambda k, dt: torch.tensor(g[k]).to(device=cuda_device, dtype=dt)

    def run(dt):
        policy = policies.DEQMPCPolicy(args, env)
        policy.model.load_state_dict(sd)
        if dt == torch.float64:
            policy.model.double()
            policy.model.init_z = lambda bsz: torch.zeros(bsz, args.hdim, dtype=torch.float64, device=cuda_device)
            fwd = policy.tracking_mpc.forward
            policy.tracking_mpc.forward = lambda *a: tuple(o.double() for o in fwd(*a))
        x, gs, ga, gm = t("x", dt), t("gt_states", dt), t("gt_actions", dt), t("mask", dt)
        trajs, dyn_res = policy(x, gs, ga, gm, qp_solve=True)
        loss, _ = policies.compute_loss(policy, gs, ga, gm, trajs, args)
        loss.backward()
        errs = [[rel(ten.detach().cpu().double(), torch.tensor(g[f"{nm}{k}"]).double()) for nm, ten in (("net", a), ("xs", b), ("us", c))]
                for k, (a, b, c) in enumerate(trajs)]
        gw = max(rel(p.grad.cpu().double(), torch.tensor(g["g_" + n]).double()) for n, p in policy.model.named_parameters())
        return errs, gw, float(loss.detach()), dyn_res

    errs, gw, loss, dyn_res = run(torch.float64)
    worst = max(max(e) for e in errs)
    print(f"policy (float64 network): worst trajectory rel err {worst:.2e} per iteration {[f'{max(e):.1e}' for e in errs]}, loss {loss:.9f} "
          f"(reference {float(g['loss']):.9f}), dyn_res {dyn_res:.9f} ({float(g['dyn_res']):.9f}), worst parameter-gradient rel err {gw:.2e}")
    # the MPC hands float32 values back (AL_mpc.py:319-320): one of them rounding the other way (our fp64 solve differs from
    # the reference's by ~1e-12) is a 6e-8 perturbation that the remaining iterations amplify (reference: 1e-12 -> 5e-7 over
    # six iterations).  Hence: early iterations to rounding, the whole chain at 1e-5, the parameter gradients at 1e-4.
    assert max(max(e) for e in errs[:3]) <= 1e-9 and worst <= 1e-5 and gw <= 1e-4
    assert abs(loss - float(g["loss"])) <= 1e-6 * abs(float(g["loss"]))
    assert abs(dyn_res - float(g["dyn_res"])) <= 1e-6 * abs(float(g["dyn_res"]))
    errs32, _, loss32, _ = run(torch.float32)
    print(f"policy (native float32 network): per iteration {[f'{max(e):.1e}' for e in errs32]}, loss {loss32:.6f}")
    assert max(errs32[0]) <= 1e-4 and max(errs32[1]) <= 1e-4


def test_graphed_train_step_equals_eager(cuda_device):
    """GraphedTrainStep: the policy forward + loss + backward captured once in a CUDA graph gives the same loss and the
    same parameter gradients as the eager step, step after step (Adam updates included)."""
    policies, env, args = _policy_setup(cuda_device, B=16)
    torch.manual_seed(1)
    pol_a = policies.DEQMPCPolicy(args, env)
    pol_b = policies.DEQMPCPolicy(args, env)
    pol_b.model.load_state_dict(pol_a.model.state_dict())
    kw = dict(device=cuda_device)
    batches = [(torch.rand(16, 2, **kw) * 2 - 1, 0.5 * torch.randn(16, 5, 2, **kw), 0.5 * torch.randn(16, 5, 1, **kw), torch.ones(16, 5, **kw))
               for _ in range(3)]
    opt_a = torch.optim.Adam(pol_a.model.parameters(), lr=1e-3)
    opt_b = torch.optim.Adam(pol_b.model.parameters(), lr=1e-3)
    eager = policies.GraphedTrainStep(pol_a, opt_a, args, batches[0], use_graph=False)
    sd0 = {k: v.clone() for k, v in pol_b.model.state_dict().items()}
    graphed = policies.GraphedTrainStep(pol_b, opt_b, args, batches[0], use_graph=True)
    pol_b.model.load_state_dict(sd0)   # the warm-up passes of the capture do not step the optimizer, but be explicit
    for bt in batches:
        la, lb = eager(*bt), graphed(*bt)
        assert abs(float(la) - float(lb)) <= 1e-5 * abs(float(la)), (float(la), float(lb))
    for (n, pa), (_, pb) in zip(pol_a.model.named_parameters(), pol_b.model.named_parameters()):
        assert rel(pb.detach().cpu().double(), pa.detach().cpu().double()) <= 1e-4, n


@pytest.mark.parametrize("name", ["cartpole1l", "cartpole2l"])
def test_envs_v1_against_reference_modules(name, cuda_device):
    """deqmpc/envs_v1.py OneLinkCartpoleDynamics / TwoLinkCartpoleDynamics (closed-form accelerations, classical RK4) against
    goldens of the reference's own modules (oracle/gen_golden_envs_v1.py): next state and both Jacobians."""
    from b200qp import envs
    g = dict(np.load(os.path.join(GOLDEN, f"dyn_envsv1_{name}.npz")))
    mod = envs.OneLinkCartpoleDynamics_jac() if name == "cartpole1l" else envs.TwoLinkCartpoleDynamics_jac()
    fwd = envs.OneLinkCartpoleDynamics() if name == "cartpole1l" else envs.TwoLinkCartpoleDynamics()
    x, u = torch.tensor(g["x"]).to(cuda_device), torch.tensor(g["u"]).to(cuda_device)
    xn, (A, B) = mod(x, u)
    assert rel(xn.cpu(), torch.tensor(g["xn"])) < 1e-13
    assert rel(fwd(x, u).cpu(), torch.tensor(g["xn"])) < 1e-13
    assert rel(A.cpu(), torch.tensor(g["A"])) < 1e-11
    assert rel(B.cpu(), torch.tensor(g["B"])) < 1e-11


def test_global_slab_mode_matches_goldens(cuda_device):
    """The code paths of problems whose state does not fit shared memory (global scratch slab: staged block Cholesky, staged
    line search, factor built in the output buffer) driven with the small goldens: B200MPC_FORCE_GLOBAL=1 in a child process
    (the switch is read once per process) must pass the same parity tests as the shared-memory mode."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, B200MPC_FORCE_GLOBAL="1")
    sel = ("test_al_mpc_matches_reference_golden or test_al_mpc_myenvs_golden or test_tracking_mpc_matches_reference_golden or "
           "test_al_mpc_float32_solver_precision or test_al_mpc_vs_oracle_seeded_large_horizon")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_mpc_parity_gpu.py"), "-q", "-m", "gpu", "-x",
                        "-k", sel, "-p", "no:cacheprovider"], cwd=root, env=env, capture_output=True, text=True, timeout=900)
    tail = (r.stdout + r.stderr)[-1500:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "no tests ran" not in r.stdout, tail
    print(r.stdout.strip().splitlines()[-1])
