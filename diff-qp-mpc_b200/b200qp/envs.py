"""Batched dynamics modules with the reference's class names and call signatures
(deqmpc/envs.py:5-54,68-82,182-233; qpth/env_dx/pendulum.py:16-84; qpth/env_dx/cartpole.py:27-96),
evaluated by the fused step / step+Jacobian kernels behind b200dyn_step / b200dyn_jac.

`dyn_spec(dx)` maps a dynamics object -- one of these classes OR the reference's own module of the
same name (also when wrapped by torch.jit.script) -- to the (env id, parameter vector) the MPC
kernels are specialised on.  An unknown dynamics object is an error: there is no path that calls
back into Python from the solver.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _params_array(params):
    return (ctypes.c_double * _lib.MPC_MAX_PARAMS)(*(list(params) + [0.0] * (_lib.MPC_MAX_PARAMS - len(params))))


def dyn_spec(dx):
    """-> (env id, params, nx, nu)"""
    if hasattr(dx, "__self__") and hasattr(dx, "__func__"):
        dx = dx.__self__          # bound method, e.g. env.dynamics_derivatives (deqmpc/policies.py:577)
    name = getattr(dx, "original_name", type(dx).__name__)
    base = name[:-4] if name.endswith("_jac") else name
    if hasattr(dx, "package"):    # deqmpc/my_envs: CartpoleDynamics / PendulumDynamics around a generated package
        from .my_envs import package_spec
        return package_spec(dx.package, dx.dt)
    if base == "PendulumDynamics":
        return _lib.ENV_PENDULUM, [float(dx.dt), float(dx.g), float(dx.m), float(dx.l)], 2, 1
    if base == "IntegratorDynamics":
        if int(dx.nx) != 2 or int(dx.nu) != 1:
            raise NotImplementedError("b200qp: IntegratorDynamics kernels are built for nx=2, nu=1")
        return _lib.ENV_INTEGRATOR, [float(dx.dt)], 2, 1
    if base == "PendulumDx":
        if hasattr(dx, "simple") and not dx.simple:
            raise NotImplementedError("b200qp: only the `simple` PendulumDx model is implemented")
        g, m, l = (float(v) for v in dx.params[:3])
        return _lib.ENV_PENDULUM_DX, [float(dx.dt), g, m, l, float(dx.max_torque)], 3, 1
    if base == "CartpoleDx":
        pr = torch.as_tensor(dx.params).detach().to(dtype=torch.float32, device="cpu")
        gravity, masscart, masspole, length = pr.unbind()
        total_mass = masspole + masscart          # float32 arithmetic, as in the reference
        polemass_length = masspole * length
        return (_lib.ENV_CARTPOLE_DX, [float(dx.dt), float(gravity), float(masscart), float(masspole), float(length),
                                       float(total_mass), float(polemass_length), float(dx.force_mag)], 5, 1)
    if base == "OneLinkCartpoleDynamics":   # deqmpc/envs_v1.py:28-82
        return _lib.ENV_CARTPOLE1L_V1, [float(dx.dt), float(dx.M), float(dx.m), float(dx.l), float(dx.g)], 4, 1
    if base == "TwoLinkCartpoleDynamics":   # deqmpc/envs_v1.py:226-310 (numeric constants are part of the formulas)
        return _lib.ENV_CARTPOLE2L_V1, [float(dx.dt)], 6, 1
    if base == "RexQuadrotor_dynamics":
        f32 = lambda t: [float(v) for v in torch.as_tensor(t).detach().to(dtype=torch.float32, device="cpu").reshape(-1)]
        mg = torch.as_tensor(dx.g).detach().to(dtype=torch.float32, device="cpu") * float(dx.m)  # float32, as self.m * self.g
        params = ([float(dx.dt), float(dx.m), float(dx.act_scale), 0.0244101, float(dx.kf), float(dx.km), float(dx.bf),
                   float(dx.motor_dist)] + f32(mg) + f32(dx.Bf) + f32(dx.J) + f32(dx.Jinv) + f32(dx.ss) + f32(dx.cd)
                  + f32(dx.cross_A))
        assert len(params) == 50
        return _lib.ENV_REX_QUADROTOR, params, 12, 4
    raise NotImplementedError(f"b200qp: no fused kernel for dynamics {name!r}; supported: PendulumDynamics, "
                              "IntegratorDynamics, PendulumDx, CartpoleDx, RexQuadrotor_dynamics (and their *_jac "
                              "variants), the deqmpc/my_envs CartpoleDynamics / PendulumDynamics and the deqmpc/envs_v1 "
                              "OneLinkCartpoleDynamics / TwoLinkCartpoleDynamics")


def _run(spec, x, u, want_jac):
    env, params, nx, nu = spec
    if not x.is_cuda:
        raise RuntimeError("b200qp dynamics run on CUDA tensors only (no CPU fallback)")
    if x.dtype not in (torch.float64, torch.float32):
        raise RuntimeError("b200qp dynamics: float64 or float32 states")
    lead = x.shape[:-1]
    xf = x.detach().reshape(-1, nx).contiguous()
    uf = u.detach().to(x.dtype).reshape(-1, nu).contiguous()
    N = xf.shape[0]
    xn = torch.empty_like(xf)
    A = torch.empty(N, nx, nx, device=x.device, dtype=x.dtype) if want_jac else None
    Bm = torch.empty(N, nx, nu, device=x.device, dtype=x.dtype) if want_jac else None
    L = _lib.lib()
    code = _lib.F64 if x.dtype == torch.float64 else _lib.F32
    st = ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
    with torch.cuda.device(x.device):
        if want_jac:
            rc = L.b200dyn_jac(env, code, _params_array(params), _p(xf), _p(uf), _p(xn), _p(A), _p(Bm), N, st)
        else:
            rc = L.b200dyn_step(env, code, _params_array(params), _p(xf), _p(uf), _p(xn), N, st)
    _lib.check(rc, "b200dyn")
    return xn.reshape(*lead, nx), A, Bm


def rollout(spec, x0, u):
    """xs (B,T,nx) with xs[:,0] = x0 and xs[:,t+1] = f(xs[:,t], u[:,t]) in ONE launch."""
    env, params, nx, nu = spec
    if not x0.is_cuda:
        raise RuntimeError("b200qp dynamics run on CUDA tensors only (no CPU fallback)")
    B, T = u.shape[0], u.shape[1]
    x0c = x0.detach().contiguous()
    uc = u.detach().to(x0.dtype).contiguous()
    xs = torch.empty(B, T, nx, device=x0.device, dtype=x0.dtype)
    code = _lib.F64 if x0.dtype == torch.float64 else _lib.F32
    st = ctypes.c_void_p(torch.cuda.current_stream(x0.device).cuda_stream)
    with torch.cuda.device(x0.device):
        rc = _lib.lib().b200dyn_rollout(env, code, _params_array(params), _p(x0c), _p(uc), _p(xs), B, T, st)
    _lib.check(rc, "b200dyn_rollout")
    return xs


class _Step(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, u, spec):
        need = x.requires_grad or u.requires_grad
        xn, A, Bm = _run(spec, x, u, need)
        if need:
            ctx.save_for_backward(A, Bm)
        ctx.shapes = (x.shape, u.shape)
        return xn

    @staticmethod
    def backward(ctx, g):
        A, Bm = ctx.saved_tensors
        gf = g.reshape(-1, 1, A.shape[1])
        gx = torch.bmm(gf, A).reshape(ctx.shapes[0])
        gu = torch.bmm(gf, Bm).reshape(ctx.shapes[1])
        return gx, gu, None


class _Dynamics(torch.nn.Module):
    def spec(self):
        return dyn_spec(self)

    def forward(self, state, action):
        return _Step.apply(state, action, self.spec())


class _DynamicsJac(_Dynamics):
    """forward(x (N,nx), u (N,nu)) -> (x_next (N,nx), (df/dx (N,nx,nx), df/du (N,nx,nu))), the
    return convention of the reference's *_jac modules (deqmpc/envs.py:74-82)."""

    def forward(self, x, u):
        xn, A, Bm = _run(self.spec(), x, u, True)
        return xn, (A, Bm)


class PendulumDynamics(_Dynamics):
    """deqmpc/envs.py:5-54"""

    def __init__(self):
        super().__init__()
        self.dt, self.max_torque, self.g, self.m, self.l = 0.05, 3.0, 10.0, 1.0, 1.0
        self.nx, self.nu = 2, 1

    def action_clip(self, action):
        return torch.clamp(action, -self.max_torque, self.max_torque)


class PendulumDynamics_jac(_DynamicsJac, PendulumDynamics):
    """deqmpc/envs.py:68-82"""


class IntegratorDynamics(_Dynamics):
    """deqmpc/envs.py:182-214"""

    def __init__(self, nx=2, nu=1, dt=0.1, max_acc=1, max_vel=1):
        super().__init__()
        self.dt, self.max_acc, self.max_vel, self.nx, self.nu = dt, max_acc, max_vel, nx, nu
        self.nq = int(nx / 2)

    def action_clip(self, action):
        return torch.clamp(action, -self.max_acc, self.max_acc)


class IntegratorDynamics_jac(_DynamicsJac, IntegratorDynamics):
    """deqmpc/envs.py:217-233"""


class _EnvDx(_Dynamics):
    def get_true_obj(self):
        """qpth/env_dx/pendulum.py:122-131 / cartpole.py:145-154"""
        q = torch.cat((self.goal_weights, self.ctrl_penalty * torch.ones(self.n_ctrl)))
        px = -torch.sqrt(self.goal_weights) * self.goal_state
        p = torch.cat((px, torch.zeros(self.n_ctrl)))
        return q, p


class PendulumDx(_EnvDx):
    """qpth/env_dx/pendulum.py:16-84 (simple model)"""

    def __init__(self, params=None, simple=True):
        super().__init__()
        self.simple = simple
        self.max_torque, self.dt, self.n_state, self.n_ctrl = 2.0, 0.05, 3, 1
        self.params = torch.tensor((10., 1., 1.)) if params is None else params
        self.goal_state = torch.Tensor([1., 0., 0.])
        self.goal_weights = torch.Tensor([1., 1., 0.1])
        self.ctrl_penalty = 0.001
        self.lower, self.upper = -2., 2.
        self.mpc_eps, self.linesearch_decay, self.max_linesearch_iter = 1e-3, 0.2, 5


class PendulumDx_jac(_DynamicsJac, PendulumDx):
    """Jacobian companion the reference lacks for env_dx: same return convention as
    deqmpc/envs.py:74-82."""


class CartpoleDx(_EnvDx):
    """qpth/env_dx/cartpole.py:27-96"""

    def __init__(self, params=None):
        super().__init__()
        self.n_state, self.n_ctrl = 5, 1
        self.params = torch.tensor((9.8, 1.0, 0.1, 0.5)) if params is None else params
        assert len(self.params) == 4
        self.force_mag = 100.
        self.dt = 0.05
        self.lower, self.upper = -self.force_mag, self.force_mag
        self.goal_state = torch.Tensor([0., 0., 1., 0., 0.])
        self.goal_weights = torch.Tensor([0.1, 0.1, 1., 1., 0.1])
        self.ctrl_penalty = 0.001
        self.mpc_eps, self.linesearch_decay, self.max_linesearch_iter = 1e-4, 0.5, 2


class CartpoleDx_jac(_DynamicsJac, CartpoleDx):
    """Jacobian companion (see PendulumDx_jac)."""


class RexQuadrotor_dynamics(_Dynamics):
    """deqmpc/rex_quadrotor.py:7-49: 12-state MRP rigid body, RK4, controls scaled by 100.  Model
    constants live in float32 tensors exactly as in the reference (they enter the fp64 arithmetic
    rounded through float32)."""

    def __init__(self, bsz=1, mass=2.0,
                 J=[[0.01566089, 0.00000318037, 0.0], [0.00000318037, 0.01562078, 0.0], [0.0, 0.0, 0.02226868]],
                 gravity=[0, 0, -9.81], motor_dist=0.28, kf=0.0244101, bf=-30.48576, km=0.00029958, bm=-0.367697,
                 quad_min_throttle=1148.0, quad_max_throttle=1832.0, ned=False, cross_A_x=0.25, cross_A_y=0.25,
                 cross_A_z=0.5, cd=[0.0, 0.0, 0.0], max_steps=100, dt=0.05, device=torch.device('cpu'), jacobian=False):
        super().__init__()
        import numpy as np
        self.m = mass
        Jn = np.array(J)
        self.J = (torch.diag(torch.FloatTensor(Jn)) if Jn.ndim == 1 else torch.FloatTensor(Jn)).unsqueeze(0)
        self.Jinv = torch.linalg.inv(self.J)
        self.g = torch.FloatTensor(gravity).unsqueeze(0)
        self.motor_dist, self.kf, self.km, self.bf, self.bm, self.bsz = motor_dist, kf, km, bf, bm, bsz
        self.Bf = torch.zeros((1, 3))
        self.Bf[0, 2] = 4 * bf
        self.quad_min_throttle, self.quad_max_throttle, self.ned = quad_min_throttle, quad_max_throttle, ned
        self.cross_A = torch.FloatTensor(np.array([cross_A_x, cross_A_y, cross_A_y])).unsqueeze(0)
        self.nx = self.state_dim = 12
        self.nu = self.control_dim = 4
        self._max_episode_steps = max_steps
        self.dt = dt
        self.act_scale = 100.0
        self.cd = torch.tensor(cd).unsqueeze(0)
        ss = torch.tensor([[1., 1, 0], [1., -1, 0], [-1., -1, 0], [-1., 1, 0]]).unsqueeze(0)
        self.ss = ss / ss.norm(dim=-1).unsqueeze(-1)
        self.device = device


class RexQuadrotor_dynamics_jac(_DynamicsJac, RexQuadrotor_dynamics):
    """deqmpc/rex_quadrotor.py:130-146"""


def _angle_normalize_2pi(x):
    """deqmpc/envs_v1.py angle_normalize_2pi"""
    return ((x) % (2 * torch.pi))


class OneLinkCartpoleDynamics(_Dynamics):
    """deqmpc/envs_v1.py:28-94: cart + one pole, closed-form accelerations, classical RK4 (dt = 0.01)"""

    def __init__(self):
        super().__init__()
        self.dt, self.max_force, self.g, self.M, self.m, self.l = 0.01, 500.0, -9.81, 0.5, 0.2, 0.5
        self.n = 1
        self.nx, self.nu = 2 * self.n + 2, 1
        self.np = self.nx // 2

    def action_clip(self, action):
        return torch.clamp(action, -self.max_force, self.max_force)

    def state_clip(self, state):
        state[..., 1:self.np] = _angle_normalize_2pi(state[..., 1:self.np])
        return state


class OneLinkCartpoleDynamics_jac(_DynamicsJac, OneLinkCartpoleDynamics):
    """Jacobian companion (forward-mode duals in the kernel); the reference has none for envs_v1 (its users are the RL /
    data-generation scripts), the return convention is that of deqmpc/envs.py:74-82."""


class TwoLinkCartpoleDynamics(_Dynamics):
    """deqmpc/envs_v1.py:226-321: the OpenOCL double cart-pole (M = 5, m1 = m2 = l1 = l2 = 1), classical RK4 (dt = 0.05)"""

    def __init__(self):
        super().__init__()
        self.dt, self.max_force, self.g, self.M, self.m1, self.m2, self.l1, self.l2 = 0.05, 5.0, 9.81, 5., 1., 1., 1., 1.
        self.n = 2
        self.nx, self.nu = 2 * self.n + 2, 1
        self.np = self.nx // 2

    def action_clip(self, action):
        return torch.clamp(action, -self.max_force, self.max_force)

    def state_clip(self, state):
        state[..., 1:self.np] = _angle_normalize_2pi(state[..., 1:self.np])
        return state


class TwoLinkCartpoleDynamics_jac(_DynamicsJac, TwoLinkCartpoleDynamics):
    """Jacobian companion (see OneLinkCartpoleDynamics_jac)."""
