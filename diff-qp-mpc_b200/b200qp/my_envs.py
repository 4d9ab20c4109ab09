"""The `deqmpc/my_envs` dynamics family behind its reference call surface:

    env = CartpoleEnv(nx=4, dt=0.05, kwargs=dict(dtype=torch.float64, device="cuda"))
    x_next = env.dynamics(state, action)                       # (bsz,nx)     dynamics.py:27-66
    A, B = env.dynamics.derivatives(state, action)             # (bsz,nx,nx), (bsz,nx,1)   dynamics.py:68-108
    x_next, (A, B) = env.dynamics_derivatives(state, action)   # dynamics.py:249-258

The reference evaluates CasADi-generated code through one torch extension per model
(`cartpole1l`, `cartpole1l_v2`, `cartpole2l`, `pendulum1l`: src/dynamics_gpu.cu) and re-assembles six
Jacobian blocks with cat/transpose; here ONE fused kernel (b200dyn_step / b200dyn_jac) returns the next
state and both Jacobians, and `Tracking_MPC` / `AL_mpc.MPC` recognise these objects (or the reference's own
`CartpoleDynamics` / `PendulumDynamics` modules) and run the model inside the fused AL solve.
float64 only, like the reference (dynamics.py:48).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .envs import _Step, _run

G = 9.81
# parameter vectors of include/b200mpc.h, without the leading dt
MODELS = {
    "pendulum1l": (_lib.ENV_PENDULUM1L, 2, [4.0, 2.0 * G]),
    "cartpole1l": (_lib.ENV_CARTPOLE1L, 4, [11.0, 1.0, 2.0, G]),
    "cartpole1l_v2": (_lib.ENV_CARTPOLE1L, 4, [0.7, 0.1, 0.05, G]),
    "cartpole2l": (_lib.ENV_CARTPOLE2L, 6, [12.0, 2.0, 1.0, 3.0, 2.0, 1.0, G]),
}


def package_spec(package, dt):
    """(env id, params, nx, nu) for a model name or for the reference's extension module of that name."""
    name = package if isinstance(package, str) else getattr(package, "__name__", str(package)).split(".")[-1]
    if name not in MODELS:
        raise NotImplementedError(f"b200qp: no fused kernel for my_envs package {name!r}; supported: {sorted(MODELS)}")
    env, nx, par = MODELS[name]
    return env, [float(dt)] + par, nx, 1


class Dynamics(torch.nn.Module):
    """deqmpc/my_envs/dynamics.py:15-108,249-258"""

    def __init__(self, nx=None, dt=0.01, kwargs=None):
        super().__init__()
        assert nx is not None
        self.nx, self.nu, self.nq, self.dt, self.kwargs = nx, 1, nx // 2, dt, kwargs

    def spec(self):
        return package_spec(self.package, self.dt)

    def _check(self, state, action):
        if state.dim() == 3:
            state, action = state.reshape(-1, self.nx), action.reshape(-1, self.nu)
        assert state.dim() == 2 and state.size(1) == self.nx
        assert action.dim() == 2 and action.size(1) == self.nu
        assert state.dtype == torch.float64
        return state, action

    def forward(self, state, action):
        lead = state.shape[:-1]
        state, action = self._check(state, action)
        return _Step.apply(state, action, self.spec()).reshape(*lead, self.nx)

    def derivatives(self, state, action):
        state, action = self._check(state, action)
        _, A, Bm = _run(self.spec(), state, action, True)
        return A, Bm

    def dynamics_derivatives(self, state, action):
        state, action = self._check(state, action)
        xn, A, Bm = _run(self.spec(), state, action, True)
        return xn, (A, Bm)


class CartpoleDynamics(Dynamics):
    """deqmpc/my_envs/cartpole.py:29-41"""

    def __init__(self, nx=None, dt=0.01, kwargs=None, version=1):
        super().__init__(nx, dt, kwargs)
        if nx == 6:
            self.package = "cartpole2l"
        elif nx == 4:
            self.package = "cartpole1l" if version == 1 else "cartpole1l_v2"
        else:
            raise NotImplementedError


class PendulumDynamics(Dynamics):
    """deqmpc/my_envs/pendulum.py:19-36 (the reference has no pendulum2l package either)"""

    def __init__(self, nx=None, dt=0.01, kwargs=None):
        super().__init__(nx, dt, kwargs)
        if nx != 2:
            raise NotImplementedError
        self.package = "pendulum1l"


class _Spaces:
    def __init__(self, low, high, shape):
        self.low, self.high, self.shape = low, high, shape


def angle_normalize_2pi(x):
    """deqmpc/utils.py: wrap to [0, 2 pi)"""
    return x % (2 * np.pi)


class _Env(torch.nn.Module):
    def action_clip(self, action):
        return torch.clamp(action, -self.u_bounds, self.u_bounds)

    def state_clip(self, state):
        state[..., 1: self.nq] = angle_normalize_2pi(state[..., 1: self.nq])
        return state

    def seed(self, seed):
        np.random.seed(seed)
        torch.manual_seed(seed)

    def step(self, action):
        action = self.action_clip(torch.as_tensor(action, **self.kwargs))
        self.state = self.state_clip(self.dynamics(self.state, action))
        self.num_steps += 1
        return self.state


class CartpoleEnv(_Env):
    """deqmpc/my_envs/cartpole.py:43-139: the attributes Tracking_MPC and the training loop read
    (`dynamics`, `dynamics_derivatives`, nx/nu/nq/dt, action_space, Qlqr/Rlqr, T, u_bounds) and `reset`."""

    def __init__(self, nx=None, dt=0.05, stabilization=False, kwargs=None):
        super().__init__()
        assert nx is not None
        self.dynamics = CartpoleDynamics(nx=nx, dt=dt, kwargs=kwargs)
        self.dynamics_derivatives = self.dynamics.dynamics_derivatives
        self.nx, self.nq, self.nu, self.dt, self.kwargs = nx, self.dynamics.nq, self.dynamics.nu, dt, kwargs
        self.spec_id = "Cartpole{}l-v0{}".format(nx // 2 - 1, "-stabilize" if stabilization else "")
        self.stabilization, self.num_successes, self.bsz, self.num_steps = stabilization, 0, 1, 0
        self.T, self.u_bounds = (300, 250.0) if nx == 6 else (200, 100.0)
        high = np.concatenate((np.full(self.nq, np.pi), np.full(self.nq, np.pi * 5)))
        self.observation_space = _Spaces(-high, high, (self.nx,))
        self.action_space = _Spaces(np.full(self.nu, -self.u_bounds), np.full(self.nu, self.u_bounds), (self.nu,))
        self.Qlqr = torch.ones(self.nx, **self.kwargs)
        self.Rlqr = torch.ones(self.nu, **self.kwargs) * 0.00000001

    def reset(self, bsz=None):
        """cartpole.py:103-139"""
        if self.stabilization:
            high = np.concatenate((np.full(self.nq, 0.05), np.full(self.nq, 0.05)))
            high[0], high[1] = 0.1, 0.1
            offset = torch.tensor([np.pi, 0.0] * self.nq, **self.kwargs)
            offset[0], offset[1] = 0.0, 0.0
            self.state = torch.tensor(np.random.uniform(low=-high, high=high), **self.kwargs) + offset
        else:
            high = np.concatenate((np.full(self.nq, np.pi), np.full(self.nq, np.pi)))
            high = high[None].repeat(self.bsz, 0)
            self.state = self.state_clip(torch.tensor(np.random.uniform(low=-high, high=high), **self.kwargs))
        self.num_successes = self.num_steps = 0
        return self.state


class PendulumEnv(_Env):
    """deqmpc/my_envs/pendulum.py:38-66"""

    def __init__(self, nx=None, dt=0.01, stabilization=False, kwargs=None):
        super().__init__()
        assert nx is not None
        self.dynamics = PendulumDynamics(nx=nx, dt=dt, kwargs=kwargs)
        self.dynamics_derivatives = self.dynamics.dynamics_derivatives
        self.nx, self.nq, self.nu, self.dt, self.kwargs = nx, self.dynamics.nq, self.dynamics.nu, dt, kwargs
        self.spec_id = "Pendulum{}l-v0{}".format(nx // 2 - 1, "-stabilize" if stabilization else "")
        self.stabilization, self.num_successes, self.u_bounds, self.num_steps = stabilization, 0, 3.0, 0
        high = np.concatenate((np.full(self.nq, np.pi), np.full(self.nq, np.pi * 5)))
        self.observation_space = _Spaces(-high, high, (self.nx, 2))
        self.action_space = _Spaces(np.full(self.nu, -self.u_bounds), np.full(self.nu, self.u_bounds), (self.nu, 2))
