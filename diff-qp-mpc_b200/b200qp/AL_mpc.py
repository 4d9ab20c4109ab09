"""al_mpc.MPC -- drop-in for the reference's qpth/AL_mpc.py:50-439 on B200.

    ctrl = MPC(n_state, n_ctrl, T, u_lower=..., u_upper=..., n_batch=B, ...)
    ctrl.reinitialize(x0, mask)
    x, u = ctrl(x0, QuadCost(C, c), dx, dx_jac)          # (B,T,nx) (B,T,nu) float32

Same constructor signature, same statefulness (warm start through `x_init`, `u_init`, `lamda_prev`,
`rho_prev`, `cost_lam_hist`, `just_initialized`), same float32 outputs, gradients w.r.t. the cost
(C, c) through the implicit backward of the last AL iteration.  The augmented-Lagrangian outer loop,
its Newton steps, the block-tridiagonal Cholesky, the 20-way line search and the batched dynamics
+ Jacobians all run inside ONE kernel launch per call (csrc/mpc_al.cuh); this file only keeps the
module state and unwraps tensors.  `dx` / `dx_jac` select the fused dynamics by type (b200qp.envs
classes or the reference's modules of the same names); arbitrary Python callables are rejected --
there is no CPU / Python fallback inside the solver.
"""
from __future__ import annotations

from enum import Enum

import torch
from torch.nn import Module

from . import al_utils
from .al_utils import ALSolve, ALState, QuadCost, LinDx  # noqa: F401
from .envs import dyn_spec, rollout as _rollout


class GradMethods(Enum):
    """qpth/AL_mpc.py:22-26"""
    AUTO_DIFF = 1
    FINITE_DIFF = 2
    ANALYTIC = 3
    ANALYTIC_CHECK = 4


def _detach_maybe(x):
    """qpth/util.py:204-207"""
    if x is None:
        return None
    return x if not x.requires_grad else x.detach()


class MPC(Module):
    """Differentiable box-constrained MPC by an augmented-Lagrangian Newton method
    (qpth/AL_mpc.py:50-195 for the constructor contract)."""

    def __init__(self, n_state, n_ctrl, T,
                 u_lower=None, u_upper=None,
                 u_init=None, x_init=None,
                 al_iter=2, verbose=0, eps=1e-7, back_eps=1e-7, n_batch=None,
                 linesearch_decay=0.2, max_linesearch_iter=10,
                 exit_unconverged=True, detach_unconverged=True, backprop=True,
                 slew_rate_penalty=None, solver_type='dense',
                 add_goal_constraint=False, x_goal=None, diag_cost=True,
                 ineqG=None, ineqh=None, dtype=torch.float64):
        super().__init__()
        assert (u_lower is None) == (u_upper is None)
        assert max_linesearch_iter > 0
        if u_lower is None:
            raise NotImplementedError("b200qp AL-MPC: control bounds are required (the reference's al_utils "
                                      "always builds the bound residuals, al_utils.py:266-271)")
        if add_goal_constraint or ineqG is not None or not diag_cost:
            raise NotImplementedError("b200qp AL-MPC: goal constraints, general inequalities and dense costs are not "
                                      "implemented (the reference's al_utils path ignores them as well)")
        self.dtype = dtype
        self.n_state, self.n_ctrl, self.T = n_state, n_ctrl, T
        self.u_lower = _detach_maybe(torch.as_tensor(u_lower).to(self.dtype))
        self.u_upper = _detach_maybe(torch.as_tensor(u_upper).to(self.dtype))
        self.x_upper = self.x_lower = None
        self.x_goal, self.ineqG, self.ineqh = x_goal, ineqG, ineqh
        self.u_init = _detach_maybe(u_init)
        self.x_init = _detach_maybe(x_init)
        self.verbose, self.eps, self.back_eps, self.n_batch = verbose, eps, back_eps, n_batch
        self.linesearch_decay, self.max_linesearch_iter = linesearch_decay, max_linesearch_iter
        self.exit_unconverged, self.detach_unconverged, self.backprop = exit_unconverged, detach_unconverged, backprop
        self.slew_rate_penalty, self.solver_type = slew_rate_penalty, solver_type
        self.add_goal_constraint, self.diag_cost, self.al_iter = add_goal_constraint, diag_cost, al_iter
        self.neq = n_state * (T - 1) + n_state
        self.nineq = n_ctrl * T * 2
        self.dyn_res_crit, self.dyn_res_factor = 1e-4, 10
        self.rho_prev = 1.0
        self.lamda_prev = torch.zeros(self.n_batch, self.neq + self.nineq).to(self.u_upper) if n_batch else None
        self.dyn_res_prev = 1000000
        self.just_initialized = True
        self._hist = None
        self._bound_cache = None
        self.status = None

    # ------------------------------------------------------------------------------------
    def forward(self, x0, cost, dx, dx_jac=None, u_init=None, x_init=None):
        """qpth/AL_mpc.py:198-252"""
        n_batch = self.n_batch if self.n_batch is not None else cost.C.size(0)
        assert cost.C.ndimension() == 4
        assert x0.ndimension() == 2 and x0.size(0) == n_batch
        spec = dyn_spec(dx)
        if dx_jac is not None and dyn_spec(dx_jac)[:2] != spec[:2]:
            raise RuntimeError("b200qp AL-MPC: dx and dx_jac describe different dynamics")
        if spec[2] != self.n_state or spec[3] != self.n_ctrl:
            raise RuntimeError("b200qp AL-MPC: dynamics sizes do not match n_state / n_ctrl")

        if u_init is not None:
            u = u_init
        elif self.u_init is None:
            u = torch.zeros(n_batch, self.T, self.n_ctrl).type_as(x0.data)
        else:
            u = self.u_init
        if u.ndimension() == 2:
            u = u.unsqueeze(0).expand(n_batch, self.T, -1).clone()
        u = u.type_as(x0.data)

        if x_init is not None:
            x = x_init
        elif self.x_init is None:
            x = self.rollout(x0, u, dx)
        else:
            x = self.x_init
        if x.ndimension() == 2:
            x = x.unsqueeze(0).expand(n_batch, self.T, -1).clone()
        x = x.type_as(x0.data)

        cost = QuadCost(cost.C.diagonal(dim1=-2, dim2=-1), cost.c)
        x, u = self.al_solve(x, u, dx, dx_jac, x0, cost, _spec=spec)
        self.x_init = x
        self.u_init = u
        return (x, u)

    def al_solve(self, x, u, dx, dx_jac, x0, cost, lamda_init=None, rho_init=None, _spec=None):
        """qpth/AL_mpc.py:254-321: one kernel launch for the whole AL loop."""
        spec = _spec if _spec is not None else dyn_spec(dx)
        B = x.shape[0]
        dev = x0.device
        lam = self.lamda_prev if lamda_init is None else lamda_init
        if lam is None:
            lam = torch.zeros(B, self.neq + self.nineq)
        rho = self.rho_prev if rho_init is None else rho_init
        if not torch.is_tensor(rho):
            rho = torch.full((B, 1), float(rho))
        state = ALState(lam.to(device=dev, dtype=self.dtype), rho.to(device=dev, dtype=self.dtype))
        state.hist = None if self.just_initialized else self._hist
        C = cost.C.to(self.dtype)
        c = cost.c.to(self.dtype)
        ul, uu = self._bounds(B, dev)
        # the reference detaches the initial trajectory (AL_mpc.py:284) and never differentiates x0
        xs, us, status = ALSolve.apply(C, c, x.detach().to(self.dtype), u.detach().to(self.dtype), x0.detach().to(self.dtype),
                                       ul, uu, state, spec, self.al_iter)
        self._hist = state.hist
        self.lamda_prev, self.rho_prev, self.status = state.lam, state.rho, status
        self.just_initialized = False
        return xs, us

    @property
    def cost_lam_hist(self):
        """[[cost_0..], [lam_0..], [rho_0..]] of the last call (qpth/AL_mpc.py:314), built on demand."""
        if self._hist is None:
            return None
        return [list(self._hist[0]), list(self._hist[1]), [r.unsqueeze(-1) for r in self._hist[2]]]

    def _bounds(self, B, dev):
        key = (B, str(dev))
        if self._bound_cache is None or self._bound_cache[0] != key:
            shape = (B, self.T, self.n_ctrl)
            f = lambda t: (t.to(dev).expand(shape) if t.dim() > 0 else t.to(dev).reshape(1, 1, 1).expand(shape)).contiguous()
            self._bound_cache = (key, f(self.u_lower), f(self.u_upper))
        return self._bound_cache[1], self._bound_cache[2]

    def rollout(self, x, actions, dynamics):
        """qpth/AL_mpc.py:398-411"""
        return _rollout(dyn_spec(dynamics), x, actions[:, :self.T])

    def reinitialize(self, x, mask):
        """qpth/AL_mpc.py:432-439"""
        self.u_init = None
        self.x_init = None
        self.rho_prev = torch.ones((self.n_batch, 1), device=x.device, dtype=x.dtype)
        self.lamda_prev = torch.zeros(self.n_batch, self.neq + self.nineq, device=x.device, dtype=x.dtype)
        self.dyn_res_prev = 1000000
        self.just_initialized = True
