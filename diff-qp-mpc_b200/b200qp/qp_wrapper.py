"""qp_wrapper.MPC -- the interior-point ("ip") MPC of the reference (qpth/qp_wrapper.py:58-715) on top of the fused
kernels: an SQP loop whose QP sub-problem over the whole horizon,

    min_tau  sum_t 1/2 tau_t' C_t tau_t + c_t' tau_t
    s.t.     x_{t+1} = F_t tau_t + f_t  (linearised dynamics),  x_0 = x_init,  u_lower <= u_t <= u_upper,

is solved by `DenseQPFunction` with the NON-linear dynamics residual as its `dyn_res` callback
(qp_wrapper.py:303-316), followed by a backtracking line search on the rolled-out cost.

What runs where: the linearisation (`dx_jac` on all (T-1) B knots at once), the residual callback (`dx` on all knots) and
the roll-outs of the line search are the fused dynamics kernels when `dx` / `dx_jac` are b200qp.envs / b200qp.my_envs
modules (any callable with the reference's interface works); the QP is b200qp's DenseQPFunction (CUDA, callbacks
between launches); the block-sparse matrices Q, A, G are assembled on the device with cached index tensors.  Same
constructor and call surface as the reference:

    ctrl = MPC(n_state, n_ctrl, T, u_lower=..., u_upper=..., qp_iter=..., n_batch=B, u_init=(T,B,nu), single_qp_solve=...)
    x, u = ctrl(x0, QuadCost(C (T,B,n_tau,n_tau), c (T,B,n_tau)), dx, dx_jac)          # (T,B,nx), (T,B,nu)

Differences from the reference, all on paths it cannot execute itself: `GradMethods.AUTO_DIFF / FINITE_DIFF`
linearisation and `slew_rate_penalty` are not implemented (the reference asserts / exits on them, qp_wrapper.py:444-450,
:527); tensors handed in need not be contiguous (the reference's `.view` at :487 fails on the transposed `u_init`
that deqmpc's Tracking_MPC passes, policies.py:637)."""
from __future__ import annotations

from collections import namedtuple
from enum import Enum

import torch
from torch.nn import Module

from . import qp

QuadCost = namedtuple("QuadCost", "C c")
LinDx = namedtuple("LinDx", "F f")
QuadCost.__new__.__defaults__ = (None,) * len(QuadCost._fields)
LinDx.__new__.__defaults__ = (None,) * len(LinDx._fields)


class GradMethods(Enum):
    AUTO_DIFF = 1
    FINITE_DIFF = 2
    ANALYTIC = 3
    ANALYTIC_CHECK = 4


def _detach(t):
    return t.detach() if torch.is_tensor(t) else t


class MPC(Module):
    def __init__(self, n_state, n_ctrl, T, u_lower=None, u_upper=None, u_zero_I=None, u_init=None, x_init=None, qp_iter=10,
                 grad_method=GradMethods.ANALYTIC, delta_u=None, verbose=0, eps=1e-7, back_eps=1e-7, n_batch=None,
                 linesearch_decay=0.2, max_linesearch_iter=10, exit_unconverged=True, detach_unconverged=True, backprop=True,
                 slew_rate_penalty=None, prev_ctrl=None, not_improved_lim=5, best_cost_eps=1e-4, solver_type='dense',
                 single_qp_solve=False, add_goal_constraint=False, x_goal=None):
        super().__init__()
        assert (u_lower is None) == (u_upper is None)
        assert max_linesearch_iter > 0
        if grad_method != GradMethods.ANALYTIC:
            raise NotImplementedError("b200qp qp_wrapper.MPC: only GradMethods.ANALYTIC (dx_jac) linearisation")
        if slew_rate_penalty is not None:
            raise NotImplementedError("b200qp qp_wrapper.MPC: slew_rate_penalty (the reference exits on it too)")
        if solver_type != 'dense':
            raise NotImplementedError("b200qp qp_wrapper.MPC: solver_type='dense' is the only one the reference implements")
        self.n_state, self.n_ctrl, self.T = n_state, n_ctrl, T
        self.u_lower, self.u_upper = _detach(u_lower), _detach(u_upper)
        self.x_goal = x_goal
        self.u_zero_I, self.u_init, self.x_init = _detach(u_zero_I), _detach(u_init), _detach(x_init)
        self.qp_iter, self.grad_method, self.delta_u, self.verbose = qp_iter, grad_method, delta_u, verbose
        self.eps, self.back_eps, self.n_batch = eps, back_eps, n_batch
        self.linesearch_decay, self.max_linesearch_iter = linesearch_decay, max_linesearch_iter
        self.exit_unconverged, self.detach_unconverged, self.backprop = exit_unconverged, detach_unconverged, backprop
        self.not_improved_lim, self.best_cost_eps = not_improved_lim, best_cost_eps
        self.slew_rate_penalty, self.prev_ctrl = slew_rate_penalty, prev_ctrl
        self.solver_type, self.single_qp_solve, self.add_goal_constraint = solver_type, single_qp_solve, add_goal_constraint
        self._idx = {}       # per-device index tensors of the block structure (qp_wrapper.py:184-210)
        self.info = {}       # iteration counts of the QP solves of the last call

    # ---------------------------------------------------------------------------------------------- structure
    def _indices(self, device):
        key = str(device)
        if key not in self._idx:
            nx, nu, T = self.n_state, self.n_ctrl, self.T
            nt = nx + nu
            ar = lambda *a: torch.arange(*a, device=device)
            r, c = torch.meshgrid(ar(nx), ar(nt), indexing='ij')                       # dynamics rows of A: [F_t] at knot t
            steps = ar(T - 1).view(-1, 1, 1)
            A_r = (r.unsqueeze(0) + nx * steps).reshape(-1)
            A_c = (c.unsqueeze(0) + nt * steps).reshape(-1)
            X_r = (ar(nx).unsqueeze(0) + nx * ar(T - 1).unsqueeze(1)).reshape(-1)       # -I on x_{t+1}
            X_c = (ar(nx).unsqueeze(0) + nt * (ar(T - 1).unsqueeze(1) + 1)).reshape(-1)
            r, c = torch.meshgrid(ar(nt), ar(nt), indexing='ij')
            stepsT = ar(T).view(-1, 1, 1)
            Q_r = (r.unsqueeze(0) + nt * stepsT).reshape(-1)
            Q_c = (c.unsqueeze(0) + nt * stepsT).reshape(-1)
            G_r = (ar(nu).unsqueeze(0) + nu * ar(T).unsqueeze(1)).reshape(-1)
            G_c = (ar(nu).unsqueeze(0) + nx + nt * ar(T).unsqueeze(1)).reshape(-1)
            self._idx[key] = (A_r, A_c, X_r, X_c, Q_r, Q_c, G_r, G_c)
        return self._idx[key]

    def compute_Ab_dense(self, F, f, x0):
        """qp_wrapper.py:639-656 (F: (T-1, B, nx, n_tau), f: (T-1, B, nx))"""
        Tm1, B, nx, nt = F.shape
        T = Tm1 + 1
        rows = (T + 1) * nx if self.add_goal_constraint else T * nx
        A = torch.zeros(B, rows, T * nt, dtype=F.dtype, device=F.device)
        b = torch.zeros(B, rows, dtype=F.dtype, device=F.device)
        A_r, A_c, X_r, X_c = self._indices(F.device)[:4]
        A[:, A_r, A_c] = F.transpose(0, 1).reshape(B, -1)
        A[:, X_r, X_c] = -1
        eye = torch.eye(nx, dtype=F.dtype, device=F.device)
        A[:, Tm1 * nx:T * nx, :nx] += eye
        if self.add_goal_constraint:
            A[:, T * nx:, -nt:-(nt - nx)] += eye
        b[:, :Tm1 * nx] = -f.transpose(0, 1).reshape(B, -1)
        b[:, Tm1 * nx:T * nx] = x0
        return A, b

    def compute_Qq_dense(self, C, c):
        """qp_wrapper.py:658-663"""
        T, B, nt, _ = C.shape
        Q = torch.zeros(B, T * nt, T * nt, dtype=C.dtype, device=C.device)
        Q_r, Q_c = self._indices(C.device)[4:6]
        Q[:, Q_r, Q_c] = C.transpose(0, 1).reshape(B, -1)
        q = c.transpose(0, 1).reshape(B, -1)
        return Q, q

    def compute_Gh_dense(self, x0):
        """qp_wrapper.py:665-680"""
        T, B, nx, nu = self.T, self.n_batch, self.n_state, self.n_ctrl
        nt = nx + nu
        opt = dict(dtype=x0.dtype, device=x0.device)
        if self.u_upper is None:
            G = torch.zeros(B, nu, T * nt, **opt)
            h = torch.ones(B, nu, **opt)
            G[:, torch.arange(nu), torch.arange(nu) + (T - 1) * nt + nx] = 1
            return G, h
        G_r, G_c = self._indices(x0.device)[6:8]
        G = torch.zeros(B, 2 * T * nu, T * nt, **opt)
        h = torch.ones(B, 2 * T * nu, **opt)
        G[:, G_r, G_c] = 1.0
        G[:, G_r + T * nu, G_c] = -1.0
        uu, ul = torch.as_tensor(self.u_upper, **opt).reshape(-1), torch.as_tensor(self.u_lower, **opt).reshape(-1)
        h[:, :T * nu] *= uu.repeat(T)[None]
        h[:, T * nu:] *= (-ul).repeat(T)[None]
        return G, h

    def compute_cost(self, xu, cost):
        """qp_wrapper.py:691-694 (xu: (B, T, n_tau))"""
        C, c = cost.C.transpose(0, 1), cost.c.transpose(0, 1)
        return 0.5 * ((xu.unsqueeze(-1) * C).sum(dim=-2) * xu).sum(dim=-1).sum(dim=-1) + (xu * c).sum(dim=-1).sum(dim=-1)

    # ---------------------------------------------------------------------------------------------- dynamics
    def rollout(self, x, actions, dynamics):
        """qp_wrapper.py:604-617 (actions: (T, B, nu) -> states (T, B, nx))"""
        xs = [x]
        for t in range(self.T - 1):
            if isinstance(dynamics, LinDx):
                xs.append(torch.bmm(dynamics.F[t], torch.cat([xs[t], actions[t]], dim=-1).unsqueeze(-1)).squeeze(-1) + dynamics.f[t])
            else:
                xs.append(dynamics(xs[t], actions[t]))
        return torch.stack(xs, 0)

    def linearize_dynamics(self, x, u, dynamics, dx_jac, diff):
        """qp_wrapper.py:482-512: all (T-1) B knots in one call of dx / dx_jac"""
        B = x.shape[1]
        _x = x[:-1].reshape(-1, self.n_state)
        _u = u[:-1].reshape(-1, self.n_ctrl)
        if not diff:
            _x, _u = _x.detach(), _u.detach()
        new_x = dynamics(_x, _u)
        if not diff:
            new_x = new_x.detach()
        R, S = dx_jac(_x, _u)[1]
        f = new_x - torch.bmm(R, _x.unsqueeze(-1)).squeeze(-1) - torch.bmm(S, _u.unsqueeze(-1)).squeeze(-1)
        f = f.view(self.T - 1, B, self.n_state)
        F = torch.cat((R.reshape(self.T - 1, B, self.n_state, self.n_state), S.reshape(self.T - 1, B, self.n_state, self.n_ctrl)), 3)
        return F, f

    def dyn_res(self, z, dx, x0):
        """qp_wrapper.py:328-349: the equality residual of the QP evaluated with the TRUE (non-linear) dynamics, in the row
        order of `compute_Ab_dense`: [x_{t+1} - f(x_t, u_t) for t < T-1 (sign: f - x_{t+1}), x_0 - x0, (x_{T-1} - x_goal)]"""
        B, T, nx, nu = self.n_batch, self.T, self.n_state, self.n_ctrl
        tau = z.reshape(B, T, nx + nu)
        xs, us = tau[..., :nx], tau[..., nx:]
        if isinstance(dx, LinDx):
            F = dx.F.permute(1, 0, 2, 3)                                   # (B, T-1, nx, n_tau)
            pred = (F * tau[:, :-1, None, :]).sum(dim=-1) + dx.f.permute(1, 0, 2)
        else:
            pred = dx(xs.reshape(-1, nx), us.reshape(-1, nu)).reshape(B, T, nx)[:, :-1]
        rows = [(pred - xs[:, 1:]).reshape(B, -1), (xs[:, 0] - x0).reshape(B, -1)]
        if self.add_goal_constraint:
            rows.append((xs[:, -1] - self.x_goal).reshape(B, -1))
        return torch.cat(rows, dim=1)

    # ---------------------------------------------------------------------------------------------- solve
    def _time_batch(self, t, full_ndim, n_batch):
        """broadcast a cost term given without its time and / or batch dimension to (T, B, ...) (qp_wrapper.py:236-254)"""
        missing = full_ndim - t.ndimension()
        if missing == 2:
            t = t[None, None]
        elif missing == 1:
            t = t[:, None]
        elif missing != 0:
            raise RuntimeError('MPC Error: Unexpected QuadCost shape.')
        return t.expand(self.T, n_batch, *t.shape[2:])

    def forward(self, x0, cost, dx, dx_jac, dx_true=None):
        self.dx_true = dx if dx_true is None else dx_true
        if not (isinstance(cost, tuple) and len(cost) == 2):
            raise NotImplementedError("b200qp qp_wrapper.MPC: QuadCost(C, c) costs only")
        C, c = cost
        if self.n_batch is None:
            if C.ndimension() != 4:
                raise RuntimeError('MPC Error: Could not infer batch size, pass in as n_batch')
            self.n_batch = C.size(1)
        B = self.n_batch
        cost = QuadCost(self._time_batch(C, 4, B), self._time_batch(c, 3, B))
        assert x0.ndimension() == 2 and x0.size(0) == B
        opt = dict(dtype=x0.dtype, device=x0.device)
        # initial controls (T, B, nu) and states (rolled out from them unless given)
        u = torch.zeros(self.T, B, self.n_ctrl, **opt) if self.u_init is None else self.u_init
        if u.ndimension() == 2:
            u = u[:, None].expand(self.T, B, -1)
        u = u.to(**opt).contiguous()
        x = self.rollout(x0, u, dx) if self.x_init is None else self.x_init
        if x.ndimension() == 2:
            x = x[:, None].expand(self.T, B, -1)
        x = x.to(**opt).contiguous()
        self.info = {"qp_iters": []}
        step = self.single_qp_ls if self.single_qp_solve else self.solve_nonlin
        x, u, _ = step(x, u, dx, dx_jac, x0, cost)
        return (x, u)

    def single_qp(self, x, u, dx, dx_jac, x0, cost):
        """qp_wrapper.py:298-326"""
        if isinstance(dx, LinDx):
            F, f = dx.F, dx.f
            if f is None:
                f = torch.zeros((self.T - 1, self.n_batch, self.n_state), dtype=x0.dtype, device=x0.device)
        else:
            F, f = self.linearize_dynamics(x, _detach(u), dx, dx_jac, diff=False)
        dyn_res_lam = lambda z: self.dyn_res(z, self.dx_true, x0)
        Q, q = self.compute_Qq_dense(cost.C, cost.c)
        A, b = self.compute_Ab_dense(F, f, x0)
        G, h = self.compute_Gh_dense(x0)
        fn = qp.DenseQPFunction()
        xhats = fn(Q, q, G, h, A, b, dyn_res_lam)
        self.info["qp_iters"].append(fn.info.get("n_iter"))
        xhats = xhats.reshape(self.n_batch, self.T, -1)
        x_hat = xhats[:, :, :self.n_state].transpose(0, 1)
        u_hat = xhats[:, :, self.n_state:].transpose(0, 1)
        cost_total = self.compute_cost(xhats, cost)
        return x_hat - x, u_hat - u, cost_total

    def line_search(self, x, u, delta_x, delta_u, dx, x0, cost):
        """qp_wrapper.py:418-436: backtracking on the ROLLED-OUT cost; every problem keeps its own alpha, the loop stops when
        the whole batch improved"""
        alpha = torch.ones(1, self.n_batch, 1, dtype=x0.dtype, device=x0.device)
        cost_total = self.compute_cost(torch.cat((x, u), dim=2).transpose(0, 1), cost)
        for _ in range(self.max_linesearch_iter):
            u_new = u + delta_u * alpha
            x_new = self.rollout(x0, u_new, dx)
            cost_total_new = self.compute_cost(torch.cat((x_new, u_new), dim=2).transpose(0, 1), cost)
            if bool((cost_total_new < cost_total).all()):
                break
            mask = (cost_total_new >= cost_total).to(x0.dtype)[None, :, None]
            alpha = alpha * self.linesearch_decay * mask + (1 - mask) * alpha
        return x_new, u_new, alpha, cost_total_new

    def single_qp_ls(self, x, u, dx, dx_jac, x0, cost):
        """qp_wrapper.py:404-415"""
        delta_x, delta_u, _ = self.single_qp(x, u, dx, dx_jac, x0, cost)
        with torch.no_grad():
            _, _, alpha, cost_total = self.line_search(x, u, delta_x, delta_u, dx, x0, cost)
        self.info["alpha"] = alpha.reshape(-1)
        return x + delta_x * alpha, u + delta_u * alpha, cost_total

    def solve_nonlin(self, x, u, dx, dx_jac, x0, cost):
        """qp_wrapper.py:352-402: qp_iter SQP steps without autograd, the per-problem best iterate, then one differentiable QP
        at the best iterate.  (The reference's stall counter is reset but never incremented, :383, so only the step-norm
        test ends the loop early; kept.)"""
        best_x = best_u = best_cost = None
        with torch.no_grad():
            for _ in range(self.qp_iter):
                u_prev = u.clone()
                delta_x, delta_u, _ = self.single_qp(x, u, dx, dx_jac, x0, cost)
                x, u, alpha, cost_total = self.line_search(x, u, delta_x, delta_u, dx, x0, cost)
                full_du_norm = (u - u_prev).norm()
                if best_cost is None:
                    best_x, best_u, best_cost = x.clone(), u.clone(), cost_total.clone()
                else:
                    better = cost_total <= best_cost + self.best_cost_eps
                    sel = better[None, :, None]
                    best_x, best_u = torch.where(sel, x, best_x), torch.where(sel, u, best_u)
                    best_cost = torch.where(better, cost_total, best_cost)
                if float(full_du_norm) < self.eps:
                    break
        x, u = best_x, best_u
        delta_x, delta_u, _ = self.single_qp(x, u, dx, dx_jac, x0, cost)
        with torch.no_grad():
            _, _, alpha, cost_total = self.line_search(x, u, delta_x, delta_u, dx, x0, cost)
        # the outputs -- and hence every gradient -- scale with this alpha; at a converged iterate (delta ~ 0) the
        # acceptance test of the line search is decided by rounding noise, in the reference as here
        self.info["alpha"] = alpha.reshape(-1)
        return x + delta_x * alpha, u + delta_u * alpha, cost_total
