"""CPU oracle for the augmented-Lagrangian MPC path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

A torch-CPU restatement of the reference algorithm (swami1995/diff-qp-mpc):

    qpth/AL_mpc.py:254-321      MPC.al_solve (warm start, AL outer loop, lambda/rho updates)
    qpth/al_utils.py:16-34      warm_start_al
    qpth/al_utils.py:37-59      merit_function
    qpth/al_utils.py:62-102     merit_grad_hessian  (dense J(B,M,N), H = diag(C) + rho J_c^T J_c)
    qpth/al_utils.py:162-318    constraint residuals / Jacobians
    qpth/al_utils.py:363-500    NewtonAL forward (4 Newton steps, Cholesky) / implicit backward
    qpth/al_utils.py:503-527    line_search_newton (20-way parallel backtracking)
    deqmpc/envs.py:5-54,182-233 pendulum / integrator dynamics (semi-implicit Euler)

It keeps the reference's dense formulation (the dense (B,N,N) Hessian and torch.linalg.cholesky_ex)
so that it is an independent check of the block-tridiagonal CUDA kernels.  Dynamics Jacobians are
analytic here (the reference obtains the same numbers by tiling + autograd).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs
may import this file.  Parity pin: `oracle/gen_golden_mpc.py` runs the REAL reference
(qpth.AL_mpc.MPC with deqmpc.envs dynamics) in the build container, asserts this restatement
agrees with it, and commits the vectors under tests/golden/mpc_*.npz.
"""
from __future__ import annotations

import math

import torch

N_LS = 20          # al_utils.py:504
NEWTON_STEPS = 4   # al_utils.py:397


# ------------------------------------------------------------------------------------ dynamics
class Pendulum:
    """deqmpc/envs.py:5-48: theta from upright, semi-implicit Euler."""
    name, nx, nu = "pendulum", 2, 1

    def __init__(self, dt=0.05, g=10.0, m=1.0, l=1.0):
        self.dt, self.g, self.m, self.l = dt, g, m, l

    def params(self):
        return [self.dt, self.g, self.m, self.l]

    def step(self, x, u):
        th, thdot = x[..., 0], x[..., 1]
        acc = (u[..., 0] + self.m * self.g * self.l * torch.sin(th)) / (self.m * self.l ** 2)
        nthdot = thdot + acc * self.dt
        nth = th + nthdot * self.dt
        return torch.stack((nth, nthdot), dim=-1)

    def jac(self, x, u):
        th = x[..., 0]
        c = self.m * self.g * self.l * torch.cos(th) / (self.m * self.l ** 2)
        dt = self.dt
        one, zero = torch.ones_like(th), torch.zeros_like(th)
        A = torch.stack((torch.stack((one + c * dt * dt, dt * one), -1),
                         torch.stack((c * dt, one), -1)), -2)
        k = 1.0 / (self.m * self.l ** 2)
        B = torch.stack((k * dt * dt * one, k * dt * one), -1).unsqueeze(-1)
        return A, B


class Integrator:
    """deqmpc/envs.py:182-214 with nx = 2 nq, nu = nq."""
    name = "integrator"

    def __init__(self, nx=2, nu=1, dt=0.1):
        self.nx, self.nu, self.dt = nx, nu, dt
        self.nq = nx // 2
        assert nu == self.nq

    def params(self):
        return [self.dt]

    def step(self, x, u):
        pos, vel = x[..., :self.nq], x[..., self.nq:]
        vel_n = vel + u * self.dt
        pos_n = pos + vel_n * self.dt
        # the reference stacks (pos_n, vel_n) on a NEW last dim and reshapes: for nq = 1 that is
        # [pos, vel]; for nq > 1 it interleaves (envs.py:196-197)
        return torch.stack((pos_n, vel_n), dim=-1).reshape(x.shape)

    def jac(self, x, u):
        nq, dt = self.nq, self.dt
        A = torch.zeros(self.nx, self.nx, dtype=x.dtype)
        B = torch.zeros(self.nx, self.nu, dtype=x.dtype)
        for i in range(nq):
            # row index of pos_n[i] / vel_n[i] in the reshaped output
            rp, rv = 2 * i, 2 * i + 1
            A[rp, i] = 1.0; A[rp, nq + i] = dt; B[rp, i] = dt * dt
            A[rv, nq + i] = 1.0; B[rv, i] = dt
        shape = x.shape[:-1]
        return A.expand(*shape, self.nx, self.nx), B.expand(*shape, self.nx, self.nu)


# ------------------------------------------------------------------------------------ pieces
def cost_value(xu, C, c):
    """al_utils.py:338-349 (diag cost)."""
    return (0.5 * (xu * C * xu).sum(-1) + (c * xu).sum(-1)).sum(-1)


def residuals(xu, x0, dyn, u_lower, u_upper):
    """al_utils.py:321-335: res = [x_{t+1} - f(x_t,u_t) (t<T-1); x_0 - x0; u-ub, lb-u per t]."""
    nx = dyn.nx
    x, u = xu[..., :nx], xu[..., nx:]
    lead = xu.shape[:-2]
    xn = dyn.step(x[..., :-1, :], u[..., :-1, :])
    eq = torch.cat(((x[..., 1:, :] - xn).reshape(*lead, -1), x[..., 0, :] - x0), dim=-1)
    ineq = torch.cat((u - u_upper, -u + u_lower), dim=-1).reshape(*lead, -1)
    return torch.cat((eq, ineq), -1), torch.cat((eq, torch.clamp(ineq, min=0)), -1)


def merit_value(xu, C, c, x0, lam, rho, dyn, u_lower, u_upper):
    """al_utils.py:37-59."""
    res, resc = residuals(xu, x0, dyn, u_lower, u_upper)
    return cost_value(xu, C, c) + 0.5 * rho[..., 0] * (resc * resc).sum(-1) + (lam * res).sum(-1)


def constraint_jac(xu, x0, dyn, u_lower, u_upper):
    """Dense J (B,M,N) and its active-set masked copy (al_utils.py:162-185,212-262,294-318)."""
    B, T, nt = xu.shape
    nx, nu = dyn.nx, dyn.nu
    x, u = xu[..., :nx], xu[..., nx:]
    A, Bm = dyn.jac(x[:, :-1], u[:, :-1])
    M, N = T * nx + 2 * T * nu, T * nt
    J = torch.zeros(B, M, N, dtype=xu.dtype)
    for t in range(T - 1):
        r = slice(t * nx, (t + 1) * nx)
        J[:, r, t * nt:t * nt + nx] = -A[:, t]
        J[:, r, t * nt + nx:(t + 1) * nt] = -Bm[:, t]
        J[:, r, (t + 1) * nt:(t + 1) * nt + nx] = torch.eye(nx, dtype=xu.dtype)
    J[:, (T - 1) * nx:T * nx, :nx] = torch.eye(nx, dtype=xu.dtype)
    eye_u = torch.eye(nu, dtype=xu.dtype)
    for t in range(T):
        r0 = T * nx + t * 2 * nu
        J[:, r0:r0 + nu, t * nt + nx:(t + 1) * nt] = eye_u
        J[:, r0 + nu:r0 + 2 * nu, t * nt + nx:(t + 1) * nt] = -eye_u
    res, resc = residuals(xu, x0, dyn, u_lower, u_upper)
    Jc = J.clone()
    act = (resc[:, T * nx:] > 0).to(xu.dtype)
    Jc[:, T * nx:] = Jc[:, T * nx:] * act[..., None]
    return res, resc, J, Jc


def merit_grad_hess(xu, C, c, x0, lam, rho, dyn, u_lower, u_upper):
    """al_utils.py:62-102."""
    B = xu.shape[0]
    res, resc, J, Jc = constraint_jac(xu, x0, dyn, u_lower, u_upper)
    grad = (C * xu + c).reshape(B, -1) + (lam[..., None] * J).sum(-2) + rho * (resc[..., None] * Jc).sum(-2)
    H = torch.diag_embed(C.reshape(B, -1)) + rho[:, :, None] * torch.bmm(Jc.transpose(1, 2), Jc)
    return grad, H


def line_search(update, xu, merit_fn, merit, x0):
    """al_utils.py:503-527."""
    nx = x0.shape[-1]
    steps = 2.0 ** (-torch.arange(N_LS, dtype=xu.dtype))
    cand = xu[None] + steps[:, None, None, None] * update[None]
    cand[:, :, 0, :nx] = x0[None]
    vals = torch.stack([merit_fn(cand[k]) for k in range(N_LS)], 0)
    best, idx = torch.min(vals, dim=0)
    bi = torch.arange(xu.shape[0])
    xn = cand[idx, bi]
    status = (best < merit).to(xu.dtype)
    out = status[:, None, None] * xn + (1 - status)[:, None, None] * xu
    return out, best, status


def newton_al(xu, x0, lam, rho, C, c, dyn, u_lower, u_upper):
    """NewtonAL.forward (al_utils.py:363-460): always 4 Newton steps; returns what backward needs."""
    B, T, nt = xu.shape
    mf = lambda z: merit_value(z, C, c, x0, lam, rho, dyn, u_lower, u_upper)
    merit = mf(xu)
    H = U = status = None
    for _ in range(NEWTON_STEPS):
        grad, H = merit_grad_hess(xu, C, c, x0, lam, rho, dyn, u_lower, u_upper)
        U, _info = torch.linalg.cholesky_ex(H)
        update = -torch.cholesky_solve(grad.reshape(B, -1, 1), U).reshape(B, T, nt)
        xu, merit, status = line_search(update, xu, mf, merit, x0)
    return xu, status, H, U


def warm_start(lam, rho, cost_start, cost_hist, lam_hist, rho_hist):
    """al_utils.py:16-34; histories are stacked newest-first."""
    idx = torch.max(cost_hist < cost_start[None], dim=0)[1]
    bi = torch.arange(lam.shape[0])
    lh = lam_hist[idx, bi]
    lam = lam * (lh.norm(p=2, dim=-1) / lam.norm(p=2, dim=-1)).unsqueeze(-1)
    return lam, rho_hist[idx, bi]


class ALState:
    """The solver state the reference keeps on the module between calls (AL_mpc.py:432-439)."""

    def __init__(self, B, M, dtype=torch.float64):
        self.lam = torch.zeros(B, M, dtype=dtype)
        self.rho = torch.ones(B, 1, dtype=dtype)
        self.hist = None  # (cost (K,B), lam (K,B,M), rho (K,B,1)) oldest-first


def al_solve(x, u, x0, C, c, dyn, u_lower, u_upper, state: ALState, al_iter=2):
    """MPC.al_solve (AL_mpc.py:254-321).  Returns fp64 (x,u), the backward context of the LAST
    NewtonAL call, and updates `state` in place."""
    nx = dyn.nx
    xu = torch.cat((x, u), dim=2)
    neq = x.shape[1] * nx
    lam, rho = state.lam, state.rho
    cost_start = cost_value(xu, C, c)
    if state.hist is not None:
        ch, lh, rh = (h.flip(0) for h in state.hist)
        lam, rho = warm_start(lam, rho, cost_start, ch, lh, rh)
    hist = [[cost_start], [lam], [rho]]
    ctx = None
    for _ in range(al_iter):
        xu, status, H, U = newton_al(xu.clone(), x0, lam, rho, C, c, dyn, u_lower, u_upper)
        ctx = dict(H=H, U=U, xu=xu, status=status)
        res, _ = residuals(xu, x0, dyn, u_lower, u_upper)
        lam = lam + rho * res
        lam = torch.cat((lam[:, :neq], torch.clamp(lam[:, neq:], min=0)), dim=1)
        hist[0].append(cost_value(xu, C, c))
        rho = rho * 10
        hist[1].append(lam)
        hist[2].append(rho)
    state.lam, state.rho = lam, rho
    state.hist = tuple(torch.stack(h, 0) for h in hist)
    return xu[..., :nx], xu[..., nx:], ctx


def al_backward(ctx, grad_xu):
    """NewtonAL.backward (al_utils.py:462-500): dC = (-H^-1 g) * x_est, dc = -H^-1 g."""
    B = grad_xu.shape[0]
    ig = -torch.cholesky_solve(grad_xu.reshape(B, -1, 1), ctx["U"]).reshape(grad_xu.shape)
    return ig * ctx["xu"], ig


def rollout(x0, u, dyn):
    """MPC.rollout (AL_mpc.py:398-411)."""
    xs = [x0]
    for t in range(u.shape[1] - 1):
        xs.append(dyn.step(xs[t], u[:, t]))
    return torch.stack(xs, 1)
