"""Tracking_MPC -- the production caller of the MPC solve (deqmpc/policies.py:560-690), on top of the
fused AL-MPC kernels.  Same constructor (`args`, `env`) and call surface:

    mpc = Tracking_MPC(args, env)
    mpc.reinitialize(x, mask)
    x_nom, u_nom = mpc(x0, xu_ref, x_ref, u_ref)       # (bsz,T,nx) (bsz,T,nu) float32

`args.solver_type == "al"` -> b200qp.AL_mpc.MPC (fused augmented-Lagrangian kernels); anything else ("ip") ->
b200qp.qp_wrapper.MPC (SQP on DenseQPFunction, deqmpc/policies.py:622-639).  `env.dynamics` /
`env.dynamics_derivatives` may be the reference's own (jit-scripted) modules or b200qp.envs / b200qp.my_envs
classes; they select the fused dynamics by class name.  (The reference's own "ip" branch raises in
qp_wrapper.py:487 -- `.view` on the transposed `u_init` it passes; here the tensor is made contiguous.)
"""
from __future__ import annotations

import torch

from . import AL_mpc as al_mpc
from . import al_utils
from . import qp_wrapper as ip_mpc


class Tracking_MPC(torch.nn.Module):
    def __init__(self, args, env):
        super().__init__()
        self.args = args
        self.nu, self.nx, self.nq, self.dt = env.nu, env.nx, getattr(env, "nq", env.nx // 2), env.dt
        self.T = args.T
        self.dyn, self.dyn_jac = env.dynamics, env.dynamics_derivatives
        self.device = args.device
        self.u_upper = torch.tensor(env.action_space.high).to(self.device)
        self.u_lower = torch.tensor(env.action_space.low).to(self.device)
        self.qp_iter, self.eps, self.warm_start, self.bsz = args.qp_iter, args.eps, args.warm_start, args.bsz
        self.dtype = torch.float64 if args.dtype == "double" else torch.float32
        if args.Q is None:
            Q = torch.ones(self.nx, dtype=self.dtype, device=self.device)
            R = torch.ones(self.nu, dtype=self.dtype, device=self.device)
        else:
            Q, R = args.Q.to(self.device), args.R.to(self.device)
        qr = torch.cat([Q, R], dim=0).to(self.dtype)
        self.Qdiag = qr.repeat(self.bsz, self.T, 1)                       # what the kernels consume
        self.Q = torch.diag(qr).repeat(self.bsz, self.T, 1, 1)            # the reference's attribute
        self.u_init = torch.randn(self.bsz, self.T, self.nu, dtype=self.dtype, device=self.device)
        self.x_init = None
        self.single_qp_solve = self.qp_iter == 1
        if args.solver_type == "al":
            self.ctrl = al_mpc.MPC(self.nx, self.nu, self.T, u_lower=self.u_lower, u_upper=self.u_upper,
                                   exit_unconverged=False, eps=1e-5, n_batch=self.bsz, backprop=False, verbose=0,
                                   u_init=self.u_init, solver_type="dense", dtype=self.dtype)
        else:  # deqmpc/policies.py:622-639
            self.ctrl = ip_mpc.MPC(self.nx, self.nu, self.T, u_lower=self.u_lower.to(self.dtype), u_upper=self.u_upper.to(self.dtype),
                                   qp_iter=self.qp_iter, exit_unconverged=False, eps=1e-5, n_batch=self.bsz, backprop=False,
                                   verbose=0, u_init=self.u_init.transpose(0, 1).contiguous(),
                                   grad_method=ip_mpc.GradMethods.ANALYTIC, solver_type="dense",
                                   single_qp_solve=self.single_qp_solve)

    def forward(self, x0, xu_ref, x_ref, u_ref):
        """deqmpc/policies.py:641-664"""
        if self.args.solver_type == "al":
            xu_ref = torch.cat([x_ref, u_ref], dim=-1)
            if self.x_init is None:
                self.x_init = self.ctrl.x_init = x_ref
                self.u_init = self.ctrl.u_init = u_ref
            self.compute_p(xu_ref)
            cost = al_utils.QuadCost(self.Q, self.p)
            nominal_states, nominal_actions = self.ctrl(x0, cost, self.dyn, self.dyn_jac)
        else:
            self.compute_p(xu_ref)
            cost = ip_mpc.QuadCost(self.Q.transpose(0, 1), self.p.transpose(0, 1))
            self.ctrl.u_init = self.u_init.transpose(0, 1).contiguous()
            nominal_states, nominal_actions = self.ctrl(x0, cost, self.dyn, self.dyn_jac)
            nominal_states, nominal_actions = nominal_states.transpose(0, 1), nominal_actions.transpose(0, 1)
        self.u_init = nominal_actions.clone().detach()
        return nominal_states, nominal_actions

    def compute_p(self, x_ref):
        """p = -Q x_ref with Q diagonal (deqmpc/policies.py:666-679)"""
        self.p = -(self.Qdiag * x_ref)
        return self.p

    def reinitialize(self, x, mask):
        """deqmpc/policies.py:681-686"""
        self.u_init = torch.randn(self.bsz, self.T, self.nu, dtype=x.dtype, device=x.device)
        self.x_init = None
        if hasattr(self.ctrl, "reinitialize"):  # qp_wrapper.MPC keeps no solver state (the reference calls it regardless and raises)
            self.ctrl.reinitialize(x, mask)
