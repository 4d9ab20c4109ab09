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


# ------------------------------------------------------------------------------------------------------------------
# DEQMPCPolicy: the network loop around Tracking_MPC (deqmpc/policies.py:190-529).  The network itself is a small torch
# MLP (plumbing: cuBLAS GEMMs + LayerNorm); what this package adds is that the whole chain
#     deq_iter x (DEQLayer -> Tracking_MPC (fused AL solve) ) -> loss -> backward (implicit MPC adjoints)
# is free of host synchronisation and can be captured ONCE in a CUDA graph (`GraphedTrainStep`): at the training shapes of
# deqmpc/run.sh the step is launch-bound (hundreds of tiny kernels around six fused solves).
class DEQLayer(torch.nn.Module):
    """deqmpc/policies.py:190-425, `layer_type == "mlp"` (the reference's "gcn" variant cannot be constructed: it reads
    `self.num_groups`, which nothing sets, policies.py:395-398) and `deq_out_type` 1 / 2 (the only ones `setup_output_layer`
    defines, policies.py:404-410).  Parameter names match the reference, so its checkpoints load unchanged."""

    def __init__(self, args, env):
        super().__init__()
        self.args = args
        self.nu, self.nx, self.nq, self.dt, self.T = env.nu, env.nx, args.nq, env.dt, args.T
        self.hdim, self.layer_type, self.out_type = args.hdim, args.layer_type, args.deq_out_type
        if self.layer_type != "mlp":
            raise NotImplementedError("b200qp DEQLayer: layer_type='mlp' only")
        if self.out_type not in (1, 2):
            raise NotImplementedError("b200qp DEQLayer: deq_out_type 1 or 2 (the reference defines no output layer for the others)")
        self.in_dim = self.nx + self.nx * (self.T - 1)
        self.inp_layer = torch.nn.Sequential(torch.nn.Linear(self.in_dim, self.hdim), torch.nn.LayerNorm(self.hdim))
        self.fcdeq1 = torch.nn.Linear(self.hdim, self.hdim)
        self.lndeq1 = torch.nn.LayerNorm(self.hdim)
        self.reludeq1 = torch.nn.ReLU()
        self.fcdeq2 = torch.nn.Linear(self.hdim, self.hdim)
        self.lndeq2 = torch.nn.LayerNorm(self.hdim)
        self.reludeq2 = torch.nn.ReLU()
        self.lndeq3 = torch.nn.LayerNorm(self.hdim)
        self.out_dim = self.nx * (self.T - 1) if self.out_type == 1 else self.nx + self.nx * (self.T - 1)
        self.out_layer = torch.nn.Sequential(torch.nn.Linear(self.hdim, self.out_dim))

    def init_z(self, bsz):
        return torch.zeros(bsz, self.hdim, dtype=torch.float32, device=self.args.device)

    def deq_layer(self, x, z):
        z = self.lndeq1(self.reludeq1(self.fcdeq1(z)))
        return self.lndeq3(self.reludeq2(z + self.lndeq2(x + self.fcdeq2(z))))

    def forward(self, x, z):
        z_out = self.deq_layer(self.inp_layer(x), z)
        steps = self.T - 1 if self.out_type == 1 else self.T
        dx_ref = self.out_layer(z_out).view(-1, steps, self.nx)
        vel_ref = dx_ref[..., self.nq:]
        dx_ref = dx_ref[..., :self.nq] * self.dt
        return torch.cat([dx_ref + x[:, None, :self.nq], vel_ref], dim=-1), z_out


class DEQMPCPolicy(torch.nn.Module):
    """deqmpc/policies.py:426-529: `deq_iter` rounds of (network proposes a reference, MPC tracks it, the MPC solution is
    fed back)."""

    def __init__(self, args, env):
        super().__init__()
        self.args = args
        self.nu, self.nx, self.nq, self.T, self.dt = env.nu, env.nx, args.nq, args.T, env.dt
        self.device = args.device
        self.deq_iter = args.deq_iter
        self.model = DEQLayer(args, env).to(self.device)
        self.out_type = args.policy_out_type
        self.tracking_mpc = Tracking_MPC(args, env)

    def forward(self, x, x_gt, u_gt, mask, iter=0, qp_solve=True, lastqp_solve=False):
        bsz = x.shape[0]
        x_ref = torch.cat([x] * self.T, dim=-1).detach().clone()
        nominal_actions = torch.zeros((bsz, self.T, self.nu), device=self.device)
        z = self.model.init_z(bsz)
        trajs = []
        if self.args.solver_type == "al":
            self.tracking_mpc.reinitialize(x, mask[:, :, None])
        for _ in range(self.deq_iter):
            x_ref, z = self.model(x_ref, z)
            if self.model.out_type == 1:
                x_ref = torch.cat([x[:, None, :], x_ref.view(-1, self.T - 1, self.nx)], dim=1)
            else:
                x_ref = x_ref.view(-1, self.T, self.nx)
            xu_ref = torch.cat([x_ref, nominal_actions], dim=-1)
            x_ref_tr, u_ref_tr = x_ref, nominal_actions
            nominal_states = x_ref
            if qp_solve:
                nominal_states, nominal_actions = self.tracking_mpc(x, xu_ref, x_ref_tr, u_ref_tr)
            trajs.append((x_ref, nominal_states, nominal_actions))
            x_ref = nominal_states.reshape(bsz, -1).detach().clone()
        dyn_res = self.tracking_mpc.dyn(x_ref.view(-1, self.nx).double(), u_gt.reshape(-1, self.nu).double()).view(bsz, -1).norm(dim=1).mean()
        if not (x.is_cuda and torch.cuda.is_current_stream_capturing()):
            dyn_res = dyn_res.item()   # the reference's host read; skipped while a CUDA graph is being captured
        if lastqp_solve:
            nominal_states, nominal_actions = self.tracking_mpc(x, xu_ref, x_ref_tr, u_ref_tr)
            trajs[-1] = (trajs[-1][0], nominal_states, nominal_actions)
        return trajs, dyn_res


def add_loss_based_on_out_type(policy, out_type, gt_states, gt_actions, gt_mask, nominal_states, nominal_actions):
    """deqmpc/policies.py:819-833"""
    loss = 0.0
    if out_type in (0, 2):
        loss = loss + torch.abs((nominal_actions - gt_actions) * gt_mask[:, :, None]).sum(dim=-1).mean()
    if out_type in (1, 2):
        loss = loss + torch.abs((nominal_states - gt_states) * gt_mask[:, :, None]).sum(dim=-1).mean()
    if out_type == 3:
        loss = loss + torch.abs((nominal_states[..., :policy.nq] - gt_states[..., :policy.nq]) * gt_mask[:, :, None]).sum(dim=-1).mean()
    return loss


def compute_loss(policy, gt_states, gt_actions, gt_mask, trajs, args):
    """deqmpc/policies.py:787-808,836-844 (the DEQ / DEQ-MPC branches: every DEQ iteration is supervised)"""
    out_type = policy.out_type if args.en_qp_solve else 1
    loss = 0.0
    for (_, nominal_states, nominal_actions) in trajs:
        loss = loss + add_loss_based_on_out_type(policy, out_type, gt_states, gt_actions, gt_mask, nominal_states, nominal_actions)
    loss_end = add_loss_based_on_out_type(policy, out_type, gt_states, gt_actions, gt_mask, nominal_states, nominal_actions)
    return loss, loss_end


class GraphedTrainStep:
    """One imitation-learning step of deqmpc/train.py:135-175 -- policy forward (deq_iter network passes and MPC solves), loss,
    backward through the implicit MPC adjoints -- captured ONCE in a CUDA graph and replayed, followed by ONE flat-bucket
    all-reduce of the DEQLayer gradients (the place train.py's data-parallel all-reduce belongs, between backward and
    optimizer.step) and the optimizer step.

        step = GraphedTrainStep(policy, optimizer, args, example_batch, process_group=None)
        loss = step(x0, gt_states, gt_actions, gt_mask)          # tensors are copied into the graph's static inputs
    """

    def __init__(self, policy, optimizer, args, example, process_group=None, use_graph=True):
        from . import dist as bdist
        self.policy, self.opt, self.args, self.group, self._dist = policy, optimizer, args, process_group, bdist
        self.static = [t.clone() for t in example]
        self.params = [p for p in policy.model.parameters()]
        self.graph = None
        self.loss = None
        self.use_graph = use_graph
        if use_graph:
            self._capture()

    def _fwd_bwd(self):
        x0, gs, ga, gm = self.static
        trajs, _ = self.policy(x0, gs, ga, gm, qp_solve=self.args.en_qp_solve)
        loss, _ = compute_loss(self.policy, gs, ga, gm, trajs, self.args)
        loss.backward()
        return loss.detach()

    def _capture(self):
        torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):  # warm-up: lazy initialisations, workspace allocations, .grad buffers
                self.opt.zero_grad(set_to_none=False)
                self._fwd_bwd()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        self.opt.zero_grad(set_to_none=False)
        with torch.cuda.graph(self.graph, stream=side):
            for p in self.params:
                if p.grad is not None:
                    p.grad.zero_()
            self.loss = self._fwd_bwd()

    def __call__(self, x0, gt_states, gt_actions, gt_mask):
        for dst, src in zip(self.static, (x0, gt_states, gt_actions, gt_mask)):
            dst.copy_(src)
        if self.graph is not None:
            self.graph.replay()
            loss = self.loss
        else:
            self.opt.zero_grad(set_to_none=False)
            loss = self._fwd_bwd()
        if self.group is not None:
            self._dist.allreduce_gradients(self.params, self.group)
        self.opt.step()
        return loss
