"""Generate tests/golden/ipmpc_*.npz from the REAL reference's interior-point MPC, qpth.qp_wrapper.MPC
(qp_wrapper.py:58-715: SQP loop, DenseQPFunction with the non-linear dynamics residual as dyn_res, line search), on the
reference's deqmpc/my_envs dynamics (its CasADi-generated C compiled into oracle/_ref/, analytic derivatives).
The jit-scripted envs of deqmpc/envs.py cannot be used here: their Jacobian modules call autograd.grad on inputs the
wrapper has detached (qp_wrapper.py:497-501) and raise.  Build container only.  TEST INFRASTRUCTURE, NOT PRODUCT."""
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore")

import gen_golden_myenvs as GM  # noqa: E402  (puts the reference on sys.path, builds oracle/_ref)

# name -> (model, B, T, qp_iter)
CASES = {
    "ipmpc_pendulum1l_B8_T5_single": ("pendulum1l", 8, 5, 1),
    "ipmpc_pendulum1l_B8_T5_sqp3": ("pendulum1l", 8, 5, 3),
    "ipmpc_cartpole1l_B4_T10_single": ("cartpole1l", 4, 10, 1),
    "ipmpc_cartpole1l_B4_T10_sqp3": ("cartpole1l", 4, 10, 3),
    # BASELINE configs[2]'s horizon: the QP has nz = 100, neq = 80, nineq = 40 (KKT order 260); appended last so that the
    # earlier cases keep their random inputs
    "ipmpc_cartpole1l_B4_T20_single": ("cartpole1l", 4, 20, 1),
}


def main():
    import subprocess
    subprocess.check_call(["make", "-C", HERE])
    from qpth import qp_wrapper as ip_mpc
    gold = os.path.join(ROOT, "tests", "golden")
    rs = np.random.RandomState(7)
    for case, (name, B, T, qp_iter) in CASES.items():
        d = GM.reference_dynamics(name)
        nx, nu = 2 * GM.MO.NQ[name], 1
        um = GM.UMAX[name]
        x0 = torch.tensor(np.concatenate([rs.uniform(-1.0, 1.0, (B, nx // 2)), rs.uniform(-0.5, 0.5, (B, nx // 2))], 1))
        u_init = torch.tensor(0.1 * rs.randn(T, B, nu))
        Cd = np.array([1.0] * nx + [1e-2] * nu)
        c0 = torch.tensor(0.1 * rs.randn(T, B, nx + nu))

        def run(pert):
            ctrl = ip_mpc.MPC(nx, nu, T, u_lower=-um * torch.ones(nu, dtype=torch.float64), u_upper=um * torch.ones(nu, dtype=torch.float64),
                              qp_iter=qp_iter, exit_unconverged=False, eps=1e-5, n_batch=B, backprop=False, verbose=0,
                              u_init=u_init.clone() + pert, grad_method=ip_mpc.GradMethods.ANALYTIC, solver_type="dense",
                              single_qp_solve=(qp_iter == 1))
            C = torch.diag(torch.tensor(Cd)).repeat(T, B, 1, 1).requires_grad_(True)
            c = c0.clone().requires_grad_(True)
            alphas = []
            ls = ctrl.line_search
            ctrl.line_search = lambda *a: (lambda r: (alphas.append(r[2].clone()), r)[1])(ls(*a))
            xs, us = ctrl(x0, ip_mpc.QuadCost(C, c), d, d.dynamics_derivatives)
            (xs.sum() + 2 * us.sum()).backward()
            return xs.detach(), us.detach(), C.grad.clone(), c.grad.clone(), alphas[-1].reshape(B)

        xs, us, dC, dc, alpha = run(0.0)
        # Conditioning of the reference's OWN gradient, measured: the SQP loop takes discrete decisions (line-search
        # acceptance, best-iterate selection, which PDIPM iterate is "best"), and the final QP is solved at a point where
        # constraints can be weakly active.  A problem whose reference gradient moves by more than 1e-7 when u_init moves
        # by 1e-10 is a knife-edge: its gradient is not a property of the inputs, and the parity test skips it (x, u of
        # every problem are compared regardless).
        _, _, dC2, dc2, _ = run(1e-10)
        mv = lambda a, b: (a - b).transpose(0, 1).reshape(B, -1).norm(dim=1) / (b.transpose(0, 1).reshape(B, -1).norm(dim=1) + 1e-300)
        stable = (mv(dC2, dC) < 1e-7) & (mv(dc2, dc) < 1e-7)
        np.savez_compressed(os.path.join(gold, f"{case}.npz"), x0=x0.numpy(), u_init=u_init.numpy(), Cd=Cd, c=c0.numpy(),
                            dt=GM.DT[name], umax=um, out_x=xs.numpy(), out_u=us.numpy(), dC=dC.numpy(), dc=dc.numpy(),
                            grad_stable=stable.numpy(), alpha=alpha.numpy())
        # The outputs are x + alpha (x_hat - x) with alpha the step of the LAST line search (qp_wrapper.py:399-401,
        # :418-436), so every gradient is alpha_j times the adjoint of the last QP.  In SQP mode that line search runs at
        # a converged point (delta ~ 0): whether `cost_new < cost` holds is rounding noise, and alpha_j ends up an
        # arbitrary power of linesearch_decay (measured on the CUDA path: same gradient direction to 1e-15, factors
        # 0.2^k).  The golden therefore records alpha and the parity test compares gradient / alpha.
        print(f"  last line-search alpha per problem: {alpha.numpy()}")
        C = types.SimpleNamespace(grad=dC); c = types.SimpleNamespace(grad=dc)
        print(f"  gradient-stable problems: {int(stable.sum())} of {B}")
        print(f"{case}: |x| {float(xs.norm()):.6f} |u| {float(us.norm()):.6f} |dC| {float(C.grad.norm()):.4f} |dc| {float(c.grad.norm()):.4f}",
              flush=True)


if __name__ == "__main__":
    main()
