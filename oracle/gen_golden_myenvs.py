"""Generate tests/golden/dyn_myenvs_<model>.npz and tests/golden/mpc_<model>_*.npz from the REAL reference:
the reference's CasADi-generated C (compiled by oracle/Makefile into oracle/_ref/) driven through the
reference's own deqmpc/my_envs/dynamics.py `Dynamics` module and qpth.AL_mpc.MPC.  Also asserts that the
oracle port (oracle/myenvs_oracle.py PortPackage + Dynamics) agrees with the reference to rounding.
Build container only (needs /root/reference).  TEST INFRASTRUCTURE, NOT PRODUCT."""
import os
import subprocess
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "_stubs"))
for p in ("/root/reference", "/root/reference/deqmpc", "/root/reference/deqmpc/my_envs"):
    sys.path.append(p)
warnings.filterwarnings("ignore")

from oracle import myenvs_oracle as MO  # noqa: E402

DT = {"pendulum1l": 0.05, "cartpole1l": 0.05, "cartpole1l_v2": 0.05, "cartpole2l": 0.03}   # train.py:99-105
UMAX = {"pendulum1l": 3.0, "cartpole1l": 100.0, "cartpole1l_v2": 10.0, "cartpole2l": 250.0}


def torch_package(name):
    """What `import cartpole1l` gives the reference (src/dynamics.cpp:40-54), on the reference's own C."""
    ref = MO.RefPackage(name)
    mod = types.SimpleNamespace(__name__=name)
    mod.dynamics = lambda q, qd, tau, h: tuple(torch.from_numpy(a) for a in ref.dynamics(q.numpy(), qd.numpy(), tau.numpy(), h.numpy()))
    mod.derivatives = lambda q, qd, tau, h: tuple(torch.from_numpy(a) for a in ref.derivatives(q.numpy(), qd.numpy(), tau.numpy(), h.numpy()))
    return mod


def reference_dynamics(name):
    import dynamics as ref_dyn   # /root/reference/deqmpc/my_envs/dynamics.py
    d = ref_dyn.Dynamics(nx=2 * MO.NQ[name], dt=DT[name], kwargs=dict(dtype=torch.float64))
    d.package = torch_package(name)
    return d


def sample(name, rs, N):
    nq = MO.NQ[name]
    x = np.concatenate([rs.uniform(-np.pi, np.pi, (N, nq)), rs.uniform(-4, 4, (N, nq))], 1)
    u = rs.uniform(-UMAX[name], UMAX[name], (N, 1))
    return x, u


def main():
    subprocess.check_call(["make", "-C", HERE])
    from qpth import AL_mpc as al_mpc, al_utils
    rs = np.random.RandomState(0)
    gold = os.path.join(ROOT, "tests", "golden")
    for name in MO.NQ:
        d = reference_dynamics(name)
        x, u = sample(name, rs, 64)
        xn, (A, B) = d.dynamics_derivatives(torch.tensor(x), torch.tensor(u))
        port = MO.Dynamics(MO.PortPackage(name), 2 * MO.NQ[name], DT[name])
        pxn, (pA, pB) = port.dynamics_derivatives(x, u)
        err = max(np.abs(pxn - xn.numpy()).max() / np.abs(xn.numpy()).max(), np.abs(pA - A.numpy()).max() / np.abs(A.numpy()).max(),
                  np.abs(pB - B.numpy()).max() / np.abs(B.numpy()).max())
        assert err < 1e-12, (name, err)
        np.savez_compressed(os.path.join(gold, f"dyn_myenvs_{name}.npz"), x=x, u=u, dt=DT[name], xn=xn.numpy(),
                            A=A.numpy(), B=B.numpy())
        print(f"{name}: |xn| {float(xn.norm()):.6f} |A| {float(A.norm()):.6f} |B| {float(B.norm()):.6f}  port-vs-reference {err:.2e}")

    # AL-MPC on the generated dynamics (the production configuration: deqmpc/policies.py:560-661 with
    # my_envs/cartpole.py CartpoleEnv, Qlqr = 1, Rlqr = 1e-8): cold call, warm call, backward
    # (the T = 20 case is BASELINE configs[2]'s horizon; appended last so that the earlier cases keep their inputs)
    for name, Bsz, T in (("cartpole1l", 8, 10), ("cartpole2l", 4, 8), ("pendulum1l", 8, 6), ("cartpole1l", 8, 20)):
        d = reference_dynamics(name)
        nx, nu = 2 * MO.NQ[name], 1
        x0 = torch.tensor(np.concatenate([rs.uniform(-1.0, 1.0, (Bsz, nx // 2)), rs.uniform(-0.5, 0.5, (Bsz, nx // 2))], 1))
        u_init = torch.tensor(rs.uniform(-1.0, 1.0, (Bsz, T, nu)))
        Cd = torch.tensor(np.tile(np.array([1.0] * nx + [1e-8 if "cartpole" in name else 1e-2] * nu), (Bsz, T, 1)))
        xref = torch.tensor(0.1 * rs.randn(Bsz, T, nx + nu))
        ul, uu = -UMAX[name] * torch.ones(nu, dtype=torch.float64), UMAX[name] * torch.ones(nu, dtype=torch.float64)
        ctrl = al_mpc.MPC(nx, nu, T, u_lower=ul, u_upper=uu, exit_unconverged=False, eps=1e-5, n_batch=Bsz, backprop=False,
                          verbose=0, u_init=u_init.clone(), solver_type="dense", dtype=torch.float64)
        ctrl.reinitialize(x0, None)
        ctrl.u_init = u_init.clone()
        out = {}
        for call in range(2):
            Cfull = torch.diag_embed(Cd).clone().requires_grad_(True)
            c = (-(Cd * xref)).clone().requires_grad_(True)
            xs, us = ctrl(x0, al_utils.QuadCost(Cfull, c), d, d.dynamics_derivatives)
            (xs.sum() + us.sum()).backward()
            out.update({f"out_x{call}": xs.detach().numpy(), f"out_u{call}": us.detach().numpy(),
                        f"out_lam{call}": ctrl.lamda_prev.numpy().copy(), f"out_rho{call}": ctrl.rho_prev.numpy().copy(),
                        f"out_dC{call}": Cfull.grad.diagonal(dim1=-2, dim2=-1).numpy().copy(), f"out_dc{call}": c.grad.numpy().copy()})
            print(f"{name} MPC call {call}: |x| {float(xs.norm()):.6f} |u| {float(us.norm()):.6f} |lam| {float(ctrl.lamda_prev.norm()):.6f}"
                  f" |dc| {float(c.grad.norm()):.6f}")
        np.savez_compressed(os.path.join(gold, f"mpc_{name}_B{Bsz}_T{T}.npz"), x0=x0.numpy(), u_init=u_init.numpy(), Cd=Cd.numpy(),
                            xref=xref.numpy(), dt=DT[name], umax=UMAX[name], **out)


if __name__ == "__main__":
    main()
