"""Generate tests/golden/mpc_*.npz from the REAL reference (run in the build container only).

    python oracle/gen_golden_mpc.py            # needs /root/reference

For every case it (1) runs the unmodified reference qpth.AL_mpc.MPC with the reference's own
dynamics modules (deqmpc/envs.py) on CPU -- `reinitialize`, a cold call, then a warm-started call,
each followed by loss.backward() with loss = x.sum() + u.sum() (SURVEY.md section 8d cfg 2) --
(2) asserts that oracle/mpc_oracle.py reproduces it, and (3) writes inputs + outputs as a small
.npz.  TEST INFRASTRUCTURE, NOT PRODUCT.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle import mpc_oracle as MO  # noqa: E402

# name -> (env, B, T, seed, n_calls, u_bound)
CASES = {
    "pend_B16_T5": ("pendulum", 16, 5, 0, 2, 3.0),
    "pend_B64_T10": ("pendulum", 64, 10, 1, 2, 3.0),
    "pend_B8_T33": ("pendulum", 8, 33, 4, 1, 3.0),
    "integ_B8_T6": ("integrator", 8, 6, 2, 2, 1.0),
    "pend_tight_B16_T8": ("pendulum", 16, 8, 3, 3, 0.5),
}


def make_case(name):
    env, B, T, seed, n_calls, ub = CASES[name]
    rs = np.random.RandomState(seed)
    dyn = MO.Pendulum() if env == "pendulum" else MO.Integrator()
    nx, nu = dyn.nx, dyn.nu
    x0 = np.stack([rs.uniform(-np.pi, np.pi, B), rs.uniform(-1, 1, B)], 1)
    if env == "integrator":
        x0 = rs.uniform(-1, 1, (B, nx))
    u_init = rs.randn(B, T, nu)
    qd = np.array([10.0, 1.0, 0.01]) if env == "pendulum" else np.array([1.0, 1.0, 0.1])
    Cd = np.tile(qd, (B, T, 1)) * rs.uniform(0.5, 1.5, (B, 1, 1))
    xref = [0.1 * rs.randn(B, T, nx + nu) for _ in range(n_calls)]
    t = lambda a: torch.tensor(a, dtype=torch.float64)
    return dict(env=env, dyn=dyn, B=B, T=T, nx=nx, nu=nu, n_calls=n_calls, ub=ub, x0=t(x0), u_init=t(u_init),
                Cd=t(Cd), xref=[t(a) for a in xref])


def run_reference(case):
    sys.path.insert(0, os.path.join(HERE, "_stubs"))
    for p in ("/root/reference", "/root/reference/deqmpc"):
        if p not in sys.path:
            sys.path.append(p)
    from qpth import AL_mpc as al_mpc, al_utils
    import envs as RE
    B, T, nx, nu = case["B"], case["T"], case["nx"], case["nu"]
    if case["env"] == "pendulum":
        dx, dxj = RE.PendulumDynamics(), RE.PendulumDynamics_jac()
    else:
        dx, dxj = RE.IntegratorDynamics(), RE.IntegratorDynamics_jac()
    ub = case["ub"] * torch.ones(nu, dtype=torch.float64)
    ctrl = al_mpc.MPC(nx, nu, T, u_lower=-ub, u_upper=ub, exit_unconverged=False, eps=1e-5, n_batch=B, backprop=False,
                      verbose=0, u_init=case["u_init"].clone(), solver_type="dense", dtype=torch.float64)
    ctrl.reinitialize(case["x0"], None)
    ctrl.u_init = case["u_init"].clone()  # reinitialize() drops it; Tracking_MPC re-seeds it (policies.py:648-649)
    outs = []
    for k in range(case["n_calls"]):
        Cfull = torch.diag_embed(case["Cd"]).clone().requires_grad_(True)
        xr = case["xref"][k]
        c = (-(case["Cd"] * xr)).clone().requires_grad_(True)
        x, u = ctrl(case["x0"], al_utils.QuadCost(Cfull, c), dx, dxj)
        (x.sum() + u.sum()).backward()
        outs.append(dict(x=x.detach().clone(), u=u.detach().clone(), lam=ctrl.lamda_prev.clone(),
                         rho=ctrl.rho_prev.clone(), dC=Cfull.grad.diagonal(dim1=-2, dim2=-1).clone(), dc=c.grad.clone()))
    return outs


def run_oracle(case):
    B, T, nx, nu = case["B"], case["T"], case["nx"], case["nu"]
    dyn = case["dyn"]
    ub = case["ub"] * torch.ones(nu, dtype=torch.float64)
    st = MO.ALState(B, T * nx + 2 * T * nu)
    u = case["u_init"].clone()
    x = None
    outs = []
    for k in range(case["n_calls"]):
        C = case["Cd"]
        c = -(C * case["xref"][k])
        if x is None:
            x = MO.rollout(case["x0"], u, dyn)
        xs, us, ctx = MO.al_solve(x.double(), u.double(), case["x0"], C, c, dyn, -ub, ub, st)
        xf, uf = xs.float(), us.float()
        g = torch.ones(B, T, nx + nu, dtype=torch.float64)
        dC, dc = MO.al_backward(ctx, g)
        outs.append(dict(x=xf, u=uf, lam=st.lam.clone(), rho=st.rho.clone(), dC=dC, dc=dc))
        x, u = xf, uf  # the reference warm-starts from its float32 outputs (AL_mpc.py:250-251,319-320)
    return outs


def main():
    outdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    for name in CASES:
        case = make_case(name)
        ref = run_reference(case)
        ora = run_oracle(case)
        worst = 0.0
        for k, (r, o) in enumerate(zip(ref, ora)):
            for key in r:
                a, b = r[key].double(), o[key].double()
                err = ((a - b).norm() / max(b.norm().item(), 1e-300)).item()
                worst = max(worst, err)
                assert err < 1e-9, f"oracle differs from the reference: case={name} call={k} key={key} rel={err:.3e}"
        save = dict(x0=case["x0"].numpy(), u_init=case["u_init"].numpy(), Cd=case["Cd"].numpy(), ub=np.float64(case["ub"]),
                    n_calls=np.int64(case["n_calls"]))
        for k, r in enumerate(ref):
            save[f"xref{k}"] = case["xref"][k].numpy()
            for key, v in r.items():
                save[f"out_{key}{k}"] = v.numpy()
        np.savez_compressed(os.path.join(outdir, f"mpc_{name}.npz"), **save)
        print(f"{name}: {case['n_calls']} calls, oracle vs reference worst rel err {worst:.2e}", flush=True)


if __name__ == "__main__":
    main()
