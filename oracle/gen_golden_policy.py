"""Generate tests/golden/policy_integrator_B8_T5.npz from the REAL reference's DEQMPCPolicy (deqmpc/policies.py:426-529: six
DEQ iterations of DEQLayer -> Tracking_MPC on the IntegratorEnv, the configuration of deqmpc/run.sh), its loss
(policies.py:800-808) and the gradients of the network parameters.  The network weights are saved with the golden so that
the test loads the same parameters.  Build container only.  TEST INFRASTRUCTURE, NOT PRODUCT."""
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "_stubs"))
for p in ("/root/reference", "/root/reference/deqmpc"):
    sys.path.append(p)
warnings.filterwarnings("ignore")


def make_args(B, T, device):
    return types.SimpleNamespace(T=T, bsz=B, dtype="double", solver_type="al", nq=1, hdim=128, layer_type="mlp", deq_out_type=1,
                                 policy_out_type=1, kernel_width=3, pooling="mean", deq_iter=6, qp_iter=1, eps=1e-2, warm_start=True,
                                 device=device, deq=True, en_qp_solve=True)


def main():
    import envs as RE
    import policies as RP
    env = RE.IntegratorEnv()
    B, T = 8, 5
    args = make_args(B, T, torch.device("cpu"))
    args.Q, args.R = env.Qlqr, env.Rlqr
    torch.manual_seed(0)
    policy = RP.DEQMPCPolicy(args, env)
    # The reference runs the network in float32 (the MPC hands float32 trajectories back, qpth/AL_mpc.py:319-320).  The chain
    # is chaotic at that noise level: the AL solve takes 8 Newton steps with a 20-way discrete line search (al_utils.py:503-527)
    # and is fed back six times; measured on the reference itself, moving the input state by 1e-7 relative moves the
    # trajectories of DEQ iteration 2 by 1e-2 and those of iteration 5 by 3e-1 (float32 or float64 network alike), while 1e-12
    # moves them by < 5e-7.  Parity of the CHAIN is therefore only meaningful with the float32 noise removed: the golden is
    # produced with the network cast to float64 and the MPC outputs cast back to float64 (the only change to the reference:
    # two dtype casts); the GPU test runs the same variant.  Iterations 0-1 of the native float32 policy are checked as well.
    policy.model.double()
    policy.model.init_z = lambda bsz: torch.zeros(bsz, args.hdim, dtype=torch.float64)
    _fwd = policy.tracking_mpc.forward
    policy.tracking_mpc.forward = lambda *a: tuple(o.double() for o in _fwd(*a))
    rs = np.random.RandomState(5)
    x = torch.tensor(np.stack([rs.uniform(-2, 2, B), rs.uniform(-1, 1, B)], 1), dtype=torch.float64)
    gt_states = torch.tensor(0.5 * rs.randn(B, T, 2), dtype=torch.float64)
    gt_actions = torch.tensor(0.5 * rs.randn(B, T, 1), dtype=torch.float64)
    mask = torch.ones(B, T, dtype=torch.float64)
    trajs, dyn_res = policy(x, gt_states, gt_actions, mask, qp_solve=True)
    loss, loss_end = RP.compute_loss(policy, gt_states, gt_actions, mask, trajs, args)
    loss.backward()
    save = dict(x=x.numpy(), gt_states=gt_states.numpy(), gt_actions=gt_actions.numpy(), mask=mask.numpy(), loss=float(loss),
                loss_end=float(loss_end), dyn_res=float(dyn_res))
    for k, (a, b, c) in enumerate(trajs):
        save.update({f"net{k}": a.detach().numpy(), f"xs{k}": b.detach().numpy(), f"us{k}": c.detach().numpy()})
    for n, p in policy.model.named_parameters():
        save["w_" + n] = p.detach().numpy()
        save["g_" + n] = p.grad.numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "policy_integrator_B8_T5.npz"), **save)
    print("policy golden: loss", float(loss), "dyn_res", dyn_res, "|x5|", float(trajs[-1][1].norm()),
          "grad norm", float(sum(p.grad.norm() ** 2 for p in policy.model.parameters()) ** 0.5))


if __name__ == "__main__":
    main()
