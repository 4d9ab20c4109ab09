"""CPU suite: the torch-only parts of b200qp/policies.py (DEQLayer, the imitation losses) against the golden of the real
reference policy (oracle/gen_golden_policy.py): the first network pass, before any MPC solve, and the loss evaluated on the
reference's own trajectories.  (The DEQ-MPC loop itself needs the CUDA library: tests/test_mpc_parity_gpu.py.)"""
import os
import types

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _args():
    return types.SimpleNamespace(T=5, bsz=8, dtype="double", solver_type="al", nq=1, hdim=128, layer_type="mlp", deq_out_type=1,
                                 policy_out_type=1, kernel_width=3, pooling="mean", deq_iter=6, qp_iter=1, eps=1e-2, warm_start=True,
                                 device=torch.device("cpu"), deq=True, en_qp_solve=True)


def test_deq_layer_first_pass_and_loss_match_reference():
    from b200qp import policies
    g = dict(np.load(os.path.join(GOLDEN, "policy_integrator_B8_T5.npz")))
    env = types.SimpleNamespace(nu=1, nx=2, dt=0.1)
    args = _args()
    model = policies.DEQLayer(args, env)
    model.load_state_dict({k[2:]: torch.tensor(v) for k, v in g.items() if k.startswith("w_")})
    model.double()
    x = torch.tensor(g["x"])
    x_ref = torch.cat([x] * args.T, dim=-1)
    out, z = model(x_ref, torch.zeros(8, args.hdim, dtype=torch.float64))
    net0 = torch.cat([x[:, None, :], out.view(-1, args.T - 1, 2)], dim=1)
    assert (net0 - torch.tensor(g["net0"])).abs().max() <= 1e-12
    # the loss of policies.py:800-808 on the reference's own trajectories
    trajs = [(torch.tensor(g[f"net{k}"]), torch.tensor(g[f"xs{k}"]), torch.tensor(g[f"us{k}"])) for k in range(6)]
    pol = types.SimpleNamespace(out_type=1, nq=1)
    loss, loss_end = policies.compute_loss(pol, torch.tensor(g["gt_states"]), torch.tensor(g["gt_actions"]), torch.tensor(g["mask"]), trajs, args)
    assert abs(float(loss) - float(g["loss"])) <= 1e-12 * abs(float(g["loss"]))
    assert abs(float(loss_end) - float(g["loss_end"])) <= 1e-12 * abs(float(g["loss_end"]))


def test_deq_layer_rejects_what_the_reference_cannot_build():
    import pytest
    from b200qp import policies
    env = types.SimpleNamespace(nu=1, nx=2, dt=0.1)
    a = _args(); a.layer_type = "gcn"
    with pytest.raises(NotImplementedError):
        policies.DEQLayer(a, env)
    a = _args(); a.deq_out_type = 3
    with pytest.raises(NotImplementedError):
        policies.DEQLayer(a, env)
