"""Generate tests/golden/datagen_{pendulum,cartpole}.npz from the REAL reference: deqmpc/datagen.py:358-408
`sample_trajectory` on a synthetic merged expert data set (seeded numpy RNG), followed by deqmpc/utils.py:256-288
`unnormalize_states_pendulum` / `unnormalize_states_cartpole_nlink` as train.py:143-148 applies them.
Build container only.  TEST INFRASTRUCTURE, NOT PRODUCT."""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "_stubs"))
for p in ("/root/reference", "/root/reference/deqmpc"):
    sys.path.append(p)
warnings.filterwarnings("ignore")


def synthetic(kind, rs):
    """10 trajectories of 50 rows; angles wrapped into [-pi, pi) so that windows cross the wrap; mask 0 on the last row"""
    rows_s, rows_a, rows_m = [], [], []
    nx = 2 if kind == "pendulum" else 4
    for _ in range(10):
        th = rs.uniform(-np.pi, np.pi) + np.cumsum(rs.uniform(0.05, 0.6, 50)) * rs.choice([-1, 1])
        th = (th + np.pi) % (2 * np.pi) - np.pi
        if kind == "pendulum":
            s = np.stack([th, rs.randn(50)], 1)
        else:
            s = np.stack([rs.randn(50), th, rs.randn(50) * 1.2, rs.randn(50)], 1)
        rows_s.append(s); rows_a.append(rs.randn(50, 1)); m = np.ones(50); m[-1] = 0; rows_m.append(m)
    return {"state": torch.tensor(np.concatenate(rows_s), dtype=torch.float32), "action": torch.tensor(np.concatenate(rows_a), dtype=torch.float32),
            "mask": torch.tensor(np.concatenate(rows_m), dtype=torch.float32)}, nx


def main():
    import importlib
    # datagen.py imports gym / sac experts at module level; only sample_trajectory is needed
    src = open("/root/reference/deqmpc/datagen.py").read()
    start = src.index("def sample_trajectory")
    end = src.index("def test_qp_mpc")
    ns = {"np": np, "torch": torch}
    exec(compile(src[start:end], "/root/reference/deqmpc/datagen.py", "exec"), ns)   # the reference's own function text
    usrc = open("/root/reference/deqmpc/utils.py").read()
    ustart = usrc.index("def unnormalize_states_pendulum")
    nxt = usrc.find("\ndef ", usrc.index("def unnormalize_states_cartpole_nlink") + 10)
    uend = nxt if nxt > 0 else len(usrc)
    exec(compile(usrc[ustart:uend], "/root/reference/deqmpc/utils.py", "exec"), ns)
    gold = os.path.join(ROOT, "tests", "golden")
    for kind in ("pendulum", "cartpole"):
        rs = np.random.RandomState(3 if kind == "pendulum" else 4)
        data, nx = synthetic(kind, rs)
        np.random.seed(1234)
        tr = ns["sample_trajectory"](data, 64, 8)
        raw = tr["state"].clone()
        un = ns["unnormalize_states_pendulum" if kind == "pendulum" else "unnormalize_states_cartpole_nlink"](tr["state"].clone())
        np.savez_compressed(os.path.join(gold, f"datagen_{kind}.npz"), data_state=data["state"].numpy(), data_action=data["action"].numpy(),
                            data_mask=data["mask"].numpy(), state_raw=raw.numpy(), state=un.numpy(), action=tr["action"].numpy(),
                            mask=tr["mask"].numpy())
        print(kind, "windows", tuple(raw.shape), "rows changed by the un-normalisation:", int((raw != un).any(-1).sum()),
              "masked-out steps:", int((tr["mask"] == 0).sum()))


if __name__ == "__main__":
    main()
