"""Expert-data sampling on the device (b200qp/datagen.py -> b200data_sample_windows) against goldens of the reference's own
`sample_trajectory` + `unnormalize_states_*` (oracle/gen_golden_datagen.py).  float32 gathers and one subtraction: bit-exact."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("kind", ["pendulum", "cartpole"])
def test_sample_trajectory_bit_exact(kind, cuda_device):
    from b200qp import datagen
    g = dict(np.load(os.path.join(GOLDEN, f"datagen_{kind}.npz")))
    data = {"state": torch.tensor(g["data_state"]), "action": torch.tensor(g["data_action"]), "mask": torch.tensor(g["data_mask"])}
    mode = "pendulum" if kind == "pendulum" else "cartpole_nlink"
    np.random.seed(1234)
    raw = datagen.sample_trajectory(data, 64, 8, device=cuda_device)
    np.random.seed(1234)
    fused = datagen.sample_trajectory(data, 64, 8, device=cuda_device, unnormalize=mode)
    for k, ref in (("state", "state_raw"), ("action", "action"), ("mask", "mask")):
        assert torch.equal(raw[k].cpu(), torch.tensor(g[ref])), k
    assert torch.equal(fused["state"].cpu(), torch.tensor(g["state"])), "fused un-normalisation"
    assert torch.equal(fused["mask"].cpu(), torch.tensor(g["mask"]))
    assert (g["state"] != g["state_raw"]).any(), "the golden must exercise the un-normalisation"
    # the stand-alone un-normalisation (train.py:143-148 calls it on the sampled batch), in place
    fn = datagen.unnormalize_states_pendulum if kind == "pendulum" else datagen.unnormalize_states_cartpole_nlink
    s = raw["state"].clone()
    out = fn(s)
    assert out is s and torch.equal(s.cpu(), torch.tensor(g["state"]))
    # windows never start on an end-of-trajectory row; rows past the end of the data are zero
    assert bool((data["mask"][raw["start"].cpu()] != 0).all())


def test_sample_trajectory_errors(cuda_device):
    from b200qp import datagen
    data = {"state": torch.zeros(10, 2), "action": torch.zeros(10, 1), "mask": torch.zeros(10)}   # nothing admissible
    with pytest.raises(IndexError):
        datagen.sample_trajectory(data, 4, 3, device=cuda_device)
    with pytest.raises(RuntimeError):
        datagen.sample_trajectory(data, 4, 3, device="cpu")
