"""Expert-data sampling for the imitation-learning loop (deqmpc/datagen.py:323-408, deqmpc/utils.py:256-288), on the device.

`sample_trajectory(gt_trajs, bsz, T)` has the reference's signature and draws its candidate rows with the reference's own
RNG call (`np.random.randint(0, N, 2 * bsz)`, so a seeded run selects the same windows), but the expert data stays in HBM and
one launch pair gathers the windows, forms the running mask product and -- optionally, fused -- un-wraps the angles
(`unnormalize="pendulum" | "cartpole_nlink"`, which train.py:143-148 applies right after sampling).  Returned tensors live
on the device: the `.to(args.device)` of train.py:138 becomes a no-op."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib

_UNNORM = {None: 0, "pendulum": 1, "cartpole_nlink": 2}
_resident = {}   # id(gt_trajs) -> device copies (the data set is uploaded once)


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def to_device(gt_trajs, device):
    """Upload the merged expert data (datagen.py:323-355: float32 "state" (N,nx), "action" (N,nu), "mask" (N,)) once."""
    key = (id(gt_trajs), str(device))
    if key not in _resident:
        _resident[key] = {k: torch.as_tensor(gt_trajs[k]).to(device=device, dtype=torch.float32).contiguous()
                          for k in ("state", "action", "mask")}
    return _resident[key]


def sample_trajectory(gt_trajs, bsz, T, device=None, unnormalize=None):
    """deqmpc/datagen.py:358-408.  Returns {"state": (bsz,T,nx), "action": (bsz,T,nu), "mask": (bsz,T)} on the device."""
    if unnormalize not in _UNNORM:
        raise ValueError(f"unnormalize must be one of {list(_UNNORM)}")
    if device is None:
        st = gt_trajs["state"]
        device = st.device if torch.is_tensor(st) and st.is_cuda else torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("b200qp.datagen.sample_trajectory runs on CUDA devices only (no CPU fallback)")
    d = to_device(gt_trajs, device)
    N, nx = d["state"].shape
    nu = d["action"].shape[1]
    idxs = torch.from_numpy(np.random.randint(0, N, bsz * 2).astype(np.int64)).to(device, non_blocking=True)
    out_s = torch.empty(bsz, T, nx, dtype=torch.float32, device=device)
    out_a = torch.empty(bsz, T, nu, dtype=torch.float32, device=device)
    out_m = torch.empty(bsz, T, dtype=torch.float32, device=device)
    sel = torch.empty(bsz, dtype=torch.int64, device=device)
    status = torch.zeros(1, dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        rc = _lib.lib().b200data_sample_windows(_p(d["state"]), _p(d["action"]), _p(d["mask"]), N, nx, nu, _p(idxs), bsz * 2, bsz, T,
                                                _UNNORM[unnormalize], _p(sel), _p(out_s), _p(out_a), _p(out_m), _p(status),
                                                ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream))
    _lib.check(rc, "b200data_sample_windows")
    if int(status.item()) < bsz:  # the reference runs off the end of its candidate list here
        raise IndexError("sample_trajectory: fewer than bsz admissible windows among the 2 * bsz candidates")
    return {"state": out_s, "action": out_a, "mask": out_m, "start": sel}


def unnormalize_states_pendulum(nominal_states):
    """deqmpc/utils.py:256-270 on a device tensor (bsz, T, nx), in place like the reference."""
    return _unnormalize(nominal_states, 1)


def unnormalize_states_cartpole_nlink(nominal_states):
    """deqmpc/utils.py:273-288 on a device tensor (bsz, T, nx), in place like the reference."""
    return _unnormalize(nominal_states, 2)


def _unnormalize(states, mode):
    # the gather kernel with an identity window per sample: rows j*T .. j*T+T-1 of the flattened input
    if not states.is_cuda or states.dtype != torch.float32:
        raise RuntimeError("b200qp.datagen: float32 CUDA tensors only")
    bsz, T, nx = states.shape
    flat = states.contiguous().view(bsz * T, nx)
    dev = states.device
    act = torch.zeros(bsz * T, 1, dtype=torch.float32, device=dev)
    msk = torch.ones(bsz * T, dtype=torch.float32, device=dev)
    idxs = torch.arange(0, bsz * T, T, dtype=torch.int64, device=dev)
    out_s = torch.empty_like(states)
    out_a = torch.empty(bsz, T, 1, dtype=torch.float32, device=dev)
    out_m = torch.empty(bsz, T, dtype=torch.float32, device=dev)
    sel = torch.empty(bsz, dtype=torch.int64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().b200data_sample_windows(_p(flat), _p(act), _p(msk), bsz * T, nx, 1, _p(idxs), bsz, bsz, T, mode, _p(sel),
                                                _p(out_s), _p(out_a), _p(out_m), _p(status),
                                                ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(rc, "b200data_sample_windows (unnormalize)")
    states.copy_(out_s)
    return states
