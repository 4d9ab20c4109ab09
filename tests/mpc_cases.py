"""Shared helpers for the MPC parity tests: golden cases of the real reference
(oracle/gen_golden_mpc.py) and the oracle-side runner."""
import os

import numpy as np
import torch

from oracle import mpc_oracle as MO
from oracle.gen_golden_mpc import CASES  # noqa: F401

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(case):
    g = dict(np.load(os.path.join(GOLDEN_DIR, f"mpc_{case}.npz")))
    return {k: (torch.as_tensor(v) if isinstance(v, np.ndarray) and v.ndim > 0 else v) for k, v in g.items()}


def oracle_dyn(case):
    return MO.Pendulum() if CASES[case][0] == "pendulum" else MO.Integrator()


def rel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return ((a - b).norm() / max(b.norm().item(), 1e-300)).item()
