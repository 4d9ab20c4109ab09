"""ctypes binding of libb200qp.so (the C ABI in include/b200qp.h).

The product path has NO CPU fallback: if the shared library is missing, or a CUDA tensor is not
supplied, the call fails loudly.  Build with `python -c "import __graft_entry__ as g; g.build()"`
(or `make -C diff-qp-mpc_b200/csrc`).
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200qp.so")

F64, F32 = 0, 1
STATUS_DOUBLES = 8
ST_NITER, ST_BEST_MAX, ST_Q_FAIL, ST_AQA_FAIL, ST_LAUNCHES, ST_SPEC_FAIL, ST_NAN_ONSET = 0, 1, 2, 3, 4, 5, 6
MAX_ITER_CAP = 64
FLAG_DENSE = 1
FLAG_EXACT = 2
FLAG_FACTORED_GRAD = 4
HOST_SLOTS = 4
PHASE_ALL, PHASE_BEGIN, PHASE_END = -1000, -1001, -1002

EXPORTS = (
    "b200qp_workspace_bytes", "b200qp_prefactor", "b200qp_forward", "b200qp_forward_phase", "b200qp_forward_phase_cb", "b200qp_forward_cb_step", "b200qp_slot_offset", "b200qp_backward", "b200qp_kkt_solve",
    "b200qp_solve_host", "b200qp_solve_host_submit", "b200qp_solve_host_wait", "b200qp_last_cuda_error", "b200qp_version", "b200qp_profile_enable",
    "b200qp_profile_read", "b200qp_set_option",
    "b200mpc_env_dims", "b200mpc_factor_elems", "b200mpc_scratch_bytes", "b200mpc_al_solve", "b200mpc_al_backward",
    "b200dyn_step", "b200dyn_jac", "b200dyn_rollout", "b200data_sample_windows",
)

ENV_PENDULUM, ENV_INTEGRATOR, ENV_PENDULUM_DX, ENV_CARTPOLE_DX, ENV_REX_QUADROTOR = 0, 1, 2, 3, 4
ENV_PENDULUM1L, ENV_CARTPOLE1L, ENV_CARTPOLE2L = 5, 6, 7
ENV_CARTPOLE1L_V1, ENV_CARTPOLE2L_V1 = 8, 9
MPC_MAX_PARAMS = 64


class MpcProblem(ctypes.Structure):
    """b200mpc_problem_t (include/b200mpc.h)"""
    _fields_ = [
        ("B", ctypes.c_int32), ("T", ctypes.c_int32), ("env", ctypes.c_int32), ("dtype", ctypes.c_int32),
        ("al_iter", ctypes.c_int32), ("newton_steps", ctypes.c_int32), ("n_ls", ctypes.c_int32),
        ("warm", ctypes.c_int32), ("hist_len", ctypes.c_int32), ("reserved", ctypes.c_int32),
        ("params", ctypes.c_double * MPC_MAX_PARAMS),
    ]


class MpcBuffers(ctypes.Structure):
    """b200mpc_buffers_t (include/b200mpc.h)"""
    _fields_ = [(n, ctypes.c_void_p) for n in (
        "x_init", "u_init", "x0", "C", "c", "u_lower", "u_upper", "lam", "rho",
        "cost_hist_in", "lam_hist_in", "rho_hist_in", "cost_hist_out", "lam_hist_out", "rho_hist_out",
        "xu", "x", "u", "status", "factor", "scratch")]


class Problem(ctypes.Structure):
    """b200qp_problem_t"""
    _fields_ = [
        ("nb", ctypes.c_int32), ("nz", ctypes.c_int32), ("nineq", ctypes.c_int32), ("neq", ctypes.c_int32),
        ("dtype", ctypes.c_int32), ("max_iter", ctypes.c_int32), ("not_improved_lim", ctypes.c_int32),
        ("flags", ctypes.c_int32), ("eps", ctypes.c_double),
        ("sQ", ctypes.c_int64), ("sp", ctypes.c_int64), ("sG", ctypes.c_int64), ("sh", ctypes.c_int64),
        ("sA", ctypes.c_int64), ("sb", ctypes.c_int64), ("kkt_reg", ctypes.c_double),
    ]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA library has not been built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` at the repo root. "
            "There is no CPU fallback for this path.")
    L = ctypes.CDLL(LIB_PATH)
    vp, pp = ctypes.c_void_p, ctypes.POINTER(Problem)
    L.b200qp_workspace_bytes.restype = ctypes.c_size_t
    L.b200qp_workspace_bytes.argtypes = [pp]
    L.b200qp_forward.restype = ctypes.c_int
    L.b200qp_forward.argtypes = [pp] + [vp] * 13
    L.b200qp_forward_phase.restype = ctypes.c_int
    L.b200qp_forward_phase.argtypes = [pp, ctypes.c_int] + [vp] * 13
    L.b200qp_forward_phase_cb.restype = ctypes.c_int
    L.b200qp_forward_phase_cb.argtypes = [pp, ctypes.c_int] + [vp] * 15
    L.b200qp_forward_cb_step.restype = ctypes.c_int
    L.b200qp_forward_cb_step.argtypes = [pp, ctypes.c_int] + [vp] * 3
    L.b200qp_slot_offset.restype = ctypes.c_size_t
    L.b200qp_slot_offset.argtypes = [pp]
    L.b200qp_backward.restype = ctypes.c_int
    L.b200qp_backward.argtypes = [pp] + [vp] * 13
    L.b200qp_prefactor.restype = ctypes.c_int
    L.b200qp_prefactor.argtypes = [pp] + [vp] * 6
    L.b200qp_kkt_solve.restype = ctypes.c_int
    L.b200qp_kkt_solve.argtypes = [pp, ctypes.c_int] + [vp] * 14
    L.b200qp_solve_host.restype = ctypes.c_int
    L.b200qp_solve_host.argtypes = [pp] + [vp] * 18
    L.b200qp_solve_host_submit.restype = ctypes.c_int
    L.b200qp_solve_host_submit.argtypes = [ctypes.c_int, pp] + [vp] * 18
    L.b200qp_solve_host_wait.restype = ctypes.c_int
    L.b200qp_solve_host_wait.argtypes = [ctypes.c_int]
    L.b200data_sample_windows.restype = ctypes.c_int
    L.b200data_sample_windows.argtypes = [vp, vp, vp, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_int, vp, vp, vp, vp, vp, vp]
    L.b200qp_profile_enable.restype = None
    L.b200qp_profile_enable.argtypes = [ctypes.c_int]
    L.b200qp_profile_read.restype = ctypes.c_int
    L.b200qp_profile_read.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int), ctypes.c_int]
    mp, bp = ctypes.POINTER(MpcProblem), ctypes.POINTER(MpcBuffers)
    ip = ctypes.POINTER(ctypes.c_int)
    L.b200mpc_env_dims.restype = ctypes.c_int
    L.b200mpc_env_dims.argtypes = [ctypes.c_int, ip, ip]
    L.b200mpc_factor_elems.restype = ctypes.c_size_t
    L.b200mpc_factor_elems.argtypes = [mp]
    L.b200mpc_scratch_bytes.restype = ctypes.c_size_t
    L.b200mpc_scratch_bytes.argtypes = [mp]
    L.b200mpc_al_solve.restype = ctypes.c_int
    L.b200mpc_al_solve.argtypes = [mp, bp, vp]
    L.b200mpc_al_backward.restype = ctypes.c_int
    L.b200mpc_al_backward.argtypes = [mp] + [vp] * 6
    dparr = ctypes.POINTER(ctypes.c_double)
    L.b200dyn_step.restype = ctypes.c_int
    L.b200dyn_step.argtypes = [ctypes.c_int, ctypes.c_int, dparr, vp, vp, vp, ctypes.c_int64, vp]
    L.b200dyn_jac.restype = ctypes.c_int
    L.b200dyn_jac.argtypes = [ctypes.c_int, ctypes.c_int, dparr, vp, vp, vp, vp, vp, ctypes.c_int64, vp]
    L.b200dyn_rollout.restype = ctypes.c_int
    L.b200dyn_rollout.argtypes = [ctypes.c_int, ctypes.c_int, dparr, vp, vp, vp, ctypes.c_int64, ctypes.c_int32, vp]
    L.b200qp_set_option.restype = ctypes.c_int
    L.b200qp_set_option.argtypes = [ctypes.c_char_p, ctypes.c_int]
    L.b200qp_last_cuda_error.restype = ctypes.c_char_p
    L.b200qp_version.restype = ctypes.c_char_p
    _lib = L
    return L


def set_option(name, value):
    """Process-wide tuning knob of the library (include/b200qp.h: b200qp_set_option)."""
    rc = lib().b200qp_set_option(name.encode(), int(value))
    if rc:
        raise ValueError(f"b200qp: unknown option {name!r}")


def check(rc, what):
    if rc == 0:
        return
    msg = {-1: "invalid argument", -2: "CUDA error", -3: "problem too large for the kernels"}.get(rc, "error")
    detail = lib().b200qp_last_cuda_error().decode() if rc == -2 else ""
    raise RuntimeError(f"b200qp: {what} failed ({rc}: {msg}) {detail}")
