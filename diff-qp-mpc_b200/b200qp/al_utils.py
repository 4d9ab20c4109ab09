"""Types that appear in the reference's MPC signatures (qpth/al_utils.py:8-13) and the autograd
node of the augmented-Lagrangian solve.

`ALSolve` stands where the reference chains `al_iter` x `NewtonAL.apply` (qpth/al_utils.py:363-500,
qpth/AL_mpc.py:282-310): the whole outer loop is ONE kernel launch (b200mpc_al_solve); backward is
the implicit step  -H^-1 g  with the block factor the forward saved (b200mpc_al_backward), which
is what the reference's last NewtonAL.backward computes (earlier AL iterations are detached in the
reference too, AL_mpc.py:284).
"""
from __future__ import annotations

import ctypes
from collections import namedtuple

import torch

from . import _lib

QuadCost = namedtuple("QuadCost", "C c")
LinDx = namedtuple("LinDx", "F f")
QuadCost.__new__.__defaults__ = (None,) * len(QuadCost._fields)
LinDx.__new__.__defaults__ = (None,) * len(LinDx._fields)

NEWTON_STEPS = 4   # qpth/al_utils.py:397
N_LINESEARCH = 20  # qpth/al_utils.py:504


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else ctypes.c_void_p(0)


def _dtype_code(dtype):
    if dtype == torch.float64:
        return _lib.F64
    if dtype == torch.float32:
        return _lib.F32
    raise RuntimeError(f"b200qp: unsupported solver dtype {dtype} (float64 or float32)")


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class ALState:
    """Solver state the reference keeps on the module between calls (qpth/AL_mpc.py:316-318,432-439)."""

    def __init__(self, lam, rho):
        self.lam, self.rho = lam, rho   # (B,M), (B,1)
        self.hist = None                # (cost (K,B), lam (K,B,M), rho (K,B)), oldest first
        self.status = None


class ALSolve(torch.autograd.Function):
    """(C_diag, c) -> (x float32, u float32, status); every other argument is data."""

    @staticmethod
    def forward(ctx, C, c, x_init, u_init, x0, u_lower, u_upper, state, spec, al_iter):
        dev, dtype = C.device, C.dtype
        if not C.is_cuda:
            raise RuntimeError("b200qp AL-MPC runs on CUDA tensors only (no CPU fallback); got " + str(dev))
        env, params, nx, nu = spec
        B, T, nt = C.shape
        assert nt == nx + nu
        M = T * nx + 2 * T * nu
        L = _lib.lib()
        warm = 0 if state.hist is None else 1
        K = 0 if state.hist is None else state.hist[0].shape[0]
        pa = (ctypes.c_double * _lib.MPC_MAX_PARAMS)(*(list(params) + [0.0] * (_lib.MPC_MAX_PARAMS - len(params))))
        prob = _lib.MpcProblem(B, T, env, _dtype_code(dtype), int(al_iter), NEWTON_STEPS, N_LINESEARCH, warm, K, 0, pa)
        opt = dict(device=dev, dtype=dtype)
        lam = state.lam.to(**opt).contiguous().clone()
        rho = state.rho.to(**opt).reshape(B).contiguous().clone()
        hist_out = (torch.empty(al_iter + 1, B, **opt), torch.empty(al_iter + 1, B, M, **opt),
                    torch.empty(al_iter + 1, B, **opt))
        xu = torch.empty(B, T, nt, **opt)
        x = torch.empty(B, T, nx, device=dev, dtype=torch.float32)
        u = torch.empty(B, T, nu, device=dev, dtype=torch.float32)
        status = torch.empty(B, **opt)
        fe = L.b200mpc_factor_elems(ctypes.byref(prob))
        if fe == 0:
            raise RuntimeError("b200qp: unsupported MPC problem size")
        factor = torch.empty(B, fe, **opt)
        sb = L.b200mpc_scratch_bytes(ctypes.byref(prob))
        scratch = torch.empty(sb, dtype=torch.uint8, device=dev) if sb else None
        keep = [t.detach().to(**opt).contiguous() for t in (x_init, u_init, x0, C, c, u_lower, u_upper)]
        hin = state.hist if warm else (None, None, None)
        hin = tuple(None if h is None else h.to(**opt).contiguous() for h in hin)
        buf = _lib.MpcBuffers(*[_p(t) for t in keep], _p(lam), _p(rho), _p(hin[0]), _p(hin[1]), _p(hin[2]),
                              _p(hist_out[0]), _p(hist_out[1]), _p(hist_out[2]), _p(xu), _p(x), _p(u), _p(status),
                              _p(factor), _p(scratch))
        with torch.cuda.device(dev):
            rc = L.b200mpc_al_solve(ctypes.byref(prob), ctypes.byref(buf), _stream(dev))
        _lib.check(rc, "b200mpc_al_solve")
        state.lam, state.rho, state.hist, state.status = lam, rho.reshape(B, 1), hist_out, status
        ctx.prob, ctx.factor, ctx.xu, ctx.nx = prob, factor, xu, nx
        ctx.mark_non_differentiable(status)
        return x, u, status

    @staticmethod
    def backward(ctx, gx, gu, _gs):
        xu, factor = ctx.xu, ctx.factor
        B, T, nt = xu.shape
        nx = ctx.nx
        g = torch.zeros(B, T, nt, device=xu.device, dtype=xu.dtype)
        if gx is not None:
            g[..., :nx] = gx.to(xu.dtype)
        if gu is not None:
            g[..., nx:] = gu.to(xu.dtype)
        dC, dc = torch.empty_like(xu), torch.empty_like(xu)
        L = _lib.lib()
        with torch.cuda.device(xu.device):
            rc = L.b200mpc_al_backward(ctypes.byref(ctx.prob), _p(factor), _p(xu), _p(g), _p(dC), _p(dc),
                                       _stream(xu.device))
        _lib.check(rc, "b200mpc_al_backward")
        return (dC, dc) + (None,) * 8
