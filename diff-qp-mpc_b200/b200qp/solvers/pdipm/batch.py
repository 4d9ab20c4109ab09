"""b200qp.solvers.pdipm.batch -- the reference's solver-level call surface
(qpth/solvers/pdipm/batch.py) on top of the C ABI:

    pre_factor_kkt(Q, G, A)                        :377-428  -> b200qp_prefactor
    factor_kkt(S_LU, R, d)                         :434-469  -> records d (the m x m factor is rebuilt
                                                                in shared memory by every solve)
    solve_kkt(Q_LU, d, G, A, S_LU, rx, rs, rz, ry) :351-374  -> b200qp_kkt_solve(prefactor=0)
    factor_solve_kkt(Q, D, G, A, rx, rs, rz, ry)   :315-348  -> b200qp_kkt_solve(prefactor=1)
    factor_solve_kkt_reg(Q~, D~, G, A, r.., eps)   :275-312  -> b200qp_kkt_solve in the regularised (dense) mode
    kkt_resid_reg(...)                             :229-243  -> batched products on the device
    solve_kkt_ir(..., niter)                       :245-272  -> the reference's regularise-and-refine loop on those
    forward(Q, p, G, h, A, b, Q_LU, S_LU, R, ...)  :46-208   -> b200qp_forward

The reference hands LU factor tensors between these calls; here `Q_LU`, `S_LU` and `R` are one
opaque `PreFactor` handle that owns the device workspace -- callers only ever pass them back in.
"""
from __future__ import annotations

import ctypes
from enum import Enum

import torch

from ... import _lib


class KKTSolvers(Enum):
    """qpth/solvers/pdipm/batch.py:40-43"""
    LU_FULL = 1
    LU_PARTIAL = 2
    IR_UNOPT = 3


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else ctypes.c_void_p(0)


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class PreFactor:
    """Opaque stand-in for the reference's (Q_LU, S_LU, R) triple."""

    def __init__(self, Q, G, A, flags=0, kkt_reg=0.0, defer=False):
        if not Q.is_cuda:
            raise RuntimeError("b200qp runs on CUDA tensors only (no CPU fallback); got " + str(Q.device))
        if Q.dim() != 3 or G.dim() != 3:
            raise RuntimeError("Unexpected number of dimensions.")
        nb, nineq, nz = G.shape
        neq = A.shape[1] if A is not None and A.nelement() > 0 else 0
        code = _lib.F64 if Q.dtype == torch.float64 else _lib.F32
        self.nb, self.nz, self.nineq, self.neq = nb, nz, nineq, neq
        self.Q, self.G = Q.contiguous(), G.contiguous()
        self.A = A.contiguous() if neq > 0 else None
        self.prob = _lib.Problem(nb, nz, nineq, neq, code, 20, 3, int(flags), 1e-12, nz * nz, nz, nineq * nz, nineq,
                                 neq * nz if neq else 0, neq, float(kkt_reg))
        L = _lib.lib()
        nbytes = L.b200qp_workspace_bytes(ctypes.byref(self.prob))
        if nbytes == 0:
            if flags:
                raise NotImplementedError("b200qp: the regularised KKT back-ends (solve_kkt_ir / factor_solve_kkt_reg) "
                                          "live in the fused kernels: nineq <= 128 and nz, neq + nineq <= 128")
            raise RuntimeError("b200qp: unsupported problem size")
        self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=Q.device)
        self.d = None
        if defer:  # the caller solves with prefactor=1 (one call: pre-factorisation + solve)
            return
        status = torch.zeros(_lib.STATUS_DOUBLES, dtype=torch.float64, device=Q.device)
        with torch.cuda.device(Q.device):
            rc = L.b200qp_prefactor(ctypes.byref(self.prob), _p(self.Q), _p(self.G), _p(self.A), _p(self.workspace),
                                    _p(status), _stream(Q.device))
        _lib.check(rc, "b200qp_prefactor")
        st = status.tolist()
        if st[_lib.ST_Q_FAIL] > 0:
            raise RuntimeError("qpth Error: Cannot perform LU factorization on Q. "
                               "Please make sure that your Q matrix is PSD and has a non-zero diagonal.")
        if st[_lib.ST_AQA_FAIL] > 0:
            raise RuntimeError("qpth Error: Cannot perform LU factorization on AQ^{-1}A^T. "
                               "Please make sure that your A matrix is full rank.")

    def solve(self, d, rx, rs, rz, ry, prefactor=0):
        L = _lib.lib()
        opt = dict(device=rx.device, dtype=rx.dtype)
        dx = torch.empty(self.nb, self.nz, **opt)
        ds = torch.empty(self.nb, self.nineq, **opt)
        dz = torch.empty(self.nb, self.nineq, **opt)
        dy = torch.empty(self.nb, self.neq, **opt) if self.neq > 0 else None
        t = [x.contiguous() if x is not None else None for x in (d, rx, rs, rz, ry)]
        with torch.cuda.device(rx.device):
            rc = L.b200qp_kkt_solve(ctypes.byref(self.prob), int(prefactor), _p(self.Q), _p(self.G), _p(self.A), _p(t[0]),
                                    _p(t[1]), _p(t[2]), _p(t[3]), _p(t[4]), _p(dx), _p(ds), _p(dz), _p(dy),
                                    _p(self.workspace), _stream(rx.device))
        _lib.check(rc, "b200qp_kkt_solve")
        return dx, ds, dz, dy


def pre_factor_kkt(Q, G, A):
    """-> (Q_LU, S_LU, R): three references to one PreFactor handle."""
    h = PreFactor(Q, G, A)
    return h, h, h


def factor_kkt(S_LU, R, d):
    """The reference completes the block LU in place for the current d; here the handle records d
    and the solve kernel factors T = R + diag(1/d) on chip."""
    S_LU.d = d.contiguous()


def solve_kkt(Q_LU, d, G, A, S_LU, rx, rs, rz, ry):
    return S_LU.solve(d, rx, rs, rz, ry)


def _diag_of(D):
    return torch.diagonal(D, dim1=-2, dim2=-1).contiguous() if D.dim() == 3 else D


def factor_solve_kkt(Q, D, G, A, rx, rs, rz, ry):
    return PreFactor(Q, G, A).solve(_diag_of(D), rx, rs, rz, ry)


def kkt_resid_reg(Q_tilde, D_tilde, G, A, eps, dx, ds, dz, dy, rx, rs, rz, ry):
    """Residual of the regularised KKT system (qpth/solvers/pdipm/batch.py:229-243)

        [Q~ 0 G' A'; 0 D~ I 0; G I -eps 0; A 0 0 -eps] [dx ds dz dy]' + [rx rs rz ry]'

    evaluated with batched products on the device (no solver kernel involved)."""
    mv = lambda M, v: torch.einsum("brc,bc->br", M, v)
    tv = lambda M, v: torch.einsum("brc,br->bc", M, v)
    resx = mv(Q_tilde, dx) + tv(G, dz) + rx
    if dy is not None:
        resx = resx + tv(A, dy)
    ress = mv(D_tilde, ds) + dz + rs
    resz = mv(G, dx) + ds - eps * dz + rz
    resy = mv(A, dx) - eps * dy + ry if dy is not None else None
    return resx, ress, resz, resy


def factor_solve_kkt_reg(Q_tilde, D, G, A, rx, rs, rz, ry, eps):
    """Solve the regularised system above (qpth/solvers/pdipm/batch.py:275-312) with the fused kernels.

    Eliminating ds = -(rs + dz) / d leaves the Schur system of the unregularised problem with 1/d replaced by
    1/d + eps on the inequality block and +eps on the equality block.  That is the system the kernels factor in
    their DenseQPFunction mode (B200QP_FLAG_DENSE regularises Q and A Q^-1 A' by kkt_reg); the inequality block is
    reached by solving with d' = d / (1 + eps d) and the right-hand side rs' = rs d'/d, and ds is recovered
    from its own row afterwards."""
    d = _diag_of(D)
    nb, nineq, nz = G.shape
    neq = A.shape[1] if A is not None and A.nelement() > 0 else 0
    eye = torch.eye(nz, dtype=Q_tilde.dtype, device=Q_tilde.device)
    h = PreFactor(Q_tilde - eps * eye, G, A, flags=_lib.FLAG_DENSE, kkt_reg=eps, defer=True)
    dp = d / (1.0 + eps * d)
    dx, _, dz, dy = h.solve(dp, rx, rs * dp / d, rz, ry if neq > 0 else None, prefactor=1)
    ds = (-rs - dz) / d
    return dx, ds, dz, dy


def solve_kkt_ir(Q, D, G, A, rx, rs, rz, ry, niter=1):
    """Regularise-and-refine KKT solve (qpth/solvers/pdipm/batch.py:245-272): solve with Q + eps I, D + eps I and
    -eps I in the constraint block, then `niter` steps of iterative refinement against the residual of the system
    regularised only in the constraint block -- the same loop as the reference, every solve on the GPU."""
    nb, nineq, nz = G.shape
    eps = 1e-7
    Dm = D if D.dim() == 3 else torch.diag_embed(D)
    Q_tilde = Q + eps * torch.eye(nz, dtype=Q.dtype, device=Q.device)
    D_tilde = Dm + eps * torch.eye(nineq, dtype=Q.dtype, device=Q.device)
    neq = A.shape[1] if A is not None and A.nelement() > 0 else 0
    A_ = A if neq > 0 else None
    dx, ds, dz, dy = factor_solve_kkt_reg(Q_tilde, D_tilde, G, A_, rx, rs, rz, ry, eps)
    resx, ress, resz, resy = kkt_resid_reg(Q, Dm, G, A_, eps, dx, ds, dz, dy, rx, rs, rz, ry if neq > 0 else None)
    for _ in range(niter):
        ddx, dds, ddz, ddy = factor_solve_kkt_reg(Q_tilde, D_tilde, G, A_, -resx, -ress, -resz,
                                                  -resy if resy is not None else None, eps)
        dx, ds, dz = dx + ddx, ds + dds, dz + ddz
        dy = dy + ddy if dy is not None else None
        resx, ress, resz, resy = kkt_resid_reg(Q, Dm, G, A_, eps, dx, ds, dz, dy, rx, rs, rz, ry if neq > 0 else None)
    return dx, ds, dz, dy


def forward(Q, p, G, h, A, b, Q_LU=None, S_LU=None, R=None, dyn_res=None, cost_grad=None,
            eps=1e-12, verbose=0, notImprovedLim=3, maxIter=20, solver=KKTSolvers.LU_PARTIAL):
    """-> (x, y, z, s) best iterates (qpth/solvers/pdipm/batch.py:46-208).  The pre-factorisation
    handles are accepted for signature compatibility; the fused forward recomputes it (one launch)."""
    from ...qp import QPFunction
    fn = QPFunction(eps=eps, verbose=verbose, notImprovedLim=notImprovedLim, maxIter=maxIter, check_Q_spd=False)
    with torch.no_grad():
        Qd = Q.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        z = fn(Qd, p.detach(), G.detach(), h.detach(), A.detach(), b.detach(), dyn_res, cost_grad)
    ctx = z.grad_fn
    y = ctx.nus if ctx.neq > 0 else None
    return z.detach(), y, ctx.lams, ctx.slacks
