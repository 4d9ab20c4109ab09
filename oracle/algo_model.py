"""CPU model of the *GPU* algorithm (Cholesky + explicit Q^-1 block elimination)  --  TEST
INFRASTRUCTURE / DESIGN TOOL, NOT PRODUCT.

oracle/qp_oracle.py restates the reference (partial-pivot LU, triangular solves).  The CUDA
kernels in diff-qp-mpc_b200/csrc use a different but mathematically equivalent elimination
(SPD Cholesky, explicit Q^-1 and [A;G]Q^-1 so that the only sequential work per iteration is
the m x m factor and its two solves).  This file is that elimination written with torch on the
CPU so that the design can be checked against the golden vectors (iteration counts, 1e-6
parity) before and while the kernels are written, and so that kernel intermediates can be
diffed against something when debugging.  It keeps every batch-global coupling of the reference
loop (qpth/solvers/pdipm/batch.py:46-214): shared stall counter, max best-residual, min mu and
the get_step fill value.
"""
from __future__ import annotations

import math

import torch

NAN = float("nan")


def _chol_poison(M):
    """Batched Cholesky; a problem whose factorisation fails gets an all-NaN factor."""
    L, info = torch.linalg.cholesky_ex(M)
    bad = info > 0
    if bad.any():
        L = L.clone()
        L[bad] = NAN
    return L, bad


class GpuKKT:
    def __init__(self, Q, G, A, explicit_inverse=True):
        nb, m, n = G.shape
        p = A.shape[1]
        self.n, self.m, self.p, self.nb = n, m, p, nb
        self.G, self.A = G, A
        self.explicit = explicit_inverse
        self.LQ, self.q_bad = _chol_poison(Q)
        B = torch.cat([A, G], 1)                       # (nb, p+m, n)
        self.B = B
        if explicit_inverse:
            self.Qi = torch.cholesky_inverse(self.LQ)
            self.BQi = torch.bmm(B, self.Qi)
            M = torch.bmm(self.BQi, B.transpose(1, 2))
        else:
            W = torch.linalg.solve_triangular(self.LQ, B.transpose(1, 2), upper=False)  # (nb,n,p+m)
            M = torch.bmm(W.transpose(1, 2), W)
        M = 0.5 * (M + M.transpose(1, 2))
        if p > 0:
            self.LA, _ = _chol_poison(M[:, :p, :p])
            self.V = torch.linalg.solve_triangular(self.LA, M[:, :p, p:], upper=False)  # (nb,p,m)
            self.R = M[:, p:, p:] - torch.bmm(self.V.transpose(1, 2), self.V)
        else:
            self.R = M
        self.LT = None

    def refactor(self, d):
        T = self.R + torch.diag_embed(1.0 / d)
        self.LT, bad = _chol_poison(T)
        return bad

    def _qsolve(self, v):
        if self.explicit:
            return torch.bmm(self.Qi, v.unsqueeze(2)).squeeze(2)
        return torch.cholesky_solve(v.unsqueeze(2), self.LQ).squeeze(2)

    def solve(self, d, rx, rs, rz, ry):
        p, m = self.p, self.m
        t = self._qsolve(rx)
        if self.explicit:
            hB = torch.bmm(self.BQi, rx.unsqueeze(2)).squeeze(2)
        else:
            hB = torch.bmm(self.B, t.unsqueeze(2)).squeeze(2)
        hz = hB[:, p:] + rs / d - rz
        if p > 0:
            hy = hB[:, :p] - ry
            u = torch.linalg.solve_triangular(self.LA, hy.unsqueeze(2), upper=False)
            hz = hz - torch.bmm(self.V.transpose(1, 2), u).squeeze(2)
        qz = torch.cholesky_solve(hz.unsqueeze(2), self.LT)
        if p > 0:
            qy = torch.linalg.solve_triangular(self.LA.transpose(1, 2), u - torch.bmm(self.V, qz), upper=True)
            w = -torch.cat([qy, qz], 1).squeeze(2)
        else:
            w = -qz.squeeze(2)
        if self.explicit:
            dx = -t - torch.bmm(w.unsqueeze(1), self.BQi).squeeze(1)
        else:
            dx = self._qsolve(-rx - torch.bmm(w.unsqueeze(1), self.B).squeeze(1))
        wz = w[:, p:]
        ds = (-rs - wz) / d
        return dx, ds, wz, (w[:, :p] if p > 0 else None)


def _nanmax(t):
    """torch.Tensor.max() semantics (NaN propagates)."""
    return t.max()


def _row_unfilled_min(v, dv):
    """Per-row pieces of get_step: min over entries that are NOT overwritten by the fill
    (NaN-propagating, +inf when every entry is filled) and whether any entry is filled."""
    a = -v / dv
    filled = dv > 0
    rmu = torch.where(filled, torch.full_like(a, math.inf), a).min(1)[0]
    return a, filled, rmu


def _get_step(v, dv):
    a, filled, rmu = _row_unfilled_min(v, dv)
    gmax = _nanmax(a).item()
    fill = gmax if gmax > 1.0 else 1.0
    has = filled.any(1)
    return torch.where(has, torch.minimum(rmu, torch.full_like(rmu, fill)), rmu)


def pdipm_model(Q, p, G, h, A, b, eps=1e-12, notImprovedLim=3, maxIter=20, explicit_inverse=True, xspace=False):
    nb, m, n = G.shape
    neq = A.shape[1]
    kkt = XSpaceKKT(Q, G, A) if xspace else GpuKKT(Q, G, A, explicit_inverse)
    one = torch.ones(nb, m, dtype=Q.dtype)
    kkt.refactor(one)
    x, s, z, y = kkt.solve(one, p, torch.zeros_like(one), -h, -b if neq > 0 else None)
    lo = s.min(1, keepdim=True)[0]
    s = torch.where(lo < 0, s - (lo - 1), s)
    lo = z.min(1, keepdim=True)[0]
    z = torch.where(lo < 0, z - (lo - 1), z)

    best = None
    stall = 0
    n_iter = 0
    for it in range(maxIter):
        n_iter = it + 1
        rx = torch.bmm(z.unsqueeze(1), G).squeeze(1) + torch.bmm(Q, x.unsqueeze(2)).squeeze(2) + p
        if neq > 0:
            rx = rx + torch.bmm(y.unsqueeze(1), A).squeeze(1)
            ry = torch.bmm(A, x.unsqueeze(2)).squeeze(2) - b
        else:
            ry = None
        rz = torch.bmm(G, x.unsqueeze(2)).squeeze(2) + s - h
        mu = ((s * z).sum(1) / m).abs()
        pri = rz.norm(2, 1) + (ry.norm(2, 1) if neq > 0 else 0.0)
        resids = pri + rx.norm(2, 1) + m * mu
        d = z / s
        bad = kkt.refactor(d)
        if best is None:
            best = dict(resids=resids.clone(), x=x.clone(), s=s.clone(), z=z.clone(),
                        y=y.clone() if neq > 0 else None)
            stall = 0
        else:
            better = resids < best["resids"]
            stall = 0 if better.any() else stall + 1
            best["resids"] = torch.where(better, resids, best["resids"])
            for k, v in (("x", x), ("s", s), ("z", z)) + ((("y", y),) if neq > 0 else ()):
                best[k] = torch.where(better[:, None], v, best[k])
        if stall == notImprovedLim or best["resids"].max() < eps or mu.min() > 1e32:
            break
        dxa, dsa, dza, dya = kkt.solve(d, rx, z, rz, ry)
        alpha = torch.minimum(torch.minimum(_get_step(z, dza), _get_step(s, dsa)), torch.ones(nb, dtype=Q.dtype))
        t3 = ((s + alpha[:, None] * dsa) * (z + alpha[:, None] * dza)).sum(1)
        t4 = (s * z).sum(1)
        sig = (t3 / t4) ** 3
        rs_c = ((-mu * sig)[:, None] + dsa * dza) / s
        zero_n = torch.zeros(nb, n, dtype=Q.dtype)
        zero_m = torch.zeros(nb, m, dtype=Q.dtype)
        dxc, dsc, dzc, dyc = kkt.solve(d, zero_n, rs_c, zero_m, torch.zeros(nb, neq, dtype=Q.dtype) if neq > 0 else None)
        dx, ds, dz = dxa + dxc, dsa + dsc, dza + dzc
        dy = dya + dyc if neq > 0 else None
        alpha = torch.minimum(0.999 * torch.minimum(_get_step(z, dz), _get_step(s, ds)), torch.ones(nb, dtype=Q.dtype))
        x = x + alpha[:, None] * dx
        s = s + alpha[:, None] * ds
        z = z + alpha[:, None] * dz
        if neq > 0:
            y = y + alpha[:, None] * dy
    return dict(zhat=best["x"], lams=best["z"], slacks=best["s"],
                nus=best["y"] if neq > 0 else torch.zeros(nb, 0, dtype=Q.dtype), n_iter=n_iter, kkt=kkt,
                resids=best["resids"])


def backward_model(fwd, dl_dz):
    kkt: GpuKKT = fwd["kkt"]
    nb, m, p = kkt.nb, kkt.m, kkt.p
    zhat, lams, nus, slacks = fwd["zhat"], fwd["lams"], fwd["nus"], fwd["slacks"]
    d = torch.clamp(lams, min=1e-8) / torch.clamp(slacks, min=1e-8)
    kkt.refactor(d)
    zm = torch.zeros(nb, m, dtype=zhat.dtype)
    dx, _, dlam, dnu = kkt.solve(d, dl_dz, zm, zm, torch.zeros(nb, p, dtype=zhat.dtype) if p > 0 else None)
    outer = lambda u, v: u.unsqueeze(2) * v.unsqueeze(1)
    out = dict(dp=dx, dG=outer(dlam, zhat) + outer(lams, dx), dh=-dlam,
               dQ=0.5 * (outer(dx, zhat) + outer(zhat, dx)))
    if p > 0:
        out["dA"] = outer(dnu, zhat) + outer(nus, dx)
        out["db"] = -dnu
    return out


class XSpaceKKT:
    """Reduced (x-space / normal-equations) elimination: K = Q + G^T D G is n x n, so the only
    sequential work per iteration is an n x n Cholesky (plus a p x p one when neq > 0).
    Same KKT system as GpuKKT.solve, different elimination order."""

    def __init__(self, Q, G, A, explicit_inverse=True):
        nb, m, n = G.shape
        self.n, self.m, self.p, self.nb = n, m, A.shape[1], nb
        self.Q, self.G, self.A = Q, G, A
        self.q_bad = torch.linalg.cholesky_ex(Q)[1] > 0

    def refactor(self, d):
        G, A = self.G, self.A
        K = self.Q + torch.bmm(G.transpose(1, 2) * d.unsqueeze(1), G)
        self.LK, bad = _chol_poison(K)
        if self.p > 0:
            self.W = torch.linalg.solve_triangular(self.LK, A.transpose(1, 2), upper=False)  # (nb,n,p)
            Sy = torch.bmm(self.W.transpose(1, 2), self.W)
            self.LS, bad2 = _chol_poison(Sy)
            bad = bad | bad2
        return bad

    def solve(self, d, rx, rs, rz, ry):
        G = self.G
        r1 = -rx + torch.bmm((rs - d * rz).unsqueeze(1), G).squeeze(1)
        v = torch.linalg.solve_triangular(self.LK, r1.unsqueeze(2), upper=False)
        if self.p > 0:
            rhs = torch.bmm(self.W.transpose(1, 2), v) + ry.unsqueeze(2)
            dy = torch.cholesky_solve(rhs, self.LS)
            v = v - torch.bmm(self.W, dy)
        dx = torch.linalg.solve_triangular(self.LK.transpose(1, 2), v, upper=True).squeeze(2)
        ds = -rz - torch.bmm(G, dx.unsqueeze(2)).squeeze(2)
        dz = -rs - d * ds
        return dx, ds, dz, (dy.squeeze(2) if self.p > 0 else None)
