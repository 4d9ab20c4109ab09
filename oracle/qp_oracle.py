"""CPU oracle for the batched PDIPM QP path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

This file is a torch-CPU *restatement* of the reference algorithm (swami1995/diff-qp-mpc):

    qpth/solvers/pdipm/batch.py:377-428   pre_factor_kkt   -> BlockKKT.__init__
    qpth/solvers/pdipm/batch.py:434-469   factor_kkt       -> BlockKKT.refactor
    qpth/solvers/pdipm/batch.py:351-374   solve_kkt        -> BlockKKT.solve
    qpth/solvers/pdipm/batch.py:46-208    forward (loop)   -> pdipm_solve
    qpth/solvers/pdipm/batch.py:211-214   get_step         -> _boundary_step
    qpth/solvers/pdipm/batch.py:315-348   factor_solve_kkt -> full_kkt_solve
    qpth/qp.py:24-126,129-183             QPFunction fwd/bwd -> qp_forward / qp_backward

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
leg may import it; the product path (diff-qp-mpc_b200/b200qp) never does.

Parity pin: the numerics boundary of the reference is torch.linalg (LAPACK getrf/getrs on CPU).
This restatement issues the same torch.linalg calls in the same order, so on CPU it is
BIT-IDENTICAL to the reference; `oracle/gen_golden.py` asserts that against the real reference
imported from /root/reference (run in the build container) and commits the resulting vectors
under tests/golden/.  The reference holds no golden vectors of its own for this path
(SURVEY.md section 8c), so the reference run is the pin.
"""
from __future__ import annotations

import torch


def _lu(M):
    """Partial-pivot LU, the CPU branch of lu_hack (batch.py:8-19)."""
    return torch.linalg.lu_factor(M, pivot=True)


def _mv(M, v):
    """(nb, r, c) x (nb, c) -> (nb, r), written the way the reference writes v^T M^T."""
    return torch.bmm(v.unsqueeze(1), M.transpose(1, 2)).squeeze(1)


def _vm(v, M):
    """(nb, r) x (nb, r, c) -> (nb, c)  ==  M^T v."""
    return torch.bmm(v.unsqueeze(1), M).squeeze(1)


def _lu_solve(LU, rhs_vec):
    return torch.linalg.lu_solve(LU[0], LU[1], rhs_vec.unsqueeze(2)).squeeze(2)


class BlockKKT:
    """Block-LU Schur-complement KKT factorisation with the D-independent part cached.

    S = [[A Q^-1 A^T, A Q^-1 G^T], [G Q^-1 A^T, G Q^-1 G^T + D^-1]]; the lower-right block is
    completed by `refactor(d)` once d = z/s is known (batch.py:390-396).
    """

    def __init__(self, Q, G, A):
        nb, m, n = G.shape
        p = A.size(1) if A.nelement() > 0 else 0
        self.nb, self.n, self.m, self.p = nb, n, m, p
        self.G, self.A = G, A
        self.Q_LU = _lu(Q)
        GQG = torch.bmm(G, torch.linalg.lu_solve(self.Q_LU[0], self.Q_LU[1], G.transpose(1, 2)))
        R = GQG.clone()
        piv = torch.IntTensor(range(1, 1 + p + m)).unsqueeze(0).repeat(nb, 1).type_as(Q).int()
        if p > 0:
            QiAt = torch.linalg.lu_solve(self.Q_LU[0], self.Q_LU[1], A.transpose(1, 2))
            AQA = torch.bmm(A, QiAt)
            GQA = torch.bmm(G, QiAt)
            AQA_LU = _lu(AQA)
            P_, L_, U_ = torch.lu_unpack(*AQA_LU)
            P_ = P_.type_as(AQA)
            U_inv = torch.linalg.lu_solve(AQA_LU[0], AQA_LU[1], P_.bmm(L_))
            S21 = GQA.bmm(U_inv)
            T_ = torch.linalg.lu_solve(AQA_LU[0], AQA_LU[1], GQA.transpose(1, 2))
            S12 = U_.bmm(T_)
            S22 = torch.zeros(nb, m, m).type_as(Q)
            data = torch.cat((torch.cat((AQA_LU[0], S12), 2), torch.cat((S21, S22), 2)), 1)
            piv[:, :p] = AQA_LU[1]
            R -= GQA.bmm(T_)
        else:
            data = torch.zeros(nb, m, m).type_as(Q)
        self.S_LU = [data, piv]
        self.R = R
        self._eye = torch.eye(m).repeat(nb, 1, 1).type_as(R).bool()

    def refactor(self, d):
        """Complete the block LU for the current d (batch.py:434-469, CPU branch)."""
        m, p = self.m, self.p
        T = self.R.clone()
        T[self._eye] += (1.0 / d).squeeze().view(-1)
        T_LU = _lu(T)
        old_packed = self.S_LU[1][:, -m:] - p
        oldP, _, _ = torch.lu_unpack(T_LU[0], old_packed, unpack_data=False)
        new_packed = T_LU[1]
        newP, _, _ = torch.lu_unpack(T_LU[0], new_packed, unpack_data=False)
        if p > 0:
            S21 = self.S_LU[0][:, -m:, :p]
            self.S_LU[0][:, -m:, :p] = newP.transpose(1, 2).bmm(oldP.bmm(S21))
        self.S_LU[1][:, -m:] = new_packed + p
        self.S_LU[0][:, -m:, -m:] = T_LU[0]

    def solve(self, d, rx, rs, rz, ry):
        """Block elimination (batch.py:351-374)."""
        p, G, A = self.p, self.G, self.A
        iq = _lu_solve(self.Q_LU, rx)
        if p > 0:
            h = torch.cat((_mv(A, iq) - ry, _mv(G, iq) + rs / d - rz), 1)
        else:
            h = _mv(G, iq) + rs / d - rz
        w = -_lu_solve(self.S_LU, h)
        g1 = -rx - _vm(w[:, p:], G)
        if p > 0:
            g1 -= _vm(w[:, :p], A)
        g2 = -rs - w[:, p:]
        dx = _lu_solve(self.Q_LU, g1)
        ds = g2 / d
        dz = w[:, p:]
        dy = w[:, :p] if p > 0 else None
        return dx, ds, dz, dy


def _boundary_step(v, dv):
    """get_step (batch.py:211-214): the fill value is a max over the WHOLE batch tensor."""
    a = -v / dv
    a[dv > 0] = max(1.0, a.max())
    return a.min(1)[0].squeeze()


def pdipm_solve(Q, p, G, h, A, b, kkt: BlockKKT, eps=1e-12, notImprovedLim=3, maxIter=20,
                cost_grad=None, dyn_res=None, trace=None):
    """Mehrotra predictor-corrector loop with per-problem best-iterate tracking and
    batch-global termination (batch.py:46-208, KKTSolvers.LU_PARTIAL branch).

    Returns (x, y, z, s, n_iter) where n_iter is the number of loop bodies entered
    (the reference does not return it; the GPU kernels must reproduce it).
    """
    nb, m, n = G.shape
    neq = kkt.p
    if cost_grad is None:
        cost_grad = lambda x: _mv(Q, x) + p
    if dyn_res is None:
        dyn_res = (lambda x: _mv(A, x) - b) if neq > 0 else (lambda x: 0.0)

    d = torch.ones(nb, m).type_as(Q)
    kkt.refactor(d)
    x, s, z, y = kkt.solve(d, p, torch.zeros(nb, m).type_as(Q), -h, -b if neq > 0 else None)

    lo = torch.min(s, 1)[0].view(nb, 1).repeat(1, m)
    sel = lo < 0
    s[sel] -= lo[sel] - 1
    lo = torch.min(z, 1)[0].view(nb, 1).repeat(1, m)
    sel = lo < 0
    z[sel] -= lo[sel] - 1

    best = None
    stall = 0
    n_iter = 0
    for it in range(maxIter):
        n_iter = it + 1
        rx = (_vm(y, A) if neq > 0 else 0.0) + _vm(z, G) + cost_grad(x)
        rs = z
        rz = _mv(G, x) + s - h
        ry = dyn_res(x)
        mu = torch.abs((s * z).sum(1).squeeze() / m)
        z_res = torch.norm(rz, 2, 1).squeeze()
        y_res = torch.norm(ry, 2, 1).squeeze() if neq > 0 else 0
        pri = y_res + z_res
        dual = torch.norm(rx, 2, 1).squeeze()
        resids = pri + dual + m * mu
        if trace is not None:
            trace.append(dict(pri=pri.clone() if torch.is_tensor(pri) else pri, dual=dual.clone(),
                              mu=mu.clone(), resids=resids.clone()))

        d = z / s
        try:
            kkt.refactor(d)
        except Exception:
            return best['x'], best['y'], best['z'], best['s'], n_iter

        if best is None:
            best = dict(resids=resids, x=x.clone(), z=z.clone(), s=s.clone(),
                        y=y.clone() if y is not None else None)
            stall = 0
        else:
            better = resids < best['resids']
            stall = 0 if better.sum() > 0 else stall + 1
            bn = better.repeat(n, 1).t()
            bm = better.repeat(m, 1).t()
            best['resids'][better] = resids[better]
            best['x'][bn] = x[bn]
            best['z'][bm] = z[bm]
            best['s'][bm] = s[bm]
            if neq > 0:
                bp = better.repeat(neq, 1).t()
                best['y'][bp] = y[bp]
        if stall == notImprovedLim or best['resids'].max() < eps or mu.min() > 1e32:
            return best['x'], best['y'], best['z'], best['s'], n_iter

        dx_a, ds_a, dz_a, dy_a = kkt.solve(d, rx, rs, rz, ry)

        alpha = torch.min(torch.min(_boundary_step(z, dz_a), _boundary_step(s, ds_a)),
                          torch.ones(nb).type_as(Q))
        am = alpha.repeat(m, 1).t()
        t1 = s + am * ds_a
        t2 = z + am * dz_a
        t3 = torch.sum(t1 * t2, 1).squeeze()
        t4 = torch.sum(s * z, 1).squeeze()
        sig = (t3 / t4) ** 3

        rx0 = torch.zeros(nb, n).type_as(Q)
        rs_c = ((-mu * sig).repeat(m, 1).t() + ds_a * dz_a) / s
        rz0 = torch.zeros(nb, m).type_as(Q)
        ry0 = torch.zeros(nb, neq).type_as(Q) if neq > 0 else torch.Tensor()
        dx_c, ds_c, dz_c, dy_c = kkt.solve(d, rx0, rs_c, rz0, ry0)

        dx = dx_a + dx_c
        ds = ds_a + ds_c
        dz = dz_a + dz_c
        dy = dy_a + dy_c if neq > 0 else None
        alpha = torch.min(0.999 * torch.min(_boundary_step(z, dz), _boundary_step(s, ds)),
                          torch.ones(nb).type_as(Q))
        x += alpha.repeat(n, 1).t() * dx
        s += alpha.repeat(m, 1).t() * ds
        z += alpha.repeat(m, 1).t() * dz
        y = y + alpha.repeat(neq, 1).t() * dy if neq > 0 else None

    return best['x'], best['y'], best['z'], best['s'], n_iter


def _expand(X, nb, ndim):
    """expandParam, second (winning) definition (util.py:69-75)."""
    if X.ndimension() in (0, ndim):
        return X, False
    if X.ndimension() == ndim - 1:
        return X.unsqueeze(0).expand(*([nb] + list(X.size()))), True
    raise RuntimeError("Unexpected number of dimensions.")


def _nbatch(Q, p, G, h, A, b):
    for t, dmn in zip((Q, p, G, h, A, b), (3, 2, 3, 2, 3, 2)):
        if t.ndimension() == dmn:
            return t.size(0)
    return 1


def qp_forward(Q_, p_, G_, h_, A_, b_, eps=1e-12, notImprovedLim=3, maxIter=20,
               cost_grad=None, dyn_res=None):
    """QPFunctionFn.forward (qp.py:24-126) without the autograd wrapper.

    Returns a dict with zhat, nus, lams, slacks, n_iter and the cached factorisation.
    """
    nb = _nbatch(Q_, p_, G_, h_, A_, b_)
    Q, _ = _expand(Q_, nb, 3)
    p, _ = _expand(p_, nb, 2)
    G, _ = _expand(G_, nb, 3)
    h, _ = _expand(h_, nb, 2)
    A, _ = _expand(A_, nb, 3)
    b, _ = _expand(b_, nb, 2)
    kkt = BlockKKT(Q, G, A)
    x, y, z, s, n_iter = pdipm_solve(Q, p, G, h, A, b, kkt, eps, notImprovedLim, maxIter,
                                     cost_grad, dyn_res)
    return dict(zhat=x, nus=y, lams=z, slacks=s, n_iter=n_iter, kkt=kkt)


def qp_backward(fwd, Q_, p_, G_, h_, A_, b_, dl_dzhat):
    """QPFunctionFn.backward (qp.py:129-183): adjoint KKT solve + outer-product gradients."""
    nb = _nbatch(Q_, p_, G_, h_, A_, b_)
    Q, Q_e = _expand(Q_, nb, 3)
    p, p_e = _expand(p_, nb, 2)
    G, G_e = _expand(G_, nb, 3)
    h, h_e = _expand(h_, nb, 2)
    A, A_e = _expand(A_, nb, 3)
    b, b_e = _expand(b_, nb, 2)
    kkt: BlockKKT = fwd['kkt']
    m, neq = kkt.m, kkt.p
    zhat, lams, nus, slacks = fwd['zhat'], fwd['lams'], fwd['nus'], fwd['slacks']
    d = torch.clamp(lams, min=1e-8) / torch.clamp(slacks, min=1e-8)
    kkt.refactor(d)
    dx, _, dlam, dnu = kkt.solve(d, dl_dzhat, torch.zeros(nb, m).type_as(G),
                                 torch.zeros(nb, m).type_as(G),
                                 torch.zeros(nb, neq).type_as(G) if neq > 0 else torch.Tensor())
    outer = lambda u, v: u.unsqueeze(2).bmm(v.unsqueeze(1))
    dps = dx
    dGs = outer(dlam, zhat) + outer(lams, dx)
    dhs = -dlam
    if G_e:
        dGs = dGs.mean(0)
    if h_e:
        dhs = dhs.mean(0)
    if neq > 0:
        dAs = outer(dnu, zhat) + outer(nus, dx)
        dbs = -dnu
        if A_e:
            dAs = dAs.mean(0)
        if b_e:
            dbs = dbs.mean(0)
    else:
        dAs, dbs = None, None
    dQs = 0.5 * (outer(dx, zhat) + outer(zhat, dx))
    if Q_e:
        dQs = dQs.mean(0)
    if p_e:
        dps = dps.mean(0)
    return dict(dQ=dQs, dp=dps, dG=dGs, dh=dhs, dA=dAs, db=dbs)


def full_kkt_solve(Q, D, G, A, rx, rs, rz, ry):
    """factor_solve_kkt (batch.py:315-348): one-shot LU of H=diag(Q,D) and its Schur complement.
    Used by the KKT self-consistency tests (reference test.py:222-247)."""
    nb, m, n = G.shape
    neq = A.size(1) if A.nelement() > 0 else 0
    H = torch.zeros(nb, n + m, n + m).type_as(Q)
    H[:, :n, :n] = Q
    H[:, -m:, -m:] = D
    eye = torch.eye(m).type_as(Q).repeat(nb, 1, 1)
    if neq > 0:
        A_ = torch.cat([torch.cat([G, eye], 2),
                        torch.cat([A, torch.zeros(nb, neq, m).type_as(Q)], 2)], 1)
        g_ = torch.cat([rx, rs], 1)
        h_ = torch.cat([rz, ry], 1)
    else:
        A_ = torch.cat([G, eye], 2)
        g_ = torch.cat([rx, rs], 1)
        h_ = rz
    H_LU = _lu(H)
    HiAt = torch.linalg.lu_solve(H_LU[0], H_LU[1], A_.transpose(1, 2))
    Hig = _lu_solve(H_LU, g_)
    S_LU = _lu(torch.bmm(A_, HiAt))
    t_ = _mv(A_, Hig) - h_
    w_ = -_lu_solve(S_LU, t_)
    t_ = -g_ - _vm(w_, A_)
    v_ = _lu_solve(H_LU, t_)
    return v_[:, :n], v_[:, n:], w_[:, :m], (w_[:, m:] if neq > 0 else None)


def random_qp(nb, nz, nineq, neq=0, seed=0, dtype=torch.float64, well_conditioned=False):
    """Random strictly-convex QP batch, the generator of prof-linear.py:64-75 (npr.seed(seed)):
    L~U[0,1), Q=LL^T+1e-3 I, G~N, z0~N, s0~U, p~N, h=G z0+s0, A~N, b=A z0.
    `well_conditioned` draws L~N(0,1) instead (SURVEY.md section 8d secondary variant)."""
    import numpy as np
    rs = np.random.RandomState(seed)
    L = rs.randn(nb, nz, nz) if well_conditioned else rs.rand(nb, nz, nz)
    Q = np.matmul(L, L.transpose((0, 2, 1))) + 1e-3 * np.eye(nz, nz)
    G = rs.randn(nb, nineq, nz)
    z0 = rs.randn(nb, nz)
    s0 = rs.rand(nb, nineq)
    p = rs.randn(nb, nz)
    h = np.matmul(G, z0[:, :, None])[:, :, 0] + s0
    A = rs.randn(nb, neq, nz)
    b = np.matmul(A, z0[:, :, None])[:, :, 0]
    return tuple(torch.tensor(a, dtype=dtype) for a in (Q, p, G, h, A, b))


# ------------------------------------------------------------------------------------------------
# DenseQPFunction: full-KKT LU variant (qpth/qp.py:187-271, qpth/solvers/pdipm/batch_LU.py)
# ------------------------------------------------------------------------------------------------
DENSE_KKT_EPS = 1e-7  # batch_LU.py:40


def dense_kkt_matrix(Q, G, A):
    """K of qp.py:195-215: unknowns [x (n), s (m), z (m), y (p)], the s-row is [0, diag(z), diag(s), 0]
    (its diagonals are rewritten every iteration)."""
    nb, m, n = G.shape
    p = A.shape[1]
    Z = lambda r, c: torch.zeros(nb, r, c, dtype=Q.dtype)
    I = torch.eye(m, dtype=Q.dtype)[None].repeat(nb, 1, 1)
    GT, AT = G.transpose(1, 2), A.transpose(1, 2)
    K1 = torch.cat([Q, GT * 0, GT, AT], dim=-1)
    K2 = torch.cat([G * 0, I, I, Z(m, p)], dim=-1)
    K3 = torch.cat([G, I, Z(m, m), Z(m, p)], dim=-1)
    K4 = torch.cat([A, Z(p, m), Z(p, m), Z(p, p)], dim=-1)
    return torch.cat([K1, K2, K3, K4], dim=-2)


def dense_solve_kkt(K, Ktilde, rx, rs, rz, ry, niter=1):
    """batch_LU.py:212-244: LU of the regularised matrix + one step of iterative refinement."""
    n, m, p = rx.size(1), rz.size(1), ry.size(1)
    r = -torch.cat((rx, rs, rz, ry), 1)
    K_LU = torch.linalg.lu_factor(Ktilde)
    l = torch.linalg.lu_solve(*K_LU, r.clone().unsqueeze(-1)).squeeze(-1)
    res = r - K.bmm(l.unsqueeze(-1)).squeeze(-1)
    for _ in range(niter):
        d = torch.linalg.lu_solve(*K_LU, res.clone().unsqueeze(-1)).squeeze(-1)
        l = l + d
        res = r - K.bmm(l.unsqueeze(-1)).squeeze(-1)
    return l[:, :n], l[:, n:n + m], l[:, n + m:n + 2 * m], l[:, n + 2 * m:n + 2 * m + p]


def _dense_step(v, dv):
    """batch_LU.py:204-210 (get_step with the dv == 0 -> 1 rule)."""
    a = -v / dv
    a[dv == 0] = 1.0
    a[dv > 0] = max(1.0, a.max())
    return a.min(1)[0].squeeze()


def dense_forward(Q, p, G, h, A, b, eps=1e-12, notImprovedLim=3, maxIter=20, cost_grad=None, dyn_res=None):
    """DenseQPFunction forward: batch_LU.forward (batch_LU.py:29-201).  `dyn_res` / `cost_grad` are this fork's
    residual callbacks (batch_LU.py:88-97); None = the canonical Ax - b / Qx + p.  Returns best iterates and the best K."""
    nb, m, n = G.shape
    neq = A.shape[1]
    K = dense_kkt_matrix(Q, G, A)
    zi = torch.arange(n, n + m)
    si_r, si_c = torch.arange(n, n + m), torch.arange(n + m, n + 2 * m)
    K[:, si_r, si_c] = torch.ones(nb, m, dtype=Q.dtype)
    K[:, zi, zi] = torch.ones(nb, m, dtype=Q.dtype)
    I = torch.ones(nb, K.shape[1], dtype=Q.dtype)
    I[:, n + m:] *= -1
    Ktilde = K + DENSE_KKT_EPS * torch.diag_embed(I)
    x, s, z, y = dense_solve_kkt(K, Ktilde, p, torch.zeros(nb, m, dtype=Q.dtype), -h, -b)
    x, s, z, y = x.clone(), s.clone(), z.clone(), y.clone()
    M = torch.min(s, 1)[0][:, None].repeat(1, m)
    sel = M < 0
    s[sel] -= M[sel] - 1
    M = torch.min(z, 1)[0][:, None].repeat(1, m)
    sel = M < 0
    z[sel] -= M[sel] - 1
    best = None
    stall = 0
    n_iter = 0
    GT, AT = G.transpose(1, 2), A.transpose(1, 2)
    for i in range(maxIter):
        n_iter = i + 1
        if cost_grad is None:
            rx = ((AT.bmm(y.unsqueeze(-1)).squeeze(-1) if neq > 0 else 0.) + GT.bmm(z.unsqueeze(-1)).squeeze(-1)
                  + Q.bmm(x.unsqueeze(-1)).squeeze(-1) + p)
        else:
            rx = ((AT.bmm(y.unsqueeze(-1)).squeeze(-1) if neq > 0 else 0.) + GT.bmm(z.unsqueeze(-1)).squeeze(-1)
                  + cost_grad(x))
        rs = s * z
        rz = G.bmm(x.unsqueeze(-1)).squeeze(-1) + s - h
        ry = dyn_res(x) if dyn_res is not None else A.bmm(x.unsqueeze(-1)).squeeze(-1) - b
        mu = torch.abs((s * z).sum(1).squeeze() / m)
        z_resid = torch.norm(rz, 2, 1).squeeze()
        y_resid = torch.norm(ry, 2, 1).squeeze() if neq > 0 else 0
        resids = y_resid + z_resid + torch.norm(rx, 2, 1).squeeze() + m * mu
        K[:, zi, zi] = z
        K[:, si_r, si_c] = s
        Ktilde[:, zi, zi] = z + DENSE_KKT_EPS
        Ktilde[:, si_r, si_c] = s
        if best is None:
            best = dict(resids=resids, x=x.clone(), z=z.clone(), s=s.clone(), y=y.clone(), K=K.clone())
            stall = 0
        else:
            sel = resids < best["resids"]
            stall = 0 if sel.sum() > 0 else stall + 1
            best["resids"][sel] = resids[sel]
            best["x"][sel] = x[sel]
            best["z"][sel] = z[sel]
            best["s"][sel] = s[sel]
            best["K"][sel] = K[sel]
            if neq > 0:
                best["y"][sel] = y[sel]
        if stall == notImprovedLim or best["resids"].max() < eps or mu.min() > 1e32:
            break
        dx_a, ds_a, dz_a, dy_a = dense_solve_kkt(K, Ktilde, rx, rs, rz, ry)
        alpha = torch.min(torch.min(_dense_step(z, dz_a), _dense_step(s, ds_a)), torch.ones(nb, dtype=Q.dtype))
        an = alpha.repeat(m, 1).t()
        t3 = torch.sum((s + an * ds_a) * (z + an * dz_a), 1).squeeze()
        t4 = torch.sum(s * z, 1).squeeze()
        sig = (t3 / t4) ** 3
        rs_c = (-mu * sig).repeat(m, 1).t() + ds_a * dz_a
        dx_c, ds_c, dz_c, dy_c = dense_solve_kkt(K, Ktilde, torch.zeros(nb, n, dtype=Q.dtype), rs_c,
                                                 torch.zeros(nb, m, dtype=Q.dtype), torch.zeros(nb, neq, dtype=Q.dtype))
        dx, ds, dz = dx_a + dx_c, ds_a + ds_c, dz_a + dz_c
        dy = dy_a + dy_c if neq > 0 else None
        alpha = torch.min(0.999 * torch.min(_dense_step(z, dz), _dense_step(s, ds)), torch.ones(nb, dtype=Q.dtype))
        x = x + alpha.repeat(n, 1).t() * dx
        s = s + alpha.repeat(m, 1).t() * ds
        z = z + alpha.repeat(m, 1).t() * dz
        if neq > 0:
            y = y + alpha.repeat(neq, 1).t() * dy
    return dict(zhat=best["x"], nus=best["y"], lams=best["z"], slacks=best["s"], K=best["K"], n_iter=n_iter)


def dense_backward(fwd, dl_dzhat):
    """qp.py:235-268: adjoint solve with the (unregularised) best K, outer-product gradients."""
    K = fwd["K"]
    nb = K.shape[0]
    zhat, lams, nus = fwd["zhat"], fwd["lams"], fwd["nus"]
    m, neq = lams.shape[1], nus.shape[1]
    zm = torch.zeros(nb, m, dtype=K.dtype)
    dx, _, dlam, dnu = dense_solve_kkt(K, K, dl_dzhat, zm, zm, torch.zeros(nb, neq, dtype=K.dtype))
    outer = lambda u, v: u.unsqueeze(2) * v.unsqueeze(1)
    return dict(dQ=0.5 * (outer(dx, zhat) + outer(zhat, dx)), dp=dx, dG=outer(dlam, zhat) + outer(lams, dx), dh=-dlam,
                dA=outer(dnu, zhat) + outer(nus, dx), db=-dnu)
