"""QPFunction -- drop-in for the reference's qpth/qp.py:19-184 on B200.

    QPFunction(eps, verbose, notImprovedLim, maxIter, solver, check_Q_spd)(Q, p, G, h, A, b
                                                                         [, dyn_res, cost_grad])

Same factory signature, same autograd contract (gradients for the six tensors, None for the two
callables, `.mean(0)` for parameters shared across the batch), same error strings.  The whole
forward (pre-factorisation + Mehrotra loop with the reference's batch-global termination) and
the adjoint backward run in hand-written sm_100a kernels behind the C ABI of include/b200qp.h;
this file only unwraps tensors.  There is no CPU path: non-CUDA inputs raise.
"""
from __future__ import annotations

import ctypes
from enum import Enum

import torch
from torch.autograd import Function

from . import _lib
from .util import expandParam, extract_nBatch

INACC_ERR = """
--------
qpth warning: Returning an inaccurate and potentially incorrect solution.

Some residual is large.
Your problem may be infeasible or difficult.

You can try using the CVXPY solver to see if your problem is feasible
and you can use the verbose option to check the convergence status of
our solver while increasing the number of iterations.

Advanced users:
You can also try to enable iterative refinement in the solver:
https://github.com/locuslab/qpth/issues/6
--------
"""


class QPSolvers(Enum):
    """qpth/qp.py:14-16"""
    PDIPM_BATCHED = 1
    CVXPY = 2


def _dtype_code(t):
    if t.dtype == torch.float64:
        return _lib.F64
    if t.dtype == torch.float32:
        return _lib.F32
    raise RuntimeError(f"b200qp: unsupported dtype {t.dtype} (float64 or float32)")


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else ctypes.c_void_p(0)


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _prep(X, nBatch, nDim, ref):
    """Contiguous tensor + batch stride (0 when the parameter is shared across the batch)."""
    if X.device != ref.device or X.dtype != ref.dtype:
        raise RuntimeError("b200qp: all QP parameters must share Q's device and dtype")
    per = 1
    for s in X.shape[-(nDim - 1):]:
        per *= s
    if X.ndimension() == nDim:
        if X.size(0) != nBatch:
            raise RuntimeError("b200qp: inconsistent batch sizes")
        return X.contiguous(), per, False
    return X.contiguous(), 0, True


class _Plan:
    """Sizes + the C problem descriptor + workspace for one call."""

    def __init__(self, Q_, p_, G_, h_, A_, b_, eps, notImprovedLim, maxIter, flags=0, kkt_reg=0.0):
        if not Q_.is_cuda:
            raise RuntimeError("b200qp.QPFunction runs on CUDA tensors only (no CPU fallback); got " + str(Q_.device))
        nBatch = extract_nBatch(Q_, p_, G_, h_, A_, b_)
        neq_zero = A_.nelement() == 0
        # shape checks exactly as the reference's expandParam would raise them
        for X, d in ((Q_, 3), (p_, 2), (G_, 3), (h_, 2)) + (() if neq_zero else ((A_, 3), (b_, 2))):
            expandParam(X, nBatch, d)
        nineq, nz = G_.shape[-2], G_.shape[-1]
        neq = 0 if neq_zero else A_.shape[-2]
        assert neq > 0 or nineq > 0
        if nineq == 0:
            raise RuntimeError("b200qp: nineq == 0 is not supported (the reference's loop divides by nineq)")
        self.nBatch, self.nz, self.nineq, self.neq = nBatch, nz, nineq, neq
        self.Q, sQ, self.Q_e = _prep(Q_, nBatch, 3, Q_)
        self.p, sp, self.p_e = _prep(p_, nBatch, 2, Q_)
        self.G, sG, self.G_e = _prep(G_, nBatch, 3, Q_)
        self.h, sh, self.h_e = _prep(h_, nBatch, 2, Q_)
        if neq > 0:
            self.A, sA, self.A_e = _prep(A_, nBatch, 3, Q_)
            self.b, sb, self.b_e = _prep(b_, nBatch, 2, Q_)
        else:
            self.A, sA, self.A_e, self.b, sb, self.b_e = None, 0, False, None, 0, False
        if maxIter > _lib.MAX_ITER_CAP:
            raise RuntimeError(f"b200qp: maxIter > {_lib.MAX_ITER_CAP} is not supported")
        self.prob = _lib.Problem(nBatch, nz, nineq, neq, _dtype_code(Q_), int(maxIter), int(notImprovedLim), int(flags),
                                 float(eps), sQ, sp, sG, sh, sA, sb, float(kkt_reg))
        L = _lib.lib()
        nbytes = L.b200qp_workspace_bytes(ctypes.byref(self.prob))
        if nbytes == 0:
            raise RuntimeError("b200qp: unsupported problem size")
        self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=Q_.device)


_I64_MIN = -(1 << 63)


def _forward_exact_sharded(L, plan, bufs, status, device, group):
    """The forward with the batch sharded over the ranks of `group`, bit-for-bit what the reference
    does on the UNSHARDED batch: its termination test and get_step fill are whole-batch reductions
    (qpth/solvers/pdipm/batch.py:127-131,141,213), here one 64-byte slot per iteration whose fields
    are unsigned MAX reductions -- all-reduced across the ranks between iteration launches."""
    import torch.distributed as dist
    args = [_ptr(plan.Q), _ptr(plan.p), _ptr(plan.G), _ptr(plan.h), _ptr(plan.A), _ptr(plan.b)] + [_ptr(t) for t in bufs] + \
           [_ptr(plan.workspace), _ptr(status), _stream(device)]
    pr = ctypes.byref(plan.prob)
    _lib.check(L.b200qp_forward_phase(pr, _lib.PHASE_BEGIN, *args), "b200qp_forward_phase(begin)")
    off = L.b200qp_slot_offset(pr)
    max_iter = plan.prob.max_iter
    slots = plan.workspace[off: off + 64 * max_iter].view(max_iter, 64)
    for it in range(max_iter):
        _lib.check(L.b200qp_forward_phase(pr, it, *args), "b200qp_forward_phase(iter)")
        keys = slots[it, :32].view(torch.int64)      # best_max, ~mu_min, amax_z, amax_s: unsigned order
        flags = slots[it, 32:].view(torch.int32)     # improved + NaN flags (0/1)
        red = torch.cat((keys ^ _I64_MIN, flags.to(torch.int64)))  # signed order == unsigned order after the flip
        dist.all_reduce(red, op=dist.ReduceOp.MAX, group=group)
        keys.copy_(red[:4] ^ _I64_MIN)
        flags.copy_(red[4:].to(torch.int32))
    _lib.check(L.b200qp_forward_phase(pr, _lib.PHASE_END, *args), "b200qp_forward_phase(end)")
    # factorisation-failure counts of ALL ranks: every rank must raise (or not) together, otherwise the ranks that carry
    # on hang in the next collective (the shared-parameter mean of backward)
    fails = status[_lib.ST_Q_FAIL:_lib.ST_AQA_FAIL + 1]
    dist.all_reduce(fails, op=dist.ReduceOp.SUM, group=group)


def _global_mean(local_mean, local_nb, group):
    """Mean over the GLOBAL batch from per-rank means (shared parameters, qpth/qp.py:160-178)."""
    import torch.distributed as dist
    buf = torch.cat([(local_mean * float(local_nb)).reshape(-1),
                     torch.tensor([float(local_nb)], dtype=local_mean.dtype, device=local_mean.device)])
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return (buf[:-1] / buf[-1]).reshape(local_mean.shape)


def QPFunction(eps=1e-12, verbose=0, notImprovedLim=3,
               maxIter=20, solver=QPSolvers.PDIPM_BATCHED,
               check_Q_spd=True, process_group=None):
    """Factory with the reference's signature (qpth/qp.py:19-21).

    `process_group` (extension): when given, the batch handed to each rank is treated as a shard of
    ONE global batch and the solve reproduces the reference on that global batch exactly (the
    batch-global termination test and step fill are all-reduced once per iteration, 96 bytes);
    without it every rank solves its shard like an independent reference call."""
    info = {}
    exact_group = process_group

    class QPFunctionFn(Function):
        @staticmethod
        def forward(ctx, Q_, p_, G_, h_, A_, b_, dyn_res=None, cost_grad=None):
            if solver != QPSolvers.PDIPM_BATCHED:
                raise NotImplementedError("b200qp: only QPSolvers.PDIPM_BATCHED is implemented (CVXPY is a "
                                          "per-instance CPU solver outside the hot path)")
            plan = _Plan(Q_, p_, G_, h_, A_, b_, eps, notImprovedLim, maxIter)
            if check_Q_spd:
                # qpth/qp.py:82-86 raises when some eigenvalue of Q has a non-positive real part.  For a symmetric Q that
                # is exactly "the LDL^T of the pre-factorisation meets a non-positive pivot" (checked below from the
                # status block, no extra work).  A NON-symmetric Q can have positive pivots and still fail the
                # reference's test, so it gets the reference's own criterion (batched eigvals on the device).
                Qc = plan.Q
                if not torch.equal(Qc, Qc.transpose(-1, -2)):
                    ev = torch.linalg.eigvals(Qc)
                    if not bool(torch.all(ev.real > 0)):
                        raise RuntimeError('Q is not SPD.')
            cb_cg, cb_ry = _classify_callbacks(plan, dyn_res, cost_grad)
            with_cb = cb_cg is not None or cb_ry is not None
            if with_cb and exact_group is not None:
                raise NotImplementedError("b200qp: non-canonical dyn_res / cost_grad callbacks together with process_group")
            if with_cb:
                plan.prob.flags |= _lib.FLAG_EXACT  # one launch per iteration, forward and backward on the same kernels
            if exact_group is not None:
                # the sharded mode IS the one-launch-per-iteration route; the flag keeps the forward and the backward of this
                # problem on the same kernels (bit-for-bit equal to the unsharded exact route)
                plan.prob.flags |= _lib.FLAG_EXACT
            L = _lib.lib()
            nb, nz, nineq, neq = plan.nBatch, plan.nz, plan.nineq, plan.neq
            opt = dict(dtype=Q_.dtype, device=Q_.device)
            zhats = torch.empty(nb, nz, **opt)
            lams = torch.empty(nb, nineq, **opt)
            slacks = torch.empty(nb, nineq, **opt)
            nus = torch.empty(nb, neq, **opt)
            status = torch.empty(_lib.STATUS_DOUBLES, dtype=torch.float64, device=Q_.device)
            with torch.cuda.device(Q_.device):
                if exact_group is not None:
                    _forward_exact_sharded(L, plan, (zhats, lams, nus, slacks), status, Q_.device, exact_group)
                    rc = 0
                elif with_cb:
                    _forward_callbacks(L, plan, (zhats, lams, nus, slacks), status, Q_.device, cb_cg, cb_ry)
                    rc = 0
                else:
                    rc = L.b200qp_forward(ctypes.byref(plan.prob), _ptr(plan.Q), _ptr(plan.p), _ptr(plan.G), _ptr(plan.h),
                                          _ptr(plan.A), _ptr(plan.b), _ptr(zhats), _ptr(lams), _ptr(nus), _ptr(slacks),
                                          _ptr(plan.workspace), _ptr(status), _stream(Q_.device))
            _lib.check(rc, "b200qp_forward")
            st = status.tolist()  # one small D2H read; also surfaces asynchronous kernel faults
            info["exact_rerun"] = False
            if st[_lib.ST_SPEC_FAIL] != 0 and exact_group is None:
                # the resident route met a step-fill situation it never speculates (include/b200qp.h,
                # B200QP_FLAG_EXACT): same problem, same workspace, one launch per iteration
                plan.prob.flags |= _lib.FLAG_EXACT
                with torch.cuda.device(Q_.device):
                    rc = L.b200qp_forward(ctypes.byref(plan.prob), _ptr(plan.Q), _ptr(plan.p), _ptr(plan.G), _ptr(plan.h),
                                          _ptr(plan.A), _ptr(plan.b), _ptr(zhats), _ptr(lams), _ptr(nus), _ptr(slacks),
                                          _ptr(plan.workspace), _ptr(status), _stream(Q_.device))
                _lib.check(rc, "b200qp_forward (exact route)")
                st = status.tolist()
                info["exact_rerun"] = True
            info.update(n_iter=int(st[_lib.ST_NITER]), best_resid_max=st[_lib.ST_BEST_MAX],
                        launches=int(st[_lib.ST_LAUNCHES]), nan_onset=int(st[_lib.ST_NAN_ONSET]))
            if st[_lib.ST_Q_FAIL] > 0:
                if check_Q_spd:
                    raise RuntimeError('Q is not SPD.')
                raise RuntimeError("qpth Error: Cannot perform LU factorization on Q. "
                                   "Please make sure that your Q matrix is PSD and has a non-zero diagonal.")
            if st[_lib.ST_AQA_FAIL] > 0:
                raise RuntimeError("qpth Error: Cannot perform LU factorization on AQ^{-1}A^T. "
                                   "Please make sure that your A matrix is full rank.")
            if st[_lib.ST_BEST_MAX] > 1. and verbose >= 0:
                print(INACC_ERR)
            ctx.plan = plan
            ctx.neq, ctx.nineq, ctx.nz = neq, nineq, nz
            ctx.nus, ctx.lams, ctx.slacks = nus, lams, slacks
            ctx.save_for_backward(zhats, Q_, p_, G_, h_, A_, b_)
            return zhats

        @staticmethod
        def backward(ctx, dl_dzhat):
            zhats = ctx.saved_tensors[0]
            plan = ctx.plan
            L = _lib.lib()
            nb, nz, nineq, neq = plan.nBatch, plan.nz, plan.nineq, plan.neq
            opt = dict(dtype=zhats.dtype, device=zhats.device)
            gz = dl_dzhat.contiguous()
            dQ = torch.empty(nb, nz, nz, **opt)
            dp = torch.empty(nb, nz, **opt)
            dG = torch.empty(nb, nineq, nz, **opt)
            dh = torch.empty(nb, nineq, **opt)
            dA = torch.empty(nb, neq, nz, **opt) if neq > 0 else None
            db = torch.empty(nb, neq, **opt) if neq > 0 else None
            with torch.cuda.device(zhats.device):
                rc = L.b200qp_backward(ctypes.byref(plan.prob), _ptr(zhats), _ptr(ctx.lams), _ptr(ctx.nus),
                                       _ptr(ctx.slacks), _ptr(gz), _ptr(dQ), _ptr(dp), _ptr(dG), _ptr(dh), _ptr(dA),
                                       _ptr(db), _ptr(plan.workspace), _stream(zhats.device))
            _lib.check(rc, "b200qp_backward")
            # parameters shared across the batch get the MEAN over it (qpth/qp.py:160-178); over the
            # GLOBAL batch when the call is a shard of one (process_group)
            mean0 = (lambda t: t.mean(0)) if exact_group is None else (lambda t: _global_mean(t.mean(0), nb, exact_group))
            if plan.Q_e:
                dQ = mean0(dQ)
            if plan.p_e:
                dp = mean0(dp)
            if plan.G_e:
                dG = mean0(dG)
            if plan.h_e:
                dh = mean0(dh)
            if neq > 0:
                if plan.A_e:
                    dA = mean0(dA)
                if plan.b_e:
                    db = mean0(db)
            return (dQ, dp, dG, dh, dA, db, None, None)

    def apply(Q, p, G, h, A, b, dyn_res=None, cost_grad=None):
        return QPFunctionFn.apply(Q, p, G, h, A, b, dyn_res, cost_grad)

    apply.info = info
    apply.Function = QPFunctionFn
    return apply


DENSE_KKT_EPS = 1e-7  # qpth/solvers/pdipm/batch_LU.py:40


def DenseQPFunction(bsz=1, eps=1e-12, verbose=0, notImprovedLim=3, maxIter=20):
    """Drop-in for the reference's full-KKT variant (qpth/qp.py:187-271 + qpth/solvers/pdipm/batch_LU.py):

        DenseQPFunction(...)(Q, p, G, h, A, b, dyn_res[, cost_grad]) -> zhat

    Same iteration as the reference -- every KKT solve is a solve with K + 1e-7 diag(+I,+I,-I,-I)
    followed by one refinement step against K, unscaled complementarity residual, the dv == 0 rule
    of its get_step, backward with the best iterate's K without clamping -- but the 2 x LU of the
    (nz + 2 nineq + neq)^2 matrix per solve is replaced by the same Schur-complement kernels as
    QPFunction (b200qp_forward with B200QP_FLAG_DENSE), which solve the identical regularised system.
    All six parameters are batched (as in the reference).  `dyn_res` / `cost_grad` that are the canonical
    Ax - b / Qx + p (checked on a probe point) stay inside the fused kernels; any other callback -- the reference's
    MPC callers pass the non-linear dynamics residual -- is evaluated between launches (`_forward_callbacks`)."""
    info = {}

    class Solver(Function):
        @staticmethod
        def forward(ctx, Q, p, G, h, A, b, dyn_res=None, cost_grad=None):
            if Q.dim() != 3 or G.dim() != 3 or A.dim() != 3:
                raise RuntimeError("b200qp.DenseQPFunction: batched (3-D) Q, G, A are required, as in the reference")
            plan = _Plan(Q, p, G, h, A, b, eps, notImprovedLim, maxIter, flags=_lib.FLAG_DENSE, kkt_reg=DENSE_KKT_EPS)
            cb_cg, cb_ry = _classify_callbacks(plan, dyn_res, cost_grad)
            L = _lib.lib()
            nb, nz, nineq, neq = plan.nBatch, plan.nz, plan.nineq, plan.neq
            opt = dict(dtype=Q.dtype, device=Q.device)
            zhats, lams, slacks = torch.empty(nb, nz, **opt), torch.empty(nb, nineq, **opt), torch.empty(nb, nineq, **opt)
            nus = torch.empty(nb, neq, **opt)
            status = torch.empty(_lib.STATUS_DOUBLES, dtype=torch.float64, device=Q.device)
            with torch.cuda.device(Q.device):
                if cb_cg is not None or cb_ry is not None:
                    _forward_callbacks(L, plan, (zhats, lams, nus, slacks), status, Q.device, cb_cg, cb_ry)
                    rc = 0
                else:
                    rc = L.b200qp_forward(ctypes.byref(plan.prob), _ptr(plan.Q), _ptr(plan.p), _ptr(plan.G), _ptr(plan.h),
                                          _ptr(plan.A), _ptr(plan.b), _ptr(zhats), _ptr(lams), _ptr(nus), _ptr(slacks),
                                          _ptr(plan.workspace), _ptr(status), _stream(Q.device))
            if rc == -3:
                raise NotImplementedError("b200qp.DenseQPFunction: this problem size is outside the fused kernels "
                                          "(nineq <= 128 and nz, neq + nineq <= 128)")
            _lib.check(rc, "b200qp_forward (dense)")
            st = status.tolist()
            info.update(n_iter=int(st[_lib.ST_NITER]), best_resid_max=st[_lib.ST_BEST_MAX])
            if st[_lib.ST_Q_FAIL] > 0 or st[_lib.ST_AQA_FAIL] > 0:
                raise RuntimeError("b200qp.DenseQPFunction: the regularised KKT matrix is singular (Q must be PSD "
                                   "and A full rank)")
            ctx.plan = plan
            ctx.nus, ctx.lams, ctx.slacks = nus, lams, slacks
            ctx.save_for_backward(zhats)
            return zhats

        @staticmethod
        def backward(ctx, dl_dzhat):
            zhats, = ctx.saved_tensors
            plan = ctx.plan
            L = _lib.lib()
            nb, nz, nineq, neq = plan.nBatch, plan.nz, plan.nineq, plan.neq
            opt = dict(dtype=zhats.dtype, device=zhats.device)
            gz = dl_dzhat.contiguous()
            dQ, dp = torch.empty(nb, nz, nz, **opt), torch.empty(nb, nz, **opt)
            dG, dh = torch.empty(nb, nineq, nz, **opt), torch.empty(nb, nineq, **opt)
            dA = torch.empty(nb, neq, nz, **opt) if neq > 0 else None
            db = torch.empty(nb, neq, **opt) if neq > 0 else None
            # the adjoint system uses the UNregularised K of the best iterate (qp.py:248-252): redo
            # the d-independent pre-factorisation without the regularisation, then solve
            # (same flags => same workspace layout as the forward; only the regularisation is dropped)
            pr0 = _lib.Problem.from_buffer_copy(plan.prob)
            pr0.kkt_reg = 0.0
            assert L.b200qp_workspace_bytes(ctypes.byref(pr0)) == plan.workspace.numel()
            with torch.cuda.device(zhats.device):
                rc = L.b200qp_prefactor(ctypes.byref(pr0), _ptr(plan.Q), _ptr(plan.G), _ptr(plan.A), _ptr(plan.workspace),
                                        ctypes.c_void_p(0), _stream(zhats.device))
                _lib.check(rc, "b200qp_prefactor (dense backward)")
                rc = L.b200qp_backward(ctypes.byref(plan.prob), _ptr(zhats), _ptr(ctx.lams), _ptr(ctx.nus), _ptr(ctx.slacks),
                                       _ptr(gz), _ptr(dQ), _ptr(dp), _ptr(dG), _ptr(dh), _ptr(dA), _ptr(db),
                                       _ptr(plan.workspace), _stream(zhats.device))
            _lib.check(rc, "b200qp_backward (dense)")
            return (dQ, dp, dG, dh, dA, db, None, None)

    def apply(Q, p, G, h, A, b, dyn_res=None, cost_grad=None):
        return Solver.apply(Q, p, G, h, A, b, dyn_res, cost_grad)

    apply.info = info
    apply.Function = Solver
    return apply


def _classify_callbacks(plan, dyn_res, cost_grad):
    """This fork evaluates `cost_grad(x)` in place of Qx+p and `dyn_res(x)` in place of Ax-b at the top of every
    iteration (qpth/solvers/pdipm/batch.py:93-102, batch_LU.py:88-97); its MPC callers pass the NON-linear dynamics
    residual (qpth/qp_wrapper.py:303-316, sl1qp_mpc.py:312-320).  A callback that IS the canonical linear form
    (checked on a probe point) stays inside the fused kernels -- including the resident several-iterations-per-launch
    route; any other callback is evaluated by the caller between launches (`_forward_callbacks`).
    Returns (cost_grad or None, dyn_res or None): the callbacks that must really be called."""
    if dyn_res is None and cost_grad is None:
        return None, None
    nb, nz = plan.nBatch, plan.nz
    g = torch.Generator(device="cpu").manual_seed(1234)
    x = torch.randn(nb, nz, generator=g, dtype=torch.float64).to(device=plan.Q.device, dtype=plan.Q.dtype)
    tol = 1e-9 if plan.Q.dtype == torch.float64 else 1e-4

    def close(a, b):
        return a.shape == b.shape and torch.allclose(a, b, rtol=tol, atol=tol * (1 + float(b.abs().max()) if b.numel() else 1.0))

    keep_cg, keep_ry = None, None
    with torch.no_grad():
        if cost_grad is not None:
            Q = plan.Q if not plan.Q_e else plan.Q.unsqueeze(0).expand(nb, nz, nz)
            p = plan.p if not plan.p_e else plan.p.unsqueeze(0).expand(nb, nz)
            want = torch.bmm(Q, x.unsqueeze(2)).squeeze(2) + p
            if not close(cost_grad(x), want):
                keep_cg = cost_grad
        if dyn_res is not None and plan.neq > 0:
            A = plan.A if not plan.A_e else plan.A.unsqueeze(0).expand(nb, plan.neq, nz)
            b = plan.b if not plan.b_e else plan.b.unsqueeze(0).expand(nb, plan.neq)
            want = torch.bmm(A, x.unsqueeze(2)).squeeze(2) - b
            if not close(dyn_res(x), want):
                keep_ry = dyn_res
    return keep_cg, keep_ry


def _forward_callbacks(L, plan, bufs, status, device, cost_grad, dyn_res):
    """The forward with caller-evaluated residual callbacks (include/b200qp.h, b200qp_forward_cb_step /
    b200qp_forward_phase_cb): one launch per iteration; between the step of iteration it-1 and the body of iteration it
    the iterate x is handed to `cost_grad` / `dyn_res` on the current stream, exactly where the reference calls them."""
    base = [_ptr(plan.Q), _ptr(plan.p), _ptr(plan.G), _ptr(plan.h), _ptr(plan.A), _ptr(plan.b)] + [_ptr(t) for t in bufs] + \
           [_ptr(plan.workspace), _ptr(status)]
    st = _stream(device)
    pr = ctypes.byref(plan.prob)
    _lib.check(L.b200qp_forward_phase(pr, _lib.PHASE_BEGIN, *base, st), "b200qp_forward_phase(begin)")
    nb, nz, neq = plan.nBatch, plan.nz, plan.neq
    dt = plan.Q.dtype
    x = torch.empty(nb, nz, dtype=dt, device=device)
    for it in range(plan.prob.max_iter):
        _lib.check(L.b200qp_forward_cb_step(pr, it, _ptr(x), _ptr(plan.workspace), st), "b200qp_forward_cb_step")
        with torch.no_grad():
            cg = cost_grad(x).to(dt).reshape(nb, nz).contiguous() if cost_grad is not None else None
            ry = dyn_res(x).to(dt).reshape(nb, neq).contiguous() if dyn_res is not None else None
        _lib.check(L.b200qp_forward_phase_cb(pr, it, *base, _ptr(cg), _ptr(ry), st), "b200qp_forward_phase_cb")
    _lib.check(L.b200qp_forward_phase(pr, _lib.PHASE_END, *base, st), "b200qp_forward_phase(end)")
