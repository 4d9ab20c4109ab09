"""Shape helpers that callers of the reference import from `qpth.util` (qpth/util.py:22-101): same names,
argument meaning and error text, written against `_RANKS` -- the one table of what a batched parameter looks
like -- instead of per-function copies.  Everything here is host-side shape logic on torch tensors; the
batched products are one-line einsum forms (they are only used by callers around the solver, the solver
itself never materialises them)."""
import torch

# name -> rank of the BATCHED form; one rank less means "shared across the batch"
_RANKS = dict(Q=3, p=2, G=3, h=2, A=3, b=2)
_ORDER = ("Q", "p", "G", "h", "A", "b")


def get_sizes(G, A=None):
    """(nineq, nz, neq, nBatch) of a QP given G (m, n) or (nb, m, n) and optionally A; neq is None when A is not
    given and 0 for an empty A (qpth/util.py:22-33)."""
    if G.dim() not in (2, 3):
        raise RuntimeError("Unexpected number of dimensions.")
    nineq, nz = G.shape[-2], G.shape[-1]
    nBatch = G.shape[0] if G.dim() == 3 else 1
    neq = None if A is None else (A.shape[1] if A.nelement() > 0 else 0)
    return nineq, nz, neq, nBatch


def expandParam(X, nBatch, nDim):
    """Broadcast view of a parameter that is shared across the batch: returns (X', was_expanded).  The reference
    defines this twice (qpth/util.py:36-42 and :69-75); the second definition -- no escape for empty tensors -- is
    the one in effect, and the one mirrored here."""
    rank = X.ndimension()
    if rank == nDim or rank == 0:
        return X, False
    if rank + 1 == nDim:
        return X.unsqueeze(0).expand(nBatch, *X.shape), True
    raise RuntimeError("Unexpected number of dimensions.")


def extract_nBatch(Q, p, G, h, A, b):
    """Batch size = leading dimension of the first parameter (in the order Q, p, G, h, A, b) that comes in its
    batched rank; 1 when every parameter is shared (qpth/util.py:45-51)."""
    for name, X in zip(_ORDER, (Q, p, G, h, A, b)):
        if X.ndimension() == _RANKS[name]:
            return X.shape[0]
    return 1


def bger(x, y):
    """batched outer product (nb, r), (nb, c) -> (nb, r, c)"""
    return torch.einsum("br,bc->brc", x, y)


def bmv(X, y):
    """batched matrix-vector product (nb, r, c), (nb, c) -> (nb, r)"""
    return torch.einsum("brc,bc->br", X, y)


def bquad(x, Q):
    """batched quadratic form x' Q x -> (nb,)"""
    return torch.einsum("br,brc,bc->b", x, Q, x)


def bdot(x, y):
    """batched inner product -> (nb,)"""
    return (x * y).sum(-1)


def bdiag(d):
    """(nb, k) -> (nb, k, k) diagonal matrices"""
    if d.ndimension() != 2:
        raise AssertionError("bdiag expects a (nBatch, k) tensor")
    return torch.diag_embed(d)
