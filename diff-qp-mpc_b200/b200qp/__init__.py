"""b200qp -- B200-native batched PDIPM QP solve behind the reference's (qpth) call surface.

Import names mirror qpth/__init__.py: `b200qp.qp.QPFunction`, `b200qp.qp.QPSolvers`,
`b200qp.solvers.pdipm.batch.{pre_factor_kkt,factor_kkt,solve_kkt}`, `b200qp.util`.
"""
from . import util  # noqa: F401
from . import qp  # noqa: F401
from .qp import QPFunction, QPSolvers  # noqa: F401
