"""b200qp -- B200-native batched PDIPM QP solve and augmented-Lagrangian MPC behind the reference's
(qpth) call surface.

Import names mirror qpth/__init__.py: `b200qp.qp.QPFunction`, `b200qp.qp.QPSolvers`,
`b200qp.AL_mpc.MPC`, `b200qp.al_utils.{QuadCost,LinDx}`, `b200qp.envs.*Dynamics`, `b200qp.util`.
"""
from . import util  # noqa: F401
from . import qp  # noqa: F401
from . import al_utils  # noqa: F401
from . import AL_mpc  # noqa: F401
from . import AL_mpc as al_mpc  # noqa: F401
from . import envs  # noqa: F401
from . import my_envs  # noqa: F401
from .qp import QPFunction, DenseQPFunction, QPSolvers  # noqa: F401
