"""Small dense host-side algebra around the device passes: fixed-point accelerators (DIIS, CDIIS),
BFGS updates, the active-set QP of GISA and the interior-point method of the convex programmes."""

from .cdiis import cdiis  # noqa: F401
from .cp import cp  # noqa: F401
from .diis import diis, lstsq_solver_dyn, lstsq_spsolver  # noqa: F401
from .quasi_newton import bfgs  # noqa: F401
from .qp import solve_qp_simplex  # noqa: F401
