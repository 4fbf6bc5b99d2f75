"""Small dense host-side accelerators around the device fixed-point map (DIIS, CDIIS, BFGS)."""

from .cdiis import cdiis  # noqa: F401
from .diis import diis, lstsq_solver_dyn, lstsq_spsolver  # noqa: F401
from .quasi_newton import bfgs  # noqa: F401
from .qp import solve_qp_simplex  # noqa: F401
