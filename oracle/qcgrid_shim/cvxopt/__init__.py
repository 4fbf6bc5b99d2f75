"""ORACLE ONLY: stand-in for the absent third-party ``cvxopt`` (reference alisa.py:27, 127-175,
glisa.py:505-550) so that the reference's convex-programme solvers run in the build container.

The reference uses two things: ``cvxopt.matrix`` as a dense container and ``cvxopt.solvers.cp``
on one fixed structure,

    min f(x)   s.t.   -x <= 0 (optional),   1^T x = N,

with f smooth and convex (strictly convex for linearly independent basis functions: the minimiser
is unique).  ``solvers.cp`` is answered here by SciPy's ``trust-constr`` with the exact Hessian of
the reference's own objective callback and tolerances at rounding level, followed by a few
projected Newton steps on the active set it identified (polish to the KKT point).  This is
deliberately NOT the product's algorithm (horton_part_b200/algo/cp.py is a primal-dual
interior-point method written independently): agreement between the two pins the product up to
the choice of convex solver, which is all that can be pinned without the package
(SURVEY.md section 8c).  Any other constraint structure raises, as the missing package would.
"""

import numpy as np

__oracle_shim__ = True  # the product refuses modules carrying this mark (utils.optional_package)


class matrix(np.ndarray):
    """Dense column-major-looking container with the three constructors the reference uses:
    ``matrix(ndarray)`` (1-D becomes a column), ``matrix(scalar, (rows, cols))``."""

    def __new__(cls, x, size=None):
        if np.isscalar(x):
            if size is None:
                raise TypeError("matrix(scalar) needs a size")
            arr = np.full(size, float(x))
        else:
            arr = np.array(x, dtype=float)
            if arr.ndim == 1:
                arr = arr.reshape(-1, 1)
            if size is not None:
                arr = arr.reshape(size, order="F")
        return arr.view(cls)


def _absent(*_args, **_kwargs):
    raise ImportError("cvxopt is not installed in this image; only solvers.cp on the LISA structure is restated")


def _cp(F, G=None, h=None, A=None, b=None, dims=None, kktsolver=None, verbose=False, options=None):
    from scipy.optimize import Bounds, LinearConstraint, minimize

    _, x0 = F()
    x0 = np.asarray(x0, dtype=float).ravel()
    n = x0.size
    bounded = G is not None
    if bounded and not (np.array_equal(np.asarray(G), -np.identity(n)) and not np.asarray(h).any()):
        raise ImportError("cvxopt shim: only G = -I, h = 0 is restated")
    if A is None or not np.array_equal(np.asarray(A), np.ones((1, n))) or np.asarray(b).size != 1:
        raise ImportError("cvxopt shim: only A = 1^T is restated")
    total = float(np.asarray(b).ravel()[0])
    one = [1.0]

    def fun(x):
        f, df = F(matrix(x))
        return float(f), np.asarray(df, dtype=float).ravel()

    def hess(x):
        return np.asarray(F(matrix(x), one)[2], dtype=float).reshape(n, n)

    start = np.maximum(x0, 1e-6) if bounded else x0
    res = minimize(
        fun, start, jac=True, hess=hess, method="trust-constr",
        bounds=Bounds(np.zeros(n), np.full(n, np.inf), keep_feasible=True) if bounded else None,
        constraints=[LinearConstraint(np.ones((1, n)), total, total)],
        options={"gtol": 1e-13, "xtol": 1e-15, "barrier_tol": 1e-15, "maxiter": 5000, "verbose": 0},
    )  # fmt: skip
    x = np.asarray(res.x, dtype=float)

    # polish: Newton steps on the free set {x_i above the barrier's floor}, bounds released when
    # their multiplier has the wrong sign; converges in 2-3 steps from the trust-constr point
    active = np.zeros(n, dtype=bool)
    if bounded:
        active = x < 1e-7 * max(1.0, abs(total))
        x = np.where(active, 0.0, x)
    for _ in range(50):
        _, g = fun(x)
        H = hess(x)
        free = np.flatnonzero(~active)
        m = free.size
        kkt = np.zeros((m + 1, m + 1))
        kkt[:m, :m] = H[np.ix_(free, free)]
        kkt[:m, m] = kkt[m, :m] = 1.0
        # Newton step for grad_F + nu = 0, sum x = N
        rhs = np.concatenate([-g[free], [total - x.sum()]])
        try:
            sol = np.linalg.solve(kkt, rhs)
        except np.linalg.LinAlgError:
            sol = np.linalg.lstsq(kkt, rhs, rcond=None)[0]
        dx, nu = sol[:m], sol[m]
        step = 1.0
        if bounded and (dx < 0).any():
            neg = dx < 0
            step = min(1.0, float(np.min(-x[free][neg] / dx[neg])))
        x[free] += step * dx
        if bounded and step < 1.0:
            hit = free[np.argmin(np.where(dx < 0, x[free], np.inf))]
            x[hit] = 0.0
            active[hit] = True
            continue
        if bounded and active.any():
            mult = fun(x)[1] + nu  # must be >= 0 on the active bounds
            worst = np.argmin(np.where(active, mult, np.inf))
            if mult[worst] < -1e-12:
                active[worst] = False
                continue
        if np.abs(dx).max(initial=0.0) <= 1e-15 * max(1.0, np.abs(x).max()):
            break
    status = "optimal" if res.status in (1, 2) or res.success else "unknown"
    return {"status": status, "x": matrix(x), "scipy": res}


class _Solvers:
    options = {}
    cp = staticmethod(_cp)
    qp = staticmethod(_absent)


solvers = _Solvers()
spmatrix = _absent
log = _absent
div = _absent
mul = _absent
