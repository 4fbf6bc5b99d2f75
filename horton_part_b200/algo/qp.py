"""Active-set solver for the small strictly convex quadratic programmes of GISA.

GISA's per-atom update (reference gisa.py:348-421) is

    min_c  1/2 c^T P c + q^T c     s.t.  c >= 0,  sum(c) = N_a

with P the (positive definite, K x K, K <= ~12) overlap matrix of the normalised Gaussians and
q = -2 int rho_a g_k.  The reference hands this to the third-party ``qpsolvers`` front end
(``quadprog`` by default: Goldfarb-Idnani); neither is in this image.  Because the problem is
strictly convex its minimiser is unique, so any exact method gives the same answer up to rounding.
This module implements a primal active-set method on the equality-constrained KKT systems

    [ P_FF  1 ] [ c_F ]   [ -q_F ]
    [ 1^T   0 ] [ -nu ] = [  N   ]

over the free set F, releasing the bound with the most negative multiplier
mu_j = (P c + q)_j - nu and blocking on the first bound hit along the step.
"""

from __future__ import annotations

import numpy as np

__all__ = ["solve_qp_simplex"]


def _kkt_on(P, q, total, free):
    n = len(free)
    kkt = np.zeros((n + 1, n + 1))
    kkt[:n, :n] = P[np.ix_(free, free)]
    kkt[:n, n] = kkt[n, :n] = 1.0
    rhs = np.concatenate([-q[free], [total]])
    try:
        sol = np.linalg.solve(kkt, rhs)
    except np.linalg.LinAlgError:
        sol = np.linalg.lstsq(kkt, rhs, rcond=None)[0]
    return sol[:n], -sol[n]


def solve_qp_simplex(P, q, total, tol=1e-12, maxiter=None):
    """argmin 1/2 x^T P x + q^T x  over  {x >= 0, sum x = total};  P symmetric positive definite.

    Returns the minimiser (exact zeros on the active bounds).  Raises RuntimeError if the
    active-set iteration does not terminate (cannot happen for strictly convex P in exact
    arithmetic; the cap guards against cycling from rounding)."""
    P = np.asarray(P, dtype=float)
    q = np.asarray(q, dtype=float).ravel()
    n = q.size
    if total < 0:
        raise ValueError("the population constraint must be non-negative")
    x = np.full(n, total / n)
    free = list(range(n))
    scale = max(1.0, float(np.abs(q).max(initial=0.0)))
    for _ in range(maxiter or 20 * (n + 1)):
        target, nu = _kkt_on(P, q, total, free)
        step = target - x[free]
        # largest feasible fraction of the step
        shrinking = step < 0
        if shrinking.any():
            ratios = np.where(shrinking, x[free] / np.where(shrinking, -step, 1.0), np.inf)
            k = int(np.argmin(ratios))
            alpha = min(1.0, float(ratios[k]))
        else:
            k, alpha = -1, 1.0
        x[free] = x[free] + alpha * step
        if alpha < 1.0:  # hit a bound: fix that coefficient at zero
            x[free[k]] = 0.0
            del free[k]
            continue
        x[free] = target
        bound = [j for j in range(n) if j not in free]
        if not bound:
            return x
        mu = (P @ x + q)[bound] - nu
        worst = int(np.argmin(mu))
        if mu[worst] >= -tol * scale:
            return x
        free.append(bound[worst])
        free.sort()
    raise RuntimeError("active-set QP did not terminate")
