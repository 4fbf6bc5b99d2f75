"""ORACLE ONLY: stand-in for the absent third-party ``qpsolvers`` (reference gisa.py:26, 409-419,
algo/diis.py:27, 256).

GISA's call has one fixed structure: minimise 1/2 x^T P x + q^T x subject to -x <= 0 and
sum(x) = b with P positive definite, so the minimiser is unique.  It is found here by brute force,
independently of the product's active-set solver: every support set F is tried, the
equality-constrained problem is solved on F, and the feasible candidate with the lowest objective
wins (K <= 12 basis functions per atom: at most 4,095 small linear solves).  Any other constraint
structure raises, as the missing package would.
"""

import itertools

import numpy as np

__oracle_shim__ = True  # the product refuses modules carrying this mark (utils.optional_package)


def solve_qp(P, q, G=None, h=None, A=None, b=None, solver=None, initvals=None, **_options):
    P = np.asarray(P, float)
    q = np.asarray(q, float).ravel()
    n = q.size
    simplex = (
        G is not None and np.array_equal(np.asarray(G), -np.identity(n))
        and h is not None and not np.asarray(h).any()
        and A is not None and np.array_equal(np.asarray(A), np.ones((1, n)))
        and b is not None and np.asarray(b).size == 1
    )  # fmt: skip
    if not simplex or n > 14:
        raise ImportError("qpsolvers is not installed in this image; only GISA's simplex QP is restated")
    total = float(np.asarray(b).ravel()[0])
    best = None
    for r in range(1, n + 1):
        for F in itertools.combinations(range(n), r):
            F = list(F)
            kkt = np.zeros((r + 1, r + 1))
            kkt[:r, :r] = P[np.ix_(F, F)]
            kkt[:r, r] = kkt[r, :r] = 1.0
            try:
                sol = np.linalg.solve(kkt, np.concatenate([-q[F], [total]]))
            except np.linalg.LinAlgError:
                continue
            if (sol[:r] < -1e-13 * max(1.0, total)).any():
                continue
            x = np.zeros(n)
            x[F] = np.maximum(sol[:r], 0.0)
            f = 0.5 * x @ P @ x + q @ x
            if best is None or f < best[0]:
                best = (f, x)
    return None if best is None else best[1]
