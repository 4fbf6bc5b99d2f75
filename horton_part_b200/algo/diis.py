"""Anderson-Pulay DIIS on a fixed-point map g(x) = x (reference: algo/diis.py:39-136, 139-236).

The map itself is the expensive part (for gLISA one device pass over the molecular grid per call);
this module only holds the subspace bookkeeping: a window of the last ``diis_size`` residuals
r_i = g(x_i) - x_i, the bordered Gram system

    [ (R R^T + (R R^T)^T)/2   -1 ] [ c ]   [  0 ]
    [          -1^T            0 ] [ l ] = [ -1 ]

and the extrapolation x~ = sum_i c_i v_i over the stored vectors (v = x for version "P", followed
by one more application of g; v = g(x) for version "A").
"""

from __future__ import annotations

import warnings
from collections import deque

import numpy as np

from ..core.logging import get_print_func

__all__ = ["diis", "lstsq_spsolver", "lstsq_solver_dyn"]


def _bordered_system(residues):
    m = len(residues)
    R = np.asarray(residues)
    gram = np.einsum("ip,jp->ij", R, R)
    B = np.zeros((m + 1, m + 1))
    B[:m, :m] = (gram + gram.T) / 2
    B[m, :m] = B[:m, m] = -1
    rhs = np.zeros(m + 1)
    rhs[m] = -1
    return B, rhs


def lstsq_spsolver(propars_list, residues_list):
    """Extrapolated vector from the bordered system, solved with SuperLU like the reference
    (``scipy.sparse.linalg.spsolve``); a singular system falls back to the newest vector."""
    from scipy.sparse import csc_matrix
    from scipy.sparse.linalg import spsolve

    B, rhs = _bordered_system(residues_list)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sol = spsolve(csc_matrix(B), rhs)
    out = np.einsum("i,ip->p", sol[:-1], np.asarray(propars_list))
    if np.isnan(out).any():
        print("DIIS: singular matrix.")
        return propars_list[-1]
    return out


def lstsq_solver_dyn(propars_list, residues_list):
    """Variant that drops the oldest vectors until the bordered matrix has no eigenvalue below
    1e-14 in magnitude (algo/diis.py:171-214)."""
    from scipy.linalg import eigh

    n = len(propars_list) + 1
    if n == 2:
        return propars_list[-1]
    B, rhs = _bordered_system(residues_list)
    first = 0
    while first < n - 1:
        w, v = eigh(B[first:, first:])
        small = int(np.sum(np.abs(w) < 1e-14))
        if small == 0:
            sol = (v * 1 / w) @ (v.T @ rhs[first:])
            out = np.einsum("i,ip->p", sol[:-1], np.asarray(propars_list[first:]))
            if (out < -1e-8).any():
                warnings.warn("Use result from the last iteration due to negative parameters found!")
            print(f"Updated size of DIIS subspace: {len(sol[:-1])}")
            return out
        first += small
    warnings.warn("Linear dependence found in DIIS error vectors.")
    print("real DIIS size: 1")
    return propars_list[-1]


def diis(x0, func, threshold, maxiter=1000, diis_size=8, version="P", lstsq_solver=None,
         conv_func=None, verbose=False, logger=None):  # fmt: skip
    """Returns ``(x, niter, history_x)``; raises RuntimeError after ``maxiter`` iterations.

    ``conv_func(residual, x_new, x_old)`` replaces the default convergence measure ||r||.  As in
    the reference the default measure is tested on the residual of the iterate *before* the update
    while the returned vector is the updated one.
    """
    solve = lstsq_solver or lstsq_spsolver
    say = get_print_func(logger, verbose)
    say("            Iter.    dRMS      ")
    say("            -----    ------    ")
    residues = deque(maxlen=diis_size)
    vectors = deque(maxlen=diis_size)
    history_x = []
    x = x0
    for it in range(maxiter):
        gx = func(x)
        r = gx - x
        previous = x.copy()
        residues.append(r)
        if version == "P":
            snapshot = x.copy()
            history_x.append(snapshot)
            vectors.append(snapshot)
            x = func(solve(list(vectors), list(residues)))
        else:
            vectors.append(gx)
            x = solve(list(vectors), list(residues))
            history_x.append(x)
        if conv_func is None:
            measure = np.linalg.norm(r)
            if measure < threshold:
                return x, it + 1, history_x
        else:
            measure = conv_func(r, x, previous)
        say(f"           {it:<4}    {measure:.6E}")
        if measure < threshold:
            return x, it + 1, history_x
    raise RuntimeError("Error: not converge!")
