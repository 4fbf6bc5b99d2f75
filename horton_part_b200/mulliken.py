"""Mulliken operators in an atom-centred basis (host NumPy; no grid, not on the GPU path).

Counterpart of the reference's ``mulliken.py`` (:20-99), kept so that the package inventory is
complete: for every atom the overlap matrix is restricted to the ROWS of that atom's basis
functions and symmetrised, ``P_a = (M_a S + S M_a) / 2`` with ``M_a`` the diagonal 0/1 mask of the
atom; the Mulliken population is ``sum(dm * P_a)``.  Shell types follow the reference's
convention: ``t >= 0`` a Cartesian shell with (t+1)(t+2)/2 functions, ``t <= -2`` a pure shell with
2|t|+1 functions.
"""

from __future__ import annotations

import numpy as np

__all__ = ["get_shell_nbasis", "basis_centers", "partition_mulliken", "get_mulliken_operators"]


def get_shell_nbasis(shell_type):
    """Number of basis functions of a shell type (mulliken.py:28-44); -1 is not a valid type."""
    if shell_type > 0:
        return (shell_type + 1) * (shell_type + 2) // 2
    if shell_type == -1:
        return -1
    return -2 * shell_type + 1


def basis_centers(shell_types, shell_maps, nbasis=None):
    """Centre index of every basis function, from the per-shell types and centres."""
    counts = [get_shell_nbasis(t) for t in shell_types]
    centers = np.repeat(np.asarray(shell_maps, dtype=np.int64), counts)
    if nbasis is not None and centers.size != nbasis:
        # the reference's slice assignments ignore a mismatch; functions beyond the shells belong to no atom
        centers = np.concatenate([centers, np.full(max(nbasis - centers.size, 0), -1)])[:nbasis]
    return centers


def partition_mulliken(operator, nbasis, shell_types, shell_maps, index):
    """In place: zero the rows of the functions that are not on atom ``index``, then symmetrise
    (mulliken.py:47-76)."""
    other = basis_centers(shell_types, shell_maps, nbasis) != index
    operator[other] = 0.0
    operator[:] = 0.5 * (operator + operator.T)


def get_mulliken_operators(overlap, ncenter, shell_types, shell_maps):
    """The list of Mulliken operators, one per centre (mulliken.py:79-99)."""
    overlap = np.asarray(overlap)
    centers = basis_centers(shell_types, shell_maps, len(overlap))
    operators = []
    for a in range(ncenter):
        rows = np.where((centers == a)[:, None], overlap, 0.0)
        operators.append(0.5 * (rows + rows.T))
    return operators
