"""Generalised MBIS (GMBIS): MBIS shell structure and initial guess, NLIS update with a fixed order
n per shell.  Counterpart of the reference's ``gmbis.py`` (:36-156)."""

from __future__ import annotations

import numpy as np

from .mbis import get_nshell
from .nlis import NLISWPart

__all__ = ["GMBISWPart", "get_initial_gmbis_propars"]

_SHELL_CAPACITY = np.array([2.0, 8.0, 8.0, 18.0, 18.0, 32.0, 32.0])


def get_initial_gmbis_propars(number, exp_n_dict, logger=None):
    """[N, S, n] per shell with MBIS' populations/exponents (gmbis.py:36-72)."""
    nshell = get_nshell(number)
    propars = np.zeros(3 * nshell, float)
    s_first = 2.0 * number
    ratio = (2.0 / s_first) ** (1.0 / (nshell - 1)) if nshell > 1 else 1.0
    for k in range(nshell):
        propars[3 * k] = _SHELL_CAPACITY[k]
        propars[3 * k + 1] = s_first * ratio**k
        propars[3 * k + 2] = exp_n_dict[(number, k)] if (number, k) in exp_n_dict else 1.0
    propars[-3] = number - propars[:-3:3].sum()
    return propars


class GMBISWPart(NLISWPart):
    """Generalize Minimal Basis Iterative Stockholder (MBIS)"""

    name = "gmbis"
    _scheme_label = "Generalize Minimal Basis Iterative Stockholder (GMBIS)"

    def __init__(self, coordinates, numbers, pseudo_numbers, grid, moldens, spindens=None, lmax=3,
                 logger=None, threshold=1e-6, maxiter=500, inner_threshold=1e-8, exp_n_dict=1.0,
                 grid_type=1, **kwargs):  # fmt: skip
        super().__init__(coordinates, numbers, pseudo_numbers, grid, moldens, spindens, lmax=lmax,
                         logger=logger, threshold=threshold, maxiter=maxiter,
                         inner_threshold=inner_threshold, exp_n_dict=exp_n_dict, nshell_dict={},
                         grid_type=grid_type, **kwargs)  # fmt: skip

    def _initial_atom_propars(self, number):
        return get_initial_gmbis_propars(number, self._exp_n_dict, logger=self.logger)

    def _atom_nshell(self, number):
        return int(get_nshell(number))
