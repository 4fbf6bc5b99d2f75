"""BFGS update of the inverse Hessian (reference: algo/quasi_newton.py:25-49).

With y = g_new - g_old, rho = 1/(s.y):

    H+ = H + (s.y + y.H.y) rho^2 s s^T - rho (H y s^T + s y^T H)

evaluated in this (expanded) form because the stockholder Newton drivers compare iteration counts
with the reference and the product form (I - rho s y^T) H (I - rho y s^T) + rho s s^T rounds
differently.
"""

from __future__ import annotations

import numpy as np

__all__ = ["bfgs"]


def bfgs(df, s, olddf, oldH):
    y = df - olddf
    sy = s @ y
    yHy = np.einsum("i,ij,j->", y, oldH, y)
    ss = np.outer(s, s)
    cross = oldH @ np.outer(y, s) + np.outer(s, y) @ oldH
    return oldH + (sy + yHy) * ss / sy**2 - cross / sy
