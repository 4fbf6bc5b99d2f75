"""Commutator-style DIIS with restarts / adaptive depth on a fixed-point map g(x) = x.

Reference: algo/cdiis.py:31-357 (R-CDIIS, AD-CDIIS, FD-CDIIS and plain Roothaan iteration).
All variants extrapolate x~ = sum_i c_i x_i over a window of iterates, with the coefficients from
the least-squares problem  min || r_ref + S gamma ||  solved through a QR factorisation of the
matrix S of residual differences:

    R-CDIIS    s_k = r_{k+1} - r_oldest, rhs = r_oldest; the window restarts (keeps `minrestart`
               iterates) when the new column is nearly in the span of the old ones:
               tau ||s|| > ||s - Q1 Q1^T s||
    AD-CDIIS   s_k = r_{k+1} - r_k, rhs = r_k; iterates whose residual is more than 1/delta times
               the newest one are dropped
    FD-CDIIS   s_k = r_{k+1} - r_k, rhs = r_k; fixed window of `diis_size`

The fixed-point map is evaluated on the device by the caller; this module is bookkeeping on
vectors of length npar.  ``modeQR="full"`` keeps the reference's incremental ``qr_insert`` /
``qr_delete`` updates (their rounding decides restarts in borderline cases).
"""

from __future__ import annotations

import numpy as np
import scipy.linalg as sla

from ..core.logging import get_print_func

__all__ = ["cdiis"]

_MODES = ("R-CDIIS", "AD-CDIIS", "FD-CDIIS", "Roothaan")


class _Window:
    """Iterates, residuals and the residual-difference matrix S (one column per step)."""

    def __init__(self, x, r):
        self.x = [x]
        self.r = [r]
        self.S = None

    def push_column(self, s, fresh):
        col = np.reshape(s, (-1, 1))
        self.S = col if fresh else np.hstack((self.S, col))

    def keep_last(self, n):
        self.S = self.S[:, -n:]
        self.r = self.r[-n:]
        self.x = self.x[-n:]

    def drop(self, indices):
        self.S = np.delete(self.S, indices, axis=1)
        for i in sorted(indices, reverse=True):
            self.r.pop(i)
            self.x.pop(i)

    def drop_oldest(self, mk):
        self.S = self.S[:, 1 : mk + 1]
        self.x.pop(0)
        self.r.pop(0)


def _coefficients(mode, gamma, mk):
    c = np.zeros(mk + 1)
    if mode == "R-CDIIS":
        c[0] = 1.0 - np.sum(gamma)
        c[1:] = gamma[:mk]
    else:
        c[0] = -gamma[0]
        for i in range(1, mk):
            c[i] = gamma[i - 1] - gamma[i]
        c[mk] = 1.0 - np.sum(c[0:mk])
    return c


def cdiis(x0, func, threshold, maxiter=1000, modeQR="full", mode="R-CDIIS", diis_size=5, param=0.1,
          minrestart=1, slidehole=False, logger=None, verbose=False):  # fmt: skip
    """Returns ``(conv, niter, rnormlist, mklist, cnormlist, xlast, history_x)``."""
    if mode not in _MODES:
        raise RuntimeError(f"Unknown mode: {mode}")
    say = get_print_func(logger, verbose)
    extrapolate = mode != "Roothaan"
    say(f"CDIIS-type fixed-point acceleration, mode {mode}")
    if mode in ("R-CDIIS", "AD-CDIIS"):
        say(f"  parameter: {param}")

    x = x0
    npar = len(x)
    if npar < diis_size:
        diis_size = npar
        say(f"  window reduced to the number of parameters ({npar})")

    r = func(x) - x
    win = _Window(x, r)
    history_x = [x]
    rnormlist, mklist, cnormlist = [], [], []
    mk, nbiter = 0, 1
    filled_once = False  # FD-CDIIS: the window has been full before
    refactor = True  # R-CDIIS: Q, R must be rebuilt from scratch
    Q = R = Q1 = None
    xlast = x

    while np.linalg.norm(win.r[-1]) > threshold and nbiter < maxiter:
        rnormlist.append(np.linalg.norm(r))
        mklist.append(mk)
        say(f"  iteration {nbiter}: depth {mk}, ||r|| = {np.linalg.norm(win.r[-1])}")

        if mk > 0 and extrapolate:
            S = win.S
            if modeQR == "economic" or mode == "AD-CDIIS":
                Q, R = sla.qr(S, mode="economic")
            elif modeQR == "full" and mode == "R-CDIIS":
                if mk == 1 or refactor:
                    refactor = False
                    Q, R = sla.qr(S)
                else:
                    Q, R = sla.qr_insert(Q, R, S[:, -1], mk - 1, "col")
            elif modeQR == "full" and mode == "FD-CDIIS":
                if mk == 1:
                    Q, R = sla.qr(S)
                elif mk < diis_size:
                    Q, R = sla.qr_insert(Q, R, S[:, -1], mk - 1, "col")
                else:
                    if filled_once:
                        Q, R = sla.qr_delete(Q, R, 0, which="col")
                    Q, R = sla.qr_insert(Q, R, S[:, -1], mk - 1, "col")
                    filled_once = True
            Q1 = Q[:, 0:mk]
            ref = win.r[0] if mode == "R-CDIIS" else win.r[-1]
            rhs = -np.dot(Q.T, np.reshape(ref, (-1, 1)))
            gamma = sla.solve_triangular(R[0:mk, 0:mk], rhs[0:mk], lower=False).flatten()
            c = _coefficients(mode, gamma, mk)
            x_tilde = np.zeros_like(x)
            for i in range(mk + 1):
                x_tilde = x_tilde + c[i] * win.x[i]
            cnormlist.append(np.linalg.norm(c, np.inf))
        else:
            x_tilde = x.copy()
            cnormlist.append(1.0)

        x = func(x_tilde)
        win.x.append(x)
        r = func(x) - x
        say(f"    ||r_next|| = {np.linalg.norm(r)}")
        if mode in ("AD-CDIIS", "FD-CDIIS"):
            s = r - win.r[-1]
        elif mode == "R-CDIIS":
            s = r - win.r[0]
        else:
            s = r.copy()
        win.r.append(r)
        win.push_column(s, fresh=(mk == 0 or not extrapolate))

        if mode == "R-CDIIS":
            if mk > 0:
                last = win.S[:, -1]
                lhs = param * np.linalg.norm(last)
                out_of_span = np.linalg.norm(last - np.dot(Q1, np.dot(Q1.T, last)))
                say(f"    tau*||s|| = {lhs}  vs  ||s - Q Q^T s|| = {out_of_span}")
                if lhs > out_of_span:
                    mk = minrestart - 1
                    win.keep_last(minrestart)
                    refactor = True
                else:
                    mk += 1
            else:
                mk += 1
        elif mode == "AD-CDIIS":
            mk += 1
            newest = np.linalg.norm(win.r[-1])
            out = []
            for i in range(0, mk - 1):
                if newest < param * np.linalg.norm(win.r[i]):
                    out.append(i)
                elif not slidehole:
                    break
            if out:
                mk -= len(out)
                say(f"    dropped iterates {out}")
                win.drop(out)
            if mk == npar + 1:
                win.drop_oldest(mk)
                mk -= 1
        elif mode == "FD-CDIIS":
            if mk == diis_size:
                win.drop_oldest(mk)
            if mk < diis_size:
                mk += 1

        nbiter += 1
        xlast = x
        history_x.append(x.copy())

    conv = not (np.linalg.norm(win.r[-1]) > threshold and nbiter == maxiter)
    return conv, nbiter - 1, rnormlist, mklist, cnormlist, xlast, history_x
