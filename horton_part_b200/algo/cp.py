"""Primal-dual interior-point method for the smooth convex programmes of aLISA and gLISA.

The reference hands

    min_x  f(x)      s.t.   G x <= h,   A x = b

(f = int rho ln(rho / sum_k x_k g_k), G = -I, h = 0, A = 1^T, b = N; alisa.py:127-175,
glisa.py:505-550) to the third-party ``cvxopt.solvers.cp``, which is not in this image.  f is
strictly convex on the feasible set whenever the basis functions are linearly independent on the
grid, so the minimiser is unique and any exact method returns the same parameters up to its
stopping tolerance.  This module is such a method, written for the shapes that occur here (a few
to ~1,500 unknowns, dense Hessians that are expensive to evaluate):

* infeasible-start primal-dual Newton steps on the perturbed KKT system

      grad f + G^T z + A^T y = 0,   G x + s = h,   A x = b,   s o z = sigma mu 1,   s, z > 0

  reduced to the (n + p) x (n + p) symmetric system in (dx, dy);
* the centring parameter sigma = (mu_aff / mu)^3 comes from an affine-scaling predictor that reuses
  the factorisation (one Hessian per iteration);
* the step keeps (s, z) strictly positive (fraction 0.99 to the boundary) and back-tracks on the
  norm of the perturbed KKT residual, of which the step is the exact Newton direction.  Trial
  points where f is not finite (a promolecule that went non-positive when G is absent) are
  rejected by the same back-tracking.

The objective protocol is cvxopt's, with NumPy arrays instead of ``cvxopt.matrix``:
``F()`` -> ``(0, x0)``; ``F(x)`` -> ``(f, Df)``; ``F(x, z)`` -> ``(f, Df, z[0] * H)``.
Option names follow cvxopt (``abstol``, ``reltol``, ``feastol``, ``maxiters``, ``show_progress``).
The default gap tolerances are tighter than cvxopt's (1e-14 / 1e-13 instead of 1e-7 / 1e-6): the
method converges superlinearly, so the last digits cost two or three iterations, and the Slater
basis sets are ill-conditioned enough that a gap of 1e-10 still leaves 4e-3 in single coefficients.
"""

from __future__ import annotations

import numpy as np
from scipy.linalg import lu_factor, lu_solve

__all__ = ["cp", "cp_simplex_batched"]


def _as_matrix(M, n):
    if M is None:
        return np.zeros((0, n))
    M = np.asarray(M, dtype=float)
    return M.reshape(-1, n)


def _max_step(v, dv):
    """Largest alpha with v + alpha dv >= 0 (v > 0); inf when no component decreases."""
    neg = dv < 0
    if not neg.any():
        return np.inf
    return float(np.min(-v[neg] / dv[neg]))


class _Newton:
    """Factorised reduced KKT matrix  [[H + G^T diag(z/s) G, A^T], [A, 0]]  at one iterate."""

    def __init__(self, H, G, A, s, z):
        n, p = H.shape[0], A.shape[0]
        self.G, self.A, self.s, self.z, self.n = G, A, s, z, n
        K = np.zeros((n + p, n + p))
        K[:n, :n] = H
        if G.shape[0]:
            K[:n, :n] += G.T @ ((z / s)[:, None] * G)
        K[:n, n:] = A.T
        K[n:, :n] = A
        shift = 0.0
        while True:
            try:
                with np.errstate(all="ignore"):
                    self.lu = lu_factor(K, check_finite=True)
                if np.all(np.abs(np.diag(self.lu[0])) > 1e-300):
                    break
            except (ValueError, np.linalg.LinAlgError):
                pass
            # singular Hessian on the feasible directions: proximal shift
            shift = 1e-12 * max(1.0, np.abs(np.diag(H)).max()) if shift == 0.0 else 100 * shift
            K[np.arange(n), np.arange(n)] += shift
            if shift > 1e6:
                raise np.linalg.LinAlgError("KKT matrix cannot be factorised")

    def solve(self, r_d, r_p, r_e, r_c):
        """Direction for residuals (dual, primal inequality, primal equality, complementarity)."""
        G, s, z, n = self.G, self.s, self.z, self.n
        rhs = np.concatenate([-r_d, -r_e])
        if G.shape[0]:
            rhs[:n] -= G.T @ ((z * r_p - r_c) / s)
        sol = lu_solve(self.lu, rhs)
        dx, dy = sol[:n], sol[n:]
        if G.shape[0]:
            ds = -r_p - G @ dx
            dz = (-r_c - z * ds) / s
        else:
            ds = dz = np.zeros(0)
        return dx, dy, ds, dz


def cp(F, G=None, h=None, A=None, b=None, options=None):
    """Solve  min f(x)  s.t.  G x <= h, A x = b  for a smooth convex f.

    Returns a dict with cvxopt's keys: ``status`` ("optimal" or "unknown"), ``x``, ``y`` (equality
    multipliers), ``zl`` / ``sl`` (inequality multipliers / slacks), ``gap``, ``primal objective``,
    ``iterations`` and the final residual measures.
    """
    opts = {"abstol": 1e-14, "reltol": 1e-13, "feastol": 1e-9, "maxiters": 100, "show_progress": 0, "printer": print}
    opts.update(options or {})
    say = opts["printer"]
    _, x0 = F()
    x = np.array(x0, dtype=float).ravel().copy()
    n = x.size
    G = _as_matrix(G, n)
    A = _as_matrix(A, n)
    mi, p = G.shape[0], A.shape[0]
    h = np.zeros(0) if mi == 0 else np.asarray(h, dtype=float).ravel()
    b = np.zeros(0) if p == 0 else np.asarray(b, dtype=float).ravel()
    one = np.ones(1)

    f, g = F(x)
    g = np.asarray(g, dtype=float).ravel()
    if not np.isfinite(f) or not np.isfinite(g).all():
        raise ValueError("the objective is not finite at the starting point")
    y = np.zeros(p)
    if mi:
        s = h - G @ x
        s = np.maximum(s, 1e-3 * max(1.0, float(np.abs(x).max())))
        z = 0.1 * max(1.0, float(np.abs(g).max())) * s.mean() / s
    else:
        s = z = np.zeros(0)
    scale_d = max(1.0, float(np.linalg.norm(g)))
    scale_p = max(1.0, float(np.abs(h).max()) if mi else 1.0)
    scale_e = max(1.0, float(np.abs(b).max()) if p else 1.0)

    def residuals(x, y, s, z, g):
        r_d = g + (G.T @ z if mi else 0.0) + (A.T @ y if p else 0.0)
        r_p = G @ x + s - h if mi else np.zeros(0)
        r_e = A @ x - b if p else np.zeros(0)
        return r_d, r_p, r_e

    status, it = "unknown", 0
    info = {}
    for it in range(int(opts["maxiters"]) + 1):
        r_d, r_p, r_e = residuals(x, y, s, z, g)
        gap = float(s @ z) if mi else 0.0
        pres = max(float(np.abs(r_p).max()) / scale_p if mi else 0.0, float(np.abs(r_e).max()) / scale_e if p else 0.0)
        dres = float(np.linalg.norm(r_d)) / scale_d
        relgap = gap / abs(f) if f != 0 else np.inf
        info = {"gap": gap, "relative gap": relgap, "primal infeasibility": pres, "dual infeasibility": dres}
        if opts["show_progress"]:
            say(f"{it:3d}: f={f: .10e} gap={gap:.2e} pres={pres:.2e} dres={dres:.2e}")
        feasible = pres <= opts["feastol"] and dres <= opts["feastol"]
        if mi and feasible and (gap <= opts["abstol"] or relgap <= opts["reltol"]):
            status = "optimal"
            break
        if mi and it == int(opts["maxiters"]):
            break  # (without inequalities the stopping rule needs one more Newton direction)
        _, _, H = F(x, one)
        H = np.asarray(H, dtype=float).reshape(n, n)
        newton = _Newton(H, G, A, s, z)
        mu = gap / mi if mi else 0.0
        sigma = 0.0
        if mi:
            # affine-scaling predictor -> centring parameter
            _, _, ds_a, dz_a = newton.solve(r_d, r_p, r_e, s * z)
            a_s, a_z = min(1.0, _max_step(s, ds_a)), min(1.0, _max_step(z, dz_a))
            mu_aff = float((s + a_s * ds_a) @ (z + a_z * dz_a)) / mi
            sigma = min(1.0, max(1e-8, (mu_aff / mu) ** 3)) if mu > 0 else 0.0
        target = sigma * mu
        r_c = s * z - target
        dx, dy, ds, dz = newton.solve(r_d, r_p, r_e, r_c)
        if not mi:
            # no inequalities, hence no duality gap: the Newton decrement dx^T H dx / 2 estimates
            # f(x) - min f and takes its place in the stopping rule
            gap = 0.5 * abs(float(r_d @ dx))
            relgap = gap / abs(f) if f != 0 else np.inf
            info.update({"gap": gap, "relative gap": relgap})
            if feasible and (gap <= opts["abstol"] or relgap <= opts["reltol"]):
                status = "optimal"
                with np.errstate(all="ignore"):
                    fn, gn = F(x + dx)  # the step is already computed: take it if it is sound
                if np.isfinite(fn) and fn <= f + 1e-12 * abs(f):
                    x, y, f, g = x + dx, y + dy, fn, np.asarray(gn, dtype=float).ravel()
                break

        def merit(r_d, r_p, r_e, s, z):
            return np.sqrt(r_d @ r_d + r_p @ r_p + r_e @ r_e + float(np.sum((s * z - target) ** 2)))

        if it == int(opts["maxiters"]):
            break
        m0 = merit(r_d, r_p, r_e, s, z)
        alpha = 1.0
        if mi:
            alpha = min(1.0, 0.99 * _max_step(s, ds), 0.99 * _max_step(z, dz))
        accepted = False
        for _ in range(60):
            xn, yn, sn, zn = x + alpha * dx, y + alpha * dy, s + alpha * ds, z + alpha * dz
            with np.errstate(all="ignore"):
                fn, gn = F(xn)
            gn = np.asarray(gn, dtype=float).ravel()
            if np.isfinite(fn) and np.isfinite(gn).all():
                if merit(*residuals(xn, yn, sn, zn, gn), sn, zn) <= (1.0 - 0.01 * alpha) * m0:
                    accepted = True
                    break
            alpha *= 0.5
            if alpha < 1e-12:
                break
        if not accepted:
            # rounding floor of the residual norm: the iterate cannot be improved any further.  It is
            # accepted when it is feasible and the gap is already below single precision.
            if feasible and (gap <= 1e-8 or relgap <= 1e-7):
                status = "optimal"
            break
        x, y, s, z, f, g = xn, yn, sn, zn, fn, gn

    return {
        "status": status, "x": x, "y": y, "zl": z, "sl": s, "primal objective": float(f),
        "iterations": it, **info,
    }  # fmt: skip


def cp_simplex_batched(bs, rho, weights, x0, density_cutoff, options=None):
    """The aLISA programme for a BATCH of atoms with identical shapes, all atoms advanced together:

        min_c  sum_i w_i rho_i ln(rho_i / (c . g)_i)     s.t.  c >= 0,  sum c = sum_i w_i rho_i

    ``bs`` (B, K, N) basis functions, ``rho`` / ``weights`` (B, N), ``x0`` (B, K).  Same method, same
    arithmetic and same stopping rule as :func:`cp` with G = -I, h = 0, A = 1^T (objective, masks and
    derivatives as in the reference's ``obj_func``, alisa.py:141-160, with ``compute_quantities``,
    utils.py:198-252), but every dense operation is one stacked NumPy call over the atoms that are
    still iterating -- the per-atom Python loop is what limits the host plug-in path on large systems
    (2,000 atoms: ~10 s per outer iteration one by one, ~0.1 s batched).

    Returns ``(x (B, K), optimal (B,) bool, iterations (B,) int)``.
    """
    opts = {"abstol": 1e-14, "reltol": 1e-13, "feastol": 1e-9, "maxiters": 100}
    opts.update({k: v for k, v in (options or {}).items() if k in opts})
    bs = np.asarray(bs, dtype=float)
    rho = np.asarray(rho, dtype=float)
    w = np.asarray(weights, dtype=float)
    nb, K, _ = bs.shape
    pop = np.einsum("bn,bn->b", w, rho)
    wr = w * rho

    def objective(idx, x, second=False):
        """f, grad (and Hessian) for the atoms ``idx`` at coefficients ``x`` (len(idx), K)."""
        g_, rho_ = bs[idx], rho[idx]
        pro = np.einsum("bk,bkn->bn", x, g_)
        sick = (rho_ < density_cutoff) | (pro < density_cutoff)
        with np.errstate(all="ignore"):
            ratio = np.where(sick, 0.0, rho_ / np.where(sick, 1.0, pro))
            ln_ratio = np.where(sick, 0.0, np.log(np.where(sick, 1.0, ratio)))
        f = np.einsum("bn,bn->b", wr[idx], ln_ratio)
        first = w[idx] * ratio
        grad = -np.einsum("bn,bkn->bk", first, g_)
        if not second:
            return f, grad
        with np.errstate(all="ignore"):
            scnd = np.where(sick, 0.0, first / np.where(sick, 1.0, pro))
        return f, grad, np.einsum("bn,bkn,bln->bkl", scnd, g_, g_)

    x = np.array(x0, dtype=float).reshape(nb, K).copy()
    everyone = np.arange(nb)
    f, g = objective(everyone, x)
    if not (np.isfinite(f).all() and np.isfinite(g).all()):
        raise ValueError("the objective is not finite at the starting point")
    y = np.zeros(nb)
    s = np.maximum(x, 1e-3 * np.maximum(1.0, np.abs(x).max(axis=1))[:, None])
    z = (0.1 * np.maximum(1.0, np.abs(g).max(axis=1)) * s.mean(axis=1))[:, None] / s
    scale_d = np.maximum(1.0, np.linalg.norm(g, axis=1))
    scale_e = np.maximum(1.0, np.abs(pop))
    optimal = np.zeros(nb, dtype=bool)
    iterations = np.zeros(nb, dtype=int)
    active = np.ones(nb, dtype=bool)

    def max_step(v, dv):
        with np.errstate(all="ignore"):
            ratio = np.where(dv < 0, -v / np.where(dv < 0, dv, -1.0), np.inf)
        return ratio.min(axis=1)

    for it in range(int(opts["maxiters"]) + 1):
        idx = np.flatnonzero(active)
        if idx.size == 0:
            break
        xa, ya, sa, za, ga, fa = x[idx], y[idx], s[idx], z[idx], g[idx], f[idx]
        r_d = ga - za + ya[:, None]
        r_p = sa - xa
        r_e = xa.sum(axis=1) - pop[idx]
        gap = np.einsum("bk,bk->b", sa, za)
        pres = np.maximum(np.abs(r_p).max(axis=1), np.abs(r_e) / scale_e[idx])
        dres = np.linalg.norm(r_d, axis=1) / scale_d[idx]
        with np.errstate(all="ignore"):
            relgap = np.where(fa != 0, gap / np.abs(fa), np.inf)
        feasible = (pres <= opts["feastol"]) & (dres <= opts["feastol"])
        done = feasible & ((gap <= opts["abstol"]) | (relgap <= opts["reltol"]))
        iterations[idx] = it
        optimal[idx[done]] = True
        active[idx[done]] = False
        if it == int(opts["maxiters"]):
            break
        keep = ~done
        if not keep.any():
            break
        idx, xa, ya, sa, za, fa = idx[keep], xa[keep], ya[keep], sa[keep], za[keep], fa[keep]
        r_d, r_p, r_e, gap, feasible, relgap = r_d[keep], r_p[keep], r_e[keep], gap[keep], feasible[keep], relgap[keep]
        m = idx.size

        _, _, H = objective(idx, xa, second=True)
        kkt = np.zeros((m, K + 1, K + 1))
        kkt[:, :K, :K] = H
        kkt[:, np.arange(K), np.arange(K)] += za / sa
        kkt[:, :K, K] = 1.0
        kkt[:, K, :K] = 1.0

        def solve(r_c):
            rhs = np.empty((m, K + 1))
            rhs[:, :K] = -r_d + (za * r_p - r_c) / sa
            rhs[:, K] = -r_e
            sol = np.linalg.solve(kkt, rhs[:, :, None])[:, :, 0]
            dx, dy = sol[:, :K], sol[:, K]
            ds = -r_p + dx
            dz = (-r_c - za * ds) / sa
            return dx, dy, ds, dz

        mu = gap / K
        _, _, ds_a, dz_a = solve(sa * za)
        a_s = np.minimum(1.0, max_step(sa, ds_a))
        a_z = np.minimum(1.0, max_step(za, dz_a))
        mu_aff = np.einsum("bk,bk->b", sa + a_s[:, None] * ds_a, za + a_z[:, None] * dz_a) / K
        with np.errstate(all="ignore"):
            sigma = np.where(mu > 0, np.minimum(1.0, np.maximum(1e-8, (mu_aff / mu) ** 3)), 0.0)
        target = sigma * mu
        dx, dy, ds, dz = solve(sa * za - target[:, None])

        def merit(r_d, r_p, r_e, s_, z_):
            return np.sqrt((r_d * r_d).sum(axis=1) + (r_p * r_p).sum(axis=1) + r_e * r_e
                           + ((s_ * z_ - target[:, None]) ** 2).sum(axis=1))  # fmt: skip

        m0 = merit(r_d, r_p, r_e, sa, za)
        alpha = np.minimum(1.0, np.minimum(0.99 * max_step(sa, ds), 0.99 * max_step(za, dz)))
        accepted = np.zeros(m, dtype=bool)
        xn, yn, sn, zn = xa.copy(), ya.copy(), sa.copy(), za.copy()
        fn, gn = fa.copy(), g[idx].copy()
        for _ in range(60):
            todo = np.flatnonzero(~accepted & (alpha >= 1e-12))
            if todo.size == 0:
                break
            a = alpha[todo][:, None]
            xt, st, zt = xa[todo] + a * dx[todo], sa[todo] + a * ds[todo], za[todo] + a * dz[todo]
            yt = ya[todo] + alpha[todo] * dy[todo]
            with np.errstate(all="ignore"):
                ft, gt = objective(idx[todo], xt)
            finite = np.isfinite(ft) & np.isfinite(gt).all(axis=1)
            rd_t = gt - zt + yt[:, None]
            rp_t = st - xt
            re_t = xt.sum(axis=1) - pop[idx[todo]]
            tgt = target[todo][:, None]
            mt = np.sqrt((rd_t * rd_t).sum(axis=1) + (rp_t * rp_t).sum(axis=1) + re_t * re_t
                         + ((st * zt - tgt) ** 2).sum(axis=1))  # fmt: skip
            with np.errstate(all="ignore"):
                ok = finite & (mt <= (1.0 - 0.01 * alpha[todo]) * m0[todo])
            hit = todo[ok]
            xn[hit], yn[hit], sn[hit], zn[hit], fn[hit], gn[hit] = xt[ok], yt[ok], st[ok], zt[ok], ft[ok], gt[ok]
            accepted[hit] = True
            alpha[todo[~ok]] *= 0.5
        # atoms whose residual norm cannot be lowered any further: rounding floor (see cp)
        stuck = ~accepted
        if stuck.any():
            fine = stuck & feasible & ((gap <= 1e-8) | (relgap <= 1e-7))
            optimal[idx[fine]] = True
            active[idx[stuck]] = False
        x[idx], y[idx], s[idx], z[idx], f[idx], g[idx] = xn, yn, sn, zn, fn, gn
    return x, optimal, iterations
