"""Host plug-in solvers of aLISA for the per-atom 1-D problem on the radial grid.

These follow the reference's solver plug-in signature (alisa.py:1304-1335)

    solver(bs_funcs (K, N), rho (N,), propars (K,), points, weights (N,), threshold, logger,
           density_cutoff, negative_cutoff, population_cutoff, **solver_options) -> propars

and operate on what the device hands back per atom: the spherical average of w_a*rho on the
atom's radial grid (N = nrad ~ 150 values) and the K tabulated basis functions.  Everything that
touches the molecular grid (promolecule, weights, projection) stays in the CUDA kernels; what runs
here is the small dense algebra the reference also delegates to LAPACK / SciPy:

    solver_diis, solver_cdiis              fixed-point acceleration     (alisa.py:460-540, 1047-1127)
    solver_newton / m_newton / quasi_newton                             (alisa.py:543-938)
    solver_trust_region                    SciPy trust-constr           (alisa.py:941-1044)
    solver_sc, solver_sc_1_iter            host versions of the device kernels, for callers that
                                           use the functions directly   (alisa.py:193-353)

    solver_cvxopt, solver_sc_plus_cvxopt   the convex programme (alisa.py:67-190, 356-457) through
                                           the built-in interior-point method of algo/cp.py (the
                                           third-party ``cvxopt`` package is not in this image;
                                           ``engine="cvxopt"`` selects it where it is installed)
"""

from __future__ import annotations

import numpy as np

from .algo import bfgs, cdiis, diis
from .utils import (
    check_pro_atom_parameters_neg_pars,
    check_pro_atom_parameters_non_neg_pars,
    compute_quantities,
    optional_package,
)

__all__ = [
    "solver_sc",
    "solver_sc_1_iter",
    "solver_diis",
    "solver_cdiis",
    "solver_newton",
    "solver_m_newton",
    "solver_quasi_newton",
    "solver_trust_region",
    "solver_cvxopt",
    "solver_cvxopt_batched",
    "solver_sc_plus_cvxopt",
    "HOST_SOLVERS",
]


def _fixed_point_map(bs_funcs, rho, weights, density_cutoff):
    def g(x):
        pro_shells, _, _, ratio, _ = compute_quantities(rho, x, bs_funcs, density_cutoff, do_ln_ratio=False)
        return np.einsum("ip,p->i", pro_shells * ratio, weights)

    return g


def solver_sc(bs_funcs, rho, propars, points, weights, threshold, logger, density_cutoff,
              negative_cutoff, population_cutoff, niter_print=1, max_niter_inner=100000):  # fmt: skip
    """c_k <- int rho c_k g_k / pro until sqrt(int (pro_old - pro)^2) < threshold."""
    g = _fixed_point_map(bs_funcs, rho, weights, density_cutoff)
    oldpro = None
    pop = np.einsum("i,i", weights, rho)
    for irep in range(int(max_niter_inner)):
        _, pro, _, _, _ = compute_quantities(rho, propars, bs_funcs, density_cutoff, do_ratio=False, do_ln_ratio=False)
        propars[:] = g(propars)
        change = 1e100
        if oldpro is not None:
            err = oldpro - pro
            change = np.sqrt(np.einsum("i,i,i", weights, err, err))
        if irep % int(niter_print) == 0:
            logger.debug(f"            {irep+1:<4}    {change:.3e}")
        if change < threshold:
            check_pro_atom_parameters_non_neg_pars(
                propars, basis_functions=np.asarray(bs_funcs), logger=logger, total_population=float(pop),
                negative_cutoff=negative_cutoff, population_cutoff=population_cutoff)  # fmt: skip
            return propars
        oldpro = pro
    logger.warning("Warning: Inner iteration is not converge!")
    return propars


def solver_sc_1_iter(bs_funcs, rho, propars, points, weights, threshold, logger, density_cutoff,
                     negative_cutoff, population_cutoff):  # fmt: skip
    return _fixed_point_map(bs_funcs, rho, weights, density_cutoff)(propars)


def solver_diis(bs_funcs, rho, propars, points, weights, threshold, logger, density_cutoff,
                negative_cutoff, population_cutoff, check_mono=True, **diis_options):  # fmt: skip
    g = _fixed_point_map(bs_funcs, rho, weights, density_cutoff)
    new_propars, _, _ = diis(propars, g, threshold, logger=logger, **diis_options)
    check_pro_atom_parameters_neg_pars(
        new_propars, bs_funcs, logger=logger, negative_cutoff=negative_cutoff,
        population_cutoff=population_cutoff, total_population=float(np.einsum("i,i", weights, rho)),
        check_monotonicity=check_mono)  # fmt: skip
    return new_propars


def solver_cdiis(bs_funcs, rho, propars, points, weights, threshold, logger, density_cutoff,
                 negative_cutoff, population_cutoff, check_mono=True, **cdiis_options):  # fmt: skip
    g = _fixed_point_map(bs_funcs, rho, weights, density_cutoff)
    conv, _, _, _, _, xlast, _ = cdiis(propars, g, threshold, logger=logger, **cdiis_options)
    if not conv:
        raise RuntimeError("Not converged!")
    check_pro_atom_parameters_neg_pars(
        xlast, basis_functions=bs_funcs, total_population=float(np.einsum("i,i", weights, rho)),
        logger=logger, negative_cutoff=negative_cutoff, population_cutoff=population_cutoff,
        check_monotonicity=check_mono)  # fmt: skip
    return xlast


def _general_newton(bs_funcs, rho, propars, points, weights, threshold, logger, density_cutoff,
                    negative_cutoff, population_cutoff, mode="bfgs", tau=1.0, linspace_size=20,
                    check_mono=False, niter=10000):  # fmt: skip
    """Newton / modified Newton / BFGS on  min int rho ln(rho/pro) + int pro  (alisa.py:753-938).

    gradient_k = -int w rho g_k / pro,  hessian_kj = int w rho g_k g_j / pro^2,
    step = solve(hessian, -1 - gradient); "modified" and "bfgs" back-track over
    linspace(tau, 0, linspace_size) until the pro-atom stays above ``negative_cutoff``
    (and decays monotonically when ``check_mono``)."""
    from scipy.linalg import solve

    def admissible(c):
        _, pro, _, _, _ = compute_quantities(rho, c, bs_funcs, density_cutoff, do_ratio=False, do_ln_ratio=False)
        ok = (pro > negative_cutoff).all()
        if ok and check_mono:
            ok = (pro[:-1] - pro[1:] > negative_cutoff).all()
        return ok

    def take_step(delta, c):
        if mode == "exact":
            return c + delta, delta
        for x in np.linspace(tau, 0, linspace_size):
            step = x * delta
            if admissible(c + step):
                return c + step, step
        raise RuntimeError("Line search failed!")

    if mode not in ("exact", "modified", "bfgs"):
        raise NotImplementedError
    pop = np.einsum("i,i", weights, rho)
    oldpro = olddf = oldH = step = None
    change = 1e100
    for irep in range(niter):
        _, pro, sick, ratio, _ = compute_quantities(rho, propars, bs_funcs, density_cutoff)
        if oldpro is not None:
            err = oldpro - pro
            change = np.sqrt(np.einsum("i,i,i", weights, err, err))
        logger.debug(f"            {irep+1:<4}    {change:.3e}")
        if change < threshold:
            check_pro_atom_parameters_neg_pars(
                propars, total_population=float(pop), pro_atom_density=np.asarray(pro), logger=logger,
                negative_cutoff=negative_cutoff, population_cutoff=population_cutoff,
                check_monotonicity=check_mono)  # fmt: skip
            return propars
        integrand = bs_funcs * ratio
        df = -np.einsum("kp,p->k", integrand, weights)
        if mode == "bfgs":
            H = np.linalg.inv(np.identity(len(propars))) if irep == 0 else bfgs(df, step, olddf, oldH)
            delta = H @ (-1 - df)
        else:
            H = None
            with np.errstate(all="ignore"):
                second = integrand / pro
            second[:, sick] = 0.0
            hess = np.einsum("kp, jp, p->kj", second, bs_funcs, weights)
            delta = solve(hess, -1 - df, assume_a="sym")
        propars[:], step = take_step(delta, propars)
        oldpro, olddf, oldH = pro, df, H
    raise RuntimeError("Inner loop: Newton does not converge!")


def solver_newton(bs_funcs, rho, propars, points, weights, threshold, logger, density_cutoff,
                  negative_cutoff, population_cutoff, **newton_options):  # fmt: skip
    return _general_newton(bs_funcs, rho, propars, points, weights, threshold, logger, density_cutoff,
                           negative_cutoff, population_cutoff, mode="exact", **newton_options)  # fmt: skip


def solver_m_newton(bs_funcs, rho, propars, points, weights, threshold, logger, density_cutoff,
                    negative_cutoff, population_cutoff, **newton_options):  # fmt: skip
    return _general_newton(bs_funcs, rho, propars, points, weights, threshold, logger, density_cutoff,
                           negative_cutoff, population_cutoff, mode="modified", **newton_options)  # fmt: skip


def solver_quasi_newton(bs_funcs, rho, propars, points, weights, threshold, logger, density_cutoff,
                        negative_cutoff, population_cutoff, **newton_options):  # fmt: skip
    return _general_newton(bs_funcs, rho, propars, points, weights, threshold, logger, density_cutoff,
                           negative_cutoff, population_cutoff, mode="bfgs", **newton_options)  # fmt: skip


def solver_trust_region(bs_funcs, rho, propars, points, weights, threshold, logger, density_cutoff,
                        negative_cutoff, population_cutoff, explicit_constr=True,
                        trust_region_options=None):  # fmt: skip
    """SciPy ``trust-constr`` with SR1 Hessian updates, bounds 0 <= c <= 200 and (optionally) the
    population as a linear equality constraint (alisa.py:941-1044).  Note that the reference
    passes ``options={...}.update(...)`` = None to SciPy, i.e. SciPy's defaults; kept."""
    from scipy.optimize import SR1, LinearConstraint, minimize

    nprim = len(propars)
    pop = np.einsum("i,i", weights, rho)
    constraint = LinearConstraint(np.ones((1, nprim)), pop, pop) if explicit_constr else None

    def objective(x=None):
        _, pro, _, ratio, ln_ratio = compute_quantities(rho, x, bs_funcs, density_cutoff)
        f = np.einsum("i,i,i", weights, rho, ln_ratio)
        df = -np.einsum("j,j,ij->i", weights, ratio, bs_funcs)
        if not explicit_constr:
            f += np.einsum("i,i", weights, pro) - pop
            df += 1
        return f, df

    result = minimize(objective, x0=propars, method="trust-constr", jac=True, bounds=[(0.0, 200)] * nprim,
                      constraints=constraint, hess=SR1(), options=None)  # fmt: skip
    if not result.success:
        raise RuntimeError("Convergence failure.")
    c = np.asarray(result["x"]).flatten()
    check_pro_atom_parameters_non_neg_pars(
        c, bs_funcs, total_population=float(pop), logger=logger, negative_cutoff=negative_cutoff,
        population_cutoff=population_cutoff)  # fmt: skip
    return c


def solver_cvxopt(bs_funcs, rho, propars, points, weights, threshold, logger, density_cutoff,
                  negative_cutoff, population_cutoff, allow_neg_params=False, engine=None,
                  **cvxopt_options):  # fmt: skip
    """The convex programme of alisa.py:67-190:
    min int rho ln(rho/pro)  s.t.  sum c = pop  (and c >= 0 unless ``allow_neg_params``).

    The reference solves it with the third-party ``cvxopt.solvers.cp``.  As with GISA's
    ``qpsolvers`` call (gisa.py), the programme is handed to that package when it is installed;
    it is not in this image, and then -- or always with ``engine="builtin"`` -- the interior-point
    method of ``algo/cp.py`` runs on the same objective, gradient and Hessian callbacks (the
    minimiser is unique, see there).  ``engine="cvxopt"`` insists on the package and raises
    ImportError when it is missing.  Remaining keyword arguments are solver options under cvxopt's
    names (``feastol``, ``abstol``, ``reltol``, ``maxiters``, ``show_progress``); without any,
    ``feastol=threshold`` as in the reference."""
    import logging

    if engine not in (None, "builtin", "cvxopt"):
        raise ValueError(f"unknown engine {engine!r}: None, 'builtin' or 'cvxopt'")
    cvxopt = None if engine == "builtin" else optional_package("cvxopt")
    if engine == "cvxopt" and cvxopt is None:
        raise ImportError("engine='cvxopt' needs the `cvxopt` package")
    nprim = len(propars)
    pop = np.einsum("i,i", weights, rho)

    def objective(x=None, z=None):
        if x is None:
            return 0, np.array(propars, dtype=float)
        x = np.asarray(x, dtype=float).ravel()
        _, pro, sick, ratio, ln_ratio = compute_quantities(rho, x, bs_funcs, density_cutoff)
        f = np.einsum("i,i,i", weights, rho, ln_ratio)
        first = weights * ratio
        df = -np.einsum("j,ij->i", first, bs_funcs)
        if z is None:
            return f, df
        second = np.divide(first, pro, out=np.zeros_like(first), where=~sick)
        return f, df, z[0] * np.einsum("k,ik,jk->ij", second, bs_funcs, bs_funcs)

    options = dict(cvxopt_options)
    if not options:
        options = {"show_progress": 3 if logger.level <= logging.DEBUG else 0, "feastol": threshold}

    if cvxopt is None:
        from .algo.cp import cp

        G = h = None
        if not allow_neg_params:
            G, h = -np.identity(nprim), np.zeros(nprim)
        options.setdefault("printer", logger.debug)
        sol = cp(objective, G=G, h=h, A=np.ones((1, nprim)), b=np.array([pop]), options=options)
        c = sol["x"]
    else:

        def objective_cvx(x=None, z=None):
            if x is None:
                return 0, cvxopt.matrix(propars[:])
            res = objective(x, z)
            out = (res[0], cvxopt.matrix(res[1].reshape((1, nprim))))
            return out if z is None else out + (cvxopt.matrix(res[2]),)

        G = h = None
        if not allow_neg_params:
            G = -cvxopt.matrix(np.identity(nprim))
            h = cvxopt.matrix(0.0, (nprim, 1))
        sol = cvxopt.solvers.cp(objective_cvx, G=G, h=h, A=cvxopt.matrix(1.0, (1, nprim)),
                                b=cvxopt.matrix(pop, (1, 1)), options=options)  # fmt: skip
        c = np.asarray(sol["x"]).flatten()
    if sol["status"] != "optimal":
        logger.error("CVXOPT not converged!")
        return None
    check_pro_atom_parameters_non_neg_pars(
        c, basis_functions=np.asarray(bs_funcs), logger=logger, total_population=float(pop),
        negative_cutoff=negative_cutoff, population_cutoff=population_cutoff)  # fmt: skip
    return c


def solver_cvxopt_batched(problems, threshold, logger, density_cutoff, negative_cutoff, population_cutoff,
                          **cvxopt_options):  # fmt: skip
    """``solver_cvxopt`` (built-in engine, non-negative coefficients) for many atoms at once.

    ``problems`` is a list of ``(bs_funcs, rho, propars, points, weights)`` tuples, one per atom, as
    the per-atom plug-in receives them.  Atoms with equal shapes (same element on the same radial
    grid) are stacked and advanced together by ``algo.cp.cp_simplex_batched``; the result of every
    atom equals the per-atom call up to the rounding of the stacked LAPACK solves.  Returns the list
    of new coefficient vectors (None where the programme did not converge, as ``solver_cvxopt``)."""
    from .algo.cp import cp_simplex_batched

    options = dict(cvxopt_options) or {"feastol": threshold}
    groups = {}
    for i, (bs, rho, *_rest) in enumerate(problems):
        groups.setdefault((np.shape(bs), np.shape(rho)), []).append(i)
    out = [None] * len(problems)
    for members in groups.values():
        bs = np.stack([np.asarray(problems[i][0], dtype=float) for i in members])
        rho = np.stack([np.asarray(problems[i][1], dtype=float) for i in members])
        x0 = np.stack([np.asarray(problems[i][2], dtype=float) for i in members])
        w = np.stack([np.asarray(problems[i][4], dtype=float) for i in members])
        x, optimal, _ = cp_simplex_batched(bs, rho, w, x0, density_cutoff, options)
        for j, i in enumerate(members):
            if not optimal[j]:
                logger.error("CVXOPT not converged!")
                continue
            check_pro_atom_parameters_non_neg_pars(
                x[j], basis_functions=bs[j], logger=logger, total_population=float(np.einsum("i,i", w[j], rho[j])),
                negative_cutoff=negative_cutoff, population_cutoff=population_cutoff)  # fmt: skip
            out[i] = x[j]
    return out


def solver_sc_plus_cvxopt(bs_funcs, rho, propars, points, weights, threshold, logger, density_cutoff,
                          negative_cutoff, population_cutoff, sc_iter_limit=1000, **cvxopt_options):  # fmt: skip
    """Self-consistent iterations, falling back to the convex programme when ``sc_iter_limit``
    iterations do not converge (alisa.py:356-457)."""
    g = _fixed_point_map(bs_funcs, rho, weights, density_cutoff)
    oldpro = None
    for irep in range(sc_iter_limit):
        _, pro, _, _, _ = compute_quantities(rho, propars, bs_funcs, density_cutoff, do_ratio=False, do_ln_ratio=False)
        propars[:] = g(propars)
        change = 1e100
        if oldpro is not None:
            err = oldpro - pro
            change = np.sqrt(np.einsum("i,i,i->", weights, err, err))
        logger.debug(f"            {irep+1:<4}    {change:.3e}")
        if change < threshold:
            check_pro_atom_parameters_non_neg_pars(
                propars, logger=logger, total_population=float(np.einsum("i,i", weights, rho)),
                negative_cutoff=negative_cutoff, population_cutoff=population_cutoff)  # fmt: skip
            return propars
        oldpro = pro
    logger.warning("Inner iteration is not converge! Using LISA-I scheme.")
    return solver_cvxopt(bs_funcs, rho, propars=propars, points=points, weights=weights, threshold=threshold,
                         logger=logger, density_cutoff=density_cutoff, negative_cutoff=negative_cutoff,
                         population_cutoff=population_cutoff, **cvxopt_options)  # fmt: skip


#: reference solver name -> host plug-in
HOST_SOLVERS = {
    "cvxopt": solver_cvxopt,
    "sc-plus-convex": solver_sc_plus_cvxopt,
    "diis": solver_diis,
    "cdiis": solver_cdiis,
    "newton": solver_newton,
    "m-newton": solver_m_newton,
    "quasi-newton": solver_quasi_newton,
    "trust-region": solver_trust_region,
}
