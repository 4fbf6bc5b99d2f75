"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU (NumPy) restatement of the reference's stockholder-iteration hot path, used only by tests/,
``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference`` legs as the
checker and CPU baseline.  The product package ``horton_part_b200`` never imports it.

Pinning: every scheme function below is checked in tests/test_oracle_golden.py against outputs of
the UNMODIFIED reference (/root/reference/src/horton_part run here through oracle/qcgrid_shim;
vectors in tests/golden/, generator oracle/gen_golden.py), including the reference's own golden
charges (tests/test_wpart.py:90-102).  Each function cites the reference lines it restates; the
arithmetic ORDER follows the reference (sequential atom order in the promolecule, the 1e-100
offsets, clip, masks) because the iteration count depends on it.

Paths are relative to /root/reference/src/horton_part.
"""

from __future__ import annotations

import numpy as np
from scipy.special import gamma as _gamma

DENSITY_CUTOFF = 1e-15  # data/constants.yaml


# ----------------------------------------------------------------------------------------------
# grid-level helpers (qc-grid semantics)
# ----------------------------------------------------------------------------------------------
def distances(points, center):
    """core/base.py:634  np.linalg.norm(points - R_a, axis=1)."""
    return np.linalg.norm(points - center, axis=1)


def shell_average(atgrid, values):
    """qc-grid AtomGrid.spherical_average evaluated at its own knots (mbis.py:179-184,
    gisa.py:287-289): per-shell sums of f*w, divided by r^2 w_rad, zero at r<1e-8, over 4 pi."""
    prod = values * atgrid.weights
    idx = np.asarray(atgrid.indices)
    sums = np.array([prod[idx[i] : idx[i + 1]].sum() for i in range(len(idx) - 1)])
    r, w = atgrid.rgrid.points, atgrid.rgrid.weights
    with np.errstate(divide="ignore", invalid="ignore"):
        sums /= r**2 * w
    sums[np.abs(r) < 1e-8] = 0.0
    return sums / (4.0 * np.pi)


def entropy(molw, rho, rho0, cutoff=DENSITY_CUTOFF):
    """core/stockholder.py:145-151."""
    sick = (rho0 < cutoff) | (rho < cutoff)
    with np.errstate(all="ignore"):
        ratio = np.divide(rho, rho0, out=np.zeros_like(rho), where=~sick)
        ln_ratio = np.log(ratio, out=np.zeros_like(rho), where=~sick)
    return np.einsum("i,i,i", molw, rho, ln_ratio)


def stockholder_weights(grid, proatom_fn, natom, owner_only=True):
    """core/stockholder.py:352-384 + update_pro :153-175: sequential promolecule accumulation with
    the 1e-100 offsets, then w_a = clip(rho0_a / rho0, 0, 1) on atom a's own slice
    (grid_type=1) or on the whole grid."""
    promol = np.zeros(grid.size)
    pro = []
    for a in range(natom):
        work = proatom_fn(a)
        promol += work
        promol += 1e-100
        lo, hi = grid.indices[a], grid.indices[a + 1]
        pro.append(work[lo:hi].copy() if owner_only else work.copy())
    weights = []
    for a in range(natom):
        lo, hi = grid.indices[a], grid.indices[a + 1]
        w = pro[a] / (promol[lo:hi] if owner_only else promol)
        weights.append(np.clip(w, 0, 1))
    return promol, weights


def stockholder_weights_local(grid, proatom_on_r, coords, natom, radius):
    """The reference's local-grid design (commented block core/stockholder.py:45-112): atom a is
    evaluated only on Grid.get_localgrid(R_a, radius); promolecule and the 1e-100 offsets are
    accumulated on those points only; w_a lives on the overlap of the local grid with atom a's own
    slice and is zero elsewhere."""
    promol = np.zeros(grid.size)
    local = []
    for a in range(natom):
        lo, hi = grid.indices[a], grid.indices[a + 1]
        idx, overlap, dist = local_index(grid.points, coords[a], radius, lo, hi)
        work = proatom_on_r(a, dist)
        promol[idx] += work
        promol[idx] += 1e-100
        local.append((idx[overlap] - lo, work[overlap]))
    weights = []
    for a in range(natom):
        lo, hi = grid.indices[a], grid.indices[a + 1]
        w = np.zeros(hi - lo)
        rel, pro = local[a]
        w[rel] = np.clip(pro / promol[lo:hi][rel], 0, 1)
        weights.append(w)
    return promol, weights


def local_index(points, center, radius, begin, end):
    """Row L: Grid.get_localgrid == cKDTree.query_ball_point(center, radius, p=2.0), canonicalised
    to ascending order; spec core/stockholder.py:84-112."""
    from scipy.spatial import cKDTree

    idx = np.sort(np.asarray(cKDTree(points).query_ball_point(center, radius, p=2.0), dtype=np.int64))
    overlap = (begin <= idx) & (idx < end)
    dist = np.linalg.norm(points[idx] - center, axis=1)
    return idx, overlap, dist


# ----------------------------------------------------------------------------------------------
# MBIS  (mbis.py)
# ----------------------------------------------------------------------------------------------
def mbis_nshell(number):
    return int(np.array([2, 10, 18, 36, 54, 86, 118]).searchsorted(number) + 1)  # mbis.py:36-46


def mbis_initial(number):
    """mbis.py:49-78."""
    k = mbis_nshell(number)
    p = np.zeros(2 * k)
    s0 = 2.0 * number
    ratio = (2.0 / s0) ** (1.0 / (k - 1)) if k > 1 else 1.0
    cap = [2.0, 8.0, 8.0, 18.0, 18.0, 32.0, 32.0]
    for i in range(k):
        p[2 * i] = cap[i]
        p[2 * i + 1] = s0 * ratio**i
    p[-2] = number - p[:-2:2].sum()
    return p


def mbis_proatom(par, r):
    """mbis.py:279-289 (eval_proatom) / :253-261 (get_proatom_rho)."""
    y = np.zeros(len(r))
    for k in range(len(par) // 2):
        N, S = par[2 * k : 2 * k + 2]
        y += N * S**3 * np.exp(-S * r) / (8 * np.pi)
    return y


def mbis_inner(rho, par, weights, r, threshold, cutoff=DENSITY_CUTOFF, max_inner=2000):
    """opt_mbis_propars, mbis.py:81-163. Returns (propars, inner iterations)."""
    par = par.copy()
    k = len(par) // 2
    terms = np.zeros((k, len(r)))
    oldpro = None
    for irep in range(max_inner):
        for i in range(k):
            N, S = par[2 * i], par[2 * i + 1]
            terms[i] = N * S**3 * np.exp(-S * r) / (8 * np.pi)
        pro = terms.sum(axis=0)
        sick = (rho < cutoff) | (pro < cutoff)
        with np.errstate(all="ignore"):
            ratio = rho / pro
        ratio[sick] = 0.0
        terms *= ratio
        for i in range(k):
            m0 = np.einsum("p,p->", weights, terms[i])
            m1 = np.einsum("p,p,p->", weights, terms[i], r)
            par[2 * i] = m0
            par[2 * i + 1] = 3 * m0 / m1
        if oldpro is None:
            change = 1e100
        else:
            err = oldpro - pro
            change = np.sqrt(np.einsum("p,p,p->", weights, err, err))
        if change < threshold:
            return par, irep + 1
        oldpro = pro
    return par, max_inner


def _radial_change(atgrids, ranges, fn, new, old):
    """compute_change on radial grids, core/iterstock.py:32-45."""
    msd = 0.0
    for a, g in enumerate(atgrids):
        r = g.rgrid.points
        d = fn(a, new[ranges[a] : ranges[a + 1]], r) - fn(a, old[ranges[a] : ranges[a + 1]], r)
        msd += np.einsum("i,i,i,i", g.rgrid.weights, 4 * np.pi * r**2, d, d)
    return np.sqrt(msd)


def _iterate(grid, rho, natom, pseudo, propars, ranges, proatom_on_r, update_atom, threshold, maxiter,
             cutoff=DENSITY_CUTOFF, coords=None, local_radius=None):  # fmt: skip
    """The outer loop of AbstractISAWPart.do_partitioning (core/iterstock.py:159-193) for
    grid_type=1: weights from the current propars, per-atom projection + update, entropy of the
    promolecule that was just used, change between new and old propars."""
    dist = [distances(grid.points, coords[a]) for a in range(natom)]
    charges = np.zeros(natom)
    hist = {"propars": [], "charges": [], "entropies": [], "changes": [], "inner": []}
    counter = 0
    while True:
        counter += 1
        old = propars.copy()
        if local_radius is None:
            promol, weights = stockholder_weights(
                grid, lambda a: proatom_on_r(a, propars[ranges[a] : ranges[a + 1]], dist[a]), natom
            )
        else:
            promol, weights = stockholder_weights_local(
                grid, lambda a, r: proatom_on_r(a, propars[ranges[a] : ranges[a + 1]], r), coords, natom,
                local_radius,
            )
        inner = []
        for a in range(natom):
            g = grid.atgrids[a]
            lo, hi = grid.indices[a], grid.indices[a + 1]
            sph = shell_average(g, weights[a] * rho[lo:hi])
            r = g.rgrid.points
            w4 = 4 * np.pi * r**2 * g.rgrid.weights
            new_a, nin = update_atom(a, sph, propars[ranges[a] : ranges[a + 1]].copy(), w4, r)
            propars[ranges[a] : ranges[a + 1]] = new_a
            charges[a] = pseudo[a] - np.einsum("p,p->", w4, sph)
            inner.append(nin)
        hist["propars"].append(propars.copy())
        hist["charges"].append(charges.copy())
        hist["entropies"].append(entropy(grid.weights, rho, promol, cutoff))
        hist["inner"].append(inner)
        change = _radial_change(grid.atgrids, ranges, proatom_on_r, propars, old)
        hist["changes"].append(change)
        if change < threshold or counter >= maxiter:
            break
    return {
        "niter": counter,
        "change": change,
        "charges": charges,
        "propars": propars,
        "promoldens": promol,
        "at_weights": weights,
        "history_propars": np.array(hist["propars"]),
        "history_charges": np.array(hist["charges"]),
        "history_entropies": np.array(hist["entropies"]),
        "history_changes": np.array(hist["changes"]),
        "history_inner": hist["inner"],
    }


def mbis(coords, numbers, pseudo, grid, rho, threshold=1e-6, inner_threshold=1e-8, maxiter=500,
         cutoff=DENSITY_CUTOFF, local_radius=None):  # fmt: skip
    """MBISWPart(...).do_partitioning(), grid_type=1 (mbis.py:167-203, 291-305)."""
    natom = len(numbers)
    inner_threshold = min(inner_threshold, threshold)  # core/iterstock.py:101
    ranges = [0]
    for z in numbers:
        ranges.append(ranges[-1] + 2 * mbis_nshell(z))
    propars = np.concatenate([mbis_initial(z) for z in numbers])
    return _iterate(
        grid, rho, natom, pseudo, propars, ranges,
        lambda a, par, r: mbis_proatom(par, r),
        lambda a, sph, par, w4, r: mbis_inner(sph, par, w4, r, inner_threshold, cutoff),
        threshold, maxiter, cutoff, coords, local_radius,
    )  # fmt: skip


# ----------------------------------------------------------------------------------------------
# exponential basis functions (core/basis.py) -- GISA / aLISA / gLISA / NLIS
# ----------------------------------------------------------------------------------------------
def exp_shell(n, population, alpha, r):
    """evaluate_function, core/basis.py:161-171:  c n alpha^(3/n) / (4 pi Gamma(3/n)) exp(-alpha r^n)."""
    pref = population * n * alpha ** (3 / n) / (4 * np.pi * _gamma(3 / n))
    return pref * np.exp(-alpha * r**n)


# ----------------------------------------------------------------------------------------------
# CPU baseline leg (bench.py): the reference's per-iteration dense pass on a sample of the grid
# ----------------------------------------------------------------------------------------------
def dense_weights_pass_mbis(points, owner, coords, ranges, propars, dist=None):
    """One ``update_at_weights`` of the reference (core/stockholder.py:352-384 with
    mbis.py:279-289) restricted to ``points``: per atom K+3 NumPy passes with the *cached*
    distances ``dist[a]`` the reference keeps (core/base.py:630-635), sequential promolecule
    accumulation, owner weights.  Returns (promol, owner weights, atom x point evaluations)."""
    natom = len(coords)
    if dist is None:
        dist = [distances(points, coords[a]) for a in range(natom)]
    promol = np.zeros(len(points))
    own = np.zeros(len(points))
    for a in range(natom):
        work = mbis_proatom(propars[ranges[a] : ranges[a + 1]], dist[a])
        promol += work
        promol += 1e-100
        mine = owner == a
        if mine.any():
            own[mine] = work[mine]
    w = np.clip(own / promol, 0, 1)
    return promol, w, natom * len(points)


def load_basis(kind):
    """Per-element (orders, exponents, initials) of data/gauss.json / slater.json, read from the
    re-keyed copy shipped with the product (tools/import_basis_tables.py documents the provenance)."""
    import json
    import pathlib

    path = pathlib.Path(__file__).resolve().parents[1] / "horton_part_b200" / "data" / "expbasis_tables.json"
    table = json.loads(path.read_text())[kind]
    return {int(z): {k: np.asarray(v, dtype=float) for k, v in t.items()} for z, t in table.items()}


def lisa_initial(numbers, pseudo, basis, nelec):
    """gisa.py:68-88: table initials floored at 1e-4, scaled per atom, then to nelec."""
    parts = []
    for z, pn in zip(numbers, pseudo):
        init = basis[int(z)]["initials"].copy()
        init[init < 1e-4] = 1e-4
        parts.append(init / np.sum(init) * pn)
    propars = np.concatenate(parts)
    return propars / np.sum(propars) * nelec


def lisa_proatom(tab, c, r):
    """core/basis.py:214-232: shells summed in order."""
    y = 0.0
    for k in range(len(c)):
        y = y + exp_shell(tab["orders"][k], c[k], tab["exponents"][k], r)
    return y


def lisa_sc_inner(bs, rho, c, weights, threshold, cutoff=DENSITY_CUTOFF, max_inner=100000, single=False):
    """solver_sc alisa.py:193-291 (single=True: solver_sc_1_iter :294-353) with
    compute_quantities utils.py:198-252."""
    c = c.copy()
    oldpro = None
    for irep in range(max_inner):
        shells = bs * c[:, None]
        pro = np.einsum("ij->j", shells)
        sick = (rho < cutoff) | (pro < cutoff)
        with np.errstate(all="ignore"):
            ratio = np.divide(rho, pro, out=np.zeros_like(rho), where=~sick)
        c[:] = np.einsum("p,ip->i", weights, shells * ratio)
        if single:
            return c, 1
        if oldpro is None:
            change = 1e100
        else:
            err = oldpro - pro
            change = np.sqrt(np.einsum("i,i,i", weights, err, err))
        if change < threshold:
            return c, irep + 1
        oldpro = pro
    return c, max_inner


def alisa(coords, numbers, pseudo, grid, rho, basis_func="gauss", solver="sc", threshold=1e-6,
          inner_threshold=1e-8, maxiter=500, cutoff=DENSITY_CUTOFF, max_inner=100000):  # fmt: skip
    """LinearISAWPart(solver='sc' | 'sc-1-iter').do_partitioning(), grid_type=1
    (gisa.py:281-318, alisa.py:1304-1335)."""
    natom = len(numbers)
    basis = load_basis(basis_func)
    inner_threshold = min(inner_threshold, threshold)
    ranges = [0]
    for z in numbers:
        ranges.append(ranges[-1] + len(basis[int(z)]["exponents"]))
    nelec = np.einsum("i,i", grid.weights, rho)
    propars = lisa_initial(numbers, pseudo, basis, nelec)
    bs = []
    for a in range(natom):  # gisa.py:91-106
        t = basis[int(numbers[a])]
        r = grid.atgrids[a].rgrid.points
        bs.append(np.array([exp_shell(t["orders"][k], 1.0, t["exponents"][k], r) for k in range(len(t["exponents"]))]))
    return _iterate(
        grid, rho, natom, pseudo, propars, ranges,
        lambda a, par, r: lisa_proatom(basis[int(numbers[a])], par, r),
        lambda a, sph, par, w4, r: lisa_sc_inner(bs[a], sph, par, w4, inner_threshold, cutoff, max_inner,
                                                 single=(solver == "sc-1-iter")),
        threshold, maxiter, cutoff, coords,
    )  # fmt: skip


# ----------------------------------------------------------------------------------------------
# NLIS / GMBIS  (nlis.py, gmbis.py)
# ----------------------------------------------------------------------------------------------
def nlis_initial(number, exp_n_dict, nshell_dict):
    """nlis.py:58-96."""
    k = nshell_dict.get(number, mbis_nshell(number))
    p = np.ones(3 * k)
    s0 = 2.0 * number
    ratio = (0.5 / s0) ** (1.0 / (k - 1)) if k > 1 else 1.0
    for i in range(k):
        p[3 * i] = number / k
        p[3 * i + 1] = s0 * ratio**i
        p[3 * i + 2] = exp_n_dict.get((number, i), 1.0)
    return p


def gmbis_initial(number, exp_n_dict):
    """gmbis.py:36-72."""
    k = mbis_nshell(number)
    p = np.zeros(3 * k)
    base = mbis_initial(number)
    p[0::3], p[1::3] = base[0::2], base[1::2]
    p[2::3] = [exp_n_dict.get((number, i), 1.0) for i in range(k)]
    return p


def nlis_proatom(par, r):
    """nlis.py:319-328."""
    y = np.zeros(len(r))
    for k in range(len(par) // 3):
        N, S, n = par[3 * k : 3 * k + 3]
        y += N * n * S ** (3 / n) * np.exp(-S * r**n) / (4 * np.pi * _gamma(3.0 / n))
    return y


def nlis_inner(rho, par, weights, r, threshold, cutoff=DENSITY_CUTOFF, max_inner=2000):
    """opt_nlis_propars, nlis.py:99-194."""
    par = par.copy()
    k = len(par) // 3
    terms = np.zeros((k, len(r)))
    oldpro = None
    for irep in range(max_inner):
        for i in range(k):
            S, n = par[3 * i + 1], par[3 * i + 2]
            terms[i] = n * S ** (3 / n) * np.exp(-S * r**n) / (4 * np.pi * _gamma(3.0 / n))
        pro = np.sum(terms * par[::3, None], axis=0)
        sick = (rho < cutoff) | (pro < cutoff)
        with np.errstate(all="ignore"):
            ratio = rho / pro
        ratio[sick] = 0.0
        tr = terms * ratio
        for i in range(k):
            N, n = par[3 * i], par[3 * i + 2]
            m0 = np.einsum("p,p->", weights, tr[i] * N)
            m1 = np.einsum("p,p,p->", weights, tr[i], r**n)
            par[3 * i] = m0
            par[3 * i + 1] = 1e-5 if np.isclose(m1, 0.0) else 3 / (m1 * n)
        if oldpro is None:
            change = 1e100
        else:
            err = oldpro - pro
            change = np.sqrt(np.einsum("p,p,p->", weights, err, err))
        if change < threshold:
            return par, irep + 1
        oldpro = pro
    return par, max_inner


def nlis(coords, numbers, pseudo, grid, rho, exp_n_dict=None, nshell_dict=None, gmbis_start=False,
         threshold=1e-6, inner_threshold=1e-8, maxiter=500, cutoff=DENSITY_CUTOFF):  # fmt: skip
    """NLISWPart / GMBISWPart (gmbis_start=True) .do_partitioning(), grid_type=1."""
    exp_n_dict, nshell_dict = exp_n_dict or {}, nshell_dict or {}
    natom = len(numbers)
    inner_threshold = min(inner_threshold, threshold)
    init = [gmbis_initial(int(z), exp_n_dict) if gmbis_start else nlis_initial(int(z), exp_n_dict, nshell_dict)
            for z in numbers]  # fmt: skip
    ranges = np.concatenate([[0], np.cumsum([len(p) for p in init])]).tolist()
    return _iterate(
        grid, rho, natom, pseudo, np.concatenate(init), ranges,
        lambda a, par, r: nlis_proatom(par, r),
        lambda a, sph, par, w4, r: nlis_inner(sph, par, w4, r, inner_threshold, cutoff),
        threshold, maxiter, cutoff, coords,
    )  # fmt: skip


# ----------------------------------------------------------------------------------------------
# ISA  (isa.py, spline pro-atoms of core/stockholder.py:202-350)
# ----------------------------------------------------------------------------------------------
def isa_proatom(rgrid_points, par, r):
    """get_proatom_spline + eval_spline + eval_proatom: clip negatives, not-a-knot CubicSpline
    (extrapolating), + 1e-100 (core/stockholder.py:218-219, 266-267, 300-302, 349)."""
    from scipy.interpolate import CubicSpline

    rho = par.copy()
    rho[rho < 0] = 0.0
    return CubicSpline(rgrid_points, rho, True)(r) + 1e-100


def isa(coords, numbers, pseudo, grid, rho, threshold=1e-6, maxiter=500, cutoff=DENSITY_CUTOFF):
    """ISAWPart.do_partitioning(), isa.py:93-122."""
    natom = len(numbers)
    sizes = [g.rgrid.size for g in grid.atgrids]
    ranges = np.concatenate([[0], np.cumsum(sizes)]).tolist()
    propars = np.zeros(ranges[-1])

    def update(a, sph, par, w4, r):
        # isa.py:108-110 (the spline of the spherical average is evaluated at its own knots)
        return np.clip(sph, 1e-100, np.inf), 0

    out = _iterate(
        grid, rho, natom, pseudo, propars, ranges,
        lambda a, par, r: isa_proatom(grid.atgrids[a].rgrid.points, par, r),
        update, threshold, maxiter, cutoff, coords,
    )  # fmt: skip
    return out


# ----------------------------------------------------------------------------------------------
# gLISA  (glisa.py) -- global optimisation on the molecular grid
# ----------------------------------------------------------------------------------------------
def glisa_setup(coords, numbers, pseudo, grid, rho, basis_func="gauss"):
    """init_propars (gisa.py:68-88) + evaluate_basis_functions(force_on_molgrid=True)
    (gisa.py:91-106) + eval_pro_shells (glisa.py:335-344)."""
    basis = load_basis(basis_func)
    ranges = [0]
    for z in numbers:
        ranges.append(ranges[-1] + len(basis[int(z)]["exponents"]))
    nelec = np.einsum("i,i", grid.weights, rho)
    propars = lisa_initial(numbers, pseudo, basis, nelec)
    shells = np.zeros((ranges[-1], grid.size))
    for a, z in enumerate(numbers):
        t = basis[int(z)]
        r = distances(grid.points, coords[a])
        for k in range(len(t["exponents"])):
            shells[ranges[a] + k] = exp_shell(t["orders"][k], 1.0, t["exponents"][k], r)
    return basis, ranges, propars, shells


def glisa_function_g(x, shells, rho, molw, cutoff=DENSITY_CUTOFF):
    """function_g, glisa.py:850-879."""
    rho0 = np.einsum("np,n->p", shells, x)
    sick = (rho < cutoff) | (rho0 < cutoff)
    out = np.zeros_like(x)
    for m in range(len(x)):
        with np.errstate(all="ignore"):
            integrand = rho * (shells[m] * x[m]) / rho0
        integrand[sick] = 0.0
        out[m] = np.einsum("i,i", molw, integrand)
    return out


def glisa_working_matrix(rho, rho0, shells, molw, nderiv, cutoff=DENSITY_CUTOFF):
    """_working_matrix, glisa.py:411-479: objective, gradient, Hessian."""
    sick = (rho < cutoff) | (rho0 < cutoff)
    with np.errstate(all="ignore"):
        ratio = np.divide(rho, rho0, out=np.zeros_like(rho), where=~sick)
        ln_ratio = np.log(ratio, out=np.zeros_like(ratio), where=~sick)
    f = np.einsum("i,i", molw, rho * ln_ratio)
    if nderiv == 0:
        return f
    M = shells.shape[0]
    grad = np.zeros(M)
    hess = np.zeros((M, M))
    for i in range(M):
        with np.errstate(all="ignore"):
            dfi = (rho * shells[i]) / rho0
        dfi[sick] = 0.0
        grad[i] = -np.einsum("i,i", molw, dfi)
        if nderiv > 1:
            for j in range(i, M):
                with np.errstate(all="ignore"):
                    h = dfi * shells[j] / rho0
                h[sick] = 0.0
                hess[i, j] = hess[j, i] = np.einsum("i,i", molw, h)
    return (f, grad) if nderiv == 1 else (f, grad, hess)


def glisa(coords, numbers, pseudo, grid, rho, basis_func="gauss", solver="sc", threshold=1e-6,
          maxiter=100, cutoff=DENSITY_CUTOFF):  # fmt: skip
    """GlobalLinearISAWPart(solver='sc' | 'newton').do_partitioning(), grid_type=1
    (glisa.py:248-281, 805-848, 617-803 mode='exact')."""
    from scipy.linalg import solve

    natom = len(numbers)
    basis, ranges, propars, shells = glisa_setup(coords, numbers, pseudo, grid, rho, basis_func)
    molw = grid.weights

    def proatom(a, par, r):
        return lisa_proatom(basis[int(numbers[a])], par, r)

    hist = {"propars": [], "entropies": [], "changes": []}
    it = 0
    while True:
        old = propars.copy()
        rho0 = np.einsum("np,n->p", shells, old)
        if solver == "sc":
            propars = glisa_function_g(old, shells, rho, molw, cutoff)
        else:  # exact Newton: delta = solve(H, -1 - grad); propars += delta
            _, grad, hess = glisa_working_matrix(rho, rho0, shells, molw, 2, cutoff)
            propars = old + solve(hess, -1 - grad, assume_a="sym")
        change = _radial_change(grid.atgrids, ranges, proatom, propars, old)
        hist["entropies"].append(entropy(molw, rho, rho0, cutoff))
        hist["changes"].append(change)
        hist["propars"].append(propars.copy())
        it += 1
        if change < threshold:
            break
        if solver != "sc" and it >= maxiter:
            raise RuntimeError("Not converged!")
    # update_at_weights(force_on_molgrid=True) + charges, glisa.py:267-278
    dist = [distances(grid.points, coords[a]) for a in range(natom)]
    promol, weights = stockholder_weights(
        grid, lambda a: proatom(a, propars[ranges[a] : ranges[a + 1]], dist[a]), natom, owner_only=False
    )
    charges = np.array([pseudo[a] - np.einsum("i,i", molw, weights[a] * rho) for a in range(natom)])
    return {
        "niter": it, "charges": charges, "propars": propars, "promoldens": promol,
        "at_weights": [weights[a][grid.indices[a] : grid.indices[a + 1]] for a in range(natom)],
        "history_propars": np.array(hist["propars"]), "history_entropies": np.array(hist["entropies"]),
        "history_changes": np.array(hist["changes"]),
    }  # fmt: skip
