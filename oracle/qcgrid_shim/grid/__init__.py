"""ORACLE / TEST INFRASTRUCTURE ONLY -- restatement of the subset of the third-party ``qc-grid``
package (pin in the reference: ``qc-grid>=0.0.9``, /root/reference/pyproject.toml:38) that
``horton_part`` touches.  qc-grid is NOT vendored under /root/reference and is not installed in
this image, so its published semantics are restated here from its documented behaviour and from
the reference's call sites (SURVEY.md section 8c lists them).  With this directory first on
``sys.path`` the *unmodified* reference package (/root/reference/src/horton_part) imports and
runs; ``oracle/gen_golden.py`` uses exactly that to produce tests/golden/*.

Nothing in the product package ``horton_part_b200`` may import this module.

Restated call sites (reference file:line):
  Grid.integrate                 core/base.py:139,264,298,326  core/stockholder.py:150,217
                                 core/iterstock.py:41,44       glisa.py:188,275,449,461,469,873
  AtomGrid.spherical_average     mbis.py:179   gisa.py:287   isa.py:109
  Grid.moments                   core/base.py:374-402
  AtomGrid.radial_component_splines  core/base.py:657
  Grid.get_localgrid             core/stockholder.py:87 (commented spec), tests/test_alisa.py:110
  BeckeWeights                   becke.py:107-116  scripts/generate_density.py:102-111
  radial transforms, 1-D grids   core/basis.py:372-373  tests/test_wpart.py:46-51  tests/common.py:95
"""

import numpy as np
from scipy.integrate import lebedev_rule
from scipy.interpolate import CubicSpline
from scipy.spatial import cKDTree

__all__ = [
    "Grid",
    "LocalGrid",
    "OneDGrid",
    "UniformInteger",
    "GaussChebyshev",
    "ExpRTransform",
    "PowerRTransform",
    "BeckeRTransform",
    "LinearFiniteRTransform",
    "AngularGrid",
    "AtomGrid",
    "MolGrid",
    "BeckeWeights",
    "LEBEDEV_DEGREES",
    "LEBEDEV_NPOINTS",
]

# Lebedev-Laikov rules: algebraic degree -> number of points on the sphere.
LEBEDEV_NPOINTS = {
    3: 6, 5: 14, 7: 26, 9: 38, 11: 50, 13: 74, 15: 86, 17: 110, 19: 146, 21: 170, 23: 194,
    25: 230, 27: 266, 29: 302, 31: 350, 35: 434, 41: 590, 47: 770, 53: 974, 59: 1202,
    65: 1454, 71: 1730, 77: 2030, 83: 2354, 89: 2702, 95: 3074, 101: 3470, 107: 3890,
    113: 4334, 119: 4802, 125: 5294, 131: 5810,
}  # fmt: skip
# the name horton_part imports (core/stockholder.py:26,134): degree -> size
LEBEDEV_DEGREES = dict(LEBEDEV_NPOINTS)
_SIZE_TO_DEGREE = {size: deg for deg, size in LEBEDEV_NPOINTS.items()}


class Grid:
    """Points + quadrature weights."""

    def __init__(self, points, weights):
        points = np.asarray(points, dtype=float)
        weights = np.asarray(weights, dtype=float)
        if len(points) != len(weights):
            raise ValueError("points and weights differ in length")
        self._points = points
        self._weights = weights
        self._kdtree = None

    @property
    def points(self):
        return self._points

    @property
    def weights(self):
        return self._weights

    @property
    def size(self):
        return self._weights.size

    def __getitem__(self, index):
        if np.isscalar(index):
            return self.__class__(np.array([self.points[index]]), np.array([self.weights[index]]))
        return self.__class__(self.points[index], self.weights[index])

    def integrate(self, *value_arrays):
        """sum_p w_p * prod_k f_k(p)  (one fused contraction, like qc-grid's einsum)."""
        if len(value_arrays) == 0:
            raise ValueError("No array is given to integrate.")
        for i, array in enumerate(value_arrays):
            if not isinstance(array, np.ndarray):
                raise TypeError(f"Arg {i} is {type(array)}, need Numpy Array.")
            if array.shape != (self.size,):
                raise ValueError(f"Arg {i} need to be of shape ({self.size},).")
        return np.einsum("i" + ",i" * len(value_arrays), self.weights, *value_arrays)

    def get_localgrid(self, center, radius):
        """Sub-grid of the points within ``radius`` of ``center`` (kd-tree ball query, p=2)."""
        center = np.asarray(center)
        if self._kdtree is None:
            pts = self.points.reshape(self.size, -1)
            self._kdtree = cKDTree(pts)
        if center.ndim == 0:
            center = center.reshape(1)
        if np.isinf(radius):
            indices = np.arange(self.size)
        else:
            indices = np.array(self._kdtree.query_ball_point(center, radius, p=2.0), dtype=int)
        return LocalGrid(self.points[indices], self.weights[indices], center, indices)

    def moments(self, orders, centers, func_vals, type_mom="cartesian", return_orders=False):
        """Multipole moments of ``func_vals`` about each center; result shape (L, ncenter).

        cartesian: all (nx, ny, nz) with nx+ny+nz = l for l = 0..orders, HORTON order
                   (alphabetical within one l: xx xy xz yy yz zz).
        radial:    int |r - c|^n f for n = 0..orders.
        pure:      real regular solid harmonics (Racah normalised), HORTON-2 order
                   C_l0, C_l1, S_l1, C_l2, S_l2, ...
        """
        func_vals = np.asarray(func_vals)
        centers = np.atleast_2d(centers)
        cols = []
        for c in centers:
            d = self.points - c
            if type_mom == "cartesian":
                pows = _cartesian_powers(orders)
                basis = np.prod(d[None, :, :] ** pows[:, None, :], axis=2)
            elif type_mom == "radial":
                r = np.linalg.norm(d, axis=1)
                basis = r[None, :] ** np.arange(orders + 1)[:, None]
            elif type_mom == "pure":
                basis = _solid_harmonics(orders, d)
            else:
                raise ValueError(f"unknown type_mom {type_mom}")
            cols.append(np.einsum("lp,p,p->l", basis, func_vals, self.weights))
        result = np.array(cols).T
        if return_orders:
            return result, (_cartesian_powers(orders) if type_mom == "cartesian" else None)
        return result


class LocalGrid(Grid):
    def __init__(self, points, weights, center, indices=None):
        super().__init__(points, weights)
        self._center = center
        self._indices = indices

    @property
    def center(self):
        return self._center

    @property
    def indices(self):
        return self._indices


def _cartesian_powers(lmax):
    rows = []
    for l in range(lmax + 1):
        for nx in range(l, -1, -1):
            for ny in range(l - nx, -1, -1):
                rows.append((nx, ny, l - nx - ny))
    return np.array(rows, dtype=int)


def _solid_harmonics(lmax, d):
    """Real regular solid harmonics R_lm(x,y,z), Racah normalisation, via the standard
    (z, r^2) recursion on (C_mm, S_mm); rows ordered C_00; C_10 C_11 S_11; C_20 C_21 S_21 ..."""
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    r2 = x * x + y * y + z * z
    # sectoral terms, un-normalised:  A_m + i B_m = (x + i y)^m
    A = [np.ones_like(x)]
    B = [np.zeros_like(x)]
    for m in range(1, lmax + 1):
        A.append(x * A[m - 1] - y * B[m - 1])
        B.append(x * B[m - 1] + y * A[m - 1])
    # Pi_l^m(z, r^2): r^(l-m) d^m P_l / d(cos)^m evaluated at z/r  (polynomial in z, r2)
    Pi = {}
    for m in range(lmax + 1):
        dfact = 1.0
        for k in range(1, 2 * m, 2):
            dfact *= k  # (2m-1)!!
        Pi[(m, m)] = dfact * np.ones_like(z)
        if m + 1 <= lmax:
            Pi[(m + 1, m)] = (2 * m + 1) * z * Pi[(m, m)]
        for l in range(m + 2, lmax + 1):
            Pi[(l, m)] = ((2 * l - 1) * z * Pi[(l - 1, m)] - (l + m - 1) * r2 * Pi[(l - 2, m)]) / (l - m)
    from math import factorial, sqrt

    rows = []
    for l in range(lmax + 1):
        rows.append(Pi[(l, 0)])
        for m in range(1, l + 1):
            norm = sqrt(2.0 * factorial(l - m) / factorial(l + m))
            rows.append(norm * Pi[(l, m)] * A[m])
            rows.append(norm * Pi[(l, m)] * B[m])
    return np.array(rows)


class OneDGrid(Grid):
    def __init__(self, points, weights, domain=None):
        super().__init__(points, weights)
        self._domain = domain

    @property
    def domain(self):
        return self._domain


class UniformInteger(OneDGrid):
    """x_i = i, w_i = 1, i = 0..n-1."""

    def __init__(self, npoints):
        super().__init__(np.arange(npoints, dtype=float), np.ones(npoints), (0, np.inf))


class GaussChebyshev(OneDGrid):
    """Gauss-Chebyshev (first kind) nodes on [-1, 1], ascending, with the 1/sqrt(1-x^2) measure
    folded into the weights so that the rule integrates plain f(x) dx."""

    def __init__(self, npoints):
        x, w = np.polynomial.chebyshev.chebgauss(npoints)
        w = w * np.sqrt(1.0 - x**2)
        super().__init__(x[::-1], w[::-1], (-1, 1))


class _RTransform:
    def transform_1d_grid(self, oned_grid):
        x = oned_grid.points
        return OneDGrid(self.transform(x), self.deriv(x) * oned_grid.weights, self.codomain)

    codomain = (0, np.inf)


class ExpRTransform(_RTransform):
    """r = rmin * exp(alpha x), alpha = ln(rmax/rmin)/b."""

    def __init__(self, rmin, rmax, b=None):
        self.rmin, self.rmax, self.b = rmin, rmax, b
        self.alpha = np.log(rmax / rmin) / b

    def transform(self, x):
        return self.rmin * np.exp(x * self.alpha)

    def deriv(self, x):
        return self.transform(x) * self.alpha


class PowerRTransform(_RTransform):
    """r = rmin * (x+1)^p, p = ln(rmax/rmin)/ln(b+1)."""

    def __init__(self, rmin, rmax, b=None):
        self.rmin, self.rmax, self.b = rmin, rmax, b
        self.power = (np.log(rmax) - np.log(rmin)) / np.log(b + 1)

    def transform(self, x):
        return self.rmin * np.power(x + 1, self.power)

    def deriv(self, x):
        return self.power * self.rmin * np.power(x + 1, self.power - 1)


class LinearFiniteRTransform(_RTransform):
    """r = (rmax-rmin)/2 * (1+x) + rmin on x in [-1, 1]."""

    def __init__(self, rmin, rmax):
        self.rmin, self.rmax = rmin, rmax

    def transform(self, x):
        return (1 + x) * (self.rmax - self.rmin) / 2 + self.rmin

    def deriv(self, x):
        return np.ones_like(x) * (self.rmax - self.rmin) / 2


class BeckeRTransform(_RTransform):
    """Becke's map of [-1, 1] to [rmin, inf): r = R (1+x)/(1-x) + rmin."""

    def __init__(self, rmin, R, trim_inf=True):
        self.rmin, self.R, self.trim_inf = rmin, R, trim_inf

    def transform(self, x):
        with np.errstate(divide="ignore"):
            r = self.R * (1 + x) / (1 - x) + self.rmin
        if self.trim_inf:
            r = np.clip(r, None, 1e16)
        return r

    def deriv(self, x):
        with np.errstate(divide="ignore"):
            d = 2 * self.R / ((1 - x) ** 2)
        if self.trim_inf:
            d = np.clip(d, None, 1e16)
        return d


class AngularGrid(Grid):
    """Lebedev-Laikov rule on the unit sphere, weights summing to 4 pi."""

    _cache = {}

    def __init__(self, degree=None, size=None):
        if degree is None:
            degree = _SIZE_TO_DEGREE[int(size)]
        if degree not in AngularGrid._cache:
            x, w = lebedev_rule(int(degree))
            AngularGrid._cache[degree] = (np.ascontiguousarray(x.T), np.array(w))
        pts, wts = AngularGrid._cache[degree]
        self.degree = degree
        super().__init__(pts, wts)


class AtomGrid(Grid):
    """Radial x angular product grid: shell i holds AngularGrid(degrees[i]) scaled by r_i; the
    weights are w_ang * w_rad_i * r_i^2; ``indices`` are the shell boundaries."""

    def __init__(self, rgrid, *, degrees=None, sizes=None, center=None, rotate=0):
        nshell = rgrid.size
        if degrees is None:
            if sizes is None:
                raise ValueError("degrees or sizes is needed")
            sizes = [int(s) for s in np.atleast_1d(sizes)]
            degrees = [_SIZE_TO_DEGREE[s] for s in sizes]
        else:
            degrees = [int(d) for d in np.atleast_1d(degrees)]
        if len(degrees) == 1:
            degrees = degrees * nshell
        if len(degrees) != nshell:
            raise ValueError("need one angular degree per radial point")
        if rotate not in (0, False):
            raise NotImplementedError("rotated atomic grids are not part of the shim")
        self._rgrid = rgrid
        self._degrees = degrees
        self._center = np.zeros(3) if center is None else np.asarray(center, dtype=float)
        blocks_p, blocks_w, bounds = [], [], [0]
        for r_i, w_i, deg in zip(rgrid.points, rgrid.weights, degrees):
            ang = AngularGrid(degree=deg)
            blocks_p.append(ang.points * r_i)
            blocks_w.append(ang.weights * w_i * r_i**2)
            bounds.append(bounds[-1] + ang.size)
        self._indices = np.array(bounds)
        super().__init__(np.vstack(blocks_p) + self._center, np.concatenate(blocks_w))

    @classmethod
    def from_pruned(cls, rgrid, radius, *, sectors_r, sectors_degree=None, sectors_size=None,
                    center=None, rotate=0):  # fmt: skip
        sectors_r = np.asarray(sectors_r) * radius
        if sectors_degree is None:
            sectors_degree = [_SIZE_TO_DEGREE[int(s)] for s in sectors_size]
        which = np.searchsorted(sectors_r, rgrid.points)
        degrees = np.asarray(sectors_degree)[which]
        return cls(rgrid, degrees=degrees, center=center, rotate=rotate)

    rgrid = property(lambda self: self._rgrid)
    degrees = property(lambda self: self._degrees)
    center = property(lambda self: self._center)
    indices = property(lambda self: self._indices)
    n_shells = property(lambda self: len(self._degrees))
    l_max = property(lambda self: int(np.max(self._degrees)))

    def get_shell_grid(self, index, r_sq=True):
        lo, hi = self._indices[index], self._indices[index + 1]
        w = self.weights[lo:hi]
        if not r_sq:
            w = w / self._rgrid.points[index] ** 2
        return Grid(self.points[lo:hi], w)

    def integrate_angular_coordinates(self, func_vals):
        """Per-shell sums of f*w with the radial factor r_i^2 w_rad_i divided out again."""
        prod = func_vals * self.weights
        shell_sums = np.array(
            [np.sum(prod[..., self._indices[i] : self._indices[i + 1]], axis=-1)
             for i in range(self.n_shells)]
        )  # fmt: skip
        shell_sums = np.moveaxis(shell_sums, 0, -1)
        with np.errstate(divide="ignore", invalid="ignore"):
            shell_sums /= self._rgrid.points**2 * self._rgrid.weights
        shell_sums[..., np.abs(self._rgrid.points) < 1e-8] = 0.0
        return shell_sums

    def spherical_average(self, func_vals):
        """(1/4pi) * angular integral per shell, returned as a cubic spline over the radial nodes."""
        f_radial = self.integrate_angular_coordinates(func_vals)
        f_radial /= 4.0 * np.pi
        return CubicSpline(x=self._rgrid.points, y=f_radial)

    def radial_component_splines(self, func_vals):
        """Splines of the real-spherical-harmonic components f_lm(r), HORTON-2 order, l <= l_max/2."""
        lmax = self.l_max // 2
        d = self.points - self._center
        r = np.linalg.norm(d, axis=1)
        with np.errstate(divide="ignore", invalid="ignore"):
            unit = np.where(r[:, None] > 0, d / r[:, None], 0.0)
        # real spherical harmonics Y_lm = sqrt((2l+1)/4pi) * R_lm(unit vector)
        ylm = _solid_harmonics(lmax, unit)
        norms = np.concatenate([[np.sqrt((2 * l + 1) / (4 * np.pi))] * (2 * l + 1) for l in range(lmax + 1)])
        ylm = ylm * norms[:, None]
        comps = self.integrate_angular_coordinates(ylm * func_vals[None, :])  # (nlm, nrad)
        return [CubicSpline(x=self._rgrid.points, y=c) for c in comps]


# Bragg-Slater radii in angstrom (Slater 1964); H uses 0.25 A as in qc-grid.  Missing noble gases
# are filled by callers (scripts/generate_density.py:104-109 does that for He, Ne, Ar, Kr, Xe).
_BRAGG_ANGSTROM = {
    1: 0.25, 3: 1.45, 4: 1.05, 5: 0.85, 6: 0.70, 7: 0.65, 8: 0.60, 9: 0.50, 11: 1.80, 12: 1.50,
    13: 1.25, 14: 1.10, 15: 1.00, 16: 1.00, 17: 1.00, 19: 2.20, 20: 1.80, 21: 1.60, 22: 1.40,
    23: 1.35, 24: 1.40, 25: 1.40, 26: 1.40, 27: 1.35, 28: 1.35, 29: 1.35, 30: 1.35, 31: 1.30,
    32: 1.25, 33: 1.15, 34: 1.15, 35: 1.15,
}  # fmt: skip
_ANGSTROM = 1.0 / 0.52917721092


class BeckeWeights:
    """Becke's fuzzy-cell weights with atomic-size adjustment (a_ij clipped to +-0.45) and a
    polynomial switch iterated ``order`` times."""

    def __init__(self, radii=None, order=3):
        self._radii = {z: r * _ANGSTROM for z, r in _BRAGG_ANGSTROM.items()}
        if radii:
            self._radii.update(radii)
        self._order = order

    def _cell_functions(self, points, atcoords, atnums):
        radii = np.array([self._radii[int(z)] for z in atnums])
        chi = radii[:, None] / radii[None, :]
        u = (chi - 1) / (chi + 1)
        a = np.clip(u / (u**2 - 1), -0.45, 0.45)
        rab = np.linalg.norm(atcoords[:, None, :] - atcoords[None, :, :], axis=-1)
        dist = np.linalg.norm(points[:, None, :] - atcoords[None, :, :], axis=-1)  # (np, natom)
        with np.errstate(divide="ignore", invalid="ignore"):
            mu = (dist[:, :, None] - dist[:, None, :]) / rab[None, :, :]
        nu = mu + a[None, :, :] * (1 - mu**2)
        for _ in range(self._order):
            nu = 1.5 * nu - 0.5 * nu**3
        s = 0.5 * (1 - nu)
        s[np.isnan(s)] = 1.0
        return np.prod(s, axis=-1)  # (np, natom)

    def compute_atom_weight(self, points, atcoords, atnums, select, cutoff=None):
        out = np.empty(len(points))
        for lo in range(0, len(points), 4096):
            cells = self._cell_functions(points[lo : lo + 4096], atcoords, atnums)
            out[lo : lo + 4096] = cells[:, select] / cells.sum(axis=-1)
        return out

    def generate_weights(self, points, atcoords, atnums, select=None, pt_ind=None):
        out = np.empty(len(points))
        if pt_ind is None:
            return self.compute_atom_weight(points, atcoords, atnums, select)
        for iatom in range(len(atnums)):
            lo, hi = pt_ind[iatom], pt_ind[iatom + 1]
            out[lo:hi] = self.compute_atom_weight(points[lo:hi], atcoords, atnums, iatom)
        return out

    def __call__(self, points, atcoords, atnums, pt_ind):
        return self.generate_weights(points, atcoords, atnums, pt_ind=pt_ind)


class MolGrid(Grid):
    """Concatenation of atomic grids, weights multiplied by atom-in-molecule (Becke) weights."""

    def __init__(self, atnums, atgrids, aim_weights, store=False):
        atnums = np.asarray(atnums)
        points = np.vstack([g.points for g in atgrids])
        atom_weights = np.concatenate([g.weights for g in atgrids])
        self._indices = np.concatenate([[0], np.cumsum([g.size for g in atgrids])])
        self._atcoords = np.array([g.center for g in atgrids])
        self._atnums = atnums
        if callable(aim_weights):
            aim_weights = aim_weights(points, self._atcoords, atnums, self._indices)
        aim_weights = np.asarray(aim_weights, dtype=float)
        if aim_weights.shape != atom_weights.shape:
            raise ValueError("aim_weights has the wrong size")
        self._aim_weights = aim_weights
        self._atweights = atom_weights
        self._atgrids = list(atgrids) if store else None
        super().__init__(points, atom_weights * aim_weights)

    @classmethod
    def from_size(cls, atnums, atcoords, size, rgrid=None, aim_weights=None, rotate=0, store=False):
        atgrids = [AtomGrid(rgrid, sizes=[size], center=c, rotate=rotate) for c in np.asarray(atcoords)]
        return cls(atnums, atgrids, aim_weights, store=store)

    indices = property(lambda self: self._indices)
    aim_weights = property(lambda self: self._aim_weights)
    atgrids = property(lambda self: self._atgrids)
    atcoords = property(lambda self: self._atcoords)
    atweights = property(lambda self: self._atweights)
    atnums = property(lambda self: self._atnums)

    def get_atomic_grid(self, index):
        if self._atgrids is None:
            raise ValueError("Atomic grids were not stored (store=False).")
        return self._atgrids[index]

    def __getitem__(self, index):
        return self.get_atomic_grid(index)
