"""Host-side molecular integration grids in the layout the partitioning classes consume.

The reference receives its grids from the third-party ``qc-grid`` package (un-vendored; call
sites: /root/reference/src/horton_part/core/base.py:26,296,461,621-634,
scripts/generate_density.py:102-111, scripts/partition_density.py:65-85).  The partitioning
classes only duck-type a handful of attributes:

    MolGrid : points (Npts,3)  weights (Npts,)  size  indices (natom+1,)  atgrids  aim_weights
    AtomGrid: points  weights  size  indices (nshell+1,)  degrees  rgrid  center  l_max
    OneDGrid: points  weights  size

so a real qc-grid ``MolGrid`` can be passed unchanged; this module exists because qc-grid is not
available on the GPU box and the benchmark / tests need to build the same grids there.  Grid
*construction* is input preparation, not the hot path: everything here is NumPy; the Lebedev
tables come from ``scipy.integrate.lebedev_rule``.  The 2,000-atom benchmark grid is built from
one shared per-element template, so construction is O(Npts) memory traffic only.
"""

from __future__ import annotations

import numpy as np

__all__ = [
    "OneDGrid",
    "UniformInteger",
    "GaussChebyshev",
    "ExpRTransform",
    "PowerRTransform",
    "BeckeRTransform",
    "AtomGrid",
    "MolGrid",
    "BeckeWeights",
    "DeviceBeckeWeights",
    "lebedev_size_to_degree",
    "LEBEDEV_DEGREES",
]

_LEB = (
    (3, 6), (5, 14), (7, 26), (9, 38), (11, 50), (13, 74), (15, 86), (17, 110), (19, 146),
    (21, 170), (23, 194), (25, 230), (27, 266), (29, 302), (31, 350), (35, 434), (41, 590),
    (47, 770), (53, 974), (59, 1202), (65, 1454), (71, 1730), (77, 2030), (83, 2354),
    (89, 2702), (95, 3074), (101, 3470), (107, 3890), (113, 4334), (119, 4802), (125, 5294),
    (131, 5810),
)  # fmt: skip
LEBEDEV_DEGREES = {deg: n for deg, n in _LEB}
_DEGREE_OF_SIZE = {n: deg for deg, n in _LEB}
_sphere_rules = {}


def lebedev_size_to_degree(size: int) -> int:
    return _DEGREE_OF_SIZE[int(size)]


def _sphere_rule(degree: int):
    """(npt,3) unit vectors and weights (sum = 4 pi) of the Lebedev rule of this degree."""
    rule = _sphere_rules.get(degree)
    if rule is None:
        from scipy.integrate import lebedev_rule

        xyz, w = lebedev_rule(int(degree))
        rule = _sphere_rules[degree] = (np.ascontiguousarray(xyz.T), np.ascontiguousarray(w))
    return rule


class _PointSet:
    def __init__(self, points, weights):
        self.points = np.ascontiguousarray(points, dtype=np.float64)
        self.weights = np.ascontiguousarray(weights, dtype=np.float64)
        if self.points.shape[0] != self.weights.shape[0]:
            raise ValueError("points and weights differ in length")

    @property
    def size(self):
        return self.weights.shape[0]

    def integrate(self, *arrays):
        """Quadrature of the point-wise product of ``arrays`` (same contraction as qc-grid)."""
        if not arrays:
            raise ValueError("No array is given to integrate.")
        for k, a in enumerate(arrays):
            if not isinstance(a, np.ndarray):
                raise TypeError(f"Arg {k} is {type(a)}, need Numpy Array.")
            if a.shape != (self.size,):
                raise ValueError(f"Arg {k} need to be of shape ({self.size},).")
        return np.einsum("i" + ",i" * len(arrays), self.weights, *arrays)


class OneDGrid(_PointSet):
    def __init__(self, points, weights, domain=None):
        super().__init__(points, weights)
        self.domain = domain


class UniformInteger(OneDGrid):
    def __init__(self, npoints):
        super().__init__(np.arange(npoints, dtype=np.float64), np.ones(npoints), (0, np.inf))


class GaussChebyshev(OneDGrid):
    def __init__(self, npoints):
        k = np.arange(npoints, 0, -1)
        theta = (2 * k - 1) * np.pi / (2 * npoints)
        nodes = np.cos(theta)  # ascending in x
        # plain-measure weights: (pi/n) * sqrt(1 - x^2)
        super().__init__(nodes, (np.pi / npoints) * np.sqrt(1.0 - nodes * nodes), (-1, 1))


class _RadialMap:
    def transform_1d_grid(self, oned):
        return OneDGrid(self.transform(oned.points), self.deriv(oned.points) * oned.weights, (0, np.inf))


class ExpRTransform(_RadialMap):
    def __init__(self, rmin, rmax, b):
        self.rmin, self.rmax, self.b = rmin, rmax, b

    def transform(self, x):
        return self.rmin * np.exp(x * (np.log(self.rmax / self.rmin) / self.b))

    def deriv(self, x):
        return self.transform(x) * (np.log(self.rmax / self.rmin) / self.b)


class PowerRTransform(_RadialMap):
    def __init__(self, rmin, rmax, b):
        self.rmin, self.rmax, self.b = rmin, rmax, b
        self.power = (np.log(rmax) - np.log(rmin)) / np.log(b + 1)

    def transform(self, x):
        return self.rmin * np.power(x + 1, self.power)

    def deriv(self, x):
        return self.power * self.rmin * np.power(x + 1, self.power - 1)


class BeckeRTransform(_RadialMap):
    def __init__(self, rmin, R):
        self.rmin, self.R = rmin, R

    def transform(self, x):
        with np.errstate(divide="ignore"):
            return np.minimum(self.R * (1 + x) / (1 - x) + self.rmin, 1e16)

    def deriv(self, x):
        with np.errstate(divide="ignore"):
            return np.minimum(2 * self.R / ((1 - x) ** 2), 1e16)


class _AtomTemplate:
    """Origin-centred atomic grid for one (radial grid, angular degrees) pair, shared by atoms."""

    def __init__(self, rgrid, degrees):
        r, wr = rgrid.points, rgrid.weights
        sizes = np.array([LEBEDEV_DEGREES[d] for d in degrees])
        self.indices = np.concatenate([[0], np.cumsum(sizes)])
        n = int(self.indices[-1])
        self.offsets = np.empty((n, 3))
        self.weights = np.empty(n)
        for i, d in enumerate(degrees):
            unit, wang = _sphere_rule(d)
            lo, hi = self.indices[i], self.indices[i + 1]
            self.offsets[lo:hi] = unit * r[i]
            self.weights[lo:hi] = wang * wr[i] * r[i] ** 2


class AtomGrid(_PointSet):
    """Radial x Lebedev product grid around ``center``; weights = w_ang * w_rad * r^2."""

    def __init__(self, rgrid, *, degrees=None, sizes=None, center=None, rotate=0, _template=None):
        if rotate not in (0, False):
            raise NotImplementedError("rotated atomic grids are not supported")
        if degrees is None:
            if sizes is None:
                raise ValueError("degrees or sizes is needed")
            degrees = [lebedev_size_to_degree(s) for s in np.atleast_1d(sizes)]
        degrees = [int(d) for d in np.atleast_1d(degrees)]
        if len(degrees) == 1:
            degrees = degrees * rgrid.size
        if len(degrees) != rgrid.size:
            raise ValueError("need one angular degree per radial point")
        tpl = _template if _template is not None else _AtomTemplate(rgrid, degrees)
        self.rgrid = rgrid
        self.degrees = degrees
        self.center = np.zeros(3) if center is None else np.asarray(center, dtype=np.float64)
        self.indices = tpl.indices
        super().__init__(tpl.offsets + self.center, tpl.weights)

    @property
    def n_shells(self):
        return len(self.degrees)

    @property
    def l_max(self):
        return max(self.degrees)

    def integrate_angular_coordinates(self, func_vals):
        prod = func_vals * self.weights
        sums = np.add.reduceat(prod, self.indices[:-1])
        with np.errstate(divide="ignore", invalid="ignore"):
            sums = sums / (self.rgrid.points**2 * self.rgrid.weights)
        sums[np.abs(self.rgrid.points) < 1e-8] = 0.0
        return sums

    def spherical_average(self, func_vals):
        from scipy.interpolate import CubicSpline

        return CubicSpline(self.rgrid.points, self.integrate_angular_coordinates(func_vals) / (4 * np.pi))


# Bragg-Slater radii (angstrom); noble gases are absent, as in Becke's scheme.
_BRAGG_SLATER = {
    1: 0.25, 3: 1.45, 4: 1.05, 5: 0.85, 6: 0.70, 7: 0.65, 8: 0.60, 9: 0.50, 11: 1.80, 12: 1.50,
    13: 1.25, 14: 1.10, 15: 1.00, 16: 1.00, 17: 1.00, 19: 2.20, 20: 1.80, 21: 1.60, 22: 1.40,
    23: 1.35, 24: 1.40, 25: 1.40, 26: 1.40, 27: 1.35, 28: 1.35, 29: 1.35, 30: 1.35, 31: 1.30,
    32: 1.25, 33: 1.15, 34: 1.15, 35: 1.15,
}  # fmt: skip
_BOHR_PER_ANGSTROM = 1.0 / 0.52917721092


class BeckeWeights:
    """Becke fuzzy-cell weights (host NumPy, O(natom^2 Npts): small molecules only)."""

    def __init__(self, radii=None, order=3):
        self._radii = {z: v * _BOHR_PER_ANGSTROM for z, v in _BRAGG_SLATER.items()}
        if radii:
            self._radii.update(radii)
        self._order = order

    def __call__(self, points, atcoords, atnums, pt_ind):
        atcoords = np.asarray(atcoords, dtype=np.float64)
        R = np.array([self._radii[int(z)] for z in atnums])
        chi = R[:, None] / R[None, :]
        u = (chi - 1) / (chi + 1)
        a = np.clip(u / (u * u - 1), -0.45, 0.45)
        rab = np.sqrt(((atcoords[:, None, :] - atcoords[None, :, :]) ** 2).sum(-1))
        out = np.empty(len(points))
        for owner in range(len(atnums)):
            for lo in range(int(pt_ind[owner]), int(pt_ind[owner + 1]), 8192):
                hi = min(lo + 8192, int(pt_ind[owner + 1]))
                dist = np.sqrt(((points[lo:hi, None, :] - atcoords[None, :, :]) ** 2).sum(-1))
                with np.errstate(divide="ignore", invalid="ignore"):
                    mu = (dist[:, :, None] - dist[:, None, :]) / rab[None]
                nu = mu + a[None] * (1 - mu * mu)
                for _ in range(self._order):
                    nu = 1.5 * nu - 0.5 * nu**3
                s = 0.5 * (1 - nu)
                s[np.isnan(s)] = 1.0
                cell = s.prod(axis=-1)
                out[lo:hi] = cell[:, owner] / cell.sum(axis=-1)
        return out


class DeviceBeckeWeights(BeckeWeights):
    """Becke weights computed by the CUDA kernel ``hp_becke_weights`` (same formula as
    :class:`BeckeWeights`; one thread per point, negligible cells pruned against the nearest atom).
    Use as the ``aim_weights`` callable of :class:`MolGrid` for systems where the O(natom^2 Npts)
    NumPy version is out of reach."""

    def __init__(self, radii=None, order=3, device=None, chunk=1 << 24):
        super().__init__(radii, order)
        self._device, self._chunk = device, chunk

    def __call__(self, points, atcoords, atnums, pt_ind):
        import torch

        from . import _lib
        from .core.device import require_cuda, stream_ptr, to_device

        dev = require_cuda(self._device)
        atcoords = np.ascontiguousarray(atcoords, dtype=np.float64)
        natom = len(atnums)
        R = np.array([self._radii[int(z)] for z in atnums])
        chi = R[:, None] / R[None, :]
        u = (chi - 1) / (chi + 1)
        aab = np.clip(u / (u * u - 1), -0.45, 0.45)
        rab = np.sqrt(((atcoords[:, None, :] - atcoords[None, :, :]) ** 2).sum(-1))
        with np.errstate(divide="ignore"):
            inv_rab = 1.0 / rab
        d_inv, d_aab = to_device(inv_rab, dev), to_device(aab, dev)
        d_xyz = to_device(atcoords, dev)
        d_off = to_device(np.asarray(pt_ind, dtype=np.int64), dev)
        out = np.empty(len(points))
        for lo in range(0, len(points), self._chunk):
            hi = min(lo + self._chunk, len(points))
            n = hi - lo
            aos = to_device(points[lo:hi], dev, np.float64)
            px, py, pz, w = (torch.empty(n, dtype=torch.float64, device=dev) for _ in range(4))
            _lib.call("hp_split_points", aos, n, px, py, pz, stream_ptr(dev))
            _lib.call("hp_becke_weights", n, px, py, pz, lo, natom, d_xyz, d_off, d_inv, d_aab,
                      int(self._order), w, stream_ptr(dev))  # fmt: skip
            out[lo:hi] = w.cpu().numpy()
        return out


class MolGrid(_PointSet):
    """Atom-block concatenation of atomic grids; ``weights = atomic weights * aim_weights``."""

    def __init__(self, atnums, atgrids, aim_weights, store=False):
        self.atnums = np.asarray(atnums)
        self.atcoords = np.array([g.center for g in atgrids])
        self.indices = np.concatenate([[0], np.cumsum([g.size for g in atgrids])])
        points = np.concatenate([g.points for g in atgrids])
        atweights = np.concatenate([g.weights for g in atgrids])
        if callable(aim_weights):
            aim_weights = aim_weights(points, self.atcoords, self.atnums, self.indices)
        self.aim_weights = np.asarray(aim_weights, dtype=np.float64)
        if self.aim_weights.shape != atweights.shape:
            raise ValueError("aim_weights has the wrong size")
        self.atweights = atweights
        self.atgrids = list(atgrids) if store else None
        super().__init__(points, atweights * self.aim_weights)

    @classmethod
    def from_size(cls, atnums, atcoords, size, rgrid, aim_weights, rotate=0, store=False):
        degrees = [lebedev_size_to_degree(size)] * rgrid.size
        tpl = _AtomTemplate(rgrid, degrees)
        atgrids = [
            AtomGrid(rgrid, degrees=degrees, center=c, rotate=rotate, _template=tpl)
            for c in np.asarray(atcoords, dtype=np.float64)
        ]
        return cls(atnums, atgrids, aim_weights, store=store)

    def get_atomic_grid(self, index):
        if self.atgrids is None:
            raise ValueError("Atomic grids were not stored (store=False).")
        return self.atgrids[index]

    __getitem__ = get_atomic_grid
