"""Exponential basis functions of the GISA / aLISA / gLISA pro-atoms (host-side helper).

API-compatible with the reference's ``core/basis.py`` (``evaluate_function`` :85-186,
``ExpBasisFuncHelper`` :241-327, ``load_params`` :42-82).  The helper only provides *parameters*
and small radial-grid evaluations (K x nrad per atom, once per job); evaluation on the molecular
grid happens inside the CUDA kernels from the (A, alpha, n) shell table.

    g(r) = c * n * alpha^(3/n) / (4 pi Gamma(3/n)) * exp(-alpha r^n)

n = 2: Gaussians (``gauss`` table), n = 1: Slater functions (``slater`` table).
"""

from __future__ import annotations

import json
import pathlib

import numpy as np
from scipy.special import gamma

__all__ = [
    "BasisFuncHelper",
    "AnalyticBasisFuncHelper",
    "ExpBasisFuncHelper",
    "NumericBasisFuncHelper",
    "evaluate_function",
    "load_params",
    "shell_norm",
]

_TABLES = pathlib.Path(__file__).resolve().parents[1] / "data" / "expbasis_tables.json"


def shell_norm(n, alpha):
    """Normalisation n alpha^(3/n) / (4 pi Gamma(3/n)) of exp(-alpha r^n) (core/basis.py:161)."""
    n = np.asarray(n, dtype=float)
    alpha = np.asarray(alpha, dtype=float)
    return n * alpha ** (3 / n) / (4 * np.pi * gamma(3 / n))


def evaluate_function(n, population, alpha, r, nderiv=0, axis=None):
    """Value (and radial derivative) of population * N(alpha, n) * exp(-alpha r^n).

    Scalar (n, alpha) give one function on ``r``; 1-D arrays give one row per function, summed
    over ``axis`` if requested.  The derivative follows the reference literally:
    ``df = -n r^(n-1) f`` (core/basis.py:173)."""
    if np.any(np.asarray(n) <= 0):
        raise ValueError("n must be positive")
    if np.any(np.asarray(alpha) < 0):
        raise ValueError("alpha must be non-negative")
    if not isinstance(r, np.ndarray):
        raise ValueError("r must be a numpy array")
    prefactor = population * n * alpha ** (3 / n) / (4 * np.pi * gamma(3 / n))
    many = not np.isscalar(n)
    if many:
        assert not (np.isscalar(alpha) or np.isscalar(population))
        prefactor, alpha, n = prefactor[:, None], alpha[:, None], n[:, None]
        r = r[None, :]
    f = prefactor * np.exp(-alpha * r**n)
    df = -n * r ** (n - 1) * f if nderiv > 0 else None
    if many and axis is not None:
        f = f.sum(axis=axis)
        if df is not None:
            df = df.sum(axis=axis)
    if nderiv == 0:
        return f
    if nderiv == 1:
        return f, df
    raise NotImplementedError


def load_params(filename, extension="json"):
    """Read {Z: [orders, exponents, initials]} from a JSON or YAML file (the reference's format)."""
    assert extension.lower() in ("json", "yaml"), "Format must be 'json' or 'yaml'."
    with open(filename) as fh:
        if extension.lower() == "json":
            data = json.load(fh)
        else:
            import yaml

            data = yaml.safe_load(fh)
    orders, exps, inits = {}, {}, {}
    for number, rows in data.items():
        z = int(number)
        orders[z], exps[z], inits[z] = (np.asarray(rows[i]) for i in range(3))
    return orders, exps, inits


class BasisFuncHelper:
    def __init__(self, initials, *args, **kwargs):
        self._initials = initials

    initials = property(lambda self: self._initials)

    def get_nshell(self, number):
        raise NotImplementedError

    def get_initial(self, number, ishell=None):
        return np.asarray(self.initials[number]) if ishell is None else self.initials[number][ishell]

    def compute_proshell_dens(self, number, ishell, population, points, nderiv=0):
        raise NotImplementedError

    def compute_proatom_dens(self, number, populations, points, nderiv=0):
        """Sum of the atom's shells with the given populations (sequential in the shell index)."""
        if nderiv not in (0, 1):
            raise NotImplementedError if nderiv > 1 else RuntimeError(
                "The argument `nderiv` should only be 0 or 1."
            )
        y = d = 0.0
        for k in range(self.get_nshell(number)):
            res = self.compute_proshell_dens(number, k, populations[k], points, nderiv)
            if nderiv == 0:
                y += res
            else:
                y += res[0]
                d += res[1]
        return y if nderiv == 0 else (y, d)


class AnalyticBasisFuncHelper(BasisFuncHelper):
    pass


class ExpBasisFuncHelper(AnalyticBasisFuncHelper):
    """Per-element (order, exponent, initial coefficient) tables."""

    def __init__(self, exponents_orders, exponents=None, initials=None):
        super().__init__(initials)
        self._orders = exponents_orders
        self._exponents = exponents

    orders = property(lambda self: self._orders)
    exponents = property(lambda self: self._exponents)

    def get_nshell(self, number):
        return len(self.exponents[number])

    def get_order(self, number, ishell=None):
        return self.orders[number] if ishell is None else self.orders[number][ishell]

    def get_exponent(self, number, ishell=None):
        return np.asarray(self.exponents[number]) if ishell is None else self.exponents[number][ishell]

    def compute_proshell_dens(self, number, ishell, population, points, nderiv=0):
        return evaluate_function(
            self.get_order(number, ishell), population, self.get_exponent(number, ishell), points, nderiv
        )

    @classmethod
    def from_function_type(cls, func_type="gauss"):
        assert func_type in ("gauss", "slater")
        table = json.loads(_TABLES.read_text())[func_type]
        orders = {int(z): np.asarray(t["orders"]) for z, t in table.items()}
        exps = {int(z): np.asarray(t["exponents"]) for z, t in table.items()}
        inits = {int(z): np.asarray(t["initials"]) for z, t in table.items()}
        return cls(orders, exps, inits)

    @classmethod
    def from_file(cls, filename):
        ext = "yaml" if str(filename).endswith(".yaml") else "json"
        orders, exps, inits = load_params(filename, extension=ext)
        for z, e in exps.items():
            if z not in inits:
                inits[z] = np.ones_like(e) / len(e)
        return cls(orders, exps, inits)

    from_yaml = from_file
    from_json = from_file


class NumericBasisFuncHelper(BasisFuncHelper):
    """Basis functions tabulated as cubic splines (core/basis.py:330-390): every shell of the
    exponential table is sampled on BeckeRTransform(1e-4, 1.5) o GaussChebyshev(nrad) and replaced by
    the not-a-knot, extrapolating ``CubicSpline`` through the samples.

    All shells of an element share one set of knots, so a pro-atom sum_k c_k S_k(r) is itself a
    piecewise cubic on those knots with coefficients sum_k c_k coef_k: that is how the device
    evaluates it (``ppoly_coefficients`` feeds ``hp_promol_weights_spline``)."""

    def __init__(self, splines_dict, initials):
        super().__init__(initials)
        self._splines_dict = splines_dict

    splines_dict = property(lambda self: self._splines_dict)

    def get_nshell(self, number):
        return len(self.splines_dict[number])

    def compute_proshell_dens(self, number, ishell, population, points, nderiv=0):
        y = population * self.splines_dict[number][ishell](points)
        if nderiv == 0:
            return y
        if nderiv == 1:
            return y, np.zeros_like(y)  # the reference returns a zero derivative (core/basis.py:356)
        raise NotImplementedError

    def get_knots(self, number):
        """Break points shared by the shells of element ``number``."""
        return np.asarray(self.splines_dict[number][0].x)

    def ppoly_coefficients(self, number):
        """(nshell, nseg, 4) SciPy PPoly coefficients, highest power first within a segment."""
        shells = self.splines_dict[number]
        return np.stack([np.ascontiguousarray(shells[k].c.T) for k in range(len(shells))])

    @classmethod
    def from_file(cls, filename, nrad=150):
        from scipy.interpolate import CubicSpline

        from ..gridlite import BeckeRTransform, GaussChebyshev

        helper = ExpBasisFuncHelper.from_file(filename)
        return cls._from_exp_helper(helper, nrad, CubicSpline, BeckeRTransform, GaussChebyshev)

    @classmethod
    def from_function_type(cls, func_type="gauss", nrad=150):
        from scipy.interpolate import CubicSpline

        from ..gridlite import BeckeRTransform, GaussChebyshev

        helper = ExpBasisFuncHelper.from_function_type(func_type)
        return cls._from_exp_helper(helper, nrad, CubicSpline, BeckeRTransform, GaussChebyshev)

    @classmethod
    def _from_exp_helper(cls, helper, nrad, CubicSpline, BeckeRTransform, GaussChebyshev):
        rgrid = BeckeRTransform(1e-4, 1.5).transform_1d_grid(GaussChebyshev(nrad))
        splines = {}
        for number, exps in helper.exponents.items():
            splines[number] = {
                k: CubicSpline(rgrid.points, helper.compute_proshell_dens(number, k, 1.0, rgrid.points), 0)
                for k in range(len(exps))
            }
        return cls(splines, helper.initials)
