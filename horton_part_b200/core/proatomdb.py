"""Database of isolated pro-atom radial densities for Hirshfeld / Hirshfeld-I (host side).

API of the reference's ``ProAtomRecord`` / ``ProAtomDB`` (core/proatomdb.py:34-447) as far as the
partitioning path uses it: records keyed by (number, charge), one radial grid per element, linear
(or geometric) combinations of charge states, spline construction, and the database preparation
helpers ``compact`` / ``normalize`` (:390-447) with the record methods they use (``compute_radii``,
``chop``, equality).  All of it is host-side set-up; the device sees PPoly coefficients only.
"""

from __future__ import annotations

import numpy as np
from scipy.interpolate import CubicHermiteSpline, CubicSpline

__all__ = ["ProAtomRecord", "ProAtomDB"]


class ProAtomRecord:
    """Radial density (and derivative) of one isolated atom or ion on a 1-D grid."""

    def __init__(self, number, charge, energy, rgrid, rho, deriv=None, pseudo_number=None, ipot_energy=None):
        self._number, self._charge, self._energy = number, charge, energy
        self._rgrid, self._rho, self._deriv = rgrid, rho, deriv
        self._pseudo_number = number if pseudo_number is None else pseudo_number
        self._ipot_energy = -energy if self.pseudo_population == 1 else ipot_energy
        self._safe = True

    number = property(lambda self: self._number)
    pseudo_number = property(lambda self: self._pseudo_number)
    charge = property(lambda self: self._charge)
    energy = property(lambda self: self._energy)
    ipot_energy = property(lambda self: self._ipot_energy)
    rho = property(lambda self: self._rho)
    deriv = property(lambda self: self._deriv)
    rgrid = property(lambda self: self._rgrid)
    safe = property(lambda self: self._safe)
    population = property(lambda self: self._number - self._charge)
    pseudo_population = property(lambda self: self._pseudo_number - self._charge)

    def update_safe(self, other):
        """Mark this record unsafe when a lower-population state of the same element is lower in
        energy (an anion not bound by the basis set); pick up the ionisation potential."""
        if other.number != self._number:
            return
        if other.population < self.population and other.energy < self.energy:
            self._safe = False
        if other.population == self.population - 1:
            self._ipot_energy = other.energy - self.energy

    def get_moment(self, order):
        return self.rgrid.integrate(self.rho, self.rgrid.points**order)

    def compute_radii(self, populations):
        """Radii (and grid indices) at which the running integral of the density reaches the
        given populations, by linear interpolation between grid points (core/proatomdb.py:154-179)."""
        radii = self.rgrid.points
        running = np.cumsum(4 * np.pi * radii**2 * self.rho * self.rgrid.weights)
        indexes = running.searchsorted(populations)
        result = []
        for pop, i in zip(populations, indexes):
            if i == len(running):
                result.append(radii[-1])
            else:
                x = (pop - running[i]) / (running[i - 1] - running[i])
                result.append(x * radii[i - 1] + (1 - x) * radii[i])
        return indexes, result

    def chop(self, npoint):
        """Keep the first ``npoint`` radial points."""
        from ..gridlite import OneDGrid

        self._rho = self._rho[:npoint]
        if self._deriv is not None:
            self._deriv = self._deriv[:npoint]
        self._rgrid = OneDGrid(self._rgrid.points[:npoint], self._rgrid.weights[:npoint])

    def __eq__(self, other):
        if not isinstance(other, ProAtomRecord):
            return NotImplemented
        same_grid = (self.rgrid.size == other.rgrid.size and np.array_equal(self.rgrid.points, other.rgrid.points)
                     and np.array_equal(self.rgrid.weights, other.rgrid.weights))  # fmt: skip
        same_deriv = (self.deriv is None and other.deriv is None) or (
            self.deriv is not None and other.deriv is not None and np.array_equal(self.deriv, other.deriv))
        return bool(
            self.number == other.number and self.charge == other.charge and self.energy == other.energy
            and same_grid and np.array_equal(self.rho, other.rho) and same_deriv
            and self.pseudo_number == other.pseudo_number and self.ipot_energy == other.ipot_energy
        )  # fmt: skip

    def __ne__(self, other):
        result = self.__eq__(other)
        return result if result is NotImplemented else not result

    __hash__ = None


class ProAtomDB:
    def __init__(self, records):
        best = {}
        for rec in records:
            key = (rec.number, rec.charge)
            if key not in best or rec.energy < best[key].energy:
                best[key] = rec
        self._records = list(best.values())
        self._map = dict(best)
        self._rgrid_map = {}
        for rec in self._records:
            known = self._rgrid_map.get(rec.number)
            if known is None:
                self._rgrid_map[rec.number] = rec.rgrid
            elif not np.allclose(known.points, rec.rgrid.points):
                raise ValueError("All proatoms of a given element must have the same radial grid.")
        for number in self.get_numbers():
            states = [self.get_record(number, q) for q in self.get_charges(number)]
            for r0 in states:
                for r1 in states:
                    r0.update_safe(r1)

    size = property(lambda self: len(self._records))

    def get_record(self, number, charge):
        return self._map[(number, charge)]

    def get_numbers(self):
        return sorted(self._rgrid_map)

    def get_charges(self, number, safe=False):
        return sorted((r.charge for r in self._records if r.number == number and (r.safe or not safe)),
                      reverse=True)  # fmt: skip

    def get_rgrid(self, number):
        return self._rgrid_map[number]

    def get_rho(self, number, parameters=0, combine="linear", do_deriv=False):
        """Density (and derivative) of one charge state (int) or of a combination {charge: coeff}."""
        if isinstance(parameters, int):
            rec = self.get_record(number, parameters)
            return (rec.rho, rec.deriv) if do_deriv else rec.rho
        if not isinstance(parameters, dict):
            raise TypeError("Could not interpret parameters argument")
        if combine not in ("linear", "geometric"):
            raise ValueError('Combine argument "%s" not supported.' % combine)
        rho = 0.0
        deriv = 0.0 if do_deriv else None
        for charge, coeff in parameters.items():
            if coeff == 0.0:
                continue
            rec = self.get_record(number, charge)
            if combine == "linear":
                rho = rho + coeff * rec.rho
                term = None if rec.deriv is None else coeff * rec.deriv
            else:
                rho = rho + coeff * np.log(rec.rho)
                term = None if rec.deriv is None else coeff * rec.deriv / rec.rho
            deriv = deriv + term if (do_deriv and term is not None and deriv is not None) else None
        if combine == "geometric":
            rho = np.exp(rho)
            if do_deriv and deriv is not None:
                deriv = rho * deriv
        if not isinstance(rho, np.ndarray):
            rho = np.zeros_like(self.get_rgrid(number).points)
            if do_deriv:
                deriv = np.zeros_like(rho)
        return (rho, deriv) if do_deriv else rho

    def get_spline(self, number, parameters=0, combine="linear"):
        rho, deriv = self.get_rho(number, parameters, combine, do_deriv=True)
        x = self.get_rgrid(number).points
        return CubicSpline(x, rho, True) if deriv is None else CubicHermiteSpline(x, rho, deriv, True)

    def compact(self, nel_lost):
        """Cut the radial grids where the tail of every *safe* state holds at most ``nel_lost``
        electrons (core/proatomdb.py:390-431); all states of an element keep a common grid."""
        from ..gridlite import OneDGrid

        for number in self.get_numbers():
            npoint = 0
            for charge in self.get_charges(number, safe=True):
                rec = self.get_record(number, charge)
                nel = rec.pseudo_number - charge
                npoint = max(npoint, int(rec.compute_radii([nel - nel_lost])[0][0]) + 1)
            for charge in self.get_charges(number):
                self.get_record(number, charge).chop(npoint)
            old = self._rgrid_map[number]
            self._rgrid_map[number] = OneDGrid(old.points[:npoint], old.weights[:npoint])

    def normalize(self):
        """Scale every density to its integer (pseudo-)population on its radial grid, in place
        (core/proatomdb.py:433-447; the integral is the reference's plain ``rgrid.integrate(rho)``)."""
        for number in self.get_numbers():
            rgrid = self.get_rgrid(number)
            for charge in self.get_charges(number):
                rec = self.get_record(number, charge)
                rec.rho[:] *= (rec.pseudo_number - charge) / rgrid.integrate(rec.rho)
