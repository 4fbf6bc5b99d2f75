"""Hirshfeld partitioning: fixed (neutral) pro-atoms from a database, one pass.

Counterpart of the reference's ``HirshfeldWPart`` (hirshfeld.py:127-193).  The pro-atom splines are
the reference's own (SciPy ``CubicHermiteSpline`` / ``CubicSpline`` through ``ProAtomDB``,
core/stockholder.py:259-269); their PPoly coefficients are uploaded and evaluated on the molecular
grid by ``hp_promol_weights_spline``.
"""

from __future__ import annotations

import numpy as np

from .core.logging import deflist
from .core.stockholder import AbstractStockholderWPart

__all__ = ["HirshfeldWPart", "check_proatomdb"]


def check_proatomdb(numbers, pseudo_numbers, proatomdb):
    """The molecule and the database must use the same effective core charges (hirshfeld_i.py)."""
    for i, (number, pseudo) in enumerate(zip(numbers, pseudo_numbers)):
        expected = proatomdb.get_record(number, 0).pseudo_number
        if expected != pseudo:
            raise ValueError(
                "The pseudo number of atom %i does not match with the proatom database (%i!=%i)"
                % (i, pseudo, expected)
            )


class DatabaseSplineMixin:
    """Device spline table fed with SciPy PPoly coefficients of database pro-atoms."""

    def _setup_spline_table(self):
        from .isa import SplineTable

        rgrids = [self.proatomdb.get_rgrid(z) for z in self.numbers]
        self._table = SplineTable(self.slab, rgrids)
        self._coef_cache = {}

    def _state_coefficients(self, number, charge):
        """(nseg, 4) PPoly coefficients of one database state, with the reference's negative-value
        fix (fix_proatom_rho) when ``self._clip_database_negatives`` is set."""
        key = (int(number), int(charge))
        if key not in self._coef_cache:
            from scipy.interpolate import CubicHermiteSpline, CubicSpline

            rho, deriv = self.proatomdb.get_rho(int(number), int(charge), do_deriv=True)
            x = self.proatomdb.get_rgrid(int(number)).points
            if self._clip_database_negatives and rho.min() < 0:
                rho = np.where(rho < 0, 0.0, rho)
                deriv = None
            spl = CubicSpline(x, rho, True) if deriv is None else CubicHermiteSpline(x, rho, deriv, True)
            self._coef_cache[key] = np.ascontiguousarray(spl.c.T)
        return self._coef_cache[key]

    def _upload_coefficients(self, per_atom):
        import torch

        flat = np.concatenate([c.ravel() for c in per_atom])
        self._table.coef.copy_(torch.from_numpy(flat))


class HirshfeldWPart(DatabaseSplineMixin, AbstractStockholderWPart):
    """Hirshfeld partitioning with Becke-Lebedev grids"""

    name = "h"
    _clip_database_negatives = True  # base eval_proatom goes through fix_proatom_rho

    def __init__(self, coordinates, numbers, pseudo_numbers, grid, moldens, proatomdb, spindens=None,
                 lmax=3, logger=None, grid_type=1, **kwargs):  # fmt: skip
        check_proatomdb(numbers, pseudo_numbers, proatomdb)
        self._proatomdb = proatomdb
        device_kw = {k: kwargs[k] for k in ("device", "comm", "local_radius") if k in kwargs}
        AbstractStockholderWPart.__init__(self, coordinates, numbers, pseudo_numbers, grid, moldens, spindens,
                                          lmax, logger, grid_type=grid_type, **device_kw)  # fmt: skip

    proatomdb = property(lambda self: self._proatomdb)

    def _init_log_scheme(self):
        self.logger.info("Initialized: %s" % self.__class__.__name__)
        deflist(self.logger, [("Scheme", "Hirshfeld"), ("Proatomic DB", self.proatomdb)])

    def get_rgrid(self, index):
        return self.proatomdb.get_rgrid(self.numbers[index])

    def get_proatom_rho(self, iatom, *args, **kwargs):
        return self.proatomdb.get_rho(self.numbers[iatom], do_deriv=True)

    def _refresh_table(self):
        if getattr(self, "_table", None) is None:
            if self.on_molgrid:
                raise NotImplementedError("Hirshfeld with grid_type 2/3 is not built yet")
            self._setup_spline_table()
            self._upload_coefficients([self._state_coefficients(z, 0) for z in self.numbers])

    def _launch_promol_weights(self, want_entropy=True):
        self._refresh_table()
        self._table.promol_weights(self.density_cutoff, True, True, want_entropy)
