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

__all__ = ["HirshfeldWPart", "check_proatomdb", "do_dispersion"]


def check_proatomdb(numbers, pseudo_numbers, proatomdb):
    """The molecule and the database must use the same effective core charges (hirshfeld_i.py)."""
    for i, (number, pseudo) in enumerate(zip(numbers, pseudo_numbers)):
        expected = proatomdb.get_record(number, 0).pseudo_number
        if expected != pseudo:
            raise ValueError(
                "The pseudo number of atom %i does not match with the proatom database (%i!=%i)"
                % (i, pseudo, expected)
            )


# C6 coefficients of the isolated atoms in atomic units (Chu & Dalgarno 2004; H: Yan et al. 1996),
# the reference data of the Tkatchenko-Scheffler rescaling (hirshfeld.py:59-103), by atomic number.
_REFERENCE_C6 = dict(zip(
    list(range(1, 39)) + [49, 50, 51, 52, 53],
    [6.499, 1.42, 1392.0, 227.0, 99.5, 46.6, 24.2, 15.6, 9.52, 6.20, 1518.0, 626.0, 528.0, 305.0, 185.0,
     134.0, 94.6, 64.2, 3923.0, 2163.0, 1383.0, 1044.0, 832.0, 602.0, 552.0, 482.0, 408.0, 373.0, 253.0,
     284.0, 498.0, 354.0, 246.0, 210.0, 162.0, 130.0, 4769.0, 3175.0, 779.0, 659.0, 492.0, 445.0, 385.0],
))  # fmt: skip


def do_dispersion(part):
    """Atoms-in-molecules C6 coefficients by volume rescaling (hirshfeld.py:48-124):
    V_a = <r^3> of the AIM density (radial moment 3 from ``do_moments``, computed on the device),
    C6_a = (V_a / V_a^free)^2 C6_a^free; -1 where no free-atom value is tabulated."""
    if part.lmax < 3:
        part.logger.warning("Skip computing dispersion coefficients because lmax=%i<3" % part.lmax)
    volumes, new1 = part._cache.load("volumes", alloc=part.natom, tags="o")
    volume_ratios, new2 = part._cache.load("volume_ratios", alloc=part.natom, tags="o")
    c6s, new3 = part._cache.load("c6s", alloc=part.natom, tags="o")
    if new1 or new2 or new3:
        part.do_moments()
        radial_moments = part._cache.load("radial_moments")
        part.logger.info("Computing atomic dispersion coefficients.")
        for i in range(part.natom):
            n = int(part.numbers[i])
            volumes[i] = radial_moments[i, 3]
            volume_ratios[i] = volumes[i] / part.proatomdb.get_record(n, 0).get_moment(3)
            c6s[i] = volume_ratios[i] ** 2 * _REFERENCE_C6[n] if n in _REFERENCE_C6 else -1


class DatabaseSplineMixin:
    """Device spline table fed with SciPy PPoly coefficients of database pro-atoms."""

    def do_dispersion(self):
        do_dispersion(self)

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

    def _atom_integrals(self, density):
        """int w_a * density.  grid_type 1 and 2 integrate every atom on its own atomic grid
        (core/base.py:287-298: the full-grid weights are cut back to the owner block); grid_type 3
        has no atomic grids and integrates w_a over the whole molecular grid, which
        ``hp_atom_weight_integrals_spline`` does without natom x Npts weight arrays."""
        if not self.only_use_molgrid:
            return super()._atom_integrals(density)
        import torch

        from . import _lib
        from .core.device import stream_ptr, to_device

        slab, t = self.slab, self._table
        if self._comm is not None:
            raise NotImplementedError("grid_type 3 populations of spline pro-atoms run on one GPU")
        dens = slab.rho if density is self._moldens else to_device(np.asarray(density, dtype=float), slab.device)
        nblk = int(_lib.call("hp_spline_integral_blocks", slab.npts))
        partial = torch.zeros(self.natom * nblk, dtype=torch.float64, device=slab.device)
        out = torch.zeros(self.natom, dtype=torch.float64, device=slab.device)
        _lib.call("hp_atom_weight_integrals_spline", slab.npts, slab.px, slab.py, slab.pz, self.natom,
                  slab.atom_xyz, t.offsets, t.knots, t.coef, float(t.proatom_offset), dens, slab.molw,
                  slab.promol, partial, out, stream_ptr(slab.device))  # fmt: skip
        return out.cpu().numpy()


class HirshfeldWPart(DatabaseSplineMixin, AbstractStockholderWPart):
    """Hirshfeld partitioning with Becke-Lebedev grids"""

    name = "h"
    _clip_database_negatives = True  # base eval_proatom goes through fix_proatom_rho

    def __init__(self, coordinates, numbers, pseudo_numbers, grid, moldens, proatomdb, spindens=None,
                 lmax=3, logger=None, grid_type=1, **kwargs):  # fmt: skip
        check_proatomdb(numbers, pseudo_numbers, proatomdb)
        self._proatomdb = proatomdb
        device_kw = {k: kwargs[k] for k in ("device", "comm", "local_radius", "device_loop") if k in kwargs}
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
            self._setup_spline_table()
            self._upload_coefficients([self._state_coefficients(z, 0) for z in self.numbers])

    def _launch_promol_weights(self, want_entropy=True):
        self._refresh_table()
        self._table.promol_weights(self.density_cutoff, True, True, want_entropy)

    def update_at_weights(self, force_on_molgrid=False):
        """One pass: promolecule and weights from the fixed database pro-atoms.  With grid_type 2/3
        the reference caches full-grid weight arrays (natom x Npts); here the cache holds every
        atom's weights on its own block and grid_type 3 populations integrate the weight function
        over the whole grid inside ``hp_atom_weight_integrals_spline``."""
        self._launch_promol_weights(want_entropy=False)
        self._publish_weights()
