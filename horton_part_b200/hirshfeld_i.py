"""Iterative Hirshfeld (Hirshfeld-I): pro-atoms interpolated between integer charge states.

Counterpart of the reference's ``HirshfeldIWPart`` (hirshfeld_i.py:36-176).  The parameters are the
atomic charges.  Each iteration mixes the database splines of floor(q) and floor(q)+1 with weights
(1-x, x) (hirshfeld_i.py:116-158) -- done on the coefficient level on the host (natom x nseg x 4
doubles), evaluated on the grid by ``hp_promol_weights_spline`` -- and integrates w_a*rho on every
atomic grid (``hp_segment_integrate``; hirshfeld_i.py:166-172).
"""

from __future__ import annotations

import numpy as np

from . import _lib
from .core.iterstock import AbstractISAWPart
from .core.logging import deflist
from .hirshfeld import DatabaseSplineMixin, check_proatomdb

__all__ = ["HirshfeldIWPart"]


class HirshfeldIWPart(DatabaseSplineMixin, AbstractISAWPart):
    """Iterative Hirshfeld partitioning with Becke-Lebedev grids"""

    name = "hi"
    _clip_database_negatives = False  # HI evaluates proatomdb.get_spline directly (hirshfeld_i.py:139)

    def __init__(self, coordinates, numbers, pseudo_numbers, grid, moldens, proatomdb, spindens=None,
                 lmax=3, logger=None, threshold=1e-6, maxiter=500, grid_type=1, **kwargs):  # fmt: skip
        check_proatomdb(numbers, pseudo_numbers, proatomdb)
        self._proatomdb = proatomdb
        device_kw = {k: kwargs[k] for k in ("device", "comm", "local_radius", "device_loop") if k in kwargs}
        AbstractISAWPart.__init__(self, coordinates, numbers, pseudo_numbers, grid, moldens, spindens, lmax,
                                  logger, threshold, maxiter, grid_type=grid_type, **device_kw)  # fmt: skip

    proatomdb = property(lambda self: self._proatomdb)

    def _init_log_scheme(self):
        self.logger.info("Initialized: %s" % self.__class__.__name__)
        deflist(
            self.logger,
            [
                ("Scheme", "Hirshfeld-I"),
                ("Convergence threshold", "%.1e" % self._threshold),
                ("Maximum iterations", self._maxiter),
                ("Proatomic DB", self._proatomdb),
            ],
        )

    def get_rgrid(self, index):
        return self.proatomdb.get_rgrid(self.numbers[index])

    def get_interpolation_info(self, i, charges=None):
        if charges is None:
            charges = self.cache.load("charges")
        target = charges[i]
        icharge = int(np.floor(target))
        return icharge, target - icharge

    def get_proatom_rho(self, iatom, charges=None, **kwargs):
        icharge, x = self.get_interpolation_info(iatom, charges)
        pseudo_pop = self.pseudo_numbers[iatom] - icharge
        number = self.numbers[iatom]
        if pseudo_pop == 1 or x == 0.0:
            return self.proatomdb.get_rho(number, {icharge: 1 - x}, do_deriv=True)
        if pseudo_pop > 1:
            return self.proatomdb.get_rho(number, {icharge: 1 - x, icharge + 1: x}, do_deriv=True)
        raise ValueError("Requesting a pro-atom with a negative (pseudo) population")

    # -- device hooks ---------------------------------------------------------------------------
    def _init_propars(self):
        import torch

        if self.on_molgrid:
            raise NotImplementedError("Hirshfeld-I with grid_type 2/3 is not built yet")
        charges = self.cache.load("charges", alloc=self.natom, tags="o")[0]
        self.cache.dump("propars", charges, tags="o")
        self._setup_spline_table()
        self._pops = torch.zeros(self.natom, dtype=torch.float64, device=self.slab.device)
        self._seg = None
        return charges

    def _refresh_table(self):
        """Mixed pro-atom coefficients for the current charges (hirshfeld_i.py:136-158)."""
        charges = self.cache.load("charges")
        per_atom = []
        for a in range(self.natom):
            icharge, x = self.get_interpolation_info(a, charges)
            pseudo_pop = self.pseudo_numbers[a] - icharge
            coef = self._state_coefficients(self.numbers[a], icharge) * (1 - x)
            if pseudo_pop > 1 and x != 0.0:
                coef = coef + self._state_coefficients(self.numbers[a], icharge + 1) * x
            elif pseudo_pop <= 0:
                raise ValueError("Requesting a pro-atom with a negative (pseudo) population")
            per_atom.append(coef)
        self._upload_coefficients(per_atom)

    def _launch_promol_weights(self, want_entropy=True):
        self._refresh_table()
        self._table.promol_weights(self.density_cutoff, True, True, want_entropy)

    def _run_iteration(self):
        """One Hirshfeld-I iteration: weights from the current charges, populations on the atomic
        grids, new charges; change on the database radial grids (host, natom x nrad doubles)."""
        import torch

        from .core.device import stream_ptr

        slab = self.slab
        dev = slab.device
        sh = slab.shard
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        self._launch_promol_weights(want_entropy=True)
        ev[1].record()
        if self._seg is None:
            self._seg = (slab.atom_point_offsets[sh.atom_lo : sh.atom_hi + 1] - slab.point_base).contiguous()
            self._scal = torch.zeros(1, dtype=torch.float64, device=dev)
        self._pops.zero_()
        _lib.call("hp_segment_integrate", sh.nlocal, self._seg, slab.atw, slab.at_w, slab.rho,
                  self._pops[sh.atom_lo : sh.atom_hi], stream_ptr(dev))  # fmt: skip
        _lib.call("hp_sum_partials", slab.npartial, slab.entropy_partials, self._scal, stream_ptr(dev))
        pack = torch.cat([self._pops, self._scal])
        if self._comm is not None:
            from .core.comm import all_reduce

            all_reduce(self._comm, pack)
        ev[2].record()
        host = pack.cpu().numpy()
        if self._state is None:
            class _Events:
                events = []

            self._state = _Events()
        self._state.events.append(ev)
        charges = self.cache.load("charges")
        old = charges.copy()
        charges[:] = self.pseudo_numbers - host[:-1]  # hirshfeld_i.py:170-172
        return float(self.compute_change(charges, old)), float(host[-1])
