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

    # -- vectorised mixing of the database states (all atoms of an element at once) --------------------
    def _element_tables(self):
        """Per element: the charges of its database states (a contiguous range), their PPoly coefficient rows,
        their radial densities, the quadrature weights 4 pi r^2 w of the element's radial grid, the atoms of
        that element and where their coefficient blocks sit in the device table."""
        tabs = getattr(self, "_hi_tables", None)
        if tabs is not None:
            return tabs
        nseg4 = np.array([4 * (self.proatomdb.get_rgrid(int(z)).size - 1) for z in self.numbers], dtype=np.int64)
        block = np.concatenate([[0], np.cumsum(nseg4)])
        tabs = {}
        for z in np.unique(self.numbers):
            z = int(z)
            charges = sorted(self.proatomdb.get_charges(z))
            if charges != list(range(charges[0], charges[-1] + 1)):
                tabs = None  # gaps in the database: keep the per-atom route
                break
            rgrid = self.proatomdb.get_rgrid(z)
            atoms = np.flatnonzero(self.numbers == z)
            tabs[z] = dict(
                first=charges[0], nstate=len(charges), atoms=atoms,
                coef=np.stack([self._state_coefficients(z, q).ravel() for q in charges]),
                rho=np.stack([self.proatomdb.get_rho(z, q) for q in charges]),
                w4=rgrid.weights * (4 * np.pi * rgrid.points**2),
                where=block[atoms][:, None] + np.arange(nseg4[atoms[0]])[None, :],
            )
        self._hi_tables = tabs if tabs is not None else False
        self._hi_flat = np.zeros(block[-1])
        return self._hi_tables

    def _mix(self, tab, charges, rows):
        """(1 - x) * rows[floor(q)] + x * rows[floor(q) + 1] for the atoms of one element, with the
        reference's rules (hirshfeld_i.py:125-132): one state only for one-electron pro-atoms and for integer
        charges.  None when a state is missing (the per-atom route then raises the reference's error)."""
        q = charges[tab["atoms"]]
        ic = np.floor(q).astype(np.int64)
        x = q - ic
        pseudo_pop = self.pseudo_numbers[tab["atoms"]] - ic
        single = (pseudo_pop == 1) | (x == 0.0)
        lo = ic - tab["first"]
        hi = np.where(single, lo, lo + 1)
        if (pseudo_pop < 1).any() or lo.min() < 0 or hi.max() >= tab["nstate"]:
            return None
        out = 0.0 + (1 - x)[:, None] * rows[lo]
        two = ~single
        if two.any():
            out[two] = out[two] + x[two, None] * rows[hi[two]]
        return out

    def _refresh_table(self):
        """Mixed pro-atom coefficients for the current charges (hirshfeld_i.py:136-158)."""
        import torch

        charges = self.cache.load("charges")
        tabs = self._element_tables()
        if tabs:
            mixed = {z: self._mix(tab, charges, tab["coef"]) for z, tab in tabs.items()}
            if all(m is not None for m in mixed.values()):
                for z, tab in tabs.items():
                    self._hi_flat[tab["where"]] = mixed[z]
                self._table.coef.copy_(torch.from_numpy(self._hi_flat))
                return
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

    def compute_change(self, propars1, propars2):
        """core/iterstock.py:32-45 for the charges as parameters: all atoms of an element at once (the database
        states are mixed as rows of one table); per-atom route of the base class when a state is missing."""
        tabs = self._element_tables()
        if tabs:
            c1, c2 = np.asarray(propars1, dtype=float), np.asarray(propars2, dtype=float)
            terms = np.zeros(self.natom)
            for tab in tabs.values():
                r1, r2 = self._mix(tab, c1, tab["rho"]), self._mix(tab, c2, tab["rho"])
                if r1 is None or r2 is None:
                    break
                delta = r1 - r2
                terms[tab["atoms"]] = np.einsum("i,ai,ai->a", tab["w4"], delta, delta)
            else:
                msd = 0.0
                for t in terms:  # atom order, as the reference accumulates
                    msd += t
                return np.sqrt(msd)
        return AbstractISAWPart.compute_change(self, propars1, propars2)

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
