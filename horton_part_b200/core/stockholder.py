"""Stockholder partitioning on the device: promolecule, owner weights, entropy.

Counterpart of the reference's ``AbstractStockholderWPart``
(/root/reference/src/horton_part/core/stockholder.py:34-403).  Where the reference loops over atoms
and makes K+3 NumPy passes over the whole grid per atom (``update_pro`` :153-175,
``update_at_weights`` :352-384), this class keeps a shell table on the GPU and runs ONE fused
kernel per outer iteration (``hp_promol_weights``).
"""

from __future__ import annotations

import numpy as np

from .. import _lib
from .base import WPart

__all__ = ["AbstractStockholderWPart"]


class AbstractStockholderWPart(WPart):
    """Common machinery of all w_a = rho0_a / rho0 schemes."""

    # subclasses set self._table (a device.ShellTable or spline table) in _init_propars

    def _init_subgrids(self):
        WPart._init_subgrids(self)
        self._log_grid_info()

    def get_wcor(self, index):
        return 1.0

    def _log_grid_info(self, nb=20):
        log = self.logger.info
        log("")
        log("=" * 80)
        log("Information of integral grids.")
        log("-" * 80)
        log(f"Grid size of molecular grid: {self.grid.size}")
        if self.local and self.natom <= 64:  # per-atom listing only for small systems
            for iatom in range(self.natom):
                log(f" Atom {iatom} ".center(80, "*"))
                atgrid = self.get_grid(iatom)
                sizes = np.diff(np.asarray(atgrid.indices))
                log(f"   |-- Radial grid size: {len(sizes)}")
                log("   |-- Angular grid sizes: ")
                for j in range(len(sizes) // nb + 1):
                    log("          " + " ".join(str(s) for s in sizes[j * nb : j * nb + nb]))
            log("-" * 80)
        log("=" * 80)
        log(" ")

    # -- hooks ---------------------------------------------------------------------------------
    def get_rgrid(self, index: int):
        raise NotImplementedError

    def get_proatom_rho(self, iatom: int, *args, **kwargs):
        raise NotImplementedError

    def _refresh_table(self):
        """Write the current pro-atom parameters into the device shell table."""
        raise NotImplementedError

    # -- the hot pass --------------------------------------------------------------------------
    def _launch_promol_weights(self, want_entropy=True):
        """Enqueue the fused promolecule / owner-weight / entropy kernel (no synchronisation)."""
        self._refresh_table()
        if self._local_radius is not None:
            if not hasattr(self._table, "local_radius"):
                raise NotImplementedError("local_radius is only available for exponential pro-atoms")
            self._table.local_radius = self._local_radius
        self._table.promol_weights(self.density_cutoff, True, True, want_entropy)

    def update_at_weights(self, force_on_molgrid=False):
        """Recompute promolecule and atomic weights from the current parameters and publish them
        as ``promoldens`` / ``at_weights_{a}`` in the cache (this call downloads the arrays; the
        iteration loop itself keeps them on the device)."""
        if self.on_molgrid or force_on_molgrid:
            raise NotImplementedError("molecular-grid weights are handled by the scheme classes")
        self._launch_promol_weights(want_entropy=False)
        self._publish_weights()

    def _publish_weights(self):
        """Download promolecule and owner weights of this rank's slab into the cache: one D2H copy
        each; the per-atom ``at_weights_{a}`` entries are views into the downloaded array."""
        from .hostmem import download

        slab = self.slab
        lo = slab.point_base
        promol_h = download(slab.promol)
        if slab.npts == self.grid.size:
            self.cache.dump("promoldens", promol_h)
        else:
            # sharded: only this rank's slice is known.  The slice is on the host now; the whole-grid array
            # (zeros elsewhere) is assembled when somebody asks for it -- touching 8 bytes x Npts of fresh
            # pages per call costs more than the rank's share of an iteration.
            from .cache import Deferred

            npts, n = self.grid.size, slab.npts
            self.cache.dump("promoldens_local", promol_h)

            def assemble():
                whole = np.zeros(npts)
                whole[lo : lo + n] = promol_h
                return whole

            self.cache.dump("promoldens", Deferred(assemble))
        at_w = download(slab.at_w)
        off = slab.atom_point_offsets_host
        self.cache.dump_many((f"at_weights_{a}", at_w[off[a] - lo : off[a + 1] - lo])
                             for a in range(slab.shard.atom_lo, slab.shard.atom_hi))  # fmt: skip

    # -- host helpers of the reference API (spline objects for user code) --------------------------
    def fix_proatom_rho(self, index, rho, deriv):
        """Clip negative parts of a tabulated pro-atom (core/stockholder.py:202-224)."""
        rgrid = self.get_rgrid(index)
        original = rgrid.integrate(rho)
        if rho.min() < 0:
            rho[rho < 0] = 0.0
            deriv = None
            error = rgrid.integrate(rho) - original
            self.logger.info("                Pro-atom not positive everywhere. Lost %.1e electrons" % error)
        return rho, deriv

    def get_proatom_spline(self, index, *args, **kwargs):
        """SciPy spline of the pro-atom's radial density on its radial grid
        (core/stockholder.py:226-269): Hermite if derivatives are available, not-a-knot otherwise."""
        from scipy.interpolate import CubicHermiteSpline, CubicSpline

        rho, deriv = self.get_proatom_rho(index, *args, **kwargs)
        rho, deriv = self.fix_proatom_rho(index, rho, deriv)
        rgrid = self.get_rgrid(index)
        if deriv is None:
            return CubicSpline(rgrid.points, rho, True)
        return CubicHermiteSpline(rgrid.points, rho, deriv, True)

    def eval_spline(self, index, spline, output, grid, label="noname"):
        """output[:] = spline(|r - R_index|) on ``grid`` (core/stockholder.py:271-302); API helper
        for user code, the partitioning itself evaluates pro-atoms in the kernels."""
        output[:] = spline(np.linalg.norm(self.coordinates[index] - grid.points, axis=1))

    def eval_proatom(self, index, output, grid):
        """Pro-atom of atom ``index`` on an arbitrary grid, + 1e-100 (core/stockholder.py:304-350)."""
        self.eval_spline(index, self.get_proatom_spline(index), output, grid, label="proatom")
        output += 1e-100
        assert np.isfinite(output).all()

    def do_prosplines(self):
        """Store the pro-atom density splines (core/stockholder.py:386-393)."""
        for index in range(self.natom):
            key = ("spline_prodensity", index)
            if key not in self.cache:
                self.logger.info("Storing proatom density spline for atom %i." % index)
                self.cache.dump(key, self.get_proatom_spline(index), tags="o")

    def _compute_entropy(self, rho, rho0):
        """Host restatement for API users (core/stockholder.py:145-151); the iteration loop gets
        the same number from the fused kernel's partial sums."""
        sick = (rho0 < self.density_cutoff) | (rho < self.density_cutoff)
        with np.errstate(all="ignore"):
            ln_ratio = np.where(sick, 0.0, np.log(np.where(sick, 1.0, rho / np.where(sick, 1.0, rho0))))
        return self._grid.integrate(rho, ln_ratio)
