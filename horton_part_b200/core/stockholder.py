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
        slab = self.slab
        lo = slab.point_base
        promol_h = slab.promol.cpu().numpy()
        if slab.npts == self.grid.size:
            self.cache.dump("promoldens", promol_h)
        else:  # sharded: only this rank's slice is known
            promol = self.cache.load("promoldens", alloc=self.grid.size)[0]
            promol[lo : lo + slab.npts] = promol_h
        at_w = slab.at_w.cpu().numpy()
        off = slab.atom_point_offsets_host
        for a in range(slab.shard.atom_lo, slab.shard.atom_hi):
            self.cache.dump(f"at_weights_{a}", at_w[off[a] - lo : off[a + 1] - lo])

    def _compute_entropy(self, rho, rho0):
        """Host restatement for API users (core/stockholder.py:145-151); the iteration loop gets
        the same number from the fused kernel's partial sums."""
        sick = (rho0 < self.density_cutoff) | (rho < self.density_cutoff)
        with np.errstate(all="ignore"):
            ln_ratio = np.where(sick, 0.0, np.log(np.where(sick, 1.0, rho / np.where(sick, 1.0, rho0))))
        return self._grid.integrate(rho, ln_ratio)

    def _atom_moments(self):
        import torch

        from .device import stream_ptr

        if not self.local:
            raise NotImplementedError("moments need atomic grids (grid_type 1 or 2)")
        slab = self.slab
        sh = slab.shard
        lmax = int(self.lmax)
        nmom = (lmax + 1) * (lmax + 2) * (lmax + 3) // 6 + (lmax + 1) ** 2 + lmax + 1
        seg = (slab.atom_point_offsets[sh.atom_lo : sh.atom_hi + 1] - slab.point_base).contiguous()
        out = torch.zeros((self.natom, nmom), dtype=torch.float64, device=slab.device)
        _lib.call("hp_atom_moments", sh.nlocal, sh.atom_lo, lmax, seg, slab.px, slab.py, slab.pz, slab.atw,
                  slab.at_w, slab.rho, slab.atom_xyz, out, stream_ptr(slab.device))  # fmt: skip
        if self._comm is not None:
            import torch.distributed as dist

            dist.all_reduce(out, group=self._comm)
        return out.cpu().numpy()

    def _atom_integrals(self, density):
        import torch

        from .device import stream_ptr, to_device

        slab = self.slab
        if density is self._moldens:
            dens = slab.rho
        else:
            dens = to_device(np.asarray(density)[slab.point_base : slab.point_base + slab.npts], slab.device)
        sh = slab.shard
        seg = slab.atom_point_offsets[sh.atom_lo : sh.atom_hi + 1] - slab.point_base
        out = torch.zeros(self.natom, dtype=torch.float64, device=slab.device)
        _lib.call("hp_segment_integrate", sh.nlocal, seg.contiguous(), slab.atw, slab.at_w, dens,
                  out[sh.atom_lo : sh.atom_hi], stream_ptr(slab.device))  # fmt: skip
        if self._comm is not None:
            import torch.distributed as dist

            dist.all_reduce(out, group=self._comm)
        return out.cpu().numpy()
