"""Iterative Stockholder Analysis (ISA): non-parametric pro-atoms.

Counterpart of the reference's ``ISAWPart`` (isa.py:37-122).  The pro-atom of atom a is the
not-a-knot cubic spline through its current spherical average on the radial grid
(core/stockholder.py:259-269), evaluated at |r_p - R_a| with extrapolation and the ``+1e-100``
offsets (:343-349); the parameters ARE the spherical averages (clipped to >= 1e-100).  All of it
runs on the device: ``hp_spline_build`` -> ``hp_promol_weights_spline`` -> ``hp_shell_project`` ->
``hp_isa_update``.
"""

from __future__ import annotations

import logging

import numpy as np

from . import _lib
from .core.iterstock import AbstractISAWPart
from .core.logging import deflist

__all__ = ["ISAWPart"]

logger = logging.getLogger(__name__)


class SplineTable:
    """Knots of every atom's radial grid (replicated on every rank) and the coefficient buffer."""

    def __init__(self, slab, rgrids, proatom_offset=1e-100):
        import torch

        from .core.device import to_device

        dev = slab.device
        sizes = [g.size for g in rgrids]
        self.offsets_host = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
        self.nknot = int(self.offsets_host[-1])
        self.offsets = to_device(self.offsets_host, dev)
        self.knots = to_device(np.concatenate([np.asarray(g.points, float) for g in rgrids]), dev)
        self.coef = torch.zeros(4 * (self.nknot - len(rgrids)), dtype=torch.float64, device=dev)
        self.work = torch.zeros(2 * self.nknot, dtype=torch.float64, device=dev)
        self.slab = slab
        #: added to every spline value: base eval_proatom does ``spline(r) + 1e-100``
        #: (core/stockholder.py:349); 0 for pro-atoms that are plain sums of tabulated shells
        self.proatom_offset = proatom_offset
        self._build_lookup(rgrids, sizes)

    def _build_lookup(self, rgrids, sizes):
        """Interval look-up tables (one per distinct knot array, shared by the atoms that use it) and
        the shared-memory tiling of the atom list for ``hp_promol_weights_spline``."""
        from .core.device import to_device

        dev = self.slab.device
        tables, pool, meta, pos = {}, [], np.zeros((len(rgrids), 3), dtype=np.int32), 0
        for a, g in enumerate(rgrids):
            x = np.ascontiguousarray(g.points, dtype=np.float64)
            if len(x) < 2:
                raise ValueError("a spline pro-atom needs at least two knots")
            key = x.tobytes()
            if key not in tables:
                nb = int(_lib.call("hp_spline_lut_size", len(x), x))
                lut, key0 = np.zeros(nb, dtype=np.uint16), np.zeros(1, dtype=np.int32)
                _lib.call("hp_spline_lut_fill", len(x), x, key0, lut)
                tables[key] = (int(key0[0]), nb, pos)
                pool.append(lut)
                pos += nb
            meta[a] = tables[key]
        # transposed inverses of the not-a-knot systems (one per distinct knot array with >= 4 knots): the
        # per-iteration spline construction is then a parallel matrix-vector product (hp_spline_build)
        self.inv_offsets = self.invT = None
        self.nknot_max = int(max(sizes))
        if min(sizes) >= 4 and self.nknot_max <= 2048:
            inv_pool, inv_pos, inv_of, where = [], 0, {}, np.zeros(len(rgrids), dtype=np.int64)
            for a, g in enumerate(rgrids):
                x = np.ascontiguousarray(g.points, dtype=np.float64)
                key = x.tobytes()
                if key not in inv_of:
                    mat = np.zeros(len(x) * len(x))
                    _lib.call("hp_spline_system_inverse", len(x), x, mat)
                    inv_of[key] = inv_pos
                    inv_pool.append(mat)
                    inv_pos += mat.size
                where[a] = inv_of[key]
            self.inv_offsets = to_device(where, dev, np.int64)
            self.invT = to_device(np.concatenate(inv_pool), dev)
        self.lut_meta = to_device(meta.ravel(), dev, np.int32)
        self.lut = to_device(np.concatenate(pool), dev, np.uint16)
        max_atoms, max_knots = np.zeros(1, np.int32), np.zeros(1, np.int32)
        _lib.call("hp_spline_tile_limits", max_atoms, max_knots)
        if max(sizes) > int(max_knots[0]):
            raise ValueError(f"radial grids with more than {int(max_knots[0])} points are not supported by the spline pass")
        tiles, a, natom = [0], 0, len(sizes)
        while a < natom:
            b, nk = a, 0
            while b < natom and b - a < int(max_atoms[0]) and nk + sizes[b] <= int(max_knots[0]):
                nk += sizes[b]
                b += 1
            tiles.append(b)
            a = b
        self.ntile = len(tiles) - 1
        self.tiles = to_device(np.asarray(tiles, dtype=np.int32), dev)

    def build(self, values, clip_negative=True):
        from .core.device import stream_ptr

        _lib.call("hp_spline_build", len(self.offsets_host) - 1, self.offsets, self.knots, values,
                  int(clip_negative), self.coef, self.work, self.inv_offsets, self.invT, self.nknot_max,
                  stream_ptr(self.slab.device))  # fmt: skip

    def promol_weights(self, density_cutoff, want_promol=True, want_weights=True, want_entropy=True,
                       proatom_offset=None, promol_offset=1e-100):  # fmt: skip
        from .core.device import stream_ptr

        s = self.slab
        if proatom_offset is None:
            proatom_offset = self.proatom_offset
        _lib.call(
            "hp_promol_weights_spline", s.npts, s.px, s.py, s.pz, s.point_base, s.natom, s.atom_xyz,
            s.atom_point_offsets, self.offsets, self.knots, self.coef, self.lut_meta, self.lut, self.ntile,
            self.tiles, float(proatom_offset), float(promol_offset), s.rho,
            s.molw, float(density_cutoff), s.promol if want_promol else None,
            s.at_w if want_weights else None, s.entropy_partials if want_entropy else None,
            stream_ptr(s.device),
        )  # fmt: skip


class ISAWPart(AbstractISAWPart):
    """Iterative Stockholder Partitioning with Becke-Lebedev grids"""

    name = "is"
    device_loop_capable = True

    def _init_log_scheme(self):
        logger.info("Initialized: %s" % self.__class__.__name__)
        deflist(
            logger,
            [
                ("Scheme", "Iterative Stockholder"),
                ("Outer loop convergence threshold", "%.1e" % self._threshold),
                ("Inner loop convergence threshold", "%.1e" % self._inner_threshold),
                ("Maximum iterations", self._maxiter),
                ("lmax", self._lmax),
            ],
        )

    def get_rgrid(self, index):
        if self.only_use_molgrid:
            raise NotImplementedError
        return self.get_grid(index).rgrid

    def get_proatom_rho(self, iatom, propars=None, **kwargs):
        if propars is None:
            propars = self.cache.load("propars")
        if self.on_molgrid:
            raise NotImplementedError
        return propars[self._ranges[iatom] : self._ranges[iatom + 1]], None

    def get_proatom_spline(self, index, *args, **kwargs):
        """SciPy spline of the current pro-atom (host helper; the kernels use the device copy)."""
        from scipy.interpolate import CubicSpline

        rho, _ = self.get_proatom_rho(index, *args, **kwargs)
        rho = np.where(rho < 0, 0.0, rho)
        return CubicSpline(self.get_rgrid(index).points, rho, True)

    def _init_propars(self):
        from .core.device import to_device

        if self.on_molgrid:
            raise NotImplementedError("ISA needs atomic grids (grid_type=1)")
        rgrids = [self.get_rgrid(a) for a in range(self.natom)]
        self._ranges = [0]
        for g in rgrids:
            self._ranges.append(self._ranges[-1] + g.size)
        propars = self.cache.load("propars", alloc=self._ranges[-1], tags="o")[0]
        slab = self.slab
        self._table = SplineTable(slab, rgrids)
        st = self._alloc_state(len(propars))
        self._par_offsets = to_device(np.asarray(self._ranges, dtype=np.int32), slab.device)
        self._pseudo = to_device(self.pseudo_numbers, slab.device, np.float64)
        self._rad_w = to_device(slab.rad_w_host, slab.device)
        return propars

    def _refresh_table(self):
        self._table.build(self._state.propars, clip_negative=True)

    def _launch_radial_update(self):
        from .core.device import stream_ptr

        slab, st = self.slab, self._state
        slab.shell_project()
        sh = slab.shard
        _lib.call("hp_isa_update", sh.nlocal, sh.atom_lo, slab.rad_offsets, slab.rad_r, self._rad_w,
                  slab.sph_avg, self._par_offsets, st.propars, self._pseudo, st.charges, st.msd,
                  stream_ptr(slab.device))  # fmt: skip

    def _finalize_propars(self):
        AbstractISAWPart._finalize_propars(self)
        slab = self.slab
        sph = np.clip(slab.sph_avg.cpu().numpy(), 1e-100, np.inf)
        ro = slab.rad_offsets_host
        for i, a in enumerate(range(slab.shard.atom_lo, slab.shard.atom_hi)):
            self.cache.dump(f"radial_points_{a}", np.clip(slab.rad_r_host[ro[i] : ro[i + 1]], 1e-100, 1e10), tags="o")
            self.cache.dump(f"spherical_average_{a}", sph[ro[i] : ro[i + 1]], tags="o")
            self.cache.dump(f"radial_weights_{a}", slab.rad_w_host[ro[i] : ro[i + 1]], tags="o")
