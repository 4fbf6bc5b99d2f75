"""Minimal Basis Iterative Stockholder (MBIS) on the B200.

Drop-in for the reference's ``MBISWPart`` (/root/reference/src/horton_part/mbis.py:206-362): same
constructor, cache keys and results.  Pro-atoms are sums of Slater shells
``N_k S_k^3 exp(-S_k r) / (8 pi)`` (mbis.py:286); the per-atom update is the MBIS fixed point
``N_k <- int rho_a term_k/pro``, ``S_k <- 3 N_k / int r rho_a term_k/pro`` iterated to
``inner_threshold`` (``opt_mbis_propars``, mbis.py:81-163), here one warp per atom in
``hp_mbis_radial_solve``.
"""

from __future__ import annotations

import logging

import numpy as np

from . import _lib
from .core.iterstock import AbstractISAWPart
from .core.logging import deflist

__all__ = ["MBISWPart", "get_nshell", "get_initial_mbis_propars"]

logger = logging.getLogger(__name__)

_NOBLE = np.array([2, 10, 18, 36, 54, 86, 118])
_SHELL_CAPACITY = np.array([2.0, 8.0, 8.0, 18.0, 18.0, 32.0, 32.0])


def get_nshell(number: int):
    """Number of MBIS shells = period of the element (mbis.py:36-46)."""
    return _NOBLE.searchsorted(number) + 1


def get_initial_mbis_propars(number: int):
    """Initial [N_1, S_1, N_2, S_2, ...]: closed inner shells, the rest in the valence shell;
    exponents geometric from 2Z down to 2 (mbis.py:49-78)."""
    nshell = get_nshell(number)
    propars = np.zeros(2 * nshell, float)
    s_first = 2.0 * number
    ratio = (2.0 / s_first) ** (1.0 / (nshell - 1)) if nshell > 1 else 1.0
    for k in range(nshell):
        propars[2 * k] = _SHELL_CAPACITY[k]
        propars[2 * k + 1] = s_first * ratio**k
    propars[-2] = number - propars[:-2:2].sum()
    return propars


def mbis_atom_work(coordinates, numbers, grid, device=None):
    """Pairs the screened dense pass evaluates per atom block, estimated from the initial MBIS
    parameters: the load-balancing weights of a sharded run (``core.device.Shard(work=...)``)."""
    from .core.device import estimate_dense_work

    per_element = {}
    for z in np.unique(numbers):
        p = get_initial_mbis_propars(int(z))
        per_element[int(z)] = (p[0::2] * p[1::2] ** 3 / (8 * np.pi), p[1::2])
    shells = [per_element[int(z)] for z in numbers]
    kinds = [(int(z), id(grid.atgrids[a].rgrid), int(grid.atgrids[a].size)) for a, z in enumerate(numbers)]
    return estimate_dense_work(coordinates, grid, shells, kinds=kinds, device=device)


class MBISWPart(AbstractISAWPart):
    """Minimal Basis Iterative Stockholder (MBIS)"""

    name = "mbis"
    max_inner = 2000  # mbis.py:123
    device_loop_capable = True

    def _init_log_scheme(self):
        logger.info("Initialized: %s" % self.__class__.__name__)
        deflist(
            logger,
            [
                ("Scheme", "Minimal Basis Iterative Stockholder (MBIS)"),
                ("Outer loop convergence threshold", "%.1e" % self._threshold),
                ("Inner loop convergence threshold", "%.1e" % self._inner_threshold),
                ("Maximum iterations", self._maxiter),
            ],
        )

    def get_rgrid(self, iatom):
        if self.only_use_molgrid:
            raise NotImplementedError
        return self.get_grid(iatom).rgrid

    def get_proatom_rho(self, iatom: int, propars=None, **kwargs):
        """Pro-atom density and radial derivative on the atom's radial grid (host; API helper)."""
        if propars is None:
            propars = self.cache.load("propars")
        r = self.radial_distances[iatom] if self.on_molgrid else self.get_rgrid(iatom).points
        y = np.zeros(len(r), float)
        d = np.zeros(len(r), float)
        mine = propars[self._ranges[iatom] : self._ranges[iatom + 1]]
        for k in range(self._nshells[iatom]):
            N, S = mine[2 * k : 2 * k + 2]
            f = N * S**3 * np.exp(-S * r) / (8 * np.pi)
            y += f
            d -= S * f
        return y, d

    # -- device hooks ---------------------------------------------------------------------------
    def _estimate_atom_work(self):
        """Pairs the screened dense pass evaluates per atom block, from the initial parameters."""
        if self.on_molgrid or self._local_radius is not None or self._grid.atgrids is None:
            return None
        return mbis_atom_work(self.coordinates, self.numbers, self._grid, self._device)

    def _init_propars(self):
        from .core.device import ShellTable, to_device

        per_element = {int(z): get_initial_mbis_propars(int(z)) for z in np.unique(self.numbers)}  # once per element
        self._nshells = [len(per_element[int(z)]) // 2 for z in self.numbers]
        self._ranges = [0]
        for k in self._nshells:
            self._ranges.append(self._ranges[-1] + 2 * k)
        propars = self.cache.load("propars", alloc=self._ranges[-1], tags="o")[0]
        propars[:] = np.concatenate([per_element[int(z)] for z in self.numbers])
        slab = self.slab
        self._table = ShellTable(slab, 1, self._nshells)  # HP_FUNCTOR_SLATER
        st = self._alloc_state(len(propars))
        st.propars.copy_(to_device(propars, slab.device))
        self._par_offsets = to_device(np.asarray(self._ranges, dtype=np.int32), slab.device)
        self._pseudo = to_device(self.pseudo_numbers, slab.device, np.float64)
        return propars

    def _refresh_table(self):
        from .core.device import stream_ptr

        t = self._table
        _lib.call("hp_table_mbis", t.nshell, self._state.propars, t.A, t.alpha, stream_ptr(self.slab.device))

    def _molgrid_shell_params(self, propars):
        import torch

        from .core.device import stream_ptr

        n = self._table.nshell
        A = torch.empty(n, dtype=torch.float64, device=propars.device)
        alpha = torch.empty_like(A)
        _lib.call("hp_table_mbis", n, propars, A, alpha, stream_ptr(self.slab.device))
        return A, alpha

    def _molgrid_apply(self, propars, s0, s1, shell_active):
        import torch

        new = propars.clone().view(-1, 2)
        new[:, 0] = torch.where(shell_active, s0, new[:, 0])                 # mbis.py:145
        new[:, 1] = torch.where(shell_active, 3.0 * s0 / s1, new[:, 1])      # mbis.py:146
        return new.view(-1)

    def _launch_radial_update(self):
        from .core.device import stream_ptr

        slab, st = self.slab, self._state
        slab.shell_project()
        sh = slab.shard
        _lib.call(
            "hp_mbis_radial_solve", sh.nlocal, sh.atom_lo, slab.rad_offsets, slab.rad_r, slab.rad_w4,
            slab.sph_avg, self._par_offsets, st.propars, self._pseudo, float(self._inner_threshold),
            float(self.density_cutoff), int(self.max_inner), slab.nrad_max, int(max(self._nshells)), st.charges, st.msd, st.niter, st.flags,
            stream_ptr(slab.device),
        )  # fmt: skip

    def _post_iteration_checks(self):
        # the reference only *warns* here (mbis.py:158,162); flags are read lazily to avoid a sync
        pass

    def _finalize_propars(self):
        AbstractISAWPart._finalize_propars(self)
        flags = self._state.flags.cpu().numpy()
        if (flags & 1).any() or getattr(self, "_molgrid_not_converged", False):
            self.logger.warning("MBIS not converged, but still go ahead!")
        if (flags & 2).any():
            self.logger.warning("The sum of propars are not equal to the atomic pop.")
        propars = self.cache.load("propars")
        ends = np.asarray(self._ranges[1:])
        valence_charges = -propars[ends - 2]
        valence_widths = 1.0 / propars[ends - 1]
        core_charges = self._cache.load("charges") - valence_charges
        self.cache.dump("core_charges", core_charges, tags="o")
        self.cache.dump("valence_charges", valence_charges, tags="o")
        self.cache.dump("valence_widths", valence_widths, tags="o")
        if self.on_molgrid:
            return
        # radial projections of the last iteration (mbis.py:185-187)
        slab = self.slab
        sph = slab.sph_avg.cpu().numpy()
        ro = slab.rad_offsets_host
        def entries():
            for i, a in enumerate(range(slab.shard.atom_lo, slab.shard.atom_hi)):
                yield f"radial_points_{a}", slab.rad_r_host[ro[i] : ro[i + 1]]
                yield f"spherical_average_{a}", sph[ro[i] : ro[i + 1]]
                yield f"radial_weights_{a}", slab.rad_w_host[ro[i] : ro[i + 1]]

        self.cache.dump_many(entries(), tags="o")
