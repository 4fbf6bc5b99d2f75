"""Non-linear ISA (NLIS): shells  N n S^(3/n) exp(-S r^n) / (4 pi Gamma(3/n))  with optimised
population N and exponent S and a fixed order n per shell.

Counterpart of the reference's ``nlis.py`` (``NLISWPart`` :197-367, ``opt_nlis_propars`` :99-194,
initial parameters :58-96).  Grid passes use the general-order functor of ``hp_promol_weights``;
the per-atom fixed point is ``hp_nlis_radial_solve``.
"""

from __future__ import annotations

import logging

import numpy as np
from scipy.special import gamma

from . import _lib
from .core.iterstock import AbstractISAWPart
from .core.logging import deflist
from .mbis import get_nshell

__all__ = ["NLISWPart", "get_nlis_nshell", "get_initial_nlis_propars"]

logger = logging.getLogger(__name__)


def get_nlis_nshell(number, nshell_dict):
    return nshell_dict.get(number, get_nshell(number))


def get_initial_nlis_propars(number, exp_n_dict, nshell_dict, logger=None):
    """[N, S, n] per shell: equal populations Z/K, exponents geometric from 2Z to 0.5, n from
    ``exp_n_dict[(Z, shell)]`` else 1 (nlis.py:58-96)."""
    nbs = get_nlis_nshell(number, nshell_dict)
    propars = np.ones(3 * nbs, float)
    s_first = 2.0 * number
    ratio = (0.5 / s_first) ** (1.0 / (nbs - 1)) if nbs > 1 else 1.0
    for k in range(nbs):
        propars[3 * k] = number / nbs
        propars[3 * k + 1] = s_first * ratio**k
        propars[3 * k + 2] = exp_n_dict[(number, k)] if (number, k) in exp_n_dict else 1.0
    return propars


class NLISWPart(AbstractISAWPart):
    """Non-Linear approximation of Iterative Stockholder (NLIS)"""

    name = "nlis"
    max_inner = 2000  # nlis.py:141
    device_loop_capable = True
    _scheme_label = "Non-Linear approximation of Iterative Stockholder (NLIS)"

    def __init__(self, coordinates, numbers, pseudo_numbers, grid, moldens, spindens=None, lmax=3,
                 logger=None, threshold=1e-6, maxiter=500, inner_threshold=1e-8, exp_n_dict=1.0,
                 nshell_dict=None, grid_type=1, **kwargs):  # fmt: skip
        self._exp_n_dict = exp_n_dict
        self._nshell_dict = nshell_dict or {}
        device_kw = {k: kwargs[k] for k in ("device", "comm", "local_radius", "device_loop") if k in kwargs}
        super().__init__(coordinates, numbers, pseudo_numbers, grid, moldens, spindens, lmax=lmax,
                         logger=logger, threshold=threshold, maxiter=maxiter,
                         inner_threshold=inner_threshold, grid_type=grid_type, **device_kw)  # fmt: skip

    def _init_log_scheme(self):
        logger.info("Initialized: %s" % self.__class__.__name__)
        deflist(
            logger,
            [
                ("Scheme", self._scheme_label),
                ("Outer loop convergence threshold", "%.1e" % self._threshold),
                ("Inner loop convergence threshold", "%.1e" % self._inner_threshold),
                ("Maximum iterations", self._maxiter),
            ],
        )

    def get_rgrid(self, iatom):
        if self.only_use_molgrid:
            raise NotImplementedError
        return self.get_grid(iatom).rgrid

    def get_proatom_rho(self, iatom, propars=None, **kwargs):
        if propars is None:
            propars = self.cache.load("propars")
        r = self.radial_distances[iatom] if self.on_molgrid else self.get_rgrid(iatom).points
        y = np.zeros(len(r), float)
        d = np.zeros(len(r), float)
        mine = propars[self._ranges[iatom] : self._ranges[iatom + 1]]
        for k in range(self._nshells[iatom]):
            N, S, n = mine[3 * k : 3 * k + 3]
            f = N * n * S ** (3 / n) * np.exp(-S * r**n) / (4 * np.pi * gamma(3.0 / n))
            y += f
            d -= N * S * n * r ** (n - 1) * f
        return y, d

    def _initial_atom_propars(self, number):
        return get_initial_nlis_propars(number, self._exp_n_dict, self._nshell_dict, logger=self.logger)

    def _atom_nshell(self, number):
        return int(get_nlis_nshell(number, self._nshell_dict))

    def _init_propars(self):
        from .core.device import ShellTable, to_device

        self._nshells = [self._atom_nshell(z) for z in self.numbers]
        if max(self._nshells) > 7:
            raise ValueError("more than 7 shells per atom are not supported")
        self._ranges = [0]
        for k in self._nshells:
            self._ranges.append(self._ranges[-1] + 3 * k)
        propars = self.cache.load("propars", alloc=self._ranges[-1], tags="o")[0]
        for a in range(self.natom):
            propars[self._ranges[a] : self._ranges[a + 1]] = self._initial_atom_propars(self.numbers[a])
        slab = self.slab
        dev = slab.device
        self._table = ShellTable(slab, 3, self._nshells)  # HP_FUNCTOR_GENERAL
        st = self._alloc_state(len(propars))
        st.propars.copy_(to_device(propars, dev))
        self._par_offsets = to_device(np.asarray(self._ranges, dtype=np.int32), dev)
        self._pseudo = to_device(self.pseudo_numbers, dev, np.float64)
        self._inv_gamma = to_device(1.0 / gamma(3.0 / propars[2::3]), dev)  # n is fixed
        return propars

    def _refresh_table(self):
        from .core.device import stream_ptr

        t = self._table
        _lib.call("hp_table_nlis", t.nshell, self._state.propars, self._inv_gamma, t.A, t.alpha, t.order,
                  stream_ptr(self.slab.device))  # fmt: skip

    def _molgrid_shell_params(self, propars):
        import torch

        from .core.device import stream_ptr

        n = self._table.nshell
        A, alpha, order = (torch.empty(n, dtype=torch.float64, device=propars.device) for _ in range(3))
        _lib.call("hp_table_nlis", n, propars, self._inv_gamma, A, alpha, order, stream_ptr(self.slab.device))
        return A, alpha

    def _molgrid_apply(self, propars, s0, s1, shell_active):
        """nlis.py:160-176 on the molecular grid: N <- m0, S <- 3/(m1 n) (1e-5 if m1 ~ 0), where the
        kernel's sums are S0 = m0 and S1 = N_old * m1."""
        import torch

        new = propars.clone().view(-1, 3)
        m1 = s1 / new[:, 0]
        s_new = torch.where(m1.abs() <= 1e-8, torch.full_like(m1, 1e-5), 3.0 / (m1 * new[:, 2]))
        new[:, 1] = torch.where(shell_active, s_new, new[:, 1])
        new[:, 0] = torch.where(shell_active, s0, new[:, 0])
        return new.view(-1)

    def _launch_radial_update(self):
        from .core.device import stream_ptr

        slab, st = self.slab, self._state
        slab.shell_project()
        sh = slab.shard
        _lib.call(
            "hp_nlis_radial_solve", sh.nlocal, sh.atom_lo, slab.rad_offsets, slab.rad_r, slab.rad_w4,
            slab.sph_avg, self._par_offsets, st.propars, self._table.offsets, self._inv_gamma, self._pseudo,
            float(self._inner_threshold), float(self.density_cutoff), int(self.max_inner), slab.nrad_max, int(max(self._nshells)), st.charges,
            st.msd, st.niter, st.flags, stream_ptr(slab.device),
        )  # fmt: skip

    def _finalize_propars(self):
        AbstractISAWPart._finalize_propars(self)
        flags = self._state.flags.cpu().numpy()
        if (flags & 1).any() or getattr(self, "_molgrid_not_converged", False):
            self.logger.warning("NLIS not converged, but still go ahead!")
        if (flags & 2).any():
            self.logger.warning("The sum of propars are not equal to the atomic pop.")
        propars = self.cache.load("propars")
        ends = np.asarray(self._ranges[1:])
        valence_charges = -propars[ends - 3]
        valence_widths = 1.0 / propars[ends - 2]
        self.cache.dump("core_charges", self._cache.load("charges") - valence_charges, tags="o")
        self.cache.dump("valence_charges", valence_charges, tags="o")
        self.cache.dump("valence_widths", valence_widths, tags="o")
        if self.on_molgrid:
            return
        slab = self.slab
        sph = slab.sph_avg.cpu().numpy()
        ro = slab.rad_offsets_host
        for i, a in enumerate(range(slab.shard.atom_lo, slab.shard.atom_hi)):
            self.cache.dump(f"radial_points_{a}", slab.rad_r_host[ro[i] : ro[i + 1]], tags="o")
            self.cache.dump(f"spherical_average_{a}", sph[ro[i] : ro[i + 1]], tags="o")
            self.cache.dump(f"radial_weights_{a}", slab.rad_w_host[ro[i] : ro[i + 1]], tags="o")
