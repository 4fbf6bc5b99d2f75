"""Iterative stockholder driver: the outer loop, convergence test and result bookkeeping.

Counterpart of the reference's ``AbstractISAWPart``
(/root/reference/src/horton_part/core/iterstock.py:61-193).  One outer iteration is

    table <- propars         (tiny kernel)
    hp_promol_weights        promolecule, owner weights, entropy partials         [the hot kernel]
    hp_shell_project         spherical averages on every atom's radial grid
    radial solve             per-atom parameter update + charge + change term     [one warp / atom]
    (NCCL all-reduce of the per-iteration state when the grid is sharded)
    hp_finish_iteration      change = sqrt(sum msd), entropy = sum partials
    one small D2H copy       [change, entropy, charges, propars]

and exactly the reference's stopping rule: ``change < threshold or counter >= maxiter``
(core/iterstock.py:187).  As in the reference the cached ``at_weights`` are those of the
*penultimate* parameters (no re-evaluation after convergence).
"""

from __future__ import annotations

import time

import numpy as np

from .. import _lib
from .cache import just_once
from .stockholder import AbstractStockholderWPart

__all__ = ["AbstractISAWPart", "IterationState"]


class IterationState:
    """Per-iteration state in ONE contiguous FP64 vector

        [ entropy | msd (natom) | charges (natom) | propars (npar) ]

    so that every run needs a single small D2H per iteration and a sharded run a single
    all-reduce: each rank zero-fills the vector, writes the entries of the atoms it owns (plus its
    partial entropy) and the SUM all-reduce then acts as an all-gather (x + 0 = x exactly, so the
    result is independent of the reduction order) while also summing the entropy."""

    def __init__(self, natom, npar, device):
        import torch

        n = natom
        self.natom, self.npar = natom, npar
        self.vec = torch.zeros(1 + 2 * n + npar, dtype=torch.float64, device=device)
        self.entropy = self.vec[0:1]
        self.msd = self.vec[1 : 1 + n]
        self.charges = self.vec[1 + n : 1 + 2 * n]
        self.propars = self.vec[1 + 2 * n :]
        self.out2 = torch.zeros(2, dtype=torch.float64, device=device)
        self.niter = torch.zeros(n, dtype=torch.int32, device=device)
        self.flags = torch.zeros(n, dtype=torch.int32, device=device)
        self.host = torch.empty(self.vec.numel() + 2, dtype=torch.float64)
        if torch.device(device).type == "cuda":
            self.host = self.host.pin_memory()
        self.events = []
        self.propars_prev = None

    def begin_sharded_update(self, par_lo, par_hi):
        """Zero everything this iteration will gather; keep the local atoms' parameters, which
        are the starting point of their solves."""
        self.propars_prev = self.propars.clone()
        self.msd.zero_()
        self.charges.zero_()
        self.propars.zero_()
        self.propars[par_lo:par_hi] = self.propars_prev[par_lo:par_hi]

    def gather(self, comm):
        import torch.distributed as dist

        dist.all_reduce(self.vec, op=dist.ReduceOp.SUM, group=comm)


class AbstractISAWPart(AbstractStockholderWPart):
    """Iterative stockholder schemes with per-atom pro-atom parameters."""

    def __init__(self, coordinates, numbers, pseudo_numbers, grid, moldens, spindens=None, lmax=3,
                 logger=None, threshold=1e-6, maxiter=500, inner_threshold=1e-8, grid_type=1,
                 **kwargs):  # fmt: skip
        self._threshold = threshold
        self._inner_threshold = inner_threshold if inner_threshold < threshold else threshold
        self._maxiter = maxiter
        self._state = None
        super().__init__(coordinates, numbers, pseudo_numbers, grid, moldens, spindens, lmax, logger,
                         grid_type=grid_type, **kwargs)  # fmt: skip

    # -- hooks of the concrete schemes ----------------------------------------------------------
    def _init_propars(self):
        """Allocate ``propars`` in the cache, build the device tables, return the host array."""
        raise NotImplementedError

    def _launch_radial_update(self):
        """Enqueue shell projection + per-atom solves for this rank's atoms; results go into the
        global device arrays ``state.propars / charges / msd``."""
        raise NotImplementedError

    def _post_iteration_checks(self):
        """Scheme-specific warnings after the state has been downloaded."""

    def compute_change(self, propars1, propars2):
        """sqrt(sum_a int 4 pi r^2 (rho0_a[propars1] - rho0_a[propars2])^2) on the radial grids
        (core/iterstock.py:32-45); host version for API users, the loop uses the device value."""
        msd = 0.0
        for index in range(self.natom):
            rho1, _ = self.get_proatom_rho(index, propars1)
            rho2, _ = self.get_proatom_rho(index, propars2)
            delta = rho1 - rho2
            rgrid = self.get_rgrid(index)
            msd += rgrid.integrate(4 * np.pi * rgrid.points**2, delta, delta)
        return np.sqrt(msd)

    # -- device state ---------------------------------------------------------------------------
    def _alloc_state(self, npar):
        self._state = IterationState(self.natom, npar, self.slab.device)
        return self._state

    def _run_iteration(self):
        """One outer iteration on the device; returns (change, entropy) after one host sync."""
        import torch

        from .device import stream_ptr

        st, slab = self._state, self.slab
        dev = slab.device
        sharded = self._comm is not None
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        self._launch_promol_weights(want_entropy=True)
        ev[1].record()
        if sharded:
            sh = slab.shard
            st.begin_sharded_update(self._ranges[sh.atom_lo], self._ranges[sh.atom_hi])
        self._launch_radial_update()
        if sharded:
            _lib.call("hp_sum_partials", slab.npartial, slab.entropy_partials, st.entropy, stream_ptr(dev))
            st.gather(self._comm)
            _lib.call("hp_finish_iteration", 1, st.entropy, self.natom, st.msd, st.out2, stream_ptr(dev))
        else:
            _lib.call("hp_finish_iteration", slab.npartial, slab.entropy_partials, self.natom, st.msd,
                      st.out2, stream_ptr(dev))  # fmt: skip
        ev[2].record()
        nv = st.vec.numel()
        st.host[:nv].copy_(st.vec, non_blocking=True)
        st.host[nv:].copy_(st.out2, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        st.events.append(ev)
        host = st.host.numpy()
        n = self.natom
        self.cache.load("propars")[:] = host[1 + 2 * n : nv]
        self.cache.load("charges", alloc=n, tags="o")[0][:] = host[1 + n : 1 + 2 * n]
        return float(host[nv]), float(host[nv + 1])

    # -- the loop -------------------------------------------------------------------------------
    def _finalize_propars(self):
        charges = self._cache.load("charges")
        dump = self.cache.dump
        dump("history_propars", np.array(self.history_propars), tags="o")
        dump("history_charges", np.array(self.history_charges), tags="o")
        dump("history_entropies", np.array(self.history_entropies), tags="o")
        dump("history_changes", np.array(self.history_changes), tags="o")
        dump("populations", self.numbers - charges, tags="o")
        dump("pseudo_populations", self.pseudo_numbers - charges, tags="o")
        dump("time_update_at_weights", np.sum(self.history_time_update_at_weights), tags="o")
        dump("time_update_propars", np.sum(self.history_time_update_propars), tags="o")

    @just_once
    def do_partitioning(self):
        new = any(("at_weights", i) not in self.cache for i in range(self.natom))
        new |= "niter" not in self.cache
        new |= "change" not in self.cache
        if not new:
            return
        t_start = time.time()
        propars = self._init_propars()
        self.logger.info("Iteration       Change      Entropy")
        counter = 0
        while True:
            counter += 1
            self.cache.dump("niter", counter, tags="o")
            change, entropy = self._run_iteration()
            self._post_iteration_checks()
            self.history_propars.append(propars.copy())
            self.history_charges.append(self.cache.load("charges").copy())
            self.history_entropies.append(entropy)
            self.history_changes.append(change)
            self.logger.info("%9i   %10.5e   %10.5e" % (counter, change, entropy))
            if change < self._threshold or counter >= self._maxiter:
                break
        self.logger.info("")
        # device-timed split of the iterations (CUDA events; the reference uses time.time())
        for ev in self._state.events:
            self.history_time_update_at_weights.append(ev[0].elapsed_time(ev[1]) * 1e-3)
            self.history_time_update_propars.append(ev[1].elapsed_time(ev[2]) * 1e-3)
        self._state.events = []
        self._publish_weights()
        self._finalize_propars()
        self.cache.dump("niter", counter, tags="o")
        self.cache.dump("change", change, tags="o")
        self.time_usage["do_partitioning_loop"] = time.time() - t_start
