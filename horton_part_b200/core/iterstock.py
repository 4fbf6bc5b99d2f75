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
        from .comm import all_reduce

        all_reduce(comm, self.vec)


class AbstractISAWPart(AbstractStockholderWPart):
    """Iterative stockholder schemes with per-atom pro-atom parameters."""

    def __init__(self, coordinates, numbers, pseudo_numbers, grid, moldens, spindens=None, lmax=3,
                 logger=None, threshold=1e-6, maxiter=500, inner_threshold=1e-8, grid_type=1,
                 device_loop=True, **kwargs):  # fmt: skip
        self._device_loop = device_loop
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

    # -- grid_type 2/3: per-atom fixed points on the molecular grid ------------------------------
    def _molgrid_shell_params(self, propars):
        """(A, alpha) device tensors of the shell table for a device parameter vector."""
        raise NotImplementedError(f"{self.name} with grid_type 2/3 is not built yet")

    def _molgrid_apply(self, propars, s0, s1, shell_active):
        """New parameter vector from the shell sums S0, S1 (only where shell_active)."""
        raise NotImplementedError

    molgrid_max_inner = 2000
    molgrid_single_update = False

    def _molgrid_pass(self, out_par, in_par, prev_par, active):
        from .device import stream_ptr

        t, s = self._table, self.slab
        if getattr(self, "_mg", None) is None:
            import torch

            lim_a, lim_s = np.zeros(1, np.int32), np.zeros(1, np.int32)
            _lib.call("hp_molgrid_update_tile_limits", lim_a, lim_s)
            ntile, tiles = t.make_tiles(int(lim_a[0]), int(lim_s[0]))
            nout = 2 * t.nshell + 2 * self.natom
            nblk = int(_lib.call("hp_molgrid_num_blocks", s.npts))
            self._mg = dict(
                ntile=ntile, tiles=tiles, nout=nout,
                partial=torch.zeros(nblk * nout, dtype=torch.float64, device=s.device),
                out=torch.zeros(nout, dtype=torch.float64, device=s.device),
                shell_atom=torch.from_numpy(np.repeat(np.arange(self.natom), t._counts)).to(s.device),
            )
        mg = self._mg
        (Ao, alo), (Ai, ali), (Ap, alp) = out_par, in_par, prev_par
        _lib.call(
            "hp_molgrid_update_pass", t.functor, s.npts, s.px, s.py, s.pz, s.natom, s.atom_xyz, t.offsets,
            Ao, alo, Ai, ali, Ap, alp, t.order, active, mg["ntile"], mg["tiles"], s.rho, s.molw, s.promol,
            float(self.density_cutoff), t.nshell, mg["partial"], mg["out"], stream_ptr(s.device),
        )  # fmt: skip
        out = mg["out"]
        if self._comm is not None:
            from .comm import all_reduce

            out = out.clone()
            all_reduce(self._comm, out)
        M = t.nshell
        return out[0 : 2 * M : 2], out[1 : 2 * M : 2], out[2 * M :: 2], out[2 * M + 1 :: 2]

    def _run_iteration_molgrid(self):
        """One outer iteration with grid_type 2/3 (mbis.py:170-173, gisa.py:257-279): the inner
        fixed point of every atom runs on the whole molecular grid, all atoms batched per pass."""
        import torch

        from .device import stream_ptr

        st, slab = self._state, self.slab
        dev = slab.device
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        self._refresh_table()
        self._table.promol_weights(self.density_cutoff, True, True, True)
        ev[1].record()
        old = st.propars.clone()
        out_par = self._molgrid_shell_params(old)
        inner, prev = old.clone(), old.clone()
        active = torch.ones(self.natom, dtype=torch.int32, device=dev)
        pop = None
        flags_not_converged = True
        max_inner = 1 if self.molgrid_single_update else int(self.molgrid_max_inner)
        for it in range(max_inner):
            s0, s1, chg, p = self._molgrid_pass(out_par, self._molgrid_shell_params(inner),
                                                self._molgrid_shell_params(prev), active)  # fmt: skip
            if pop is None:
                pop = p.clone()
            shell_active = active[self._mg["shell_atom"]].bool()
            prev = inner
            inner = self._molgrid_apply(inner, s0, s1, shell_active)
            if self.molgrid_single_update:
                flags_not_converged = False
                break
            if it > 0:
                converged = torch.sqrt(chg) < self._inner_threshold
                active = active * (~converged).to(torch.int32)
                if not bool(active.any().item()):
                    flags_not_converged = False
                    break
        self._molgrid_not_converged = flags_not_converged
        st.propars.copy_(inner)
        st.charges.copy_(torch.from_numpy(np.ascontiguousarray(self.pseudo_numbers)).to(dev) - pop)
        # compute_change on the molecular grid: chg column with in = new, prev = old outer parameters
        _, _, msd, _ = self._molgrid_pass(out_par, self._molgrid_shell_params(inner),
                                          self._molgrid_shell_params(old), None)  # fmt: skip
        st.msd.copy_(msd)
        _lib.call("hp_sum_partials", slab.npartial, slab.entropy_partials, st.entropy, stream_ptr(dev))
        if self._comm is not None:
            from .comm import all_reduce

            all_reduce(self._comm, st.entropy)
        _lib.call("hp_finish_iteration", 1, st.entropy, self.natom, st.msd, st.out2, stream_ptr(dev))
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

    def _run_iteration(self):
        """One outer iteration on the device; returns (change, entropy) after one host sync."""
        import torch

        from .device import stream_ptr

        if self.on_molgrid:
            return self._run_iteration_molgrid()

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

    # -- device-resident loop -------------------------------------------------------------------
    #: True for schemes whose iteration is nothing but kernel launches on the slab's stream (no host
    #: solver, no torch allocation, no collective): MBIS, NLIS/GMBIS, ISA, aLISA with a device solver
    device_loop_capable = False

    def _use_device_loop(self):
        """Run iterations 2..n as ONE CUDA-graph launch (csrc/hp_loop.cu: the body of a conditional WHILE
        node, convergence test on the device)?  Default: whenever the scheme allows it -- the host then
        pays one launch and one synchronisation for the whole loop instead of one D2H + sync per
        iteration, which is what bounds H2O-sized systems.  ``device_loop=False`` (constructor) or
        HP_B200_DEVICE_LOOP=0 keeps the host-driven loop; the results are identical."""
        import os

        env = os.environ.get("HP_B200_DEVICE_LOOP")
        want = self._device_loop if env is None else env != "0"
        return bool(want and self.device_loop_capable and self._comm is None and not self.on_molgrid)

    def _run_device_loop(self, done, propars):
        """Iterations done+1 .. niter on the device.  Returns (counter, change, entropy) of the last
        iteration and appends to the history lists exactly what the host loop would have appended."""
        import torch

        from .device import stream_ptr

        st, slab = self._state, self.slab
        dev = slab.device
        nv, n = st.vec.numel(), self.natom
        maxiter = int(self._maxiter)
        hist = torch.zeros((maxiter, nv + 2), dtype=torch.float64, device=dev)
        counter = torch.full((1,), done, dtype=torch.int32, device=dev)
        stamps = torch.zeros(2 * (maxiter + 1), dtype=torch.int64, device=dev)
        # stream capture is not allowed on the legacy default stream: the loop runs on a side stream
        main = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            stream = stream_ptr(dev)
            handle = np.zeros(1, dtype=np.uint64)
            _lib.call("hp_loop_stamp", None, counter, -1, 1, stamps, stream)  # "end of the previous iteration"
            _lib.call("hp_loop_begin", stream, handle)
            loop = int(handle[0])
            try:
                # the body: one iteration, exactly the launches of _run_iteration
                self._launch_promol_weights(want_entropy=True)
                _lib.call("hp_loop_stamp", loop, counter, 0, 0, stamps, stream)
                self._launch_radial_update()
                _lib.call("hp_finish_iteration", slab.npartial, slab.entropy_partials, n, st.msd, st.out2, stream)
                _lib.call("hp_loop_commit", loop, nv, st.vec, st.out2, float(self._threshold), maxiter, hist,
                          counter, stamps, stream)  # fmt: skip
                _lib.call("hp_loop_end", loop, stream)
                t0 = time.time()
                _lib.call("hp_loop_launch", loop, stream)
                side.synchronize()
                self.time_usage["device_loop"] = time.time() - t0
            finally:
                _lib.call("hp_loop_destroy", loop, stream)
        main.wait_stream(side)
        niter = int(counter.item())
        rows = hist[done:niter].cpu().numpy()
        ns = stamps.cpu().numpy().reshape(-1, 2)
        charges = self.cache.load("charges", alloc=n, tags="o")[0]
        change = entropy = None
        for k, row in enumerate(rows):
            c = done + k  # zero-based index of this iteration
            propars[:] = row[1 + 2 * n : nv]
            charges[:] = row[1 + n : 1 + 2 * n]
            change, entropy = float(row[nv]), float(row[nv + 1])
            self.history_propars.append(propars.copy())
            self.history_charges.append(charges.copy())
            self.history_entropies.append(entropy)
            self.history_changes.append(change)
            self.history_time_update_at_weights.append((ns[c, 0] - ns[c - 1, 1]) * 1e-9)
            self.history_time_update_propars.append((ns[c, 1] - ns[c, 0]) * 1e-9)
            self.logger.info("%9i   %10.5e   %10.5e" % (c + 1, change, entropy))
        return niter, change, entropy

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
            if counter == 1 and self._use_device_loop():
                # every lazily allocated buffer exists after the first iteration: the rest of the loop
                # runs on the device (same kernels, same order, same stopping rule)
                for ev in self._state.events:
                    self.history_time_update_at_weights.append(ev[0].elapsed_time(ev[1]) * 1e-3)
                    self.history_time_update_propars.append(ev[1].elapsed_time(ev[2]) * 1e-3)
                self._state.events = []
                counter, change, entropy = self._run_device_loop(counter, propars)
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
