"""Global Linear ISA (gLISA): all pro-atom coefficients optimised together on the molecular grid.

Counterpart of the reference's ``GlobalLinearISAWPart`` (glisa.py:55-1028).  The reference tabulates
every basis function on the whole grid (``pro_shells`` and ``rho*pro_shells``, two (M, Npts) arrays,
glisa.py:335-344) and then makes one NumPy pass per coefficient (``function_g`` :850-879) or per
coefficient pair (``_working_matrix`` :411-479).  Here the basis functions are regenerated inside
the kernels:

    calc_promol_dens       -> hp_promol_weights(promol_offset=0)
    function_g / gradient  -> hp_shell_moments(power=1)         I_m = int rho g_m / rho0
    Hessian                -> hp_hessian                        H_mn = int rho g_m g_n / rho0^2
    final charges          -> hp_atom_weight_integrals          (no natom x Npts weight arrays)

Solvers (every O(Npts) pass is a kernel; the host holds vectors of length M and M x M matrices):

    "sc"                            fixed point c <- g(c)                       glisa.py:805-848
    "newton" / "m-newton" / "quasi-newton"   exact / back-tracking / BFGS       glisa.py:572-803
                                    (exact Newton step: Cholesky + refinement of the M x M system on
                                    the device, host LAPACK as in the reference when H is not positive
                                    definite; BFGS algebra on the host; line-search admissibility of
                                    all step lengths in one hp_radial_valid launch)
    "diis" / "cdiis"                Anderson-Pulay acceleration of g            glisa.py:883-925, 993-1028
    "trust-region"                  SciPy trust-constr on (f, grad) from the device   glisa.py:927-987

    "cvxopt" (the default)          the convex programme through the built-in interior-point
                                    method of algo/cp.py (the third-party package the reference
                                    calls is not in this image)                 glisa.py:488-570
"""

from __future__ import annotations

import time

import numpy as np

from . import _lib, gisa
from .algo import bfgs, cdiis, diis
from .alisa import setup_bs_helper
from .core.basis import ExpBasisFuncHelper, shell_norm
from .core.cache import just_once
from .core.stockholder import AbstractStockholderWPart
from .core.logging import deflist
from .utils import fix_propars

__all__ = ["GlobalLinearISAWPart"]


class GlobalLinearISAWPart(AbstractStockholderWPart):
    name = "glisa"
    max_sc_iterations = 1_000_000  # the reference loops without a cap (glisa.py:823)

    def __init__(self, coordinates, numbers, pseudo_numbers, grid, moldens, spindens=None, lmax=3,
                 logger=None, threshold=1e-6, maxiter=500, solver="cvxopt", solver_options=None,
                 basis_func="gauss", grid_type=1, basis_type="analytic", **kwargs):  # fmt: skip
        self._maxiter = maxiter
        self._threshold = threshold
        self.basis_func = basis_func
        self._func_type = basis_func.upper() if basis_func in ("gauss", "slater") else "Customized"
        self._bs_helper = None
        self.basis_type = basis_type
        self._solver = solver
        self._solver_options = solver_options or {}
        self._ranges = []
        super().__init__(coordinates, numbers, pseudo_numbers, grid, moldens, spindens, lmax, logger,
                         grid_type, **kwargs)  # fmt: skip

    maxiter = property(lambda self: self._maxiter)
    threshold = property(lambda self: self._threshold)

    @property
    def bs_helper(self):
        return setup_bs_helper(self)

    @property
    def mol_pop(self):
        return self.nelec  # grid.integrate(moldens), glisa.py:186-188

    @property
    def propars(self):
        return self.cache.load("propars")

    def get_rgrid(self, index):
        if self.only_use_molgrid:
            self.logger.debug("rgird is not available when only_use_molgrid is `True`.")
            raise NotImplementedError
        return self.get_grid(index).rgrid

    def to_atomic_grid(self, index, data):
        if self.only_use_molgrid:
            self.logger.debug("atom grids are not available when only_use_molgrid is `True`.")
            raise NotImplementedError
        return super().to_atomic_grid(index, data)

    def get_proatom_rho(self, iatom, propars=None, **kwargs):
        return gisa.get_proatom_rho(self, iatom, propars=propars)

    def compute_change(self, propars1, propars2):
        """Host helper (radial grids); the solvers use hp_radial_change."""
        msd = 0.0
        for a in range(self.natom):
            d = self.get_proatom_rho(a, propars1)[0] - self.get_proatom_rho(a, propars2)[0]
            rgrid = self.get_rgrid(a)
            msd += rgrid.integrate(4 * np.pi * rgrid.points**2, d, d)
        return np.sqrt(msd)

    def _init_log_scheme(self):
        info = [
            ("Scheme", "Linear Iterative Stockholder"),
            ("Outer loop convergence threshold", "%.1e" % self.threshold),
            ("Using global ISA", True),
            ("Maximum outer iterations", self.maxiter),
            ("lmax", self.lmax),
            ("Solver", self._solver.__name__ if callable(self._solver) else self._solver.upper()),
            ("Basis function type", self._func_type),
        ]
        info += [(k, str(v)) for k, v in self._solver_options.items()]
        deflist(self.logger, info)
        self.logger.info(" ")

    # -- device set-up --------------------------------------------------------------------------
    def _init_propars(self):
        import torch

        from .core.device import ShellTable, to_device

        self._numeric = not isinstance(self.bs_helper, ExpBasisFuncHelper)
        if self._numeric and self.on_molgrid:
            # as in the reference: the numeric helper has no exponents for the molecular-grid bookkeeping
            raise NotImplementedError('gLISA with basis_type="numeric" needs grid_type=1')
        propars = gisa.init_propars(self)
        if not self.on_molgrid:
            gisa.evaluate_basis_functions(self)  # radial grids only: used by compute_change
        slab = self.slab
        dev = slab.device
        if self._numeric:
            self._init_numeric_tables(slab, dev)
        else:
            orders = np.concatenate([np.asarray(self.bs_helper.get_order(z), float) for z in self.numbers])
            alphas = np.concatenate([np.asarray(self.bs_helper.get_exponent(z), float) for z in self.numbers])
            functor = 2 if np.all(orders == 2.0) else (1 if np.all(orders == 1.0) else 3)
            self._table = ShellTable(slab, functor, self._nshells)
            self._table.alpha.copy_(to_device(alphas, dev))
            if functor == 3:
                self._table.order.copy_(to_device(orders, dev))
            self._norms = to_device(shell_norm(orders, alphas), dev)
        self._c = to_device(propars, dev)
        self._par_offsets = to_device(np.asarray(self._ranges, dtype=np.int32), dev)
        sh = slab.shard
        if not self.on_molgrid:
            blocks = [self.cache.load(f"bs_funcs_{a}") for a in range(sh.atom_lo, sh.atom_hi)]
            offs = np.concatenate([[0], np.cumsum([b.size for b in blocks])]).astype(np.int64)
            self._bs_offsets = to_device(offs, dev)
            self._bs_flat = to_device(np.concatenate([b.ravel() for b in blocks]), dev)
        M = len(propars)
        nblk = int(_lib.call("hp_molgrid_num_blocks", slab.npts))
        self._partial = torch.zeros(nblk * max(M, self.natom), dtype=torch.float64, device=dev)
        self._moments = torch.zeros(M, dtype=torch.float64, device=dev)
        self._msd = torch.zeros(self.natom, dtype=torch.float64, device=dev)
        self._scal = torch.zeros(2, dtype=torch.float64, device=dev)
        return propars

    def _init_numeric_tables(self, slab, dev):
        """basis_type="numeric" (core/basis.py:330-387): every shell is a tabulated spline on its element's
        knots.  The promolecule sum_m c_m S_m is ONE piecewise cubic per atom (coefficients mixed per
        evaluation, as aLISA does) for ``hp_promol_weights_spline``; the shell integrals and the Hessian read
        the per-shell coefficient blocks (``hp_shell_moments_table``, ``hp_hessian_table``)."""
        from .core.device import to_device
        from .isa import SplineTable

        class _Knots:  # the SplineTable only needs .points / .size
            def __init__(self, x):
                self.points, self.size = x, x.size

        helper = self.bs_helper
        self._table = SplineTable(slab, [_Knots(helper.get_knots(int(z))) for z in self.numbers], proatom_offset=0.0)
        self._ppoly = {int(z): helper.ppoly_coefficients(int(z)) for z in np.unique(self.numbers)}  # (K, nseg, 4)
        pool, start, pos = [], {}, 0
        for z, block in self._ppoly.items():
            start[z] = pos
            pool.append(block.ravel())
            pos += block.size
        offs = []
        for z, k in zip(self.numbers, self._nshells):
            seg = self._ppoly[int(z)].shape[1] * 4
            offs.extend(start[int(z)] + j * seg for j in range(k))
        self._shell_coef = to_device(np.concatenate(pool), dev)
        self._shell_coef_off = to_device(np.asarray(offs, dtype=np.int64), dev, np.int64)
        self._shell_offsets = to_device(np.asarray(self._ranges, dtype=np.int32), dev, np.int32)
        self._norms = None

    def _refresh_table(self):
        from .core.device import stream_ptr

        if self._numeric:
            import torch

            c = self._c.cpu().numpy()
            mixed = [np.einsum("k,ksc->sc", c[self._ranges[a] : self._ranges[a + 1]], self._ppoly[int(z)])
                     for a, z in enumerate(self.numbers)]  # fmt: skip
            self._table.coef.copy_(torch.from_numpy(np.concatenate([m.ravel() for m in mixed])))
            return
        t = self._table
        _lib.call("hp_table_scaled", t.nshell, self._c, self._norms, t.A, stream_ptr(self.slab.device))

    def _all_reduce(self, tensor):
        if self._comm is not None:
            from .core.comm import all_reduce

            all_reduce(self._comm, tensor)

    def _promol_and_entropy(self):
        """rho0 = sum_m c_m g_m on the local slab (no 1e-100 offsets: calc_promol_dens) and the
        entropy int rho ln(rho/rho0)  ->  device scalar self._scal[1]."""
        from .core.device import stream_ptr

        self._refresh_table()
        if self._numeric:
            self._table.promol_weights(self.density_cutoff, True, False, True, proatom_offset=0.0, promol_offset=0.0)
        else:
            self._table.promol_weights(self.density_cutoff, True, False, True, promol_offset=0.0)
        slab = self.slab
        _lib.call("hp_sum_partials", slab.npartial, slab.entropy_partials, self._scal[1:], stream_ptr(slab.device))

    def _shell_integrals(self, power=1):
        """I_m = int rho g_m / rho0^power over the local slab (unit-population basis functions)."""
        from .core.device import stream_ptr

        t, s = self._table, self.slab
        if self._numeric:
            M = len(self._c)
            _lib.call(
                "hp_shell_moments_table", s.npts, s.px, s.py, s.pz, s.natom, s.atom_xyz, self._shell_offsets,
                t.offsets, t.knots, t.lut_meta, t.lut, self._shell_coef_off, self._shell_coef, s.rho, s.molw,
                s.promol, float(self.density_cutoff), int(power), M, int(max(self._nshells)), self._partial,
                self._moments, stream_ptr(s.device),
            )  # fmt: skip
            return self._moments
        _lib.call(
            "hp_shell_moments", t.functor, s.npts, s.px, s.py, s.pz, s.natom, s.atom_xyz, t.offsets,
            self._norms, t.alpha, t.order, t.ntile, t.tiles, s.rho, s.molw, s.promol,
            float(self.density_cutoff), int(power), t.nshell, self._partial, self._moments,
            stream_ptr(s.device),
        )  # fmt: skip
        return self._moments

    def _device_change(self, c_new, c_old):
        from .core.device import stream_ptr

        s = self.slab
        sh = s.shard
        if self.on_molgrid:
            # compute_change on the molecular grid (core/iterstock.py:40-41): the `chg` column of
            # the update pass with in = new and prev = old coefficients
            import torch

            t = self._table
            if getattr(self, "_mg", None) is None:
                lim_a, lim_s = np.zeros(1, np.int32), np.zeros(1, np.int32)
                _lib.call("hp_molgrid_update_tile_limits", lim_a, lim_s)
                ntile, tiles = t.make_tiles(int(lim_a[0]), int(lim_s[0]))
                nout = 2 * t.nshell + 2 * self.natom
                nblk = int(_lib.call("hp_molgrid_num_blocks", s.npts))
                self._mg = dict(ntile=ntile, tiles=tiles,
                                partial=torch.zeros(nblk * nout, dtype=torch.float64, device=s.device),
                                out=torch.zeros(nout, dtype=torch.float64, device=s.device))  # fmt: skip
            mg = self._mg
            a_new, a_old = c_new * self._norms, c_old * self._norms
            _lib.call(
                "hp_molgrid_update_pass", t.functor, s.npts, s.px, s.py, s.pz, s.natom, s.atom_xyz, t.offsets,
                a_old, t.alpha, a_new, t.alpha, a_old, t.alpha, t.order, None, mg["ntile"], mg["tiles"], s.rho,
                s.molw, s.promol, float(self.density_cutoff), t.nshell, mg["partial"], mg["out"],
                stream_ptr(s.device),
            )  # fmt: skip
            self._msd.copy_(mg["out"][2 * t.nshell :: 2])
            return self._msd
        self._msd.zero_()
        _lib.call("hp_radial_change", sh.nlocal, sh.atom_lo, s.rad_offsets, s.rad_w4, self._par_offsets,
                  self._bs_offsets, self._bs_flat, c_new, c_old, self._msd, stream_ptr(s.device))  # fmt: skip
        return self._msd

    def function_g(self, x):
        """The fixed-point map g(c)_m = int rho c_m g_m / rho0[c] (glisa.py:850-879); host in/out."""
        import torch

        self._c.copy_(torch.from_numpy(np.ascontiguousarray(x, dtype=float)))
        self._promol_and_entropy()
        integrals = self._shell_integrals(1).clone()
        self._all_reduce(integrals)
        return (self._c * integrals).cpu().numpy()

    # -- solvers --------------------------------------------------------------------------------
    def _opt_propars(self, *args, **kwargs):
        if callable(self._solver):
            return self._solver(*args, **kwargs)
        if isinstance(self._solver, str):
            name = f"solver_{self._solver.replace('-', '_')}"
            if hasattr(self, name):
                return getattr(self, name)(*args, **kwargs)
            raise RuntimeError(f"Unknown solver: {name}")
        raise TypeError(f"The type of solver {type(self._solver)} is not supported.")

    def solver_sc(self, niter_print=1):
        """Self-consistent iteration c <- g(c) until the radial-grid change drops below threshold."""
        import torch

        propars = self.propars
        self.logger.info("Iteration       Change      Entropy")
        it = 0
        while True:
            c_old = self._c.clone()
            self._promol_and_entropy()
            integrals = self._shell_integrals(1)
            pack = torch.cat([integrals, self._scal[1:2]])
            self._all_reduce(pack)
            self._c.copy_(c_old * pack[:-1])
            msd = self._device_change(self._c, c_old)
            self._all_reduce(msd)
            change = float(torch.sqrt(msd.sum()).item())
            entropy = float(pack[-1].item())
            propars[:] = self._c.cpu().numpy()
            self.history_entropies.append(entropy)
            self.history_changes.append(change)
            self.history_propars.append(propars.copy())
            if (it + 1) % niter_print == 0:
                self.logger.info("%9i   %10.5e   %10.5e" % (it + 1, change, entropy))
            if change < self._threshold:
                break
            it += 1
            if it >= self.max_sc_iterations:
                raise RuntimeError("Not converged!")
        self.cache.dump("niter", it + 1, tags="o")
        return propars

    def hessian(self):
        """Dense Hessian H_mn = int rho g_m g_n / rho0^2 for the promolecule currently in
        slab.promol (device tensor, M x M)."""
        import torch

        from .core.device import stream_ptr, to_device

        t, s = self._table, self.slab
        M = len(self._c)
        if getattr(self, "_hess", None) is None:
            nbytes = int(_lib.call("hp_hessian_scratch_bytes", M))
            self._hess_scratch = torch.empty(nbytes, dtype=torch.uint8, device=s.device)
            self._hess = torch.zeros((M, M), dtype=torch.float64, device=s.device)
            shell_atom = np.repeat(np.arange(self.natom, dtype=np.int32), self._nshells)
            self._shell_atom = to_device(shell_atom, s.device)
        if self._numeric:
            _lib.call(
                "hp_hessian_table", s.npts, s.px, s.py, s.pz, s.atom_xyz, self._shell_atom, t.offsets, t.knots,
                t.lut_meta, t.lut, self._shell_coef_off, self._shell_coef, s.rho, s.molw, s.promol,
                float(self.density_cutoff), M, self._hess_scratch, self._hess_scratch.numel(), self._hess,
                stream_ptr(s.device),
            )  # fmt: skip
            return self._hess
        _lib.call(
            "hp_hessian", t.functor, s.npts, s.px, s.py, s.pz, s.atom_xyz, self._shell_atom, self._norms,
            t.alpha, t.order, s.rho, s.molw, s.promol, float(self.density_cutoff), M, self._hess_scratch,
            self._hess_scratch.numel(), self._hess, stream_ptr(s.device),
        )  # fmt: skip
        return self._hess

    def hessian_tiles(self):
        """(executed, total, points per sub-panel) of the 64 x 64 quadrant products of the last
        :meth:`hessian` call: the panel's block screening skips the quadrants whose basis functions vanish
        on a sub-panel of points.  flop executed = executed * 2 * 64 * 64 * points per sub-panel."""
        from .core.device import stream_ptr

        out = np.zeros(2, dtype=np.int64)
        ppt = np.zeros(1, dtype=np.int32)
        s = self.slab
        _lib.call("hp_hessian_tiles_executed", len(self._c), s.npts, self._hess_scratch, out[0:1], out[1:2], ppt,
                  stream_ptr(s.device))  # fmt: skip
        return int(out[0]), int(out[1]), int(ppt[0])

    def _objective(self, x=None, nderiv=1):
        """(f, grad[, hess]) of  f(c) = int rho ln(rho/rho0[c])  (glisa.py:411-479) from the device;
        x=None evaluates at the coefficients already in self._c.  Host NumPy out."""
        import torch

        if x is not None:
            self._c.copy_(torch.from_numpy(np.ascontiguousarray(x, dtype=float)))
        self._promol_and_entropy()
        if nderiv == 0:
            f = self._scal[1:2].clone()
            self._all_reduce(f)
            return float(f.item())
        pack = torch.cat([-self._shell_integrals(1), self._scal[1:2]])
        self._all_reduce(pack)
        host = pack.cpu().numpy()
        if nderiv == 1:
            return float(host[-1]), host[:-1].copy()
        hess = self.hessian().clone()
        self._all_reduce(hess)
        return float(host[-1]), host[:-1].copy(), hess.cpu().numpy()

    def _newton_direction(self, x):
        """(f, grad, delta) with delta = solve(H, -1 - grad), the Newton step of glisa.py:700-703.

        H = Gu^T Gu is a Gram matrix, so the system is solved where H already lives: Cholesky factorisation
        on the device (cuSOLVER through torch.linalg) followed by two steps of iterative refinement with the
        FP64 residual, which brings the step to the accuracy of the reference's LAPACK solve without moving
        the M x M matrix to the host (M = 1,500: 103 ms of host ``sysv`` against 3 ms here).  A matrix that
        is not numerically positive definite, or HP_B200_HOST_SOLVE=1, takes the reference's route
        (scipy.linalg.solve, assume_a="sym") on the host."""
        import os

        import torch

        if x is not None:
            self._c.copy_(torch.from_numpy(np.ascontiguousarray(x, dtype=float)))
        self._promol_and_entropy()
        pack = torch.cat([-self._shell_integrals(1), self._scal[1:2]])
        self._all_reduce(pack)
        hess = self.hessian()
        self._all_reduce(hess)
        host = pack.cpu().numpy()
        f, df = float(host[-1]), host[:-1].copy()
        if os.environ.get("HP_B200_HOST_SOLVE") != "1":
            rhs = (-1.0 - pack[:-1])[:, None]
            chol, info = torch.linalg.cholesky_ex(hess)
            if int(info.item()) == 0:
                delta = torch.cholesky_solve(rhs, chol)
                for _ in range(2):
                    delta = delta + torch.cholesky_solve(rhs - hess @ delta, chol)
                delta = delta[:, 0].cpu().numpy()
                if np.isfinite(delta).all():
                    return f, df, delta
        from scipy.linalg import solve

        try:
            return f, df, solve(hess.cpu().numpy(), -1 - df, assume_a="sym")
        except np.linalg.LinAlgError as exc:
            raise RuntimeError(exc)

    def _promol_population(self):
        """int rho0 over the molecular grid for the promolecule currently in slab.promol."""
        import torch

        from .core.device import stream_ptr

        s = self.slab
        sh = s.shard
        seg = (s.atom_point_offsets[sh.atom_lo : sh.atom_hi + 1] - s.point_base).contiguous()
        per_atom = torch.zeros(max(sh.nlocal, 1), dtype=torch.float64, device=s.device)
        _lib.call("hp_segment_integrate", sh.nlocal, seg, s.molw, s.promol, None, per_atom, stream_ptr(s.device))
        total = torch.zeros(1, dtype=torch.float64, device=s.device)
        _lib.call("hp_sum_partials", max(sh.nlocal, 1), per_atom, total, stream_ptr(s.device))
        self._all_reduce(total)
        return float(total.item())

    def _change(self, new, old):
        """compute_change(new, old) on the device from host coefficient vectors."""
        import torch

        dev = self.slab.device
        c_new = torch.from_numpy(np.ascontiguousarray(new, dtype=float)).to(dev)
        c_old = torch.from_numpy(np.ascontiguousarray(old, dtype=float)).to(dev)
        msd = self._device_change(c_new, c_old)
        self._all_reduce(msd)
        return float(torch.sqrt(msd.sum()).item())

    def _candidate_validity(self, candidates, check_mono):
        """is_promol_valid (glisa.py:283-307) for a stack of coefficient vectors in one launch:
        bool per candidate (all atoms admissible on their radial grids)."""
        import torch

        from .core.device import stream_ptr

        if self.on_molgrid:
            raise NotImplementedError("line searches with grid_type 2/3 are not built")
        s = self.slab
        sh = s.shard
        cand = torch.from_numpy(np.ascontiguousarray(candidates, dtype=float)).to(s.device)
        ncand, npar = cand.shape
        flags = torch.zeros((ncand, max(sh.nlocal, 1)), dtype=torch.int32, device=s.device)
        _lib.call("hp_radial_valid", sh.nlocal, sh.atom_lo, s.rad_offsets, self._par_offsets, self._bs_offsets,
                  self._bs_flat, cand, ncand, npar, float(self.negative_cutoff), int(bool(check_mono)), flags,
                  stream_ptr(s.device))  # fmt: skip
        bad = (flags != 0).any(dim=1).to(torch.int32)
        if self._comm is not None:
            from .core.comm import all_reduce

            all_reduce(self._comm, bad, "max")
        return ~bad.bool().cpu().numpy()

    def is_promol_valid(self, propars, check_mono):
        return bool(self._candidate_validity(np.asarray(propars, dtype=float)[None, :], check_mono)[0])

    def calc_promol_dens(self, propars):
        """rho0 = sum_m c_m g_m on this rank's slab, downloaded (API helper; the solvers keep it
        on the device)."""
        import torch

        self._c.copy_(torch.from_numpy(np.ascontiguousarray(propars, dtype=float)))
        self._promol_and_entropy()
        return self.slab.promol.cpu().numpy()

    def solver_newton(self, maxiter=100):
        """Exact Newton: full steps (glisa.py:573-575)."""
        return self._solver_general_newton(maxiter=maxiter, mode="exact")

    def solver_m_newton(self, maxiter=100, linspace_size=40, tau=1.0, linesearch_mode="valid-promol", **kwargs):
        """Newton direction with a back-tracking line search (glisa.py:577-593)."""
        return self._solver_general_newton(mode="modified", maxiter=maxiter, linesearch_mode=linesearch_mode,
                                           tau=tau, linspace_size=linspace_size, **kwargs)  # fmt: skip

    def solver_quasi_newton(self, mode="bfgs", maxiter=1000, niter_exact_newton=0,
                            linesearch_mode="valid-promol", tau=1.0, linspace_size=40, **kwargs):  # fmt: skip
        """BFGS inverse-Hessian updates, optionally started by exact Newton steps (glisa.py:595-615)."""
        assert mode in ["bfgs"]
        return self._solver_general_newton(mode=mode, maxiter=maxiter, niter_exact_newton=niter_exact_newton,
                                           linesearch_mode=linesearch_mode, tau=tau,
                                           linspace_size=linspace_size, **kwargs)  # fmt: skip

    def _frozen_mask(self, propars, delta):
        """1 for free coefficients, 0 for the most diffuse functions that sit at zero while the
        step pushes them negative (glisa.py:696-706 with utils.fix_propars)."""
        mask = np.ones_like(delta)
        if isinstance(self.bs_helper, ExpBasisFuncHelper):
            for a in range(self.natom):
                lo, hi = self._ranges[a], self._ranges[a + 1]
                for k in fix_propars(self.bs_helper.get_exponent(self.numbers[a]), propars[lo:hi], delta[lo:hi]):
                    mask[k + lo] = 0
        return mask

    def _line_search(self, mode, linesearch_mode, delta, propars, tau, linspace_size, check_mono, old_f):
        if mode == "exact":
            return propars + delta, delta
        mask = self._frozen_mask(propars, delta)
        scales = np.linspace(tau, 0, linspace_size, endpoint=False)
        steps = scales[:, None] * delta[None, :]
        candidates = propars[None, :] + mask[None, :] * steps
        valid = self._candidate_validity(candidates, check_mono)
        old_extended = None
        for j in np.flatnonzero(valid):
            if linesearch_mode == "valid-promol":
                return candidates[j], steps[j]
            if old_extended is None:  # extended KL of the current point: promolecule is still on the device
                old_extended = old_f + self._promol_population()
            f = self._objective(candidates[j], nderiv=0)
            if (f + self._promol_population()) - old_extended < self.negative_cutoff:
                return candidates[j], steps[j]
        raise RuntimeError("Line search failed!")

    def _solver_general_newton(self, mode="bfgs", maxiter=1000, niter_exact_newton=0,
                               linesearch_mode="valid-promol", tau=1.0, linspace_size=40, check_mono=False):  # fmt: skip
        """Newton-type minimisation of  int rho ln(rho/rho0) + int rho0  (glisa.py:617-803):
        step = solve(H, -1 - grad) with H exact ("exact", "modified") or BFGS-updated ("bfgs")."""
        if mode not in ("exact", "modified", "bfgs"):
            raise RuntimeError(f"Wrong Newton mode :{mode}. It should be one of ['exact', 'modified', 'bfgs']")
        if linesearch_mode not in ("valid-promol", "with-extended-kl"):
            raise RuntimeError(
                f"Wrong linesearch_mode {linesearch_mode}. It should be one of ['valid-promol', 'with-extended-kl']"
            )
        assert tau >= 0
        propars = self.propars
        pop = self.mol_pop
        H = olddf = oldH = step = None
        self.logger.info("            Iter.    Change    Entropy")
        self.logger.info("            -----    ------    -------")
        for irep in range(maxiter):
            old_propars = propars.copy()
            if mode == "bfgs":
                if irep == 0 or irep <= niter_exact_newton - 1:
                    if niter_exact_newton == 0:
                        f, df = self._objective(propars, 1)
                        hess = np.identity(len(df))
                    else:
                        f, df, hess = self._objective(propars, 2)
                    H = np.linalg.inv(hess)
                else:
                    f, df = self._objective(propars, 1)
                    H = bfgs(df, step, olddf, oldH)
                delta = H @ (-1 - df)
            else:
                f, df, delta = self._newton_direction(propars)
            pmin = float(self.slab.promol.min().item()) if self.slab.npts else 0.0
            propars[:], step = self._line_search(mode, linesearch_mode, delta, propars, tau, linspace_size,
                                                 check_mono, f)  # fmt: skip
            entropy = f  # _compute_entropy(rho, pro) is the objective at the old coefficients
            change = self._change(propars, old_propars)
            self.history_entropies.append(entropy)
            self.history_propars.append(propars.copy())
            self.history_changes.append(change)
            self.logger.info(f"            {irep+1:<4}    {change:.5e}    {entropy:.5e}")
            if change < self.threshold:
                # check_pro_atom_parameters on the promolecule of the old coefficients (glisa.py:786-796)
                if pmin < self.negative_cutoff:
                    raise RuntimeError("Negative pro-atom density found!")
                if abs(np.sum(propars) - pop) > self.population_cutoff:
                    self.logger.warning(
                        "WARNING: The sum of pro-atom parameters is not equal to reference population."
                    )
                self.cache.dump("niter", irep + 1, tags="o")
                return propars
            olddf, oldH = df, H
        raise RuntimeError("Not converged!")

    def _molgrid_l2_change(self, x, old_x):
        """sqrt(int (rho0[x] - rho0[old_x])^2) on the molecular grid (conv_func of solver_diis,
        glisa.py:886-893): two promolecule passes and one weighted reduction, all on the device."""
        import torch

        from .core.device import stream_ptr

        s = self.slab
        self._c.copy_(torch.from_numpy(np.ascontiguousarray(old_x, dtype=float)))
        self._promol_and_entropy()
        old = s.promol.clone()
        self._c.copy_(torch.from_numpy(np.ascontiguousarray(x, dtype=float)))
        self._promol_and_entropy()
        diff = s.promol - old
        sh = s.shard
        seg = (s.atom_point_offsets[sh.atom_lo : sh.atom_hi + 1] - s.point_base).contiguous()
        per_atom = torch.zeros(max(sh.nlocal, 1), dtype=torch.float64, device=s.device)
        _lib.call("hp_segment_integrate", sh.nlocal, seg, s.molw, diff, diff, per_atom, stream_ptr(s.device))
        total = per_atom.sum().reshape(1)
        self._all_reduce(total)
        return float(torch.sqrt(total).item())

    def _entropies_of(self, history):
        return [self._objective(x, nderiv=0) for x in history]

    def _final_check(self, propars, check_mono):
        """check_pro_atom_parameters(propars, pro_atom_density=rho0[propars], ...) with the
        reference's defaults (negativity of coefficients warns, negative promolecule raises)."""
        import warnings

        import torch

        self._c.copy_(torch.from_numpy(np.ascontiguousarray(propars, dtype=float)))
        self._promol_and_entropy()
        if (propars < -1e-12).any():
            warnings.warn("WARNING: Not all pro-atom parameters are positive!")
        pmin = self.slab.promol.min().reshape(1) if self.slab.npts else torch.zeros(1, device=self.slab.device)
        if self._comm is not None:
            from .core.comm import all_reduce

            all_reduce(self._comm, pmin, "min")
        if float(pmin.item()) < -1e-12:
            raise RuntimeError("Negative pro-atom density found!")
        if abs(np.sum(propars) - self.mol_pop) > 1e-4:
            self.logger.warning("WARNING: The sum of pro-atom parameters is not equal to reference population.")

    def solver_diis(self, use_dmrs=False, **diis_options):
        """DIIS on the fixed-point map (glisa.py:883-925)."""

        def conv_func(residual, x, old_x):
            return np.linalg.norm(residual) if use_dmrs else self._molgrid_l2_change(x, old_x)

        propars = self.propars
        propars[:], niter, history = diis(propars, self.function_g, self.threshold, conv_func=conv_func,
                                          verbose=True, logger=self.logger, **diis_options)  # fmt: skip
        self._final_check(propars, False)
        self.cache.dump("niter", niter, tags="o")
        self.history_entropies.extend(self._entropies_of(history[1:]))
        self.history_propars = history[1:]
        return propars

    def residual(self, x):
        return self.function_g(x) - x

    def solver_cdiis(self, **cdiis_options):
        """Restarted / adaptive-depth CDIIS on the fixed-point map (glisa.py:993-1028)."""
        conv, nbiter, rnormlist, _, _, propars, history = cdiis(
            self.propars, self.function_g, self.threshold, self.maxiter, logger=self.logger, verbose=True,
            **cdiis_options)  # fmt: skip
        if not conv:
            raise RuntimeError("Not converged!")
        self._final_check(propars, False)
        self.cache.dump("niter", nbiter, tags="o")
        self.history_entropies.extend(self._entropies_of(history[1:]))
        self.history_propars = history[1:]
        self.history_changes = rnormlist
        return propars

    def solver_cvxopt(self, allow_neg_pars=False, verbose=False, **cvxopt_options):
        """The global convex programme  min f(c)  s.t.  c >= 0 (optional), sum c = N  of
        glisa.py:488-570.  The reference hands it to the third-party ``cvxopt.solvers.cp``; here it
        goes to the interior-point method of ``algo/cp.py`` (unique minimiser, see there) with f, its
        gradient and its Hessian from the device kernels.  ``niter`` counts the distinct points at
        which the Hessian was evaluated, as in the reference."""
        from .algo.cp import cp

        propars = self.propars
        nb_par = len(propars)
        mol_pop = self.mol_pop
        history = [propars.copy()]

        def objective(x=None, z=None):
            if x is None:
                return 0, propars.copy()
            x = np.asarray(x, dtype=float).ravel()
            if z is None:
                return self._objective(x, 1)
            f, df, hess = self._objective(x, 2)
            if not np.allclose(x, history[-1]):
                history.append(x.copy())
            return f, df, z[0] * hess

        G = h = None
        if not allow_neg_pars:
            G, h = -np.identity(nb_par), np.zeros(nb_par)
        options = dict(cvxopt_options)
        options.setdefault("show_progress", bool(verbose))
        options.setdefault("printer", self.logger.info)
        sol = cp(objective, G=G, h=h, A=np.ones((1, nb_par)), b=np.array([mol_pop]), options=options)
        propars[:] = sol["x"]
        self._final_check(propars, check_mono=False)
        if sol["status"] != "optimal":
            raise RuntimeError("CVXOPT not converged!")
        self.cache.dump("niter", len(history) - 1, tags="o")
        self.history_entropies.extend(self._entropies_of(history[:-1]))
        self.history_propars = history[1:]
        return propars

    def solver_trust_region(self, allow_neg_pars=False):
        """SciPy ``trust-constr`` with SR1 updates on the extended objective
        f + int rho0 - N  (glisa.py:927-987); f and its gradient come from the device."""
        from scipy.optimize import SR1, LinearConstraint, minimize

        pars0 = self.propars
        pop = self.mol_pop
        nb_par = len(pars0)

        def cost_grad(x):
            f, df = self._objective(x, 1)
            return f + (self._promol_population() - pop), df + 1

        if allow_neg_pars:
            bounds, constraint = None, LinearConstraint(np.ones((1, nb_par)), pop, pop)
        else:
            bounds, constraint = [(0.0, 200)] * nb_par, None
        previous = []

        def callback(current, state):
            previous.append(np.array(current, copy=True))
            if len(previous) >= 2:
                change = self._molgrid_l2_change(previous[-1], previous[-2])
                return change < self._threshold and state["status"]

        result = minimize(cost_grad, pars0, method="trust-constr", jac=True, hess=SR1(), bounds=bounds,
                          constraints=constraint, callback=callback,
                          options={"gtol": self._threshold, "maxiter": self._maxiter, "verbose": 0})  # fmt: skip
        self.logger.info(f'Optimizer message: "{result.message}"')
        if not result.success:
            raise RuntimeError("Convergence failure.")
        return result.x

    # -- driver ---------------------------------------------------------------------------------
    @just_once
    def do_partitioning(self):
        new = any(f"at_weights_{i}" not in self.cache for i in range(self.natom))
        new |= "niter" not in self.cache
        if not new:
            return
        import torch

        from .core.device import stream_ptr

        self._init_propars()
        t0 = time.time()
        new_propars = self._opt_propars(**self._solver_options)
        self.history_time_update_propars.append(time.time() - t0)
        propars = self.cache.load("propars")
        propars[:] = new_propars
        self._c.copy_(torch.from_numpy(np.ascontiguousarray(propars)))

        # update_at_weights(force_on_molgrid=True) + charges (glisa.py:267-278): promolecule WITH the
        # 1e-100 offsets, N_a = int clip(rho0_a/rho0, 0, 1) rho, without natom x Npts weight arrays
        t0 = time.time()
        self._refresh_table()
        t, s = self._table, self.slab
        pops = torch.zeros(self.natom, dtype=torch.float64, device=s.device)
        if self._numeric:
            # eval_proatom of the reference adds no offset to a tabulated pro-atom (glisa.py:226-246);
            # update_pro adds 1e-100 per atom to the promolecule
            t.promol_weights(self.density_cutoff, True, True, False, proatom_offset=0.0)
            nblk = int(_lib.call("hp_spline_integral_blocks", s.npts))
            partial = torch.zeros(self.natom * nblk, dtype=torch.float64, device=s.device)
            _lib.call("hp_atom_weight_integrals_spline", s.npts, s.px, s.py, s.pz, self.natom, s.atom_xyz, t.offsets,
                      t.knots, t.coef, 0.0, s.rho, s.molw, s.promol, partial, pops, stream_ptr(s.device))  # fmt: skip
        else:
            t.promol_weights(self.density_cutoff, True, True, False)
            _lib.call(
                "hp_atom_weight_integrals", t.functor, s.npts, s.px, s.py, s.pz, s.natom, s.atom_xyz, t.offsets,
                t.A, t.alpha, t.order, t.ntile, t.tiles, s.rho, s.molw, s.promol, self._partial, pops,
                stream_ptr(s.device),
            )  # fmt: skip
        self._all_reduce(pops)
        charges = self.cache.load("charges", alloc=self.natom, tags="o")[0]
        charges[:] = self.pseudo_numbers - pops.cpu().numpy()
        self._publish_weights()  # owner-slice weights (the reference caches full-grid arrays here)
        self.history_time_update_at_weights.append(time.time() - t0)
        self._finalize_propars()

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
