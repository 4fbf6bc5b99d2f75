"""Gaussian ISA (GISA): pro-atoms expanded in fixed exponential basis functions.

Counterpart of the reference's ``gisa.py`` (``GaussianISAWPart`` :109-345, ``init_propars`` :68-88,
``evaluate_basis_functions`` :91-106, QP interface :348-421).  The per-iteration grid passes run on
the GPU; GISA's own per-atom update is a small strictly convex quadratic programme (K <= 12
unknowns) that stays on the host as in the reference: through the third-party ``qpsolvers``
package when it is installed, otherwise through the exact active-set solver in ``algo/qp.py``.
"""

from __future__ import annotations

import warnings

import numpy as np

from . import _lib
from .algo.qp import solve_qp_simplex
from .core.basis import ExpBasisFuncHelper, shell_norm
from .core.cache import just_once
from .core.iterstock import AbstractISAWPart
from .core.logging import deflist
from .utils import check_pro_atom_parameters, optional_package

__all__ = ["GaussianISAWPart", "get_proatom_rho", "init_propars", "evaluate_basis_functions"]


def get_proatom_rho(part, iatom, propars=None):
    """Pro-atom density and derivative of atom ``iatom`` on its radial grid (host helper)."""
    if propars is None:
        propars = part.cache.load("propars")
    mine = propars[part._ranges[iatom] : part._ranges[iatom + 1]]
    points = part.radial_distances[iatom] if part.on_molgrid else part.get_rgrid(iatom).points
    return part.bs_helper.compute_proatom_dens(part.numbers[iatom], mine, points, 1)


def init_propars(part):
    """Initial coefficients: table initials (floored at 1e-4), scaled per atom to its pseudo
    number and globally to the number of electrons (gisa.py:68-88)."""
    part._nshells = [part.bs_helper.get_nshell(z) for z in part.numbers]
    part._ranges = [0]
    for k in part._nshells:
        part._ranges.append(part._ranges[-1] + k)
    propars = part.cache.load("propars", alloc=part._ranges[-1], tags="o")[0]
    propars[:] = 1.0
    for a in range(part.natom):
        inits = part.bs_helper.get_initial(part.numbers[a])
        inits[inits < 1e-4] = 1e-4  # in place, as the reference does (mutates the helper's table)
        propars[part._ranges[a] : part._ranges[a + 1]] = inits / np.sum(inits) * part.pseudo_numbers[a]
    propars[:] = propars / np.sum(propars) * part.nelec
    part.initial_propars_modified = propars.copy()
    return propars


def evaluate_basis_functions(part, force_on_molgrid=False):
    """Unit-population basis functions -> cache ``bs_funcs_{a}`` (gisa.py:91-106): on each atom's
    radial grid, or -- only for host plug-in solvers with grid_type 2/3, the device solvers
    regenerate them in-kernel -- on the whole molecular grid."""
    molgrid = part.on_molgrid or force_on_molgrid
    shared = {}  # radial grids: atoms of one element on the same radial grid object share the table
    for a in range(part.natom):
        k = part._ranges[a + 1] - part._ranges[a]
        if molgrid:
            r = part.radial_distances[a]
            table = None
        else:
            rgrid = part.get_rgrid(a)
            r = rgrid.points
            table = shared.get((int(part.numbers[a]), id(rgrid)))
        if table is None:
            table = np.array([part.bs_helper.compute_proshell_dens(part.numbers[a], i, 1.0, r) for i in range(k)])
            if not molgrid:
                shared[(int(part.numbers[a]), id(rgrid))] = table
        bs = part.cache.load(f"bs_funcs_{a}", alloc=(k, r.size))[0]
        bs[:, :] = table


def expbasis_atom_work(coordinates, numbers, pseudo_numbers, grid, bs_helper, device=None):
    """Load-balancing weights of a sharded run of the exponential-basis schemes: pairs the screened
    dense pass evaluates per atom block, estimated from the table's initial coefficients (scaled per
    atom like ``init_propars``; the global normalisation to the electron count cancels in the
    screening test).  None for bases with mixed orders (no screening there)."""
    from .core.device import estimate_dense_work

    per_element = {}
    for z in np.unique(numbers):
        z = int(z)
        order = np.asarray(bs_helper.get_order(z), float)
        alpha = np.asarray(bs_helper.get_exponent(z), float)
        inits = np.maximum(np.asarray(bs_helper.get_initial(z), float), 1e-4)
        per_element[z] = (order, alpha, inits / inits.sum() * shell_norm(order, alpha))
    orders = np.concatenate([v[0] for v in per_element.values()])
    if not (np.all(orders == 2.0) or np.all(orders == 1.0)):
        return None
    shells, kinds = [], []
    for a, z in enumerate(numbers):
        _, alpha, amp = per_element[int(z)]
        shells.append((amp * float(pseudo_numbers[a]), alpha))
        kinds.append((int(z), float(pseudo_numbers[a]), id(grid.atgrids[a].rgrid), int(grid.atgrids[a].size)))
    return estimate_dense_work(coordinates, grid, shells, gaussian=bool(np.all(orders == 2.0)), kinds=kinds,
                               device=device)


def molgrid_host_update(promol, rho, points, weights, bs_funcs, ranges, propars, pseudo_numbers, opt_propars):
    """Per-atom parameter updates of ONE outer iteration on the molecular grid, through a host
    solver (gisa.py:257-279 with core/stockholder.py:352-384 and core/iterstock.py:32-45).

    ``promol`` is the promolecule the device pass built from ``propars`` (with the reference's
    1e-100 offsets); everything here is the K_a x Npts dense algebra the reference also does on
    the host for a plug-in solver.  ``opt_propars(a, bs_funcs_a, rho_a, propars_a)`` returns the new
    coefficients of atom ``a``.  Returns (new propars, charges, per-atom change terms)."""
    natom = len(ranges) - 1
    new = propars.copy()
    charges = np.zeros(natom)
    msd = np.zeros(natom)
    for a in range(natom):
        lo, hi = ranges[a], ranges[a + 1]
        bs = bs_funcs[a]
        at_weights = np.einsum("k,kp->p", propars[lo:hi], bs)
        at_weights /= promol
        np.clip(at_weights, 0, 1, out=at_weights)
        rho_a = at_weights * rho
        new[lo:hi] = opt_propars(a, bs, rho_a, propars[lo:hi].copy())
        charges[a] = pseudo_numbers[a] - np.einsum("i,i", weights, rho_a)
        delta = np.einsum("k,kp->p", new[lo:hi] - propars[lo:hi], bs)
        msd[a] = np.einsum("i,i,i", weights, delta, delta)
    return new, charges, msd


class GaussianISAWPart(AbstractISAWPart):
    name = "gisa"
    #: solver names that run as CUDA kernels (subclasses extend this)
    device_solvers = {}

    def __init__(self, coordinates, numbers, pseudo_numbers, grid, moldens, spindens=None, lmax=3,
                 logger=None, threshold=1e-6, maxiter=500, inner_threshold=1e-8,
                 radius_cutoff=np.inf, solver="quadprog", solver_options=None, grid_type=1,
                 **kwargs):  # fmt: skip
        self._solver = solver
        self._solver_options = solver_options or {}
        if not hasattr(self, "_bs_helper"):
            self._bs_helper = None
        self._ranges = None
        super().__init__(coordinates, numbers, pseudo_numbers, grid, moldens, spindens, lmax=lmax,
                         logger=logger, threshold=threshold, maxiter=maxiter,
                         inner_threshold=inner_threshold, radius_cutoff=radius_cutoff,
                         grid_type=grid_type, **kwargs)  # fmt: skip

    @property
    def bs_helper(self):
        if self._bs_helper is None:
            self._bs_helper = ExpBasisFuncHelper.from_function_type()
        return self._bs_helper

    def _init_log_scheme(self):
        deflist(
            self.logger,
            [
                ("Scheme", "Gaussian Iterative Stockholder Analysis (GISA)"),
                ("Outer loop convergence threshold", "%.1e" % self._threshold),
                ("Inner loop convergence threshold", "%.1e" % self._inner_threshold),
                ("Maximum iterations", self._maxiter),
                ("lmax", self._lmax),
                ("Solver", self._solver),
                ("Grid type", self.grid_type),
            ],
        )
        if callable(self._solver):
            warnings.warn("Customized solver is used, the argument `inner_threshold` is not used.")

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
        return get_proatom_rho(self, iatom, propars)

    # -- device hooks ---------------------------------------------------------------------------
    def _estimate_atom_work(self):
        if self.on_molgrid or self._local_radius is not None or self._grid.atgrids is None:
            return None
        if not isinstance(self.bs_helper, ExpBasisFuncHelper):
            return None  # spline basis: no screening, blocks are balanced by point count
        return expbasis_atom_work(self.coordinates, self.numbers, self.pseudo_numbers, self._grid, self.bs_helper,
                                  self._device)

    def _init_propars(self):
        import torch

        from .core.device import ShellTable, to_device

        propars = init_propars(self)
        if not self.on_molgrid:
            self._evaluate_basis_functions()
        slab = self.slab
        dev = slab.device
        self._numeric = not isinstance(self.bs_helper, ExpBasisFuncHelper)
        if self._numeric:
            # tabulated basis functions (basis_type="numeric"): the pro-atom sum_k c_k S_k(r) is one
            # piecewise cubic per atom on the element's knots; coefficients are mixed per iteration
            from .isa import SplineTable

            class _Knots:  # the SplineTable only needs .points / .size
                def __init__(self, x):
                    self.points, self.size = x, x.size

            self._table = SplineTable(slab, [_Knots(self.bs_helper.get_knots(z)) for z in self.numbers],
                                      proatom_offset=0.0)  # gisa.py:224-244: no offset on the pro-atom
            self._ppoly = {int(z): self.bs_helper.ppoly_coefficients(int(z)) for z in np.unique(self.numbers)}
        else:
            self._init_shell_table(slab, dev)
        st = self._alloc_state(len(propars))
        st.propars.copy_(to_device(propars, dev))
        self._par_offsets = to_device(np.asarray(self._ranges, dtype=np.int32), dev)
        self._pseudo = to_device(self.pseudo_numbers, dev, np.float64)
        self._molgrid_host = False
        if self.on_molgrid and (callable(self._solver) or self._solver not in self.device_solvers or self._numeric):
            # host plug-in on the molecular grid: K_a x Npts basis tables on the host, as in the reference
            # (also for tabulated basis functions, which the molecular-grid kernels do not regenerate)
            if self._comm is not None:
                raise NotImplementedError("host plug-in solvers with grid_type 2/3 run on one GPU (the "
                                          "device solvers " + str(list(self.device_solvers)) + " shard)")  # fmt: skip
            self._molgrid_host = True
            self._evaluate_basis_functions()
            return propars
        if self.on_molgrid:
            self.molgrid_single_update = bool(self.device_solvers[self._solver][1])
            opt = self.device_solvers[self._solver][0]
            self.molgrid_max_inner = int(float(self._solver_options.get(opt, 100000))) if opt else 1
            return propars
        # radial-grid basis functions of the local atoms, concatenated (K_a x nrad_a row-major)
        sh = slab.shard
        blocks = [self.cache.load(f"bs_funcs_{a}") for a in range(sh.atom_lo, sh.atom_hi)]
        offs = np.concatenate([[0], np.cumsum([b.size for b in blocks])]).astype(np.int64)
        flat = np.concatenate([b.ravel() for b in blocks]) if blocks else np.zeros(0)
        self._bs_offsets = to_device(offs, dev)
        self._bs_flat = to_device(flat, dev) if flat.size else torch.zeros(1, dtype=torch.float64, device=dev)
        self._nrad_max = int(np.max(np.diff(slab.rad_offsets_host))) if sh.nlocal else 1
        self._nshell_max = max(self._nshells[sh.atom_lo : sh.atom_hi], default=1)
        return propars

    def _init_shell_table(self, slab, dev):
        from .core.device import ShellTable, to_device

        orders = np.concatenate([np.asarray(self.bs_helper.get_order(z), float) for z in self.numbers])
        alphas = np.concatenate([np.asarray(self.bs_helper.get_exponent(z), float) for z in self.numbers])
        if np.all(orders == 2.0):
            functor = 2
        elif np.all(orders == 1.0):
            functor = 1
        else:
            functor = 3
        self._table = ShellTable(slab, functor, self._nshells)
        self._table.alpha.copy_(to_device(alphas, dev))
        if functor == 3:
            self._table.order.copy_(to_device(orders, dev))
        self._norms = to_device(shell_norm(orders, alphas), dev)

    def _refresh_table(self):
        from .core.device import stream_ptr

        if self._numeric:
            import torch

            propars = self.cache.load("propars")
            mixed = [np.einsum("k,ksc->sc", propars[self._ranges[a] : self._ranges[a + 1]], self._ppoly[int(z)])
                     for a, z in enumerate(self.numbers)]  # fmt: skip
            self._table.coef.copy_(torch.from_numpy(np.concatenate([m.ravel() for m in mixed])))
            return
        t = self._table
        _lib.call("hp_table_scaled", t.nshell, self._state.propars, self._norms, t.A, stream_ptr(self.slab.device))

    def _molgrid_shell_params(self, propars):
        return propars * self._norms, self._table.alpha

    def _molgrid_apply(self, propars, s0, s1, shell_active):
        import torch

        return torch.where(shell_active, s0, propars)  # alisa.py:268

    def _run_iteration_molgrid(self):
        if not self._molgrid_host:
            return super()._run_iteration_molgrid()
        return self._run_iteration_molgrid_host()

    def _run_iteration_molgrid_host(self):
        """One outer iteration with grid_type 2/3 and a host solver: promolecule and entropy on the
        device, the per-atom K_a x Npts problems on the host (``molgrid_host_update``)."""
        import torch

        from .core.device import stream_ptr

        st, slab = self._state, self.slab
        dev = slab.device
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        self._refresh_table()
        self._table.promol_weights(self.density_cutoff, True, True, True)
        ev[1].record()
        promol = slab.promol.cpu().numpy()
        propars = self.cache.load("propars")
        grid = self.grid
        alphas = [self.bs_helper.get_exponent(z) if isinstance(self.bs_helper, ExpBasisFuncHelper) else None
                  for z in self.numbers]  # fmt: skip

        def opt(a, bs, rho_a, start):
            return self._opt_propars(bs, rho_a, start, grid.points, grid.weights, alphas[a], self._inner_threshold)

        new, charges, msd = molgrid_host_update(
            promol, self._moldens, grid.points, grid.weights, [self.cache.load(f"bs_funcs_{a}") for a in range(self.natom)],
            self._ranges, propars, self.pseudo_numbers, opt)  # fmt: skip
        st.propars.copy_(torch.from_numpy(new).to(dev))
        st.charges.copy_(torch.from_numpy(charges).to(dev))
        st.msd.copy_(torch.from_numpy(msd).to(dev))
        _lib.call("hp_finish_iteration", slab.npartial, slab.entropy_partials, self.natom, st.msd, st.out2,
                  stream_ptr(dev))  # fmt: skip
        ev[2].record()
        out2 = st.out2.cpu().numpy()
        st.events.append(ev)
        propars[:] = new
        self.cache.load("charges", alloc=self.natom, tags="o")[0][:] = charges
        return float(out2[0]), float(out2[1])

    @property
    def device_loop_capable(self):
        """Only the device solvers keep the whole iteration on the GPU (host plug-ins need the spherical
        averages on the host every iteration; tabulated bases mix their splines on the host)."""
        return (not callable(self._solver)) and self._solver in self.device_solvers and not self._numeric

    def _launch_radial_update(self):
        self.slab.shell_project()
        if not callable(self._solver) and self._solver in self.device_solvers:
            self._launch_device_solver(self.device_solvers[self._solver])
        else:
            self._host_radial_update()

    def _launch_device_solver(self, spec):
        raise NotImplementedError

    def _host_radial_update(self):
        """Per-atom updates through a host solver (user callable or the QP interface): the
        spherical averages come back from the device (natom x nrad doubles), parameters go up."""
        import torch

        slab, st = self.slab, self._state
        sph = slab.sph_avg.cpu().numpy()
        propars = self.cache.load("propars")
        old = propars.copy()
        sh, ro = slab.shard, slab.rad_offsets_host
        charges = np.zeros(self.natom)
        msd = np.zeros(self.natom)
        batched = self._batched_host_solver()
        exp_basis = isinstance(self.bs_helper, ExpBasisFuncHelper)
        atoms = []  # (atom, bs_funcs, spherical average, radial points, 4 pi r^2 w) of the local atoms
        shell_weights = {}  # 4 pi r^2 w per distinct radial grid
        for i, a in enumerate(range(sh.atom_lo, sh.atom_hi)):
            rgrid = self.get_rgrid(a)
            if id(rgrid) not in shell_weights:
                shell_weights[id(rgrid)] = 4 * np.pi * rgrid.points**2 * rgrid.weights
            atoms.append((a, self.cache.load(f"bs_funcs_{a}"), sph[ro[i] : ro[i + 1]], rgrid.points,
                          shell_weights[id(rgrid)]))  # fmt: skip
        solved = None
        if batched is not None:  # all local atoms in stacked NumPy calls instead of one by one
            solved = batched([(bs, rho_sph, propars[self._ranges[a] : self._ranges[a + 1]].copy(), points, r_weights)
                              for a, bs, rho_sph, points, r_weights in atoms])  # fmt: skip
        for i, (a, bs, rho_sph, points, r_weights) in enumerate(atoms):
            lo, hi = self._ranges[a], self._ranges[a + 1]
            if solved is not None:
                propars[lo:hi] = solved[i]
            else:
                alphas = self.bs_helper.get_exponent(self.numbers[a]) if exp_basis else None
                propars[lo:hi] = self._opt_propars(bs, rho_sph, propars[lo:hi].copy(), points, r_weights, alphas,
                                                   self._inner_threshold)  # fmt: skip
            charges[a] = self.pseudo_numbers[a] - np.einsum("i,i", r_weights, rho_sph)
            # the atom's term of compute_change (core/iterstock.py:32-45) from the tabulated unit shells,
            # like the device solvers: rho0_a[new] - rho0_a[old] = sum_k (new - old)_k g_k
            delta = np.einsum("k,kp->p", propars[lo:hi], bs) - np.einsum("k,kp->p", old[lo:hi], bs)
            msd[a] = np.einsum("i,i,i", r_weights, delta, delta)
        dev = slab.device
        st.propars[self._ranges[sh.atom_lo] : self._ranges[sh.atom_hi]] = torch.from_numpy(
            propars[self._ranges[sh.atom_lo] : self._ranges[sh.atom_hi]]
        ).to(dev)
        st.charges[sh.atom_lo : sh.atom_hi] = torch.from_numpy(charges[sh.atom_lo : sh.atom_hi]).to(dev)
        st.msd[sh.atom_lo : sh.atom_hi] = torch.from_numpy(msd[sh.atom_lo : sh.atom_hi]).to(dev)

    def _batched_host_solver(self):
        """A callable solving the radial problems of ALL local atoms at once, or None when the
        solver has to be called atom by atom (user callables, solvers without a stacked version)."""
        return None

    def _opt_propars(self, bs_funcs, rho, propars, points, weights, alphas, threshold):
        if callable(self._solver):
            return self._solver(bs_funcs, rho, propars, points, weights, alphas, threshold, **self._solver_options)
        return opt_propars_qp_interface(bs_funcs, rho, propars, weights, alphas, self._solver, **self._solver_options)

    @just_once
    def _evaluate_basis_functions(self):
        evaluate_basis_functions(self)

    def _finalize_propars(self):
        AbstractISAWPart._finalize_propars(self)
        if self.on_molgrid:
            return
        slab = self.slab
        sph = slab.sph_avg.cpu().numpy()
        ro = slab.rad_offsets_host
        for i, a in enumerate(range(slab.shard.atom_lo, slab.shard.atom_hi)):
            self.cache.dump(f"radial_points_{a}", slab.rad_r_host[ro[i] : ro[i + 1]], tags="o")
            self.cache.dump(f"spherical_average_{a}", sph[ro[i] : ro[i + 1]], tags="o")
            self.cache.dump(f"radial_weights_{a}", slab.rad_w_host[ro[i] : ro[i + 1]], tags="o")


def opt_propars_qp_interface(bs_funcs, rho, propars, weights, alphas, solver="quadprog", **solver_options):
    """GISA's quadratic programme  min 1/2 c^T P c + q^T c,  c >= 0,  sum c = pop  (gisa.py:348-421).

    With the third-party ``qpsolvers`` package installed the call is handed to it exactly like the
    reference does.  The package is not in this image; the programme is strictly convex, hence its
    minimiser unique, and ``algo.qp.solve_qp_simplex`` (exact active-set method) returns it.
    ``solver="active-set"`` selects the built-in solver unconditionally."""
    nprim = bs_funcs.shape[0]
    s = alphas[:, None] + alphas[None, :]
    P = 2 / np.pi**1.5 * (alphas[:, None] * alphas[None, :]) ** 1.5 / s**1.5
    P = (P + P.T) / 2
    q = -2 * np.einsum("i,ni,i->n", weights, bs_funcs, rho)
    pop = np.einsum("i,i", weights, rho)
    result = None
    if solver != "active-set":
        qpsolvers = optional_package("qpsolvers")
        if qpsolvers is not None:
            result = qpsolvers.solve_qp(
                P, q, -np.identity(nprim), np.zeros((nprim, 1)), np.ones((1, nprim)), np.ones((1, 1)) * pop,
                solver=solver, initvals=np.zeros_like(propars), **solver_options,
            )  # fmt: skip
    if result is None:
        result = solve_qp_simplex(P, q, float(pop))
    check_pro_atom_parameters(result, total_population=float(pop))
    return result
