"""Linear ISA with local (per-atom) solvers: aLISA.

Counterpart of the reference's ``LinearISAWPart`` (alisa.py:1155-1335).  Built-in solvers that run
as CUDA kernels (one warp per atom, all atoms in one launch):

    "sc"         self-consistent fixed point to ``inner_threshold``        (alisa.py:193-291)
    "sc-1-iter"  a single fixed-point update per outer iteration           (alisa.py:294-353)

A callable ``solver`` keeps the reference's plug-in signature
``solver(bs_funcs, rho, propars, points, weights, threshold, logger, density_cutoff=...,
negative_cutoff=..., population_cutoff=..., **solver_options)`` (alisa.py:1304-1335) and runs on the
host on the projected radial densities (natom x nrad doubles come back from the device per outer
iteration).  The remaining built-in names of the reference ("diis", "cdiis", "newton", "m-newton",
"quasi-newton", "trust-region", "cvxopt", "sc-plus-convex") are shipped as exactly such plug-ins
(``lisa_solvers.py``): small dense algebra per atom on K x nrad arrays, while promolecule, weights
and projection stay in the CUDA kernels.
"""

from __future__ import annotations

import numpy as np

from . import _lib
from .core.basis import AnalyticBasisFuncHelper, ExpBasisFuncHelper, NumericBasisFuncHelper
from .core.logging import deflist
from .gisa import GaussianISAWPart
from .utils import optional_package
from .lisa_solvers import (  # noqa: F401  (re-exported like the reference's alisa module)
    HOST_SOLVERS,
    solver_cdiis,
    solver_cvxopt,
    solver_cvxopt_batched,
    solver_diis,
    solver_m_newton,
    solver_newton,
    solver_quasi_newton,
    solver_sc,
    solver_sc_1_iter,
    solver_sc_plus_cvxopt,
    solver_trust_region,
)

__all__ = ["LinearISAWPart", "setup_bs_helper"] + [f.__name__ for f in HOST_SOLVERS.values()] + ["solver_sc", "solver_sc_1_iter"]


def setup_bs_helper(part):
    """Resolve ``part.basis_func`` ("gauss", "slater", a file name, or a helper instance)."""
    if part._bs_helper is None:
        bf = part.basis_func
        if isinstance(bf, str):
            if bf.lower() in ("gauss", "slater"):
                part.logger.info(f"Load {bf.upper()} basis functions")
                if part.basis_type == "analytic":
                    part._bs_helper = ExpBasisFuncHelper.from_function_type(bf.lower())
                elif part.basis_type == "numeric":
                    part._bs_helper = NumericBasisFuncHelper.from_function_type(bf.lower())
                else:
                    raise RuntimeError("The bs_type should be one of analytic and numeric.")
            else:
                part.logger.info(f"Load basis functions from custom json file: {bf}")
                part._bs_helper = ExpBasisFuncHelper.from_file(bf)
        elif isinstance(bf, (AnalyticBasisFuncHelper, NumericBasisFuncHelper)):
            part._bs_helper = bf
        else:
            raise NotImplementedError("The type of basis_func should be one of string or class BasisFuncHelper.")
    return part._bs_helper


class LinearISAWPart(GaussianISAWPart):
    name = "lisa"
    # name -> (max inner iterations option, single update?)
    device_solvers = {"sc": ("max_niter_inner", False), "sc-1-iter": (None, True)}
    #: every built-in name of the reference (alisa.py:1168-1203); those not in device_solvers run
    #: as host plug-ins on the projected radial problem
    builtin_solvers = {"sc": solver_sc, "sc-1-iter": solver_sc_1_iter, **HOST_SOLVERS}

    def __init__(self, coordinates, numbers, pseudo_numbers, grid, moldens, spindens=None, lmax=3,
                 logger=None, threshold=1e-6, maxiter=500, inner_threshold=1e-8,
                 radius_cutoff=np.inf, solver="cvxopt", solver_options=None, basis_func="gauss",
                 basis_type="analytic", grid_type=1, **kwargs):  # fmt: skip
        self.basis_func = basis_func
        self._func_type = basis_func.upper() if basis_func in ("gauss", "slater") else "Customized"
        self.basis_type = basis_type
        self._bs_helper = None
        super().__init__(coordinates, numbers, pseudo_numbers, grid, moldens, spindens, lmax, logger,
                         threshold, maxiter, inner_threshold, radius_cutoff, solver, solver_options,
                         grid_type, **kwargs)  # fmt: skip
        if self.grid_type not in [1]:
            self.logger.info(
                f"The grid type is {self.grid_type} and please set `check_mono` to `False` when using aLISA+- methods."
            )

    @property
    def bs_helper(self):
        return setup_bs_helper(self)

    def _init_log_scheme(self):
        info = [
            ("Scheme", "Linear Iterative Stockholder"),
            ("Outer loop convergence threshold", "%.1e" % self._threshold),
            ("Inner loop convergence threshold", "%.1e" % self._inner_threshold),
            ("Using global ISA", False),
            ("Maximum outer iterations", self._maxiter),
            ("lmax", self._lmax),
            ("Solver", self._solver.__name__ if callable(self._solver) else self._solver.upper()),
            ("Basis function type", self._func_type),
            ("Grid type", self.grid_type),
        ]
        info += [(f"Solver options -- {k}", str(v)) for k, v in self._solver_options.items()]
        deflist(self.logger, info)
        self.logger.info(" ")

    def _launch_device_solver(self, spec):
        from .core.device import stream_ptr

        opt_name, single = spec
        max_inner = 1 if single else int(float(self._solver_options.get(opt_name, 100000)))
        slab, st = self.slab, self._state
        sh = slab.shard
        _lib.call(
            "hp_lisa_sc_radial_solve", sh.nlocal, sh.atom_lo, slab.rad_offsets, slab.rad_w4, slab.sph_avg,
            self._par_offsets, st.propars, self._bs_offsets, self._bs_flat, self._pseudo,
            float(self._inner_threshold), float(self.density_cutoff), float(self.population_cutoff),
            max_inner, int(single), self._nrad_max, self._nshell_max, st.charges, st.msd, st.niter,
            st.flags, stream_ptr(slab.device),
        )  # fmt: skip

    def _batched_host_solver(self):
        """The built-in convex programme (the default solver) has a stacked version; it is used
        unless the caller asks for the third-party engine, sign-free coefficients or the package is
        installed (then the per-atom call hands the programme to it, as the reference does)."""
        opts = self._solver_options
        if callable(self._solver) or self._solver != "cvxopt" or opts.get("allow_neg_params", False):
            return None
        if opts.get("engine") == "cvxopt" or (opts.get("engine") is None and optional_package("cvxopt") is not None):
            return None
        engine_free = {k: v for k, v in opts.items() if k not in ("engine", "allow_neg_params")}

        def solve(problems):
            return solver_cvxopt_batched(
                problems, self._inner_threshold, self.logger, density_cutoff=self.density_cutoff,
                negative_cutoff=self.negative_cutoff, population_cutoff=self.population_cutoff, **engine_free)

        return solve

    def _opt_propars(self, bs_funcs, rho, propars, points, weights, alphas, threshold):
        if callable(self._solver):
            solver = self._solver
        elif self._solver in self.builtin_solvers:
            solver = self.builtin_solvers[self._solver]
        else:
            raise NotImplementedError
        return solver(
            bs_funcs, rho, propars, points, weights, threshold, self.logger,
            density_cutoff=self.density_cutoff, negative_cutoff=self.negative_cutoff,
            population_cutoff=self.population_cutoff, **self._solver_options,
        )  # fmt: skip

    def _finalize_propars(self):
        GaussianISAWPart._finalize_propars(self)
        if not callable(self._solver) and self._solver == "sc" and not self.on_molgrid:
            flags = self._state.flags.cpu().numpy()
            if (flags & 1).any():
                self.logger.warning("Warning: Inner iteration is not converge!")
            if (flags & 2).any():
                self.logger.warning("WARNING: The sum of pro-atom parameters is not equal to reference population.")
