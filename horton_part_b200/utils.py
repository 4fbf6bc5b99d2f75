"""Constants, geometry type checks and the scheme registry.

Mirrors /root/reference/src/horton_part/utils.py:51-59 (cut-offs from data/constants.yaml),
:62-104 (``wpart_schemes``), :107-186 (``typecheck_geo``), :198-252 (``compute_quantities``, the
1-D radial helper of the host plug-in solvers) and :255-577 (validity checks, ``fix_propars``).
"""

from __future__ import annotations

import warnings

import numpy as np

__all__ = [
    "optional_package",
    "wpart_schemes",
    "typecheck_geo",
    "DENSITY_CUTOFF",
    "NEGATIVE_CUTOFF",
    "POPULATION_CUTOFF",
    "ANGSTROM",
    "compute_quantities",
    "check_pro_atom_parameters",
    "check_pro_atom_parameters_neg_pars",
    "check_pro_atom_parameters_non_neg_pars",
    "check_pars_population",
    "check_pars_negativity",
    "check_dens_negativity",
    "check_dens_monotonicity",
    "fix_propars",
]

# data/constants.yaml of the reference
DENSITY_CUTOFF = 1.0e-15
NEGATIVE_CUTOFF = -1.0e-12
POPULATION_CUTOFF = 1.0e-4
ANGSTROM = 1.889726133921252

_SCHEMES = {
    "mbis": ("mbis", "MBISWPart"),
    "is": ("isa", "ISAWPart"),
    "lisa": ("alisa", "LinearISAWPart"),
    "gisa": ("gisa", "GaussianISAWPart"),
    "glisa": ("glisa", "GlobalLinearISAWPart"),
    "nlis": ("nlis", "NLISWPart"),
    "gmbis": ("gmbis", "GMBISWPart"),
    "h": ("hirshfeld", "HirshfeldWPart"),
    "hi": ("hirshfeld_i", "HirshfeldIWPart"),
    "b": ("becke", "BeckeWPart"),
}


def optional_package(name: str):
    """The optional third-party solver package `name` (``cvxopt``, ``qpsolvers``) or None.

    The test infrastructure under oracle/ ships stand-ins with the same import names so that the
    reference can run in the build container; they mark themselves with ``__oracle_shim__`` and are
    never used by the product, whatever ``sys.path`` looks like."""
    import importlib

    try:
        module = importlib.import_module(name)
    except ImportError:
        return None
    return None if getattr(module, "__oracle_shim__", False) else module


def wpart_schemes(scheme: str):
    """Return the partitioning class registered under the reference's short name."""
    import importlib

    try:
        module, name = _SCHEMES[scheme]
    except KeyError:
        raise NotImplementedError(f"unknown scheme {scheme!r}: one of {sorted(_SCHEMES)}") from None
    try:
        return getattr(importlib.import_module(f"{__package__}.{module}"), name)
    except ModuleNotFoundError:
        raise NotImplementedError(f"scheme {scheme!r} is not built yet") from None


def typecheck_geo(coordinates=None, numbers=None, pseudo_numbers=None, need_coordinates=True,
                  need_numbers=True, need_pseudo_numbers=True):  # fmt: skip
    """Validate (coordinates, numbers, pseudo_numbers); returns [natom, *checked arrays].

    Same rules as the reference: coordinates float (natom,3), numbers int64 (natom,),
    pseudo_numbers (natom,) converted to float, defaulting to ``numbers.astype(float)``.
    """
    first = next((a for a in (coordinates, numbers, pseudo_numbers) if a is not None), None)
    if first is None:
        raise TypeError("At least one argument is required and should not be None")
    natom = len(first)

    if coordinates is None:
        if need_coordinates:
            raise TypeError("Coordinates can not be None.")
    elif coordinates.shape != (natom, 3) or not issubclass(coordinates.dtype.type, float):
        raise TypeError("The argument centers must be a float array with shape (natom,3).")

    if numbers is None:
        if need_numbers:
            raise TypeError("Numbers can not be None.")
    elif numbers.shape != (natom,) or not issubclass(numbers.dtype.type, np.int64):
        raise TypeError("The argument numbers must be a vector with length natom.")

    if pseudo_numbers is None:
        if need_pseudo_numbers:
            pseudo_numbers = numbers.astype(float)
    else:
        if pseudo_numbers.shape != (natom,):
            raise TypeError("The argument pseudo_numbers must be a vector with length natom.")
        if not issubclass(pseudo_numbers.dtype.type, float):
            pseudo_numbers = pseudo_numbers.astype(float)

    out = [natom]
    if need_coordinates:
        out.append(coordinates)
    if need_numbers:
        out.append(numbers)
    if need_pseudo_numbers:
        out.append(pseudo_numbers)
    return out


# -- helpers of the host plug-in solvers (1-D radial problems, K x nrad arrays) -------------------
def compute_quantities(density, pro_atom_params, basis_functions, density_cutoff, do_sick=True,
                       do_ratio=True, do_ln_ratio=True):  # fmt: skip
    """(pro_shells, pro_density, sick, ratio, ln_ratio) for coefficients ``pro_atom_params`` of
    the tabulated ``basis_functions`` (K, N); ratio = density / pro_density and its logarithm are
    0 where either density is below ``density_cutoff`` (utils.py:198-252)."""
    c = np.asarray(pro_atom_params).flatten()
    pro_shells = basis_functions * c[:, None]
    pro_density = np.einsum("ij->j", pro_shells)
    sick = ratio = ln_ratio = None
    if do_sick:
        sick = (density < density_cutoff) | (pro_density < density_cutoff)
    with np.errstate(all="ignore"):
        if do_ratio:
            ratio = np.divide(density, pro_density, out=np.zeros_like(density), where=~sick)
        if do_ln_ratio:
            ln_ratio = np.log(ratio, out=np.zeros_like(density), where=~sick)
    return pro_shells, pro_density, sick, ratio, ln_ratio


def _report(message, as_warn, logger, error, level="warning"):
    if not as_warn:
        raise RuntimeError(error)
    if logger is not None:
        getattr(logger, level)(message)
    else:
        warnings.warn(message)


def check_pars_population(propars, ref_pop, as_warn=True, logger=None, population_cutoff=POPULATION_CUTOFF):
    diff = np.sum(propars) - ref_pop
    if np.abs(diff) > population_cutoff:
        tag = "WARNING" if as_warn else "ERROR"
        text = (f"{tag}: The sum of pro-atom parameters is not equal to reference population.\n"
                f"{tag}: The difference is {diff:.5E}")  # fmt: skip
        _report(text, as_warn, logger, text)


def check_pars_negativity(pars, as_warn=True, logger=None, negative_cutoff=NEGATIVE_CUTOFF):
    assert np.ndim(pars) == 1
    if (pars < negative_cutoff).any():
        _report("WARNING: Not all pro-atom parameters are positive!", as_warn, logger,
                "Negative pro-atom parameters found!")  # fmt: skip


def check_dens_negativity(dens, as_warn=False, logger=None, negative_cutoff=NEGATIVE_CUTOFF):
    assert np.ndim(dens) in [1, 3]
    if (dens < negative_cutoff).any():
        _report("WARNING: Not all pro-atom density are positive!", as_warn, logger,
                "Negative pro-atom density found!")  # fmt: skip


def check_dens_monotonicity(dens, as_warn=True, logger=None, negative_cutoff=NEGATIVE_CUTOFF):
    assert np.ndim(dens) == 1
    if (dens[:-1] - dens[1:] < negative_cutoff).any():
        text = "WARNING: Pro-atom density should be monotonically decreasing."
        _report(text, as_warn, logger, "Pro-atom density should be monotonically decreasing.", level="info")
        warnings.warn(text)


def check_pro_atom_parameters(pro_atom_params, basis_functions=None, total_population=None,
                              pro_atom_density=None, check_monotonicity=True, check_negativity=True,
                              check_propars_negativity=True, logger=None,
                              negative_cutoff=NEGATIVE_CUTOFF, population_cutoff=POPULATION_CUTOFF):  # fmt: skip
    """Validity of a set of pro-atom coefficients (utils.py:304-401): negative density raises,
    negative coefficients / population mismatch warn, non-monotonic density raises if requested."""
    if np.asarray(pro_atom_params).ndim != 1:
        raise ValueError("pro_atom_params must be a 1D array")
    if basis_functions is not None and np.asarray(basis_functions).ndim != 2:
        raise ValueError("basis_functions must be a 2D array")
    if pro_atom_density is not None and np.asarray(pro_atom_density).ndim != 1:
        raise ValueError("pro_atom_density must be a 1D array")
    if check_propars_negativity:
        check_pars_negativity(pro_atom_params, negative_cutoff=negative_cutoff)
    if basis_functions is not None and pro_atom_density is None:
        if pro_atom_params.size != basis_functions.shape[0]:
            raise ValueError("Length of pro_atom_params does not match the number of basis functions")
        pro_atom_density = (basis_functions * pro_atom_params[:, None]).sum(axis=0)
    if check_negativity and pro_atom_density is not None:
        check_dens_negativity(pro_atom_density, logger=logger, negative_cutoff=negative_cutoff)
    if total_population is not None:
        check_pars_population(pro_atom_params, total_population, logger=logger, population_cutoff=population_cutoff)
    if check_monotonicity and pro_atom_density is not None:
        check_dens_monotonicity(pro_atom_density, as_warn=False, logger=logger, negative_cutoff=negative_cutoff)


def check_pro_atom_parameters_non_neg_pars(pro_atom_params, basis_functions=None, total_population=None,
                                           pro_atom_density=None, logger=None,
                                           negative_cutoff=NEGATIVE_CUTOFF,
                                           population_cutoff=POPULATION_CUTOFF):  # fmt: skip
    return check_pro_atom_parameters(
        pro_atom_params, basis_functions, total_population, pro_atom_density, check_monotonicity=False,
        check_negativity=False, check_propars_negativity=False, logger=logger,
        negative_cutoff=negative_cutoff, population_cutoff=population_cutoff,
    )  # fmt: skip


def check_pro_atom_parameters_neg_pars(pro_atom_params, basis_functions=None, total_population=None,
                                       pro_atom_density=None, check_monotonicity=True,
                                       check_negativity=True, logger=None,
                                       negative_cutoff=NEGATIVE_CUTOFF,
                                       population_cutoff=POPULATION_CUTOFF):  # fmt: skip
    return check_pro_atom_parameters(
        pro_atom_params, basis_functions, total_population, pro_atom_density,
        check_monotonicity=check_monotonicity, check_negativity=check_negativity,
        check_propars_negativity=False, logger=logger, negative_cutoff=negative_cutoff,
        population_cutoff=population_cutoff,
    )  # fmt: skip


def fix_propars(exp_array, propars, delta):
    """Indices of the most diffuse functions whose coefficient is (numerically) zero while the
    Newton step wants to push it negative: walking up from the smallest exponent, stop at the
    first function that does not qualify (utils.py:555-577)."""
    frozen = []
    for k in np.argsort(exp_array):
        if np.abs(propars[k]) < 1e-4 and delta[k] < 0.0:
            frozen.append(k)
        else:
            break
    return frozen
