"""Constants, geometry type checks and the scheme registry.

Mirrors /root/reference/src/horton_part/utils.py:51-59 (cut-offs from data/constants.yaml),
:62-104 (``wpart_schemes``) and :107-186 (``typecheck_geo``).
"""

from __future__ import annotations

import numpy as np

__all__ = [
    "wpart_schemes",
    "typecheck_geo",
    "DENSITY_CUTOFF",
    "NEGATIVE_CUTOFF",
    "POPULATION_CUTOFF",
    "ANGSTROM",
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
}


def wpart_schemes(scheme: str):
    """Return the partitioning class registered under the reference's short name."""
    import importlib

    try:
        module, name = _SCHEMES[scheme]
    except KeyError:
        raise NotImplementedError(f"scheme {scheme!r} is outside the accelerated hot path") from None
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
