"""ORACLE ONLY: ``from grid.becke import BeckeWeights`` (reference becke.py:24)."""
from . import BeckeWeights  # noqa: F401
