"""ORACLE ONLY: the back-port the reference imports (utils.py:26) maps to the stdlib."""
from importlib.resources import files  # noqa: F401
