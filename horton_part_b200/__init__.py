"""B200-native stockholder-iteration hot path behind horton-part's WPart class API.

Python owns device buffers through PyTorch and calls hand-written sm_100a CUDA kernels through the
C ABI in ``include/hp_b200.h`` (``libhp_b200.so``, built by ``__graft_entry__.build()``).
There is no CPU fallback: the classes raise if CUDA or the extension is missing.
"""

from .utils import wpart_schemes  # noqa: F401

__version__ = "0.1.0"

_LAZY = {
    "MBISWPart": "mbis",
    "ISAWPart": "isa",
    "LinearISAWPart": "alisa",
    "GaussianISAWPart": "gisa",
    "GlobalLinearISAWPart": "glisa",
    "NLISWPart": "nlis",
    "GMBISWPart": "gmbis",
    "HirshfeldWPart": "hirshfeld",
    "HirshfeldIWPart": "hirshfeld_i",
    "BeckeWPart": "becke",
}


def __getattr__(name):
    if name in _LAZY:
        import importlib

        return getattr(importlib.import_module(f"{__name__}.{_LAZY[name]}"), name)
    raise AttributeError(name)
