"""ORACLE ONLY: placeholder for the absent third-party ``qpsolvers`` (reference gisa.py:26,
algo/diis.py:27); raises on use."""


def solve_qp(*_args, **_kwargs):
    raise ImportError("qpsolvers is not installed in this image; solver unavailable in the oracle")
