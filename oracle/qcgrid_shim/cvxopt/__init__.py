"""ORACLE ONLY: placeholder for the absent third-party ``cvxopt`` (reference alisa.py:27,
glisa.py) so that the reference package imports; every entry point raises on use, so the
cvxopt-backed solvers stay 'parity unpinned' (SURVEY.md section 8c)."""


def _absent(*_args, **_kwargs):
    raise ImportError("cvxopt is not installed in this image; solver unavailable in the oracle")


class _Solvers:
    options = {}
    cp = staticmethod(_absent)
    qp = staticmethod(_absent)


solvers = _Solvers()
matrix = _absent
spmatrix = _absent
log = _absent
div = _absent
mul = _absent
