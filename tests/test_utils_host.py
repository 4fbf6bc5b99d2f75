"""Host utilities of the drop-in boundary (horton_part_b200/utils.py): argument checking, the
scheme registry, the 1-D helper of the plug-in solvers and the validity checks.  Where the
reference tree is present (the build container) the functions are compared with the reference's
own implementations on random inputs; the behavioural assertions run everywhere.  CPU only."""

import importlib
import logging
import pathlib
import sys
import warnings

import numpy as np
import pytest
from conftest import ROOT

from horton_part_b200 import utils

REF_SRC = pathlib.Path("/root/reference/src")


@pytest.fixture(scope="module")
def ref_utils():
    if not REF_SRC.is_dir():
        pytest.skip("reference tree not present on this machine")
    saved = {k: sys.modules.get(k) for k in ("grid", "cvxopt", "qpsolvers", "importlib_resources")}
    sys.path[:0] = [str(ROOT / "oracle" / "qcgrid_shim"), str(REF_SRC)]
    try:
        yield importlib.import_module("horton_part.utils")
    finally:
        del sys.path[:2]
        for k, v in saved.items():  # never leave the oracle's stand-ins importable by the product
            if v is None:
                sys.modules.pop(k, None)


def test_typecheck_geo_rules():
    xyz = np.zeros((3, 3))
    z = np.array([8, 1, 1])
    natom, c, n, p = utils.typecheck_geo(xyz, z, None)
    assert natom == 3 and c is xyz and n is z and p.dtype == float and (p == z).all()
    natom, p = utils.typecheck_geo(None, None, np.array([6, 1, 1]), need_coordinates=False, need_numbers=False)
    assert natom == 3 and p.dtype == float  # integer pseudo numbers are converted
    with pytest.raises(TypeError, match="At least one"):
        utils.typecheck_geo()
    with pytest.raises(TypeError, match="Coordinates"):
        utils.typecheck_geo(None, z, None)
    with pytest.raises(TypeError, match="Numbers"):
        utils.typecheck_geo(xyz, None, None)
    with pytest.raises(TypeError, match="float array"):
        utils.typecheck_geo(np.zeros((3, 3), dtype=int), z, None)
    with pytest.raises(TypeError, match="float array"):
        utils.typecheck_geo(np.zeros((3, 2)), z, None)
    with pytest.raises(TypeError, match="numbers"):
        utils.typecheck_geo(xyz, z.astype(np.int32), None)  # int64 only, as in the reference
    with pytest.raises(TypeError, match="numbers"):
        utils.typecheck_geo(xyz, z[:2], None)
    with pytest.raises(TypeError, match="pseudo_numbers"):
        utils.typecheck_geo(xyz, z, np.ones(2))


def test_scheme_registry_names():
    names = {"h": "HirshfeldWPart", "hi": "HirshfeldIWPart", "is": "ISAWPart", "mbis": "MBISWPart",
             "nlis": "NLISWPart", "gmbis": "GMBISWPart", "b": "BeckeWPart", "lisa": "LinearISAWPart",
             "glisa": "GlobalLinearISAWPart", "gisa": "GaussianISAWPart"}  # utils.py:62-104 of the reference
    for short, cls in names.items():
        assert utils.wpart_schemes(short).__name__ == cls
        assert utils.wpart_schemes(short).name == short
    with pytest.raises(NotImplementedError, match="unknown scheme"):
        utils.wpart_schemes("mulliken")


def test_constants_match_the_reference_yaml(ref_utils):
    assert utils.DENSITY_CUTOFF == ref_utils.DENSITY_CUTOFF == 1e-15
    assert utils.NEGATIVE_CUTOFF == ref_utils.NEGATIVE_CUTOFF == -1e-12
    assert utils.POPULATION_CUTOFF == ref_utils.POPULATION_CUTOFF == 1e-4
    assert utils.ANGSTROM == ref_utils.ANGSTROM


def test_compute_quantities_equals_the_reference(ref_utils):
    rng = np.random.default_rng(4)
    for trial in range(20):
        k, n = int(rng.integers(1, 9)), int(rng.integers(5, 200))
        bs = rng.uniform(0, 2, size=(k, n)) * np.exp(-rng.uniform(0, 40, size=(1, n)))
        c = rng.uniform(-0.2 if trial % 4 == 0 else 0.0, 3, size=k)
        rho = rng.uniform(0, 1, size=n) * np.exp(-rng.uniform(0, 50, size=n))  # includes values below the cut-off
        flags = dict(do_sick=True, do_ratio=trial % 3 != 0, do_ln_ratio=trial % 3 == 1)
        if flags["do_ln_ratio"]:
            flags["do_ratio"] = True
        mine = utils.compute_quantities(rho, c, bs, 1e-15, **flags)
        ref = ref_utils.compute_quantities(rho, c, bs, 1e-15, **flags)
        for a, b in zip(mine, ref):
            assert (a is None) == (b is None)
            if a is not None:
                np.testing.assert_array_equal(a, b)  # same NumPy expressions: bit for bit


def test_fix_propars_equals_the_reference(ref_utils):
    rng = np.random.default_rng(9)
    for _ in range(200):
        k = int(rng.integers(1, 10))
        exps = rng.permutation(10.0 ** rng.uniform(-1, 2, size=k))
        pars = np.where(rng.random(k) < 0.5, rng.uniform(0, 2e-4, size=k), rng.uniform(0, 2, size=k))
        delta = rng.normal(size=k)
        assert list(utils.fix_propars(exps, pars, delta)) == list(ref_utils.fix_propars(exps, pars, delta))
    assert utils.fix_propars(np.array([3.0, 1.0, 2.0]), np.array([1.0, 0.0, 0.0]), -np.ones(3)) == [1, 2]


def _outcome(fn, *args, **kwargs):
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        try:
            fn(*args, **kwargs)
            raised = None
        except Exception as exc:  # noqa: BLE001
            raised = (type(exc).__name__, str(exc))
    return raised, sorted(str(w.message) for w in caught)


def test_validity_checks_equal_the_reference(ref_utils):
    """Same exceptions, same messages, same warnings for good and bad coefficient sets."""
    rng = np.random.default_rng(1)
    r = np.linspace(0.01, 6, 40)
    bs = np.array([a**1.5 * np.exp(-a * r**2) for a in (0.3, 1.0, 4.0)])
    cases = [
        (np.array([1.0, 2.0, 3.0]), {}),
        (np.array([1.0, -0.5, 3.0]), {}),  # a negative coefficient
        (np.array([1.0, -5.0, 0.1]), {}),  # ... that makes the density negative / non-monotonic
        (np.array([1.0, 2.0, 3.0]), {"total_population": 6.0}),
        (np.array([1.0, 2.0, 3.0]), {"total_population": 6.1}),  # population off by more than 1e-4
        (np.array([0.0, 0.0, 1.0]), {"check_monotonicity": False}),
    ]
    for pars, extra in cases:
        for name in ("check_pro_atom_parameters", "check_pro_atom_parameters_neg_pars"):
            mine = _outcome(getattr(utils, name), pars, basis_functions=bs, **extra)
            ref = _outcome(getattr(ref_utils, name), pars, basis_functions=bs, **extra)
            assert mine == ref, (name, pars, extra)
        extra2 = {k: v for k, v in extra.items() if k != "check_monotonicity"}
        assert _outcome(utils.check_pro_atom_parameters_non_neg_pars, pars, basis_functions=bs, **extra2) == _outcome(
            ref_utils.check_pro_atom_parameters_non_neg_pars, pars, basis_functions=bs, **extra2)
    for bad in (dict(pro_atom_params=np.ones((2, 2))), dict(pro_atom_params=np.ones(3), basis_functions=np.ones(3)),
                dict(pro_atom_params=np.ones(2), basis_functions=bs)):  # fmt: skip
        assert _outcome(utils.check_pro_atom_parameters, **bad)[0] == _outcome(ref_utils.check_pro_atom_parameters, **bad)[0]
    dens = np.exp(-r) * (1 + 0.3 * np.sin(6 * r))
    for fn in ("check_dens_monotonicity", "check_dens_negativity", "check_pars_negativity"):
        for arr in (dens, -dens, np.sort(dens)[::-1].copy()):
            for as_warn in (True, False):
                assert _outcome(getattr(utils, fn), arr, as_warn=as_warn) == _outcome(getattr(ref_utils, fn), arr, as_warn=as_warn)
    for pop in (dens.sum(), dens.sum() + 1.0):
        for as_warn in (True, False):
            assert _outcome(utils.check_pars_population, dens, pop, as_warn=as_warn) == _outcome(
                ref_utils.check_pars_population, dens, pop, as_warn=as_warn)
    # with a logger the warnings go to the logger in both
    log = logging.getLogger("test_utils_host")
    assert _outcome(utils.check_pars_population, dens, 0.0, logger=log) == _outcome(ref_utils.check_pars_population, dens, 0.0, logger=log)
