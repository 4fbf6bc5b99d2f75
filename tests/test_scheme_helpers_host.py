"""Host helpers of the MBIS / NLIS / GMBIS classes (shell counts and initial parameters for every
element): the known answers of the reference's tests (tests/test_mbis.py:26-41,
tests/test_nlis.py:53-70), and -- where the reference tree is present -- equality with the
reference's own functions for Z = 1 ... 104.  CPU only."""

import importlib
import pathlib
import sys

import numpy as np
import pytest
from conftest import ROOT

from horton_part_b200 import gmbis, mbis, nlis

REF_SRC = pathlib.Path("/root/reference/src")


@pytest.fixture(scope="module")
def ref():
    if not REF_SRC.is_dir():
        pytest.skip("reference tree not present on this machine")
    saved = {k: sys.modules.get(k) for k in ("grid", "cvxopt", "qpsolvers", "importlib_resources")}
    sys.path[:0] = [str(ROOT / "oracle" / "qcgrid_shim"), str(REF_SRC)]
    try:
        yield {name: importlib.import_module(f"horton_part.{name}") for name in ("mbis", "nlis", "gmbis")}
    finally:
        del sys.path[:2]
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)


def test_mbis_known_answers():
    assert [mbis.get_nshell(z) for z in (1, 2, 3, 17, 18, 21, 44, 72, 104)] == [1, 1, 2, 3, 3, 4, 5, 6, 7]
    assert (mbis.get_initial_mbis_propars(1) == [1.0, 2.0]).all()
    assert (mbis.get_initial_mbis_propars(2) == [2.0, 4.0]).all()
    assert (mbis.get_initial_mbis_propars(3) == [2.0, 6.0, 1.0, 2.0]).all()


def test_nlis_known_answers():
    nshell = 4
    nshell_dict = {z: nshell for z in range(1, 105)}
    exp_n_dict = {(1, k): 2.0 for k in range(nshell)}
    for z in (1, 21, 44, 72, 104):
        assert nlis.get_nlis_nshell(z, nshell_dict) == nshell
    values = nlis.get_initial_nlis_propars(1, exp_n_dict, nshell_dict)
    expected = []
    for width in (2.0, 1.25992105, 0.79370053, 0.5):  # tests/test_nlis.py:53-70
        expected += [1 / nshell, width, 2.0]
    assert values == pytest.approx(expected)


def test_every_element_equals_the_reference(ref):
    for z in range(1, 105):
        assert mbis.get_nshell(z) == ref["mbis"].get_nshell(z)
        np.testing.assert_array_equal(mbis.get_initial_mbis_propars(z), ref["mbis"].get_initial_mbis_propars(z))
        for nshell_dict in ({}, {z: 3}, {z: 5}):
            assert nlis.get_nlis_nshell(z, nshell_dict) == ref["nlis"].get_nlis_nshell(z, nshell_dict)
            k = nlis.get_nlis_nshell(z, nshell_dict)
            for exp_n_dict in ({}, {(z, i): 1.0 + 0.25 * i for i in range(k)}):
                mine = nlis.get_initial_nlis_propars(z, dict(exp_n_dict), nshell_dict)
                theirs = ref["nlis"].get_initial_nlis_propars(z, dict(exp_n_dict), nshell_dict)
                np.testing.assert_array_equal(np.asarray(mine), np.asarray(theirs))
        for exp_n_dict in ({}, {(z, i): 1.5 for i in range(mbis.get_nshell(z))}):
            np.testing.assert_array_equal(np.asarray(gmbis.get_initial_gmbis_propars(z, dict(exp_n_dict))),
                                          np.asarray(ref["gmbis"].get_initial_gmbis_propars(z, dict(exp_n_dict))))
