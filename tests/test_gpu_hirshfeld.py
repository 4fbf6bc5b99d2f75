"""GPU parity of the database-spline schemes (Hirshfeld, Hirshfeld-I) against the reference's own
outputs and golden charges (tests/test_wpart.py:70-87)."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_hirshfeld_h2o(h2o, h2o_proatomdb):
    from horton_part_b200 import HirshfeldWPart

    db, gold = h2o_proatomdb
    part = HirshfeldWPart(h2o["coords"], h2o["numbers"], h2o["pseudo"], h2o["grid"], h2o["rho"], db)
    part.do_charges()
    np.testing.assert_allclose(part["charges"], gold["h/charges"], rtol=1e-8, atol=1e-10)
    assert abs(part["charges"] - np.array([-0.246171541212, 0.123092011074, 0.123079530138])).max() < 2e-3
    np.testing.assert_allclose(part["promoldens"][::97], gold["h/promoldens_sample"], rtol=1e-9)
    np.testing.assert_allclose(part["at_weights_0"][::53], gold["h/at_weights_0_sample"], rtol=1e-9, atol=1e-13)
    part.do_moments()
    assert abs(part["charges"] - part["cartesian_multipoles"][:, 0]).max() < 1e-3


def test_hirshfeld_i_h2o(h2o, h2o_proatomdb):
    from horton_part_b200 import HirshfeldIWPart

    db, gold = h2o_proatomdb
    part = HirshfeldIWPart(h2o["coords"], h2o["numbers"], h2o["pseudo"], h2o["grid"], h2o["rho"], db)
    part.do_charges()
    assert part["niter"] == int(gold["hi/niter"]) == 15
    np.testing.assert_allclose(part["charges"], gold["hi/charges"], rtol=1e-8, atol=1e-10)
    assert abs(part["charges"] - np.array([-0.4214, 0.2107, 0.2107])).max() < 2e-3
    np.testing.assert_allclose(part["history_changes"], gold["hi/history_changes"], rtol=1e-6)
    np.testing.assert_allclose(part["history_charges"], gold["hi/history_charges"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(part["history_entropies"], gold["hi/history_entropies"], rtol=1e-8)
    np.testing.assert_allclose(part["promoldens"][::97], gold["hi/promoldens_sample"], rtol=1e-9)


def test_proatomdb_pseudo_number_mismatch(h2o, h2o_proatomdb):
    from horton_part_b200 import HirshfeldWPart

    db, _ = h2o_proatomdb
    with pytest.raises(ValueError, match="pseudo number"):
        HirshfeldWPart(h2o["coords"], h2o["numbers"], np.array([6.0, 1.0, 1.0]), h2o["grid"], h2o["rho"], db)
