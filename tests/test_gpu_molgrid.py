"""grid_type 2/3: per-atom fixed points on the molecular grid (row a9) against the reference's own
runs (tests/golden: mbis_gt2, lisa_sc_gt2)."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _gold(gold, tag):
    return {k.split("/", 1)[1]: gold[k] for k in gold.files if k.startswith(tag + "/")}


def _compare(part, ref, ptol=1e-8):
    assert part["niter"] == int(ref["niter"])
    np.testing.assert_allclose(part["charges"], ref["charges"], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(part["propars"], ref["propars"], rtol=ptol, atol=1e-9)
    np.testing.assert_allclose(part["history_changes"], ref["history_changes"], rtol=1e-5)
    np.testing.assert_allclose(part["history_entropies"], ref["history_entropies"], rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(part["promoldens"][::97], ref["promoldens_sample"], rtol=1e-8)


def test_mbis_grid_type_2_h2o(h2o):
    from horton_part_b200 import MBISWPart

    part = MBISWPart(h2o["coords"], h2o["numbers"], h2o["pseudo"], h2o["grid"], h2o["rho"], grid_type=2)
    part.do_partitioning()
    _compare(part, _gold(h2o["gold"], "mbis_gt2"))
    assert part["niter"] == 27


def test_alisa_sc_grid_type_2_h2o(h2o):
    from horton_part_b200 import LinearISAWPart

    part = LinearISAWPart(h2o["coords"], h2o["numbers"], h2o["pseudo"], h2o["grid"], h2o["rho"], solver="sc",
                          grid_type=2)
    part.do_partitioning()
    _compare(part, _gold(h2o["gold"], "lisa_sc_gt2"), ptol=1e-5)
    assert part["niter"] == 22


def test_mbis_grid_type_3_equals_2(water6):
    """grid_type 3 (molecular grid only, no atomic grids needed) gives the same partitioning."""
    from horton_part_b200 import MBISWPart

    args = (water6["coords"], water6["numbers"], water6["pseudo"], water6["grid"], water6["rho"])
    p2 = MBISWPart(*args, grid_type=2, maxiter=8)
    p3 = MBISWPart(*args, grid_type=3, maxiter=8)
    p2.do_partitioning()
    p3.do_partitioning()
    np.testing.assert_allclose(p3["charges"], p2["charges"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(p3["history_changes"], p2["history_changes"], rtol=1e-12)


def test_nlis_grid_type_2_h2o(h2o):
    from horton_part_b200 import NLISWPart

    part = NLISWPart(h2o["coords"], h2o["numbers"], h2o["pseudo"], h2o["grid"], h2o["rho"], exp_n_dict={},
                     grid_type=2)
    part.do_partitioning()
    _compare(part, _gold(h2o["gold"], "nlis_gt2"))


def test_glisa_sc_grid_type_2_h2o(h2o):
    from horton_part_b200 import GlobalLinearISAWPart

    part = GlobalLinearISAWPart(h2o["coords"], h2o["numbers"], h2o["pseudo"], h2o["grid"], h2o["rho"],
                                solver="sc", grid_type=2)
    part.do_partitioning()
    _compare(part, _gold(h2o["gold"], "glisa_sc_gt2"), ptol=1e-5)
