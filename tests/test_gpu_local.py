"""Cut-off (local grid) mode: the fused kernel with `local_radius` must reproduce the reference's
local-grid semantics built from the bit-exact index (row L), and converge to the dense result."""

import numpy as np
import pytest
import stockholder_oracle as oracle

pytestmark = pytest.mark.gpu


def _mbis(case, **kw):
    from horton_part_b200 import MBISWPart

    part = MBISWPart(case["coords"], case["numbers"], case["pseudo"], case["grid"], case["rho"], **kw)
    part.do_partitioning()
    return part


def test_local_radius_matches_local_grid_oracle(make_water):
    case = make_water(12, nrad=30, nang=38, seed=3)
    for radius in (3.0, 6.5):
        part = _mbis(case, local_radius=radius, maxiter=6)
        ref = oracle.mbis(case["coords"], case["numbers"], case["pseudo"], case["grid"], case["rho"],
                          maxiter=6, local_radius=radius)
        assert part["niter"] == ref["niter"] == 6
        # same points included (bit-exact rule) => same promolecule to round-off, zeros where no atom reaches
        assert np.array_equal(part["promoldens"] == 0.0, ref["promoldens"] == 0.0)
        np.testing.assert_allclose(part["promoldens"], ref["promoldens"], rtol=1e-11)
        for a in range(12):
            assert np.array_equal(part[f"at_weights_{a}"] == 0.0, ref["at_weights"][a] == 0.0)
            np.testing.assert_allclose(part[f"at_weights_{a}"], ref["at_weights"][a], rtol=1e-11, atol=1e-300)
        np.testing.assert_allclose(part["charges"], ref["charges"], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(part["history_changes"], ref["history_changes"], rtol=1e-7)


def test_huge_radius_equals_dense_and_pairs_are_counted(make_water):
    case = make_water(12, nrad=30, nang=38, seed=3)
    dense = _mbis(case, maxiter=5)
    big = _mbis(case, maxiter=5, local_radius=1e9)
    np.testing.assert_allclose(big["charges"], dense["charges"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(big["promoldens"], dense["promoldens"], rtol=1e-13)
    assert big._table.pairs_evaluated() == 12 * case["grid"].size
    small = _mbis(case, maxiter=5, local_radius=4.0)
    assert 0 < small._table.pairs_evaluated() < 12 * case["grid"].size


def test_charges_converge_to_dense_with_radius(water6):
    dense = _mbis(water6)
    errs = []
    for radius in (8.0, 12.0, 16.0):
        part = _mbis(water6, local_radius=radius)
        errs.append(np.abs(part["charges"] - dense["charges"]).max())
    assert errs[0] > errs[2] and errs[2] < 1e-9, errs
