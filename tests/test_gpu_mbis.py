"""GPU parity: MBIS through the WPart API (C-ABI kernels) vs the reference's own outputs
(tests/golden) and the pinned oracle.  Tolerances: north_star's 1e-8 relative on charges and
parameters, identical iteration counts."""

import numpy as np
import pytest
import stockholder_oracle as oracle

pytestmark = pytest.mark.gpu

RTOL = 1e-8


def _mbis(case, **kw):
    from horton_part_b200 import MBISWPart

    part = MBISWPart(case["coords"], case["numbers"], case["pseudo"], case["grid"], case["rho"], **kw)
    part.do_partitioning()
    return part


def _compare(part, ref):
    assert part["niter"] == int(ref["niter"])
    np.testing.assert_allclose(part["charges"], ref["charges"], rtol=RTOL, atol=1e-10)
    np.testing.assert_allclose(part["propars"], ref["propars"], rtol=RTOL)
    np.testing.assert_allclose(part["history_changes"], ref["history_changes"], rtol=1e-6)
    np.testing.assert_allclose(part["history_entropies"], ref["history_entropies"], rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(part["history_charges"], ref["history_charges"], rtol=RTOL, atol=1e-10)


def _gold(gold, tag):
    return {k.split("/", 1)[1]: gold[k] for k in gold.files if k.startswith(tag + "/")}


def test_h2o_against_reference_run(h2o):
    part = _mbis(h2o)
    ref = _gold(h2o["gold"], "mbis")
    _compare(part, ref)
    assert part["niter"] == 27
    # reference's own golden charges, tests/test_wpart.py:95-102
    assert abs(part["charges"] - np.array([-0.61891067, 0.3095756, 0.30932584])).max() < 2e-3
    np.testing.assert_allclose(part["promoldens"][::97], ref["promoldens_sample"], rtol=1e-10)
    for a in range(3):
        np.testing.assert_allclose(part[f"at_weights_{a}"][::53], ref[f"at_weights_{a}_sample"], rtol=1e-10, atol=1e-300)
    # the reference evaluates a CubicSpline at its own knots: exact except round-off at the last knot
    np.testing.assert_allclose(part["spherical_average_0"], ref["spherical_average_0"], rtol=1e-10, atol=1e-25)
    for key in ("core_charges", "valence_charges", "valence_widths"):
        np.testing.assert_allclose(part[key], ref[key], rtol=RTOL)
    # tests/test_wpart.py:97-101
    assert part["charges"] == pytest.approx(part["valence_charges"] + part["core_charges"])
    assert (part["core_charges"] > 0).all() and (part["valence_charges"] < 0).all()
    assert (part["valence_widths"] > 0).all()


def test_water6_against_reference_run(water6):
    part = _mbis(water6)
    _compare(part, _gold(water6["gold"], "mbis"))
    assert abs(part["charges"] - np.tile([-0.6, 0.3, 0.3], 2)).max() < 5e-5


def test_water12_against_oracle(make_water):
    case = make_water(12, nrad=30, nang=38, seed=3)
    part = _mbis(case)
    ref = oracle.mbis(case["coords"], case["numbers"], case["pseudo"], case["grid"], case["rho"])
    _compare(part, ref)
    np.testing.assert_allclose(part["promoldens"], ref["promoldens"], rtol=1e-11)
    for a in range(12):
        np.testing.assert_allclose(part[f"at_weights_{a}"], ref["at_weights"][a], rtol=1e-11, atol=1e-300)


def test_do_charges_and_once_semantics(h2o):
    part = _mbis(h2o)
    part.do_charges()
    ref = _gold(h2o["gold"], "mbis")
    np.testing.assert_allclose(part["charges"], ref["charges"], rtol=RTOL, atol=1e-10)
    n = part["niter"]
    part.do_partitioning()  # just_once: no second run
    assert part["niter"] == n and len(part.history_changes) == n
    part.clear()
    assert "charges" not in part.cache
    keys = set(_mbis(h2o).do_all())
    for k in ("charges", "populations", "pseudo_populations", "propars", "niter", "change",
              "history_propars", "history_charges", "history_entropies", "history_changes",
              "core_charges", "valence_charges", "valence_widths", "radial_points_0",
              "spherical_average_0", "radial_weights_0", "time_update_at_weights", "time_update_propars"):
        assert k in keys, k


def test_maxiter_and_threshold(water6):
    part = _mbis(water6, maxiter=5)
    assert part["niter"] == 5
    part = _mbis(water6, threshold=1e-3)
    ref = oracle.mbis(water6["coords"], water6["numbers"], water6["pseudo"], water6["grid"], water6["rho"], threshold=1e-3)
    assert part["niter"] == ref["niter"]
    np.testing.assert_allclose(part["charges"], ref["charges"], rtol=RTOL, atol=1e-10)


def test_multipole_moments_against_reference_run(h2o):
    """do_moments (row a13) vs the reference's own do_moments run through the qc-grid shim."""
    part = _mbis(h2o)
    part.do_moments()
    ref = _gold(h2o["gold"], "mbis")
    for key in ("cartesian_multipoles", "pure_multipoles", "radial_moments"):
        scale = np.abs(ref[key]).max()
        np.testing.assert_allclose(part[key], ref[key], rtol=1e-8, atol=1e-10 * scale, err_msg=key)
    # tests/test_wpart.py:62-63
    part.do_charges()
    assert abs(part["charges"] - part["cartesian_multipoles"][:, 0]).max() < 1e-3
    assert abs(part["charges"] - part["pure_multipoles"][:, 0]).max() < 1e-3
