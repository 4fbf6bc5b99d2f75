"""basis_type="numeric": NumericBasisFuncHelper (core/basis.py:330-390 of the reference) against
values the reference itself produced (tests/golden/water6_numeric.npz, oracle/gen_golden.py::
case_numeric), and the identity the device path rests on: a pro-atom sum_k c_k S_k(r) equals ONE
piecewise cubic whose coefficients are the c-weighted sums of the shells' coefficients."""

import numpy as np
import pytest
from conftest import GOLDEN

from horton_part_b200.core.basis import NumericBasisFuncHelper

GOLD = np.load(GOLDEN / "water6_numeric.npz")


@pytest.mark.parametrize("func_type", ["gauss", "slater"])
def test_tabulated_shells_match_reference(func_type):
    helper = NumericBasisFuncHelper.from_function_type(func_type)
    r = GOLD["helper/r"]
    for z in (1, 6, 8):
        ref = GOLD[f"helper/{func_type}/{z}"]
        assert helper.get_nshell(z) == ref.shape[0]
        got = np.array([helper.compute_proshell_dens(z, k, 1.0, r) for k in range(ref.shape[0])])
        # same SciPy CubicSpline through the same samples; the radial grid agrees to 3e-14
        scale = np.abs(ref).max(axis=1, keepdims=True)
        assert np.abs(got - ref).max() <= 1e-11 * scale.max()
        np.testing.assert_allclose(got, ref, rtol=1e-8, atol=1e-12 * scale.max())
        pops = np.linspace(0.3, 1.1, ref.shape[0])
        np.testing.assert_allclose(helper.compute_proatom_dens(z, pops, r, 0), GOLD[f"helper/{func_type}/{z}/proatom"],
                                   rtol=1e-8, atol=1e-12 * scale.max())  # fmt: skip
        y, dy = helper.compute_proshell_dens(z, 0, 2.0, r, 1)
        np.testing.assert_array_equal(dy, 0.0)  # the reference's placeholder derivative
        np.testing.assert_allclose(y, 2.0 * got[0], rtol=1e-15)
    with pytest.raises(NotImplementedError):
        helper.compute_proshell_dens(8, 0, 1.0, r, 2)


def test_mixed_coefficients_are_the_proatom():
    """What _refresh_table uploads: einsum('k,ksc->sc', c, ppoly) evaluated like the kernel does
    (interval = clamp(searchsorted_right - 1), Horner in (r - x_i), extrapolating both sides)."""
    helper = NumericBasisFuncHelper.from_function_type("gauss")
    rng = np.random.default_rng(2)
    r = np.concatenate([[0.0, 1e-7], np.geomspace(1e-5, 80.0, 500)])
    for z in (1, 8):
        x = helper.get_knots(z)
        coef = helper.ppoly_coefficients(z)
        assert coef.shape == (helper.get_nshell(z), x.size - 1, 4) and np.all(np.diff(x) > 0)
        c = rng.uniform(0.0, 2.0, size=helper.get_nshell(z))
        mixed = np.einsum("k,ksc->sc", c, coef)
        seg = np.clip(np.searchsorted(x, r, side="right") - 1, 0, x.size - 2)
        d = r - x[seg]
        m = mixed[seg]
        got = ((m[:, 0] * d + m[:, 1]) * d + m[:, 2]) * d + m[:, 3]
        ref = helper.compute_proatom_dens(z, c, r, 0)
        np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-13 * np.abs(ref).max())


def test_setup_bs_helper_selects_the_numeric_helper():
    from horton_part_b200.alisa import setup_bs_helper

    class Part:
        _bs_helper = None
        basis_func = "gauss"
        basis_type = "numeric"

        class logger:
            info = staticmethod(lambda *a: None)

    assert isinstance(setup_bs_helper(Part()), NumericBasisFuncHelper)
    bad = Part()
    bad.basis_type = "tabulated"
    with pytest.raises(RuntimeError, match="analytic and numeric"):
        setup_bs_helper(bad)
    given = Part()
    given.basis_func = NumericBasisFuncHelper.from_function_type("slater", nrad=40)
    assert setup_bs_helper(given) is given.basis_func
