"""Host basis-function helper (horton_part_b200/core/basis.py) against the known answers of the
reference's own tests/core/test_basis.py:100-188 and its data tables (the exponent tables shipped in
horton_part_b200/data/expbasis_tables.json were imported from data/gauss.json / slater.json by
tools/import_basis_tables.py).  CPU only."""

import json

import numpy as np
import pytest

from horton_part_b200.core.basis import ExpBasisFuncHelper, NumericBasisFuncHelper, evaluate_function, load_params, shell_norm


def test_evaluate_function_scalar_and_vector_known_answers():
    # tests/core/test_basis.py:112-170: a unit Gaussian at r = 1 and its radial derivative
    value, deriv = np.pi ** (-1.5) * np.exp(-1), -2 * np.pi ** (-1.5) * np.exp(-1)
    r = np.array([1.0])
    f = evaluate_function(2, 1.0, 1.0, r)
    assert isinstance(f, np.ndarray) and f.shape == r.shape and np.isclose(f, value)
    assert np.isclose(evaluate_function(2, 1.0, 1.0, r, 1)[-1], deriv)
    ones = np.ones(10)
    many = evaluate_function(2 * ones, ones, ones, np.ones(100), 0)
    assert many.shape == (10, 100) and many == pytest.approx(np.full((10, 100), value))
    d = evaluate_function(2 * ones, ones, ones, np.ones(100), 1)[-1]
    assert d.shape == (10, 100) and d == pytest.approx(np.full((10, 100), deriv))
    summed = evaluate_function(2 * ones, ones, ones, np.ones(7), 0, axis=0)
    assert summed.shape == (7,) and summed == pytest.approx(10 * value)
    with pytest.raises(NotImplementedError):  # test_get_pro_a_k_raises_notimplementederror
        evaluate_function(2, 1.0, 2.0, np.array([3.0]), 2)
    with pytest.raises(ValueError):
        evaluate_function(0, 1.0, 2.0, np.array([3.0]))
    with pytest.raises(ValueError):
        evaluate_function(2, 1.0, -2.0, np.array([3.0]))
    with pytest.raises(ValueError):
        evaluate_function(2, 1.0, 2.0, [3.0])


def test_shells_are_normalised_to_their_population():
    r = np.linspace(0, 40, 400001)
    for n, alpha in ((1.0, 1.3), (2.0, 0.4), (1.5, 0.9)):
        f = evaluate_function(n, 2.5, alpha, r)
        assert abs(np.trapezoid(4 * np.pi * r**2 * f, r) - 2.5) < 1e-6
        assert np.isclose(f[0], 2.5 * shell_norm(n, alpha))


def test_tables_match_the_reference_data():
    slater = ExpBasisFuncHelper.from_function_type("slater")
    gauss = ExpBasisFuncHelper.from_function_type("gauss")
    # tests/core/test_basis.py:173-181
    assert slater.get_exponent(1) == pytest.approx([6.6, 4.62, 3.23, 2.26, 1.58, 1.1, 0.77, 1.0])
    assert slater.get_exponent(6) == pytest.approx(
        [22.8, 17.35, 13.2, 8.79, 10.83, 7.91, 5.78, 4.22, 4.51, 3.46, 2.65, 2.03, 1.56])
    # shell counts quoted in SURVEY section 8a (data/gauss.json, data/slater.json)
    assert [gauss.get_nshell(z) for z in (1, 6, 7, 8, 9, 14, 16, 17, 35)] == [4, 6, 6, 6, 6, 9, 9, 9, 12]
    assert [slater.get_nshell(z) for z in (1, 6, 7, 8, 9, 16, 17)] == [8, 13, 14, 14, 14, 21, 21]
    assert set(np.unique(np.concatenate([gauss.get_order(z) for z in (1, 8)]))) == {2}
    assert set(np.unique(np.concatenate([slater.get_order(z) for z in (1, 8)]))) == {1}
    for helper in (gauss, slater):
        for z in (1, 6, 8):
            assert len(helper.get_initial(z)) == helper.get_nshell(z)
            assert helper.get_initial(z, 0) == helper.get_initial(z)[0]


def test_load_params_json_yaml_and_missing_initials(tmp_path):
    sample = {"1": ([1, 2], [0.5, 1.5], [0.1, 0.9]), "2": ([2, 1], [1.0, 2.0], [0.2, 0.8])}
    for ext in ("json", "yaml"):
        path = tmp_path / f"basis.{ext}"
        path.write_text(json.dumps(sample))  # JSON is valid YAML, as in the reference's own test
        orders, exps, inits = load_params(path, extension=ext)
        assert set(orders) == {1, 2} and list(exps[2]) == [1.0, 2.0] and list(inits[1]) == [0.1, 0.9]
        helper = ExpBasisFuncHelper.from_file(path)
        assert helper.get_nshell(1) == 2 and helper.get_order(2, 1) == 1
    with pytest.raises(AssertionError):
        load_params(tmp_path / "basis.json", extension="toml")
    numeric = NumericBasisFuncHelper.from_file(tmp_path / "basis.json", nrad=30)
    r = np.array([0.3, 1.0, 2.5])
    exact = ExpBasisFuncHelper.from_file(tmp_path / "basis.json")
    for k in range(2):  # 30 radial points: the spline follows the shell to interpolation accuracy
        np.testing.assert_allclose(numeric.compute_proshell_dens(1, k, 1.0, r), exact.compute_proshell_dens(1, k, 1.0, r),
                                   rtol=5e-2, atol=1e-4)  # fmt: skip
    assert numeric.get_knots(1).size == 30 and numeric.ppoly_coefficients(2).shape == (2, 29, 4)


def test_proatom_density_is_the_sequential_shell_sum():
    helper = ExpBasisFuncHelper.from_function_type("gauss")
    r = np.geomspace(1e-3, 12.0, 50)
    pops = np.linspace(0.2, 1.4, helper.get_nshell(8))
    y, d = helper.compute_proatom_dens(8, pops, r, 1)
    ref_y = sum(helper.compute_proshell_dens(8, k, pops[k], r) for k in range(len(pops)))
    np.testing.assert_allclose(y, ref_y, rtol=1e-15)
    np.testing.assert_allclose(d, sum(helper.compute_proshell_dens(8, k, pops[k], r, 1)[1] for k in range(len(pops))), rtol=1e-14)
    with pytest.raises(NotImplementedError):
        helper.compute_proatom_dens(8, pops, r, 2)
