"""GPU parity on the BASELINE.json configurations themselves (not surrogates): the goldens are runs of
the UNMODIFIED reference on the real 150 x 194 grid (oracle/gen_golden.py: case_config2/3/4).

  config 2  20-atom organic-like chain, 582,000 points: MBIS, Hirshfeld, ISA at full size; Hirshfeld-I on
            the same chain with N -> O on a database promolecule (the reference's database ends at charge -1)
  config 3  aLISA `sc`, gauss and slater basis: 24-atom water cluster, 698,400 points, to convergence
  config 4  gLISA `newton` (exact Hessian) and `sc`: 12-atom peptide-like chain, 349,200 points

Bar: identical `niter`, charges 1e-8 (relative, with an absolute floor of 1e-9 for charges near 0).
"""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NRAD, NANG = 150, 194


def _gold(gold, tag):
    return {k.split("/", 1)[1]: gold[k] for k in gold.files if k.startswith(tag + "/")}


def _grid(coords, numbers):
    from horton_part_b200 import gridlite

    rgrid = gridlite.BeckeRTransform(1e-4, 1.5).transform_1d_grid(gridlite.GaussChebyshev(NRAD))
    # Becke weights on the device (hp_becke_weights, pinned against the reference's in test_gpu_becke)
    return gridlite.MolGrid.from_size(numbers, coords, NANG, rgrid, gridlite.DeviceBeckeWeights(), store=True)


def _records(z, prefix="record/"):
    from horton_part_b200 import gridlite
    from horton_part_b200.core.proatomdb import ProAtomDB, ProAtomRecord

    records = []
    for key in z.files:
        if not key.startswith(prefix):
            continue
        v = z[key]
        number, charge, energy, rmin, rmax, npoint = int(v[0]), int(v[1]), float(v[2]), v[3], v[4], int(v[5])
        rgrid = gridlite.PowerRTransform(rmin, rmax, npoint - 1).transform_1d_grid(gridlite.UniformInteger(npoint))
        records.append(ProAtomRecord(number, charge, energy, rgrid, v[6 : 6 + npoint].copy(), v[6 + npoint :].copy()))
    return ProAtomDB(records)


@pytest.fixture(scope="module")
def config2():
    from conftest import GOLDEN
    from horton_part_b200 import synthetic

    if not (GOLDEN / "config2_organic20.npz").exists():
        pytest.skip("tests/golden/config2_organic20.npz not generated yet (oracle/gen_golden.py config2)")
    z = np.load(GOLDEN / "config2_organic20.npz")
    coords, numbers = z["coordinates"], z["numbers"]
    c2, n2 = synthetic.organic_like(20, 0)
    assert np.array_equal(n2, numbers) and np.abs(c2 - coords).max() == 0.0  # the generator is the fixture
    grid = _grid(coords, numbers)
    rho = synthetic.slater_promolecule_host(grid.points, coords, numbers)
    np.testing.assert_allclose(rho[::997], z["dens_sample"], rtol=1e-13)
    np.testing.assert_allclose(grid.aim_weights[::997], z["aim_weights_sample"], rtol=1e-9, atol=1e-14)
    return dict(coords=coords, numbers=numbers, pseudo=numbers.astype(float), grid=grid, rho=rho, gold=z)


def _check_common(part, ref, grid, wtol=1e-8, patol=1e-300, wfloor=0.0, watol=1e-13):
    np.testing.assert_allclose(part["charges"], ref["charges"], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(part["promoldens"][::997], ref["promoldens_sample"], rtol=1e-8, atol=patol)
    for a in range(3):
        w = part[f"at_weights_{a}"]
        # where the promolecule is above `wfloor`: below it spline pro-atoms are rounding noise around zero
        # (cubic pieces through 1e-100) and their ratio is arbitrary -- in the reference too
        lo, hi = int(grid.indices[a]), int(grid.indices[a + 1])
        solid = part["promoldens"][lo:hi][::53] > wfloor
        np.testing.assert_allclose(w[::53][solid], ref[f"at_weights_{a}_sample"][solid], rtol=wtol, atol=watol)


def test_config2_mbis_full_size(config2):
    from horton_part_b200 import MBISWPart

    c = config2
    part = MBISWPart(c["coords"], c["numbers"], c["pseudo"], c["grid"], c["rho"])
    part.do_partitioning()
    ref = _gold(c["gold"], "mbis")
    assert part["niter"] == int(ref["niter"]) == 139
    _check_common(part, ref, c["grid"])
    np.testing.assert_allclose(part["propars"], ref["propars"], rtol=1e-8)
    np.testing.assert_allclose(part["history_changes"], ref["history_changes"], rtol=1e-6)
    np.testing.assert_allclose(part["history_entropies"], ref["history_entropies"], rtol=1e-8, atol=1e-11)
    # exact Slater promolecule: MBIS recovers the generating populations (known answer)
    from horton_part_b200.synthetic import SLATER_SHELLS

    q_true = np.array([z - sum(n for n, _ in SLATER_SHELLS[int(z)]) for z in c["numbers"]], float)
    assert np.abs(part["charges"] - q_true).max() < 2e-4


def test_config2_hirshfeld_full_size(config2):
    from horton_part_b200 import HirshfeldWPart

    c = config2
    part = HirshfeldWPart(c["coords"], c["numbers"], c["pseudo"], c["grid"], c["rho"], _records(c["gold"]))
    part.do_charges()
    _check_common(part, _gold(c["gold"], "h"), c["grid"])


def test_config2_isa_full_size(config2):
    from horton_part_b200 import ISAWPart

    c = config2
    ref = _gold(c["gold"], "isa")
    part = ISAWPart(c["coords"], c["numbers"], c["pseudo"], c["grid"], c["rho"])
    part.do_partitioning()
    # the reference itself stops at maxiter = 500 here without reaching the 1e-6 threshold
    assert part["niter"] == int(ref["niter"])
    # spline pro-atoms are numerically zero (+- 1e-14: cubic pieces through values of 1e-100) far from the
    # nuclei, where the exponential schemes still resolve 1e-300: absolute floor for the promolecule there.
    # Measured (profiles/r2_parity_report.txt): charges 1.2e-10, parameters above 1e-6 6.4e-8 relative after
    # 500 unconverged iterations.
    # After 500 iterations ISA is still moving here (the reference stops at maxiter too): the tails of the
    # tabulated pro-atoms (values below 1e-6) agree to ~1e-2 relative only, and so do the weights they
    # dominate (measured: 3.6e-6 absolute); everything that carries density agrees as above.
    _check_common(part, ref, c["grid"], patol=1e-12, wfloor=1e-9, watol=2e-5)
    np.testing.assert_allclose(part["history_changes"], ref["history_changes"], rtol=1e-5)
    np.testing.assert_allclose(part["history_entropies"], ref["history_entropies"], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(part["propars"], ref["propars"], rtol=1e-6, atol=1e-11)


def test_config2_hirshfeld_i_full_size():
    """Hirshfeld-I at config-2 size (oracle/gen_golden.py case_config2_hi: the chain with N -> O, density =
    promolecule of the database pro-atoms at fixed charges, because the reference's cached database ends at
    charge -1 and the diffuse Slater promolecule drives oxygen beyond it)."""
    from conftest import GOLDEN
    from horton_part_b200 import HirshfeldIWPart

    if not (GOLDEN / "config2_hi.npz").exists():
        pytest.skip("tests/golden/config2_hi.npz not generated yet (oracle/gen_golden.py config2_hi)")
    z = np.load(GOLDEN / "config2_hi.npz")
    ref = _gold(z, "hi")
    coords, numbers = z["coordinates"], z["numbers"]
    grid = _grid(coords, numbers)
    # the density: promolecule of the database pro-atoms at the generating charges (gen_golden.database_promolecule)
    db = _records(z)
    rho = np.zeros(grid.size)
    for R, zn, q in zip(coords, numbers, z["generating_charges"]):
        ic = int(np.floor(q))
        x = float(q - ic)
        one = (int(zn) - ic) == 1 or x == 0.0  # hirshfeld_i.py:125-132
        spline = db.get_spline(int(zn), {ic: 1 - x} if one else {ic: 1 - x, ic + 1: x})
        r = np.linalg.norm(grid.points - R, axis=1)
        rho += np.where(r <= db.get_rgrid(int(zn)).points[-1], np.clip(spline(np.minimum(r, 1e3)), 0.0, None), 0.0)
    np.testing.assert_allclose(rho[::997], z["dens_sample"], rtol=1e-12, atol=1e-300)
    part = HirshfeldIWPart(coords, numbers, numbers.astype(float), grid, rho, _records(z))
    part.do_charges()
    assert part["niter"] == int(ref["niter"])
    _check_common(part, ref, grid)
    np.testing.assert_allclose(part["history_changes"], ref["history_changes"], rtol=1e-6)
    np.testing.assert_allclose(part["history_entropies"], ref["history_entropies"], rtol=1e-8, atol=1e-11)


@pytest.fixture(scope="module")
def config3():
    from conftest import GOLDEN
    from horton_part_b200 import synthetic
    from horton_part_b200.core.basis import ExpBasisFuncHelper

    z = np.load(GOLDEN / "config3_water24.npz")
    coords, numbers = z["coordinates"], z["numbers"]
    grid = _grid(coords, numbers)
    helper = ExpBasisFuncHelper.from_function_type("gauss")
    rho = synthetic.expbasis_promolecule_host(grid.points, coords, numbers, helper, scale={8: 8.6, 1: 0.7})
    np.testing.assert_allclose(rho[::997], z["dens_sample"], rtol=1e-12)
    return dict(coords=coords, numbers=numbers, pseudo=numbers.astype(float), grid=grid, rho=rho, gold=z)


@pytest.mark.parametrize("basis,tag", [("gauss", "lisa_sc_gauss"), ("slater", "lisa_sc_slater")])
def test_config3_alisa_sc_real_grid(config3, basis, tag):
    from horton_part_b200 import LinearISAWPart

    c = config3
    ref = _gold(c["gold"], tag)
    part = LinearISAWPart(c["coords"], c["numbers"], c["pseudo"], c["grid"], c["rho"], solver="sc", basis_func=basis)
    part.do_partitioning()
    assert part["niter"] == int(ref["niter"])
    np.testing.assert_allclose(part["charges"], ref["charges"], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(part["history_changes"], ref["history_changes"], rtol=1e-5)
    np.testing.assert_allclose(part["history_entropies"], ref["history_entropies"], rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(part["promoldens"][::97], ref["promoldens_sample"], rtol=1e-8, atol=1e-300)
    # measured (profiles/r2_parity_report.txt): coefficients above 1e-6 agree to 3e-15 (gauss) / 8e-14 (slater)
    # relative although the slater table is near-degenerate (cond 3.8e6 of the weighted basis matrix on the
    # radial grid, tools/sensitivity.py): the fixed point itself is well conditioned
    np.testing.assert_allclose(part["propars"], ref["propars"], rtol=1e-9, atol=1e-12)


@pytest.fixture(scope="module")
def config4():
    from conftest import GOLDEN
    from horton_part_b200 import synthetic
    from horton_part_b200.core.basis import ExpBasisFuncHelper

    z = np.load(GOLDEN / "config4_peptide12.npz")
    coords, numbers = z["coordinates"], z["numbers"]
    grid = _grid(coords, numbers)
    helper = ExpBasisFuncHelper.from_function_type("gauss")
    rho = synthetic.expbasis_promolecule_host(grid.points, coords, numbers, helper,
                                              scale={1: 0.75, 6: 6.2, 7: 7.3, 8: 8.4})
    np.testing.assert_allclose(rho[::997], z["dens_sample"], rtol=1e-12)
    return dict(coords=coords, numbers=numbers, pseudo=numbers.astype(float), grid=grid, rho=rho, gold=z)


@pytest.mark.parametrize("solver,tag", [("newton", "glisa_newton"), ("sc", "glisa_sc")])
def test_config4_glisa_real_grid(config4, solver, tag):
    from horton_part_b200 import GlobalLinearISAWPart

    c = config4
    ref = _gold(c["gold"], tag)
    kw = dict(maxiter=60) if solver == "sc" else {}
    part = GlobalLinearISAWPart(c["coords"], c["numbers"], c["pseudo"], c["grid"], c["rho"], solver=solver, **kw)
    part.do_partitioning()
    assert part["niter"] == int(ref["niter"])
    np.testing.assert_allclose(part["charges"], ref["charges"], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(part["propars"], ref["propars"], rtol=1e-9, atol=1e-12)  # measured 9e-13
    np.testing.assert_allclose(part["history_entropies"], ref["history_entropies"], rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(part["promoldens"][::97], ref["promoldens_sample"], rtol=1e-8, atol=1e-300)
