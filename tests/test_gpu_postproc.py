"""GPU parity of the post-processing rows: Becke scheme, density decomposition,
pro-atom splines, dispersion coefficients -- against runs of the unmodified reference on water
HF/STO-3G (tests/golden/h2o_postproc.npz, oracle/gen_golden.py::case_postproc)."""

import contextlib
import io

import numpy as np
import pytest
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

GOLD = np.load(GOLDEN / "h2o_postproc.npz")


def _part(case, scheme, **kw):
    from horton_part_b200 import wpart_schemes

    return wpart_schemes(scheme)(case["coords"], case["numbers"], case["pseudo"], case["grid"], case["rho"], **kw)


def test_becke_scheme(h2o):
    part = _part(h2o, "b")
    part.do_charges()
    np.testing.assert_allclose(part["charges"], GOLD["becke/charges"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(part["populations"], GOLD["becke/populations"], rtol=1e-10)
    np.testing.assert_allclose(part["at_weights_0"][::53], GOLD["becke/at_weights_0_sample"], rtol=1e-10, atol=1e-14)
    with contextlib.redirect_stdout(io.StringIO()):
        part.do_moments()
    ref = GOLD["becke/cartesian_multipoles"]
    np.testing.assert_allclose(part["cartesian_multipoles"], ref, rtol=1e-8, atol=1e-9 * np.abs(ref).max())
    # the reference's own test of the scheme (tests/test_becke.py:47-57): clear() forgets the results
    part.clear()
    with pytest.raises(KeyError):
        part["charges"]
    part.do_charges()
    np.testing.assert_allclose(part["charges"], GOLD["becke/charges"], rtol=1e-8, atol=1e-10)
    assert abs(part["populations"].sum() - h2o["gold"]["nelec"]) < 1e-3  # Becke weights sum to one


def test_becke_radii_follow_the_paper():
    from horton_part_b200.becke import becke_radii
    from horton_part_b200.utils import ANGSTROM

    r = becke_radii(np.array([1, 6, 8, 2, 10]))
    np.testing.assert_allclose(r[:3], np.array([0.35, 0.70, 0.60]) * ANGSTROM)
    assert r[3] > 0 and r[4] > 0  # noble gases fall back to Cordero's covalent radii


def test_density_decomposition(h2o):
    part = _part(h2o, "mbis")
    with contextlib.redirect_stdout(io.StringIO()):
        part.do_density_decomposition()
    rgrid = h2o["grid"].atgrids[0].rgrid
    r_mid = np.sqrt(rgrid.points[:-1] * rgrid.points[1:])
    for a in range(3):
        dec = part.cache.load("density_decomposition", a)
        keys = sorted(dec)
        assert len(keys) == (h2o["grid"].atgrids[a].l_max // 2 + 1) ** 2 == 81
        knots = np.array([dec[k](rgrid.points) for k in keys])
        mid = np.array([dec[k](r_mid) for k in keys])
        ref_k, ref_m = GOLD[f"mbis/decomp_{a}_knots"], GOLD[f"mbis/decomp_{a}_mid"]
        scale = np.abs(ref_k).max()
        # components that vanish by symmetry are rounding noise relative to the monopole
        np.testing.assert_allclose(knots, ref_k, rtol=1e-8, atol=1e-12 * scale)
        np.testing.assert_allclose(mid, ref_m, rtol=1e-8, atol=1e-12 * scale)
        # the l=0 component is sqrt(4 pi) times the spherical average the iteration used
        np.testing.assert_allclose(knots[0], np.sqrt(4 * np.pi) * part[f"spherical_average_{a}"], rtol=1e-10,
                                   atol=1e-14 * scale)  # fmt: skip
    # second call is a no-op (just_once)
    before = part.cache.load("density_decomposition", 0)
    part.do_density_decomposition()
    assert part.cache.load("density_decomposition", 0) is before


def test_prosplines_mbis(h2o):
    part = _part(h2o, "mbis")
    part.do_partitioning()
    part.do_prosplines()
    rgrid = h2o["grid"].atgrids[0].rgrid
    r_mid = np.sqrt(rgrid.points[:-1] * rgrid.points[1:])
    for a in range(3):
        got = part.cache.load("spline_prodensity", a)(r_mid)
        np.testing.assert_allclose(got, GOLD[f"mbis/prospline_{a}_mid"], rtol=1e-7, atol=1e-14)
    # eval_proatom (API hook for user code) agrees with the kernel's pro-atom on the owner block
    out = np.zeros(h2o["grid"].atgrids[0].size)
    part.eval_proatom(0, out, h2o["grid"].atgrids[0])
    w = out / part.to_atomic_grid(0, part["promoldens"])
    np.testing.assert_allclose(np.clip(w, 0, 1)[::53], part["at_weights_0"][::53], rtol=2e-3, atol=1e-6)


def test_dispersion_hirshfeld_i(h2o, h2o_proatomdb):
    db, _ = h2o_proatomdb
    part = _part(h2o, "hi", proatomdb=db)
    with contextlib.redirect_stdout(io.StringIO()):
        part.do_dispersion()
    for key in ("radial_moments", "volumes", "volume_ratios", "c6s"):
        np.testing.assert_allclose(part[key], GOLD[f"hi/{key}"], rtol=1e-7, err_msg=key)
    part.do_prosplines()
    pts = db.get_rgrid(8).points
    got = part.cache.load("spline_prodensity", 0)(np.sqrt(pts[:-1] * pts[1:]))
    np.testing.assert_allclose(got, GOLD["hi/prospline_0_mid"], rtol=1e-7, atol=1e-14)


def test_do_all_lists_output_keys(h2o):
    part = _part(h2o, "mbis")
    with contextlib.redirect_stdout(io.StringIO()):
        keys = part.do_all()
    for expected in ("charges", "populations", "cartesian_multipoles", "pure_multipoles", "radial_moments",
                     "niter", "history_charges", ("density_decomposition", 0), ("spline_prodensity", 2)):  # fmt: skip
        assert expected in keys, expected


@pytest.mark.parametrize("basis", ["gauss", "slater"])
def test_aim_cube_arrays_on_the_device(tmp_path, basis):
    """hp_aim_on_points against the NumPy restatement of the reference's part-cube lines (oracle/cube_oracle.py,
    scripts/generate_cube.py:140-157, 213-227), then the `part-cube` program end to end."""
    import cube_oracle
    import yaml

    from horton_part_b200 import synthetic
    from horton_part_b200.core.basis import ExpBasisFuncHelper
    from horton_part_b200.scripts import generate_cube as gc

    coords, numbers = synthetic.water_cluster(6, 0)
    helper = ExpBasisFuncHelper.from_function_type(basis)
    rng = np.random.default_rng(7)
    propars = np.concatenate([np.asarray(helper.get_initial(int(z)), float) * rng.uniform(0.5, 1.5, helper.get_nshell(int(z)))
                              for z in numbers])  # fmt: skip
    grid = gc.UniformGrid.from_molecule(numbers, coords, spacing=0.45, extension=5.0)
    pts = grid.points
    density = synthetic.expbasis_promolecule_host(pts, coords, numbers, ExpBasisFuncHelper.from_function_type("gauss"),
                                                  scale={8: 8.6, 1: 0.7})  # fmt: skip
    rho0, promol, aim = gc.aim_on_points(numbers, coords, pts, density, propars, basis, chunk=50000)  # several chunks
    r0, p0, a0 = cube_oracle.aim_on_points(helper, numbers, coords, pts, density, propars)
    np.testing.assert_allclose(rho0, r0, rtol=1e-12, atol=1e-300)
    np.testing.assert_allclose(promol, p0, rtol=1e-12, atol=1e-300)
    np.testing.assert_allclose(aim, a0, rtol=1e-12, atol=1e-300)
    # the program: uniform-grid NPZ + part-dens output -> NPZ with the AIM arrays + cube files
    fn_in, fn_part, fn_out = tmp_path / "grid.npz", tmp_path / "part.npz", tmp_path / "out" / "cube.npz"
    np.savez(fn_in, atnums=numbers, atcorenums=numbers.astype(float), atcoords=coords, origin=grid.origin,
             axes=grid.axes, shape=np.array(grid.shape), density=density)
    np.savez(fn_part, history_propars=np.stack([propars * 0.9, propars]))
    cfg = tmp_path / "cube.yaml"
    cfg.write_text(yaml.safe_dump({"part-cube": {"inputs": [str(fn_in)], "partdens": [str(fn_part)],
                                                 "outputs": [str(fn_out)], "basis_func": basis}}))
    assert gc.main([str(cfg)]) == 0
    out = np.load(fn_out)
    np.testing.assert_allclose(out["aim_rho"], a0, rtol=1e-12, atol=1e-300)
    back = gc.read_cube(tmp_path / "out" / "cube_rho0_3.cube")
    np.testing.assert_allclose(back["data"], r0[3], rtol=1e-5, atol=1e-30)
    assert (tmp_path / "out" / "cube_rho_mol.cube").exists() and (tmp_path / "out" / "cube_rho0_mol.cube").exists()
