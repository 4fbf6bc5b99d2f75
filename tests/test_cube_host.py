"""AIM cube export (horton_part_b200/scripts/generate_cube.py): file format as the reference
writes it (scripts/generate_cube.py:100-137), round trip, and the AIM arrays on a uniform grid for
a promolecular density, where the weight functions reproduce the generating atoms.  CPU only: the
arrays come from the oracle's NumPy restatement here, tests/test_gpu_postproc.py checks the device kernel."""

import numpy as np
import pytest

from horton_part_b200 import synthetic
from horton_part_b200.core.basis import ExpBasisFuncHelper
from horton_part_b200.scripts import generate_cube as gc


def _water():
    coords, numbers = synthetic.water_cluster(3, 0)
    helper = ExpBasisFuncHelper.from_function_type("gauss")
    propars = np.concatenate([np.asarray(helper.get_initial(int(z)), float) / np.sum(helper.get_initial(int(z)))
                              * {8: 8.6, 1: 0.7}[int(z)] for z in numbers])  # fmt: skip
    return coords, numbers, helper, propars


def test_uniform_grid_order_and_quadrature():
    grid = gc.UniformGrid([0.0, 1.0, 2.0], np.diag([0.5, 0.25, 0.125]), (2, 3, 4))
    pts = grid.points
    assert pts.shape == (24, 3) and grid.size == 24
    np.testing.assert_allclose(pts[0], [0, 1, 2])
    np.testing.assert_allclose(pts[1], [0, 1, 2.125])  # z runs fastest
    np.testing.assert_allclose(pts[4], [0, 1.25, 2])  # then y
    np.testing.assert_allclose(pts[12], [0.5, 1, 2])  # x is the outer loop
    assert np.allclose(grid.weights, 0.5 * 0.25 * 0.125)
    coords, numbers, *_ = _water()
    box = gc.UniformGrid.from_molecule(numbers, coords, spacing=0.4, extension=3.0)
    assert (box.points.min(axis=0) <= coords.min(axis=0) - 3.0 + 1e-12).all()
    assert (box.points.max(axis=0) >= coords.max(axis=0) + 3.0 - 1e-12).all()
    with pytest.raises(ValueError):
        gc.UniformGrid([0, 0, 0], np.identity(3), (2, 0, 2))


def test_cube_file_format_and_round_trip(tmp_path):
    coords, numbers, *_ = _water()
    grid = gc.UniformGrid([-1.0, -2.0, -3.0], np.diag([0.5, 0.6, 0.7]), (3, 2, 4))
    data = np.linspace(-1e-3, 2.5e4, grid.size)
    path = tmp_path / "x.cube"
    gc.to_cube(path, numbers, numbers.astype(float), coords, grid, data)
    rows = path.read_text().splitlines()
    assert rows[0] == "Cubefile created with HORTON-PART"
    assert rows[1] == "OUTER LOOP: X, MIDDLE LOOP: Y, INNER LOOP: Z"
    assert rows[2] == "    3   -1.000000   -2.000000   -3.000000"
    assert rows[3] == "    3    0.500000    0.000000    0.000000"
    assert rows[5] == "    4    0.000000    0.000000    0.700000"
    assert rows[6].startswith("    8    8.000000") and len(rows[6].split()) == 5
    assert rows[9] == "".join(" %12.5E" % v for v in data[:6]) and len(rows) == 9 + 4  # 24 values, six per line
    back = gc.read_cube(path)
    assert back["grid"].shape == grid.shape and np.allclose(back["grid"].axes, grid.axes)
    np.testing.assert_array_equal(back["atnums"], numbers)
    np.testing.assert_allclose(back["atcoords"], coords, atol=5e-7)
    np.testing.assert_allclose(back["data"], data, rtol=1e-5, atol=1e-12)
    with pytest.raises(ValueError, match="cube"):
        gc.to_cube(tmp_path / "x.txt", numbers, numbers, coords, grid, data)
    with pytest.raises(ValueError, match="same size"):
        gc.to_cube(path, numbers, numbers, coords, grid, data[:-1])


def test_aim_arrays_of_a_promolecular_density(tmp_path):
    coords, numbers, helper, propars = _water()
    grid = gc.UniformGrid.from_molecule(numbers, coords, spacing=0.35, extension=6.0)
    pts = grid.points
    density = synthetic.expbasis_promolecule_host(pts, coords, numbers, helper, scale={8: 8.6, 1: 0.7})
    import cube_oracle  # the NumPy restatement of the reference's lines (oracle/); the product evaluates on the GPU

    arrays = cube_oracle.aim_on_points(helper, numbers, coords, pts, density, propars)
    out = gc.write_aim_cubes(str(tmp_path / "w"), numbers, numbers.astype(float), coords, grid, density, arrays=arrays)
    rho0, promol, aim = out["rho0"], out["promol"], out["aim_rho"]
    assert rho0.shape == aim.shape == (3, grid.size)
    # the density IS the promolecule of these coefficients: the AIM densities are the pro-atoms
    np.testing.assert_allclose(promol, density, rtol=1e-12, atol=1e-99)
    np.testing.assert_allclose(aim, rho0, rtol=1e-10, atol=1e-99)
    np.testing.assert_allclose(aim.sum(axis=0), density, rtol=1e-12, atol=1e-99)
    # one row against the helper directly (compute_rho0 = scripts/generate_cube.py:140-157)
    r1 = np.linalg.norm(pts - coords[1], axis=1)
    np.testing.assert_allclose(rho0[1], helper.compute_proatom_dens(1, propars[6:10], r1, 0), rtol=1e-14)
    files = sorted(p.name for p in tmp_path.iterdir())
    assert files == sorted(["w_rho_mol.cube", "w_rho0_mol.cube"] + [f"w_rho_{a}.cube" for a in range(3)]
                           + [f"w_rho0_{a}.cube" for a in range(3)])  # fmt: skip
    np.testing.assert_allclose(gc.read_cube(tmp_path / "w_rho0_2.cube")["data"], rho0[2], rtol=1e-5, atol=1e-30)
    with pytest.raises(ValueError, match="number of coefficients"):
        gc.compute_rho0(numbers, np.ones((3, 4)), propars[:-1])
    with pytest.raises(RuntimeError, match="Invalid func_type"):
        gc.compute_rho0(numbers, np.ones((3, 4)), propars, "no-such-basis")
