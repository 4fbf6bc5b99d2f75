"""The part-gen / part-dens NPZ wire format: writer/reader round trip on the
CPU, the command-line program on the GPU."""

import numpy as np
import pytest


def test_npz_round_trip(tmp_path, water6):
    from horton_part_b200.scripts.partition_density import construct_molgrid_from_dict, write_part_gen_npz

    fn = tmp_path / "water6.npz"
    write_part_gen_npz(fn, water6["coords"], water6["numbers"], water6["pseudo"], water6["grid"], water6["rho"])
    data = np.load(fn)
    for key in ("atcoords", "atnums", "atcorenums", "density", "aim_weights", "points", "weights", "atom_idxs",
                "atom0/rgrid/points", "atom0/rgrid/weights", "atom0/shell_idxs", "atom5/points", "atom5/weights"):
        assert key in data, key  # keys of scripts/generate_density.py:232-257
    grid = construct_molgrid_from_dict(data)
    g0 = water6["grid"]
    assert np.array_equal(grid.points, g0.points) and np.array_equal(grid.weights, g0.weights)
    assert np.array_equal(grid.indices, g0.indices)
    assert np.array_equal(grid.atgrids[3].weights, g0.atgrids[3].weights)
    assert np.array_equal(grid.atgrids[3].indices, g0.atgrids[3].indices)
    assert np.array_equal(grid.atgrids[3].rgrid.points, g0.atgrids[3].rgrid.points)
    assert grid.integrate(water6["rho"]) == pytest.approx(g0.integrate(water6["rho"]), rel=1e-15)
    # files without stored per-atom points are rebuilt from the radial grid + shell sizes
    slim = {k: data[k] for k in data.files if not (k.startswith("atom") and k.endswith(("/points", "/weights"))
                                                   and "rgrid" not in k) and k not in ("points", "weights", "atom_idxs")}
    grid2 = construct_molgrid_from_dict(slim)
    np.testing.assert_allclose(grid2.points, g0.points, atol=1e-13)
    np.testing.assert_allclose(grid2.weights, g0.weights, rtol=1e-13)


@pytest.mark.gpu
def test_part_dens_program(tmp_path, water6):
    import yaml

    from horton_part_b200 import MBISWPart
    from horton_part_b200.scripts.partition_density import main, write_part_gen_npz

    fn_in, fn_out = tmp_path / "dens.npz", tmp_path / "out" / "part.npz"
    write_part_gen_npz(fn_in, water6["coords"], water6["numbers"], water6["pseudo"], water6["grid"], water6["rho"])
    cfg = tmp_path / "cfg.yaml"
    cfg.write_text(yaml.safe_dump({"part-dens": {"inputs": [str(fn_in)], "outputs": [str(fn_out)], "type": "mbis",
                                                 "maxiter": 200, "log_level": "WARNING", "save": ["propars"]}}))
    assert main([str(cfg)]) == 0
    out = np.load(fn_out)
    for key in ("history_entropies", "history_charges", "history_propars", "time", "time_update_at_weights",
                "time_update_propars", "niter", "charges", "natom", "atnums", "atcorenums", "lmax", "maxiter",
                "threshold"):  # tests/scripts/test_main.py:101-119
        assert key in out, key
    # against the REFERENCE's run on the same system (tests/golden/water6_slater.npz, oracle/gen_golden.py)
    gold = water6["gold"]
    assert int(out["niter"]) == int(gold["mbis/niter"])
    np.testing.assert_allclose(out["charges"], gold["mbis/charges"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(out["save/part.cache/propars"], gold["mbis/propars"], rtol=1e-8)
    np.testing.assert_allclose(out["history_changes"], gold["mbis/history_changes"], rtol=1e-6)
    np.testing.assert_allclose(out["history_entropies"], gold["mbis/history_entropies"], rtol=1e-8, atol=1e-11)
    # ... and against the class API driven directly
    direct = MBISWPart(water6["coords"], water6["numbers"], water6["pseudo"], water6["grid"], water6["rho"], maxiter=200)
    direct.do_partitioning()
    assert int(out["niter"]) == direct["niter"]
    np.testing.assert_allclose(out["charges"], direct["charges"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(out["save/part.cache/propars"], direct["propars"], rtol=1e-13)
    assert main([str(cfg), "--skip_exist_files"]) == 0


def test_stored_atom_grid_has_lebedev_degrees(water6):
    """part-dens rebuilds atomic grids from the NPZ; do_density_decomposition reads their l_max
    (core/base.py:646-657; qc-grid AtomGrid.l_max = largest Lebedev degree)."""
    from horton_part_b200.scripts.partition_density import _Stored, _StoredAtomGrid

    g = water6["grid"].atgrids[0]
    stored = _StoredAtomGrid(g.points, g.weights, _Stored(g.rgrid.points, g.rgrid.weights), g.indices, g.center)
    assert stored.l_max == g.l_max and stored.degrees == list(g.degrees) and stored.n_shells == g.rgrid.size
    with pytest.raises(ValueError, match="Lebedev"):
        _StoredAtomGrid(g.points[:7], g.weights[:7], _Stored(g.rgrid.points[:1], g.rgrid.weights[:1]), [0, 7], g.center)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["mbis", "lisa", "nlis", "is"])
def test_part_dens_density_decomposition_job(tmp_path, water6, kind):
    """`part_job_type: do_density_decomposition` (scripts/partition_density.py:290-322): radial projections and
    the (order, exponent, population) table per atom in the output, nested `save` entries resolved."""
    import yaml

    from horton_part_b200.scripts.partition_density import main, write_part_gen_npz

    fn_in, fn_out = tmp_path / "dens.npz", tmp_path / "part.npz"
    write_part_gen_npz(fn_in, water6["coords"], water6["numbers"], water6["pseudo"], water6["grid"], water6["rho"])
    cfg = tmp_path / "cfg.yaml"
    cfg.write_text(yaml.safe_dump({"part-dens": {
        "inputs": [str(fn_in)], "outputs": [str(fn_out)], "type": kind, "maxiter": 40, "solver": "sc",
        "part_job_type": "do_density_decomposition", "log_level": "WARNING", "save": ["coordinates", "charges"]}}))
    assert main([str(cfg)]) == 0
    out = np.load(fn_out)
    natom = len(water6["numbers"])
    nrad = water6["grid"].atgrids[0].rgrid.size
    for a in range(natom):
        for key in (f"radial_points_{a}", f"spherical_average_{a}", f"radial_weights_{a}", f"bs_info_{a}"):
            assert key in out, key
        assert out[f"spherical_average_{a}"].shape == (nrad,)
    last = out["history_propars"][-1]
    if kind == "mbis":  # rows (1, S, N): oxygen has two shells
        assert out["bs_info_0"].shape == (2, 3) and np.array_equal(out["bs_info_0"][:, 2], last[[0, 2]])
    elif kind == "lisa":  # rows (order, exponent, population) of the gauss table
        assert out["bs_info_0"].shape[1] == 3 and np.all(out["bs_info_0"][:, 0] == 2.0)
        assert out["bs_info_0"][:, 2].sum() == pytest.approx(8.0 - out["charges"][0], abs=1e-3)
    elif kind == "is":
        assert out["bs_info_0"].shape == (nrad,)
    np.testing.assert_array_equal(out["save/part.coordinates"], water6["coords"])  # attribute of the object
    np.testing.assert_array_equal(out["save/part.cache/charges"], out["charges"])  # cache key
