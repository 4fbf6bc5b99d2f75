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
    direct = MBISWPart(water6["coords"], water6["numbers"], water6["pseudo"], water6["grid"], water6["rho"], maxiter=200)
    direct.do_partitioning()
    assert int(out["niter"]) == direct["niter"]
    np.testing.assert_allclose(out["charges"], direct["charges"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(out["save/part.cache/propars"], direct["propars"], rtol=1e-13)
    assert main([str(cfg), "--skip_exist_files"]) == 0
