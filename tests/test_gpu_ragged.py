"""Ragged inputs through the whole path: every atom has its own radial grid size and a pruned
angular grid (a different Lebedev order per radial shell), atom blocks are not multiples of the
1,024-point chunks, one atom's block is smaller than a single chunk, another has a single radial
shell of 6 points.  MBIS / aLISA-sc / ISA against the pinned oracle: same iteration counts, charges
to 1e-8 (north_star tolerance)."""

import numpy as np
import pytest
import stockholder_oracle as oracle

pytestmark = pytest.mark.gpu


def ragged_case():
    from horton_part_b200 import gridlite, synthetic

    coords, numbers = synthetic.water_cluster(9, seed=3)
    rng = np.random.default_rng(7)
    orders = [3, 5, 7, 9, 11, 13, 15, 17]  # Lebedev degrees with 6 ... 110 points
    atgrids = []
    for a, center in enumerate(coords):
        nrad = [37, 23, 29, 41, 5, 31, 26, 1, 33][a]
        rgrid = gridlite.BeckeRTransform(1e-4, 1.5).transform_1d_grid(gridlite.GaussChebyshev(nrad))
        if nrad == 1:
            degrees = [3]
        else:  # pruned: small angular grids near the nucleus and far out, large ones in between
            degrees = [orders[min(len(orders) - 1, int(6 * np.sin(np.pi * (i + 0.5) / nrad)) + int(rng.integers(0, 2)))]
                       for i in range(nrad)]  # fmt: skip
        atgrids.append(gridlite.AtomGrid(rgrid, degrees=degrees, center=center))
    grid = gridlite.MolGrid(numbers, atgrids, gridlite.BeckeWeights(), store=True)
    rho = synthetic.slater_promolecule_host(grid.points, coords, numbers)
    sizes = np.diff(grid.indices)
    assert sizes.min() == 6 and (sizes % 1024 != 0).all() and len(set(sizes)) == len(sizes)
    return dict(coords=coords, numbers=numbers, pseudo=numbers.astype(float), grid=grid, rho=rho)


@pytest.fixture(scope="module")
def case():
    return ragged_case()


def _args(c):
    return c["coords"], c["numbers"], c["pseudo"], c["grid"], c["rho"]


def test_mbis_on_ragged_grids(case):
    from horton_part_b200 import MBISWPart

    part = MBISWPart(*_args(case), maxiter=60)
    part.do_partitioning()
    ref = oracle.mbis(*_args(case), maxiter=60)
    assert part["niter"] == ref["niter"]
    np.testing.assert_allclose(part["charges"], ref["charges"], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(part["propars"], ref["propars"], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(part["history_changes"], ref["history_changes"], rtol=1e-6, atol=1e-12)
    for a in (4, 7):  # the 5-shell atom and the single-shell atom
        lo, hi = case["grid"].indices[a], case["grid"].indices[a + 1]
        assert part[f"at_weights_{a}"].shape == (hi - lo,)
        assert np.isfinite(part[f"at_weights_{a}"]).all()


def test_alisa_sc_on_ragged_grids(case):
    from horton_part_b200 import LinearISAWPart

    part = LinearISAWPart(*_args(case), solver="sc", maxiter=40)
    part.do_partitioning()
    ref = oracle.alisa(*_args(case), solver="sc", maxiter=40)
    assert part["niter"] == ref["niter"]
    np.testing.assert_allclose(part["charges"], ref["charges"], rtol=1e-8, atol=1e-9)


def test_populations_and_moments_on_ragged_grids(case):
    """Post-processing kernels (segment integrals, multipoles) see the same ragged blocks."""
    from horton_part_b200 import MBISWPart

    part = MBISWPart(*_args(case), maxiter=5)
    part.do_charges()
    part.do_moments()
    grid, rho = case["grid"], case["rho"]
    for a in (0, 4, 7):
        lo, hi = grid.indices[a], grid.indices[a + 1]
        w = part[f"at_weights_{a}"]
        pop = np.einsum("i,i,i", grid.atweights[lo:hi], w, rho[lo:hi])
        assert abs(part["radial_moments"][a, 0] - pop) < 1e-11 * max(1.0, abs(pop))
