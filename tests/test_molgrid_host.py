"""Host half of the grid_type 2/3 plug-in path (gisa.molgrid_host_update) against a reference run:
the promolecule, which the product gets from the device pass, is built here with NumPy so that the
per-atom K_a x Npts algebra and the convex-programme plug-in are checked on CPU.  Golden:
tests/golden/water6_convex.npz (reference LinearISAWPart(grid_type=2), default solver, its
cvxopt.solvers.cp call answered by the oracle's stand-in)."""

import logging

import numpy as np
from conftest import GOLDEN

from horton_part_b200 import gridlite, lisa_solvers, synthetic
from horton_part_b200.core.basis import ExpBasisFuncHelper
from horton_part_b200.gisa import molgrid_host_update


def test_molgrid_host_update_reproduces_reference_run():
    gold = np.load(GOLDEN / "water6_convex.npz")
    tag = "g/lisa_cvxopt_gt2"
    coords, numbers = synthetic.water_cluster(6, 0)
    rgrid = gridlite.BeckeRTransform(1e-4, 1.5).transform_1d_grid(gridlite.GaussChebyshev(40))
    grid = gridlite.MolGrid.from_size(numbers, coords, 50, rgrid, gridlite.BeckeWeights(), store=True)
    helper = ExpBasisFuncHelper.from_function_type("gauss")
    rho = synthetic.expbasis_promolecule_host(grid.points, coords, numbers, helper, scale={8: 8.6, 1: 0.7})
    pseudo = numbers.astype(float)
    ranges = np.concatenate([[0], np.cumsum([helper.get_nshell(z) for z in numbers])])
    propars = np.ones(ranges[-1])
    for a, z in enumerate(numbers):  # gisa.py:68-88
        inits = np.maximum(helper.get_initial(z), 1e-4)
        propars[ranges[a] : ranges[a + 1]] = inits / inits.sum() * pseudo[a]
    propars *= grid.integrate(rho) / propars.sum()
    bs = []
    for a, z in enumerate(numbers):
        r = np.linalg.norm(grid.points - coords[a], axis=1)
        bs.append(np.array([helper.compute_proshell_dens(z, k, 1.0, r) for k in range(helper.get_nshell(z))]))
    log = logging.getLogger("test_molgrid_host")

    def opt(a, bs_a, rho_a, start):
        return lisa_solvers.solver_cvxopt(bs_a, rho_a, start, grid.points, grid.weights, 1e-8, log, 1e-15, -1e-12,
                                          1e-4, engine="builtin")  # fmt: skip

    changes = []
    for niter in range(1, 200):
        promol = np.zeros(grid.size)
        for a in range(len(numbers)):  # core/stockholder.py:153-175
            promol += propars[ranges[a] : ranges[a + 1]] @ bs[a]
            promol += 1e-100
        propars, charges, msd = molgrid_host_update(promol, rho, grid.points, grid.weights, bs, ranges, propars,
                                                    pseudo, opt)  # fmt: skip
        changes.append(np.sqrt(msd.sum()))
        if changes[-1] < 1e-6:
            break
    assert niter == int(gold[f"{tag}/niter"])
    np.testing.assert_allclose(charges, gold[f"{tag}/charges"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(propars, gold[f"{tag}/propars"], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(changes, gold[f"{tag}/history_changes"], rtol=1e-6, atol=1e-12)
