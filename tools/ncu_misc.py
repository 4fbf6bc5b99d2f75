#!/usr/bin/env python
"""Small driver for ncu captures of the kernels the big benchmarks do not reach: the aLISA block solver
(config 3, slater table), the molecular-grid update pass (grid_type 2) and the Becke weights kernel."""
import logging
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
logging.disable(logging.INFO)

import torch  # noqa: E402

import cases  # noqa: E402

from horton_part_b200 import LinearISAWPart, gridlite, synthetic  # noqa: E402
from horton_part_b200.core.basis import ExpBasisFuncHelper  # noqa: E402

dev = torch.device("cuda:0")
helper = ExpBasisFuncHelper.from_function_type("gauss")
# Becke weights on the device for a 100-atom cluster at the real grid size
coords, numbers = synthetic.water_cluster(100, seed=0)
rgrid = gridlite.BeckeRTransform(1e-4, 1.5).transform_1d_grid(gridlite.GaussChebyshev(150))
grid = gridlite.MolGrid.from_size(numbers, coords, 194, rgrid, gridlite.DeviceBeckeWeights(), store=True)
rho = synthetic.expbasis_promolecule_host(grid.points[:10], coords, numbers, helper)  # touch the helper only
# config 3, slater table, three outer iterations: lisa_sc_block_kernel
grid2 = cases.grid_for(coords, numbers)
rho2, w2 = synthetic.expbasis_promolecule_device(grid2, coords, numbers, helper, scale={8: 8.6, 1: 0.7}, device=dev)
cases.finish_grid(grid2, w2)
part = LinearISAWPart(coords, numbers, numbers.astype(float), grid2, rho2, solver="sc", basis_func="slater", device=dev,
                      maxiter=3, device_loop=False)
part.do_partitioning()
print("config3 slater: inner iterations per atom (max)", int(part._state.niter.max().item()))
# molecular-grid update pass: aLISA sc, grid_type 2, 12 atoms on a 40 x 50 grid
c12, n12 = synthetic.water_cluster(12, seed=0)
rg = gridlite.BeckeRTransform(1e-4, 1.5).transform_1d_grid(gridlite.GaussChebyshev(40))
g12 = gridlite.MolGrid.from_size(n12, c12, 50, rg, gridlite.BeckeWeights(), store=True)
r12 = synthetic.expbasis_promolecule_host(g12.points, c12, n12, helper, scale={8: 8.6, 1: 0.7})
p12 = LinearISAWPart(c12, n12, n12.astype(float), g12, r12, solver="sc", grid_type=2, device=dev, maxiter=2)
p12.do_partitioning()
torch.cuda.synchronize()
print("done")
