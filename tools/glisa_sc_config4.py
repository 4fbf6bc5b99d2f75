"""gLISA `sc` solver on config 4 (300 atoms, M = 1,500): iterations, wall time and the per-iteration passes."""
import json
import logging
import sys
import time

import torch

sys.path.insert(0, ".")
logging.disable(logging.INFO)
from tools import cases  # noqa: E402


def main(threshold=1e-6):
    from horton_part_b200 import GlobalLinearISAWPart, synthetic
    from horton_part_b200.core.basis import ExpBasisFuncHelper

    dev = "cuda:0"
    cases.warm_up(dev)
    coords, numbers = synthetic.peptide_like(300, seed=0)
    grid = cases.grid_for(coords, numbers)
    helper = ExpBasisFuncHelper.from_function_type("gauss")
    rho, w = synthetic.expbasis_promolecule_device(grid, coords, numbers, helper, device=dev,
                                                   scale={1: 0.75, 6: 6.2, 7: 7.3, 8: 8.4})  # fmt: skip
    cases.finish_grid(grid, w)
    part = GlobalLinearISAWPart(coords, numbers, numbers.astype(float), grid, rho, solver="sc", device=dev,
                                threshold=threshold, maxiter=5000)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    part.do_partitioning()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"run": "config 4 gLISA sc", "threshold": threshold, "niter": int(part["niter"]), "seconds": dt,
                      "ms_per_iteration": 1e3 * dt / int(part["niter"]),
                      "charges_head": [float(x) for x in part["charges"][:3]]}))


main(*(float(a) for a in sys.argv[1:]))
