"""cProfile of one config-4 gLISA Newton run (host-side view: where the wall time between kernels goes)."""
import cProfile
import io
import pstats
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from tools import cases  # noqa: E402


def main():
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
    for rep in range(2):
        part = GlobalLinearISAWPart(coords, numbers, numbers.astype(float), grid, rho, solver="newton", device=dev)
        torch.cuda.synchronize()
        pr = cProfile.Profile()
        pr.enable()
        part.do_partitioning()
        torch.cuda.synchronize()
        pr.disable()
    out = io.StringIO()
    pstats.Stats(pr, stream=out).sort_stats("cumulative").print_stats(45)
    print(out.getvalue())
    print("niter", part["niter"])


main()
