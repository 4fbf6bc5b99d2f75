"""Two config-4 Hessians (the launch list of the second one is what tools/gpu_r3b.sh reads under ncu)."""
import sys

import torch

sys.path.insert(0, ".")
from tools import cases  # noqa: E402


def main(natom=300):
    from horton_part_b200 import GlobalLinearISAWPart, synthetic
    from horton_part_b200.core.basis import ExpBasisFuncHelper

    dev = "cuda:0"
    coords, numbers = synthetic.peptide_like(natom, seed=0)
    grid = cases.grid_for(coords, numbers)
    helper = ExpBasisFuncHelper.from_function_type("gauss")
    rho, w = synthetic.expbasis_promolecule_device(grid, coords, numbers, helper, device=dev,
                                                   scale={1: 0.75, 6: 6.2, 7: 7.3, 8: 8.4})  # fmt: skip
    cases.finish_grid(grid, w)
    part = GlobalLinearISAWPart(coords, numbers, numbers.astype(float), grid, rho, solver="newton", device=dev)
    part._init_propars()
    part._promol_and_entropy()
    for _ in range(2):
        part.hessian()
        torch.cuda.synchronize()
    print("quadrants", part.hessian_tiles())


main(*(int(a) for a in sys.argv[1:]))
