"""Config-4 Hessian with and without the panel's block screening (HP_B200_HESSIAN_SCREEN is read per call)."""
import os
import sys

import torch

sys.path.insert(0, ".")
from tools import cases  # noqa: E402


def main(natom=300):
    from horton_part_b200 import GlobalLinearISAWPart, synthetic
    from horton_part_b200.core.basis import ExpBasisFuncHelper

    dev = "cuda:0"
    cases.warm_up(dev)
    coords, numbers = synthetic.peptide_like(natom, seed=0)
    grid = cases.grid_for(coords, numbers)
    helper = ExpBasisFuncHelper.from_function_type("gauss")
    rho, w = synthetic.expbasis_promolecule_device(grid, coords, numbers, helper, device=dev,
                                                   scale={1: 0.75, 6: 6.2, 7: 7.3, 8: 8.4})  # fmt: skip
    cases.finish_grid(grid, w)
    part = GlobalLinearISAWPart(coords, numbers, numbers.astype(float), grid, rho, solver="newton", device=dev)
    part._init_propars()
    part._promol_and_entropy()
    part.hessian()
    torch.cuda.synchronize()
    for label, env in (("screened", {}), ("unscreened", {"HP_B200_HESSIAN_SCREEN": "0"}), ("screened again", {})):
        os.environ.pop("HP_B200_HESSIAN_SCREEN", None)
        os.environ.update(env)
        times = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            part.hessian()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        print(f"{label:36s} {min(times):8.2f} ms   tiles {part.hessian_tiles()[:2]}", flush=True)


main(*(int(a) for a in sys.argv[1:]))
