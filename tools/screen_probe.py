"""Dense pass of config 5 with atom screening on / off: ms per iteration, evaluated pairs, and the
largest difference of promolecule / charges between the two (run on the B200)."""
import logging
import os
import sys

sys.path.insert(0, ".")
import numpy as np
import torch

logging.disable(logging.INFO)
import bench
from horton_part_b200 import MBISWPart, synthetic
from horton_part_b200.core.device import Shard

natom = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
dev = torch.device("cuda:0")
coords, numbers, grid = bench.build_system(natom)
shard = Shard(natom, grid.indices, 0, 1)
rho, w_loc, lo, hi = synthetic.slater_promolecule_device(grid, coords, numbers, device=dev, shard=shard)
grid.weights[:] = grid.atweights * w_loc
res = {}
for mode in ("1", "0"):
    os.environ["HP_B200_ATOM_SCREEN"] = mode
    part = MBISWPart(coords, numbers, numbers.astype(float), grid, rho, device=dev, maxiter=6)
    part._init_propars()
    for _ in range(2):
        part._run_iteration()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        part._run_iteration()
    e1.record()
    torch.cuda.synchronize()
    res[mode] = (part.slab.promol.clone(), part["charges"].copy(), e0.elapsed_time(e1) / 4,
                 part._table.pairs_evaluated() / (natom * grid.size), part._table.shells_evaluated() / (natom * grid.size))
    print(f"atom_screen={mode}: {res[mode][2]:.2f} ms/iteration, pairs evaluated {res[mode][3]:.4f}, "
          f"shells per dense pair {res[mode][4]:.4f}", flush=True)
    del part
a, b = res["1"], res["0"]
d = (a[0] - b[0]).abs()
ulp = torch.from_numpy(np.spacing(b[0].abs().cpu().numpy())).to(dev)
print("promol: max |diff| / ulp =", float((d / ulp).max()), " points differing:", int((d > 0).sum()), "of", d.numel())
print("charges: max |diff| =", float(np.abs(a[1] - b[1]).max()))
