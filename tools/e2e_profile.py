import sys, time, logging
sys.path.insert(0, '.')
import numpy as np, torch
logging.disable(logging.INFO)
import bench
from horton_part_b200 import MBISWPart, synthetic
from horton_part_b200.core.device import Shard
dev = torch.device('cuda:0')
coords, numbers, grid = bench.build_system(2000)
shard = Shard(2000, grid.indices, 0, 1)
rho_loc, w_loc, lo, hi = synthetic.slater_promolecule_device(grid, coords, numbers, device=dev, shard=shard)
rho = rho_loc; grid.weights[:] = grid.atweights * w_loc
torch.cuda.synchronize()
def T(label, t0):
    torch.cuda.synchronize(); t = time.perf_counter(); print(f"{label:28s} {t - t0:.3f} s"); return t
for rep in range(2):
    t = time.perf_counter(); t00 = t
    part = MBISWPart(coords, numbers, numbers.astype(float), grid, rho, device=dev, maxiter=5)
    t = T("constructor", t)
    _ = part.slab
    t = T("slab upload", t)
    part._init_propars()
    t = T("init_propars", t)
    for i in range(5):
        part._run_iteration()
    t = T("5 iterations", t)
    part._publish_weights()
    t = T("publish_weights", t)
    part.history_propars.append(part.cache.load('propars').copy())
    part._finalize_propars()
    t = T("finalize", t)
    print("total", time.perf_counter() - t00)
