"""End-to-end MBIS call on config 5 with the slab uploaded in one piece / in two parts, from pageable and from
page-locked arrays: wall time of the call (3 repetitions each, after a warm-up call)."""
import logging
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
logging.disable(logging.INFO)

import torch  # noqa: E402

import bench  # noqa: E402
from horton_part_b200 import MBISWPart, synthetic  # noqa: E402
from horton_part_b200.core import hostmem  # noqa: E402


def main(natom=2000, steps=5):
    dev = torch.device("cuda:0")
    coords, numbers, grid = bench.build_system(natom)
    rho, w, lo, hi = synthetic.slater_promolecule_device(grid, coords, numbers, device=dev)
    grid.aim_weights[:] = w
    grid.weights[:] = grid.atweights * w
    pseudo = numbers.astype(float)

    def call():
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        part = MBISWPart(coords, numbers, pseudo, grid, rho, device=dev, maxiter=steps)
        part.do_partitioning()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        gpu = float(np.sum(part.history_time_update_at_weights) + np.sum(part.history_time_update_propars))
        return dt, gpu, float(part.history_time_update_at_weights[0])

    for mode in ("pageable", "pinned"):
        if mode == "pinned":
            for name in ("points", "weights", "atweights"):
                pin = hostmem.pinned_empty(getattr(grid, name).shape)
                pin[...] = getattr(grid, name)
                setattr(grid, name, pin)
            pin = hostmem.pinned_empty(rho.shape)
            pin[...] = rho
            rho = pin
        for split in ("0", "1"):
            os.environ["HP_B200_SPLIT_UPLOAD"] = split
            call()
            runs = [call() for _ in range(3)]
            print(f"{mode:9s} split={split}: e2e " + " ".join(f"{r[0]:.3f}" for r in runs) +
                  f" s; GPU time in iterations {runs[-1][1]:.3f} s; first weights pass {1e3 * runs[-1][2]:.1f} ms", flush=True)


main()
