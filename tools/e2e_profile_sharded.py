#!/usr/bin/env python
"""Where does the fixed host time of a SHARDED end-to-end call go?  Runs rank 0's share of an 8-way sharded
config-5 job on ONE GPU with a stand-in communicator whose all-reduce is a no-op (timings are representative,
results are not), under cProfile, and prints the phases + the top host functions.  VERDICT r1, weak #9."""
import cProfile
import io
import logging
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
logging.disable(logging.INFO)

import torch  # noqa: E402

import bench  # noqa: E402
from horton_part_b200 import MBISWPart, synthetic  # noqa: E402
from horton_part_b200.core.comm import HpComm  # noqa: E402
from horton_part_b200.core.device import Shard  # noqa: E402
from horton_part_b200.mbis import mbis_atom_work  # noqa: E402


class FakeComm(HpComm):
    def __init__(self, world, rank):
        self.world, self.rank, self._handle = world, rank, 0

    def all_reduce(self, tensor, op="sum"):
        return tensor


def main(world=8, steps=5):
    dev = torch.device("cuda:0")
    coords, numbers, grid = bench.build_system(2000)
    shard = Shard(2000, grid.indices, 0, world, work=mbis_atom_work(coords, numbers, grid, dev))
    rho_loc, w_loc, lo, hi = synthetic.slater_promolecule_device(grid, coords, numbers, device=dev, shard=shard)
    rho = np.zeros(grid.size)
    rho[lo:hi] = rho_loc
    grid.aim_weights[lo:hi] = w_loc
    grid.weights[lo:hi] = grid.atweights[lo:hi] * w_loc
    comm = FakeComm(world, 0)
    for rep in range(3):
        torch.cuda.synchronize()
        prof = cProfile.Profile() if rep == 2 else None
        t0 = time.perf_counter()
        if prof:
            prof.enable()
        part = MBISWPart(coords, numbers, numbers.astype(float), grid, rho, device=dev, comm=comm, maxiter=steps)
        part.do_partitioning()
        torch.cuda.synchronize()
        if prof:
            prof.disable()
        dt = time.perf_counter() - t0
        gpu = float(np.sum(part.history_time_update_at_weights) + np.sum(part.history_time_update_propars))
        print(f"rep {rep}: e2e {dt:.3f} s for {steps} iterations on rank 0 of {world} ({hi - lo} points); GPU time in iterations {gpu:.3f} s; "
              f"fixed host+copy time {dt - gpu:.3f} s")
        if prof:
            s = io.StringIO()
            pstats.Stats(prof, stream=s).sort_stats("cumulative").print_stats(28)
            print(s.getvalue()[:6000])


if __name__ == "__main__":
    main()
