#!/usr/bin/env python
"""cProfile of rank 0's end-to-end MBIS call on config 5 under a real torchrun launch (NCCL, N ranks):
where the fixed host time of a sharded call goes.  python -m torch.distributed.run --nproc-per-node N tools/e2e_profile_multi.py"""
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
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from horton_part_b200 import MBISWPart, synthetic  # noqa: E402
from horton_part_b200.core import hostmem  # noqa: E402
from horton_part_b200.core.device import Shard  # noqa: E402
from horton_part_b200.mbis import mbis_atom_work  # noqa: E402


def main(steps=5):
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    comm = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        comm = dist.group.WORLD
    coords, numbers, grid = bench.build_system(2000)
    shard = Shard(2000, grid.indices, rank, world, work=mbis_atom_work(coords, numbers, grid, dev) if world > 1 else None)
    rho_loc, w_loc, lo, hi = synthetic.slater_promolecule_device(grid, coords, numbers, device=dev, shard=shard)
    rho = np.zeros(grid.size)
    rho[lo:hi] = rho_loc
    grid.aim_weights[lo:hi] = w_loc
    grid.weights[lo:hi] = grid.atweights[lo:hi] * w_loc
    for name in ("points", "weights", "atweights"):
        pin = hostmem.pinned_empty(getattr(grid, name).shape)
        pin[...] = getattr(grid, name)
        setattr(grid, name, pin)
    pin = hostmem.pinned_empty(rho.shape)
    pin[...] = rho
    rho = pin
    pseudo = numbers.astype(float)
    for rep in range(3):
        if comm is not None:
            dist.barrier()
        torch.cuda.synchronize()
        prof = cProfile.Profile() if (rep == 2 and rank == 0) else None
        t0 = time.perf_counter()
        if prof:
            prof.enable()
        part = MBISWPart(coords, numbers, pseudo, grid, rho, device=dev, comm=comm, maxiter=steps)
        part.do_partitioning()
        torch.cuda.synchronize()
        if prof:
            prof.disable()
        dt = time.perf_counter() - t0
        gpu = float(np.sum(part.history_time_update_at_weights) + np.sum(part.history_time_update_propars))
        if rank == 0:
            print(f"rep {rep}: e2e {dt:.3f} s on rank 0 of {world}; GPU time in iterations {gpu:.3f} s; rest {dt - gpu:.3f} s", flush=True)
        del part
    if rank == 0:
        out = io.StringIO()
        pstats.Stats(prof, stream=out).sort_stats("cumulative").print_stats(40)
        print(out.getvalue())


main()
