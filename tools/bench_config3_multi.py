#!/usr/bin/env python
"""BASELINE.json config 3 (aLISA sc, 100-atom water cluster) and config 4 (gLISA newton, 300 atoms) sharded
over the ranks of a torchrun launch: strong scaling of the two configurations that exchange more than the
state vector (config 4: the 18 MB Hessian all-reduce).  Rank 0 prints one JSON line per run.

    python -m torch.distributed.run --nproc-per-node N tools/bench_config3_multi.py [3] [4]
"""
from __future__ import annotations

import json
import logging
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
logging.disable(logging.INFO)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import cases  # noqa: E402

from horton_part_b200 import GlobalLinearISAWPart, LinearISAWPart, synthetic  # noqa: E402
from horton_part_b200.core.basis import ExpBasisFuncHelper  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    comm = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        comm = dist.group.WORLD
    fd = os.dup(1)
    os.dup2(2, 1)  # NCCL banners go to stderr
    which = [a for a in sys.argv[1:] if a in ("3", "4")] or ["3"]
    helper = ExpBasisFuncHelper.from_function_type("gauss")

    def emit(obj):
        if rank == 0:
            os.write(fd, (json.dumps(obj) + "\n").encode())

    def timed(part):
        if comm is not None:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        part.do_partitioning()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if comm is not None:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return float(dt.item())

    cases.warm_up(dev)
    if "3" in which:
        coords, numbers = synthetic.water_cluster(100, seed=0)
        grid = cases.grid_for(coords, numbers)
        rho, w = synthetic.expbasis_promolecule_device(grid, coords, numbers, helper, scale={8: 8.6, 1: 0.7}, device=dev)
        cases.finish_grid(grid, w)
        for basis, maxiter in (("gauss", 500), ("slater", 50)):
            part = LinearISAWPart(coords, numbers, numbers.astype(float), grid, rho, solver="sc", basis_func=basis,
                                  device=dev, comm=comm, maxiter=maxiter)
            dt = timed(part)
            emit({"run": f"config3 aLISA sc {basis}, 100 atoms", "n_gpus": world, "niter": int(part["niter"]), "seconds": dt,
                  "ms_per_iteration": 1e3 * dt / int(part["niter"]), "charges_head": [float(x) for x in part["charges"][:3]]})
    if "4" in which:
        coords, numbers = synthetic.peptide_like(300, seed=0)
        grid = cases.grid_for(coords, numbers)
        rho, w = synthetic.expbasis_promolecule_device(grid, coords, numbers, helper, device=dev,
                                                       scale={1: 0.75, 6: 6.2, 7: 7.3, 8: 8.4})
        cases.finish_grid(grid, w)
        part = GlobalLinearISAWPart(coords, numbers, numbers.astype(float), grid, rho, solver="newton", device=dev, comm=comm)
        dt = timed(part)
        emit({"run": "config4 gLISA newton, 300 atoms, M = 1500", "n_gpus": world, "niter": int(part["niter"]), "seconds": dt,
              "seconds_per_newton_iteration": dt / int(part["niter"]), "charges_head": [float(x) for x in part["charges"][:3]]})
    if comm is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
