#!/usr/bin/env python
"""Timing of BASELINE.json configs 2-4 at full size on one B200 (evidence for DESIGN.md; the driver's
headline bench is bench.py = config 5).  Prints one JSON line per run.

    python tools/bench_configs.py [2] [3] [4]
"""

from __future__ import annotations

import json
import logging
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
logging.disable(logging.INFO)

import torch  # noqa: E402

from horton_part_b200 import gridlite, synthetic  # noqa: E402
from horton_part_b200.core.basis import ExpBasisFuncHelper  # noqa: E402

DEV = torch.device("cuda:0")


def grid_for(coords, numbers, nrad=150, nang=194):
    rgrid = gridlite.BeckeRTransform(1e-4, 1.5).transform_1d_grid(gridlite.GaussChebyshev(nrad))
    return gridlite.MolGrid.from_size(numbers, coords, nang, rgrid, np.ones(len(numbers) * nrad * nang), store=True)


def finish_grid(grid, w):
    grid.aim_weights[:] = w
    grid.weights[:] = grid.atweights * w


def timed(label, part, extra=None):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    part.do_partitioning()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    niter = int(part["niter"])
    out = {"run": label, "natom": part.natom, "npts": int(part.grid.size), "niter": niter, "seconds": dt,
           "ms_per_iteration": 1e3 * dt / niter, "iterations_per_s": niter / dt,
           "charges_head": [float(x) for x in part["charges"][:3]]}
    out.update(extra or {})
    print(json.dumps(out), flush=True)
    return out


def config2():
    """~20-atom organic, ISA (and Hirshfeld-free MBIS for reference) on a Slater promolecule."""
    from horton_part_b200 import ISAWPart, MBISWPart

    coords, numbers = synthetic.organic_like(20, seed=0)
    grid = grid_for(coords, numbers)
    rho, w, _, _ = synthetic.slater_promolecule_device(grid, coords, numbers, device=DEV)
    finish_grid(grid, w)
    pseudo = numbers.astype(float)
    timed("config2 ISA 20 atoms", ISAWPart(coords, numbers, pseudo, grid, rho, device=DEV))
    timed("config2 MBIS 20 atoms", MBISWPart(coords, numbers, pseudo, grid, rho, device=DEV))


def config3():
    """aLISA-sc, 100-atom water cluster, gauss and slater bases, Gaussian promolecule density."""
    from horton_part_b200 import LinearISAWPart

    coords, numbers = synthetic.water_cluster(100, seed=0)
    grid = grid_for(coords, numbers)
    helper = ExpBasisFuncHelper.from_function_type("gauss")
    rho, w = synthetic.expbasis_promolecule_device(grid, coords, numbers, helper, scale={8: 8.6, 1: 0.7}, device=DEV)
    finish_grid(grid, w)
    pseudo = numbers.astype(float)
    for basis, maxiter in (("gauss", 500), ("slater", 50)):
        part = LinearISAWPart(coords, numbers, pseudo, grid, rho, solver="sc", basis_func=basis, device=DEV,
                              maxiter=maxiter)
        kbar = float(np.mean([part.bs_helper.get_nshell(int(z)) for z in numbers]))
        out = timed(f"config3 aLISA-sc {basis} 100 atoms", part, {"mean_shells": kbar})
        ev = np.sum(part.history_time_update_at_weights)
        print(json.dumps({"run": f"config3 {basis} kernel", "evals_per_s_in_weights_kernel":
                          out["niter"] * 100 * grid.size / ev, "weights_kernel_s": ev}), flush=True)
    # the reference's DEFAULT aLISA solver (the convex programme): host plug-in on the projected
    # radial problems, all atoms stacked per outer iteration (lisa_solvers.solver_cvxopt_batched)
    part = LinearISAWPart(coords, numbers, pseudo, grid, rho, device=DEV)
    timed("config3 aLISA default solver (convex programme, batched host plug-in) gauss 100 atoms", part,
          {"seconds_in_host_solver": float(np.sum(part.history_time_update_propars))})


def config4():
    """gLISA on a 300-atom peptide-like chain, gauss basis (M = 1,500), Gaussian promolecule."""
    from horton_part_b200 import GlobalLinearISAWPart

    coords, numbers = synthetic.peptide_like(300, seed=0)
    grid = grid_for(coords, numbers)
    helper = ExpBasisFuncHelper.from_function_type("gauss")
    rho, w = synthetic.expbasis_promolecule_device(grid, coords, numbers, helper, device=DEV)
    finish_grid(grid, w)
    pseudo = numbers.astype(float)
    part = GlobalLinearISAWPart(coords, numbers, pseudo, grid, rho, solver="newton", device=DEV)
    out = timed("config4 gLISA newton 300 atoms", part)
    # one Hessian on its own
    part._promol_and_entropy()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    part.hessian()
    e1.record()
    torch.cuda.synchronize()
    M = part._table.nshell
    ms = e0.elapsed_time(e1)
    flop = float(M) * (M + 1) * grid.size
    print(json.dumps({"run": "config4 hessian", "M": M, "npts": int(grid.size), "ms": ms,
                      "tflops_algorithmic": flop / (ms * 1e-3) / 1e12}), flush=True)
    e0.record()
    part._shell_integrals(1)
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"run": "config4 gradient/function_g moments", "ms": e0.elapsed_time(e1),
                      "shell_evals_per_s": M * grid.size / (e0.elapsed_time(e1) * 1e-3)}), flush=True)
    part2 = GlobalLinearISAWPart(coords, numbers, pseudo, grid, rho, solver="sc", device=DEV, threshold=1e-4)
    timed("config4 gLISA sc (threshold 1e-4) 300 atoms", part2)


if __name__ == "__main__":
    which = sys.argv[1:] or ["2", "3", "4"]
    for w in which:
        {"2": config2, "3": config3, "4": config4}[w]()
