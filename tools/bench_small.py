#!/usr/bin/env python
"""Config 1 (BASELINE.json): H2O MBIS on the reference's test grid, end to end from host arrays, with the
device-resident loop and with the host-driven loop.  The reference needs 0.16 s on one CPU core
(SURVEY.md section 6).  Prints one JSON line per variant.  Inputs: tests/golden/h2o_hf_sto3g.npz."""

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

from horton_part_b200 import gridlite  # noqa: E402


def h2o_case():
    z = np.load(os.path.join(ROOT, "tests", "golden", "h2o_hf_sto3g.npz"))
    coords, numbers, pseudo = z["coordinates"], z["numbers"], z["pseudo_numbers"]
    rgrid = gridlite.ExpRTransform(5e-4, 2e1, 119).transform_1d_grid(gridlite.UniformInteger(120))
    grid = gridlite.MolGrid.from_size(numbers, coords, 110, rgrid, z["aim_weights"], store=True)
    return coords, numbers, pseudo, grid, z["dens"], z


def run(scheme, device_loop, reps=7, **kw):
    from horton_part_b200 import wpart_schemes

    coords, numbers, pseudo, grid, rho, z = h2o_case()
    times, loop = [], []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        part = wpart_schemes(scheme)(coords, numbers, pseudo, grid, rho, device_loop=device_loop, **kw)
        part.do_partitioning()
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
        loop.append(part.time_usage.get("do_partitioning_loop", float("nan")))
    out = {"run": f"config1 H2O {scheme} {kw or ''}".strip(), "device_loop": device_loop, "npts": int(grid.size),
           "niter": int(part["niter"]), "seconds_first_call": times[0], "seconds_best": min(times),
           "seconds_median": float(np.median(times)), "loop_seconds_best": min(loop),
           "ms_per_iteration_best": 1e3 * min(loop) / int(part["niter"]),
           "gpu_seconds_weights": float(np.sum(part.history_time_update_at_weights)),
           "gpu_seconds_propars": float(np.sum(part.history_time_update_propars)),
           "charges": [float(q) for q in part["charges"]], "reference_cpu_seconds": 0.16 if scheme == "mbis" else None}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    for flag in (True, False):
        run("mbis", flag)
    for flag in (True, False):
        run("is", flag)
    for flag in (True, False):
        run("lisa", flag, solver="sc")
