#!/usr/bin/env python
"""Measured conditioning behind the parity tolerances that are looser than 1e-8.

For every scheme the NumPy oracle (oracle/stockholder_oracle.py, pinned against the reference's runs) is run
twice on the 6-atom water fixture: on the density and on the density times (1 + 1e-13 * N(0,1)) -- the size of
the last-bit differences between two correct FP64 implementations of the same quadrature.  The amplification
factors  |delta x| / 1e-13  of charges, parameters and the convergence history are what a tolerance has to
absorb; the table goes into DESIGN.md section 5 and next to the tolerances in tests/.  CPU only (~1 min).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]

import stockholder_oracle as oracle  # noqa: E402

from horton_part_b200 import gridlite, synthetic  # noqa: E402

EPS = 1e-13


def water6():
    coords, numbers = synthetic.water_cluster(6, 0)
    rgrid = gridlite.BeckeRTransform(1e-4, 1.5).transform_1d_grid(gridlite.GaussChebyshev(40))
    grid = gridlite.MolGrid.from_size(numbers, coords, 50, rgrid, gridlite.BeckeWeights(), store=True)
    return coords, numbers, numbers.astype(float), grid, synthetic.slater_promolecule_host(grid.points, coords, numbers)


def main():
    coords, numbers, pseudo, grid, rho = water6()
    noise = 1.0 + EPS * np.random.default_rng(11).normal(size=rho.size)
    runs = {
        "mbis": lambda d: oracle.mbis(coords, numbers, pseudo, grid, d),
        "alisa_sc_gauss": lambda d: oracle.alisa(coords, numbers, pseudo, grid, d, basis_func="gauss"),
        "alisa_sc_slater": lambda d: oracle.alisa(coords, numbers, pseudo, grid, d, basis_func="slater"),
        "isa": lambda d: oracle.isa(coords, numbers, pseudo, grid, d, maxiter=60),
        "nlis": lambda d: oracle.nlis(coords, numbers, pseudo, grid, d),
    }
    for name, fn in runs.items():
        a, b = fn(rho), fn(rho * noise)
        pa, pb = np.asarray(a["propars"]), np.asarray(b["propars"])
        big = np.abs(pa) > 1e-6
        out = {
            "scheme": name, "niter": [int(a["niter"]), int(b["niter"])], "perturbation": EPS,
            "charges_abs": float(np.abs(a["charges"] - b["charges"]).max()),
            "charges_amplification": float(np.abs(a["charges"] - b["charges"]).max() / EPS),
            "propars_rel_max(|p|>1e-6)": float((np.abs(pa - pb)[big] / np.abs(pa)[big]).max()),
            "propars_amplification": float((np.abs(pa - pb)[big] / np.abs(pa)[big]).max() / EPS),
            "propars_abs_max": float(np.abs(pa - pb).max()),
        }
        if "history_changes" in a:
            ha, hb = np.asarray(a["history_changes"]), np.asarray(b["history_changes"])
            n = min(len(ha), len(hb))
            out["history_changes_rel_max"] = float((np.abs(ha[:n] - hb[:n]) / np.abs(ha[:n])).max())
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
