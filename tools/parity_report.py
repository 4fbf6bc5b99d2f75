#!/usr/bin/env python
"""Measured deviations of the GPU path from the reference's own runs (tests/golden/*.npz), case by case:
what the tolerances in tests/ have to absorb.  One line per case:
niter (gpu/ref), max |dq|, max relative deviation of the parameters above 1e-6, max absolute deviation of
all parameters, max relative deviation of the `change` history.  Run on the GPU box; the table is committed
under profiles/."""
from __future__ import annotations

import logging
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
logging.disable(logging.INFO)

import conftest  # noqa: E402  (fixture builders)

import horton_part_b200 as hp  # noqa: E402
from horton_part_b200 import gridlite, synthetic  # noqa: E402
from horton_part_b200.core.basis import ExpBasisFuncHelper  # noqa: E402

GOLD = conftest.GOLDEN


def gold(z, tag):
    return {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(tag + "/")}


def report(label, part, ref):
    q, qr = part["charges"], ref["charges"]
    line = f"{label:44s} niter {int(part['niter']):4d}/{int(ref['niter']) if 'niter' in ref else -1:4d}  max|dq| {np.abs(q - qr).max():8.1e}"
    if "propars" in ref:
        p, pr = np.asarray(part["propars"]), np.asarray(ref["propars"])
        big = np.abs(pr) > 1e-6
        line += f"  propars rel(|p|>1e-6) {(np.abs(p - pr)[big] / np.abs(pr)[big]).max():8.1e}  abs {np.abs(p - pr).max():8.1e}"
    if "history_changes" in ref:
        h, hr = np.asarray(part["history_changes"]), np.asarray(ref["history_changes"])
        n = min(len(h), len(hr))
        line += f"  changes rel {(np.abs(h[:n] - hr[:n]) / np.abs(hr[:n])).max():8.1e}"
    if "history_entropies" in ref:
        h, hr = np.asarray(part["history_entropies"]), np.asarray(ref["history_entropies"])
        n = min(len(h), len(hr))
        line += f"  entropy abs {np.abs(h[:n] - hr[:n]).max():8.1e}"
    print(line, flush=True)


def run(cls, case, **kw):
    part = getattr(hp, cls)(case["coords"], case["numbers"], case["pseudo"], case["grid"], case["rho"], **kw)
    part.do_partitioning()
    return part


def main():
    h2o = conftest._h2o_case(gridlite)
    w6 = conftest._water_case(gridlite, 6, 40, 50, gold=np.load(GOLD / "water6_slater.npz"))
    w6g = conftest._water_case(gridlite, 6, 40, 50, gold=np.load(GOLD / "water6_gauss.npz"))
    helper = ExpBasisFuncHelper.from_function_type("gauss")
    w6g["rho"] = synthetic.expbasis_promolecule_host(w6g["grid"].points, w6g["coords"], w6g["numbers"], helper,
                                                     scale={8: 8.6, 1: 0.7})
    print("== water HF/STO-3G, reference's test grid (config 1)")
    for tag, cls, kw in (("mbis", "MBISWPart", {}), ("isa", "ISAWPart", {}), ("lisa_sc_gauss", "LinearISAWPart", dict(solver="sc")),
                         ("lisa_sc_slater", "LinearISAWPart", dict(solver="sc", basis_func="slater")),
                         ("nlis", "NLISWPart", dict(exp_n_dict={})), ("gmbis", "GMBISWPart", dict(exp_n_dict={})),
                         ("glisa_sc", "GlobalLinearISAWPart", dict(solver="sc")),
                         ("mbis_gt2", "MBISWPart", dict(grid_type=2)), ("lisa_sc_gt2", "LinearISAWPart", dict(solver="sc", grid_type=2))):  # fmt: skip
        report(f"h2o {tag}", run(cls, h2o, **kw), gold(h2o["gold"], tag))
    print("== 6-atom water cluster, Slater promolecule, 40 x 50 grid")
    for tag, cls, kw in (("mbis", "MBISWPart", {}), ("isa", "ISAWPart", dict(maxiter=60)), ("lisa_sc_gauss", "LinearISAWPart", dict(solver="sc")),
                         ("nlis", "NLISWPart", dict(exp_n_dict={})), ("glisa_sc", "GlobalLinearISAWPart", dict(solver="sc"))):  # fmt: skip
        report(f"water6 {tag}", run(cls, w6, **kw), gold(w6["gold"], tag))
    print("== 6-atom water cluster, Gaussian promolecule")
    for tag, cls, kw in (("glisa_newton", "GlobalLinearISAWPart", dict(solver="newton")), ("glisa_sc", "GlobalLinearISAWPart", dict(solver="sc")),
                         ("lisa_sc_gauss", "LinearISAWPart", dict(solver="sc"))):  # fmt: skip
        report(f"water6g {tag}", run(cls, w6g, **kw), gold(w6g["gold"], tag))
    # real-grid configurations (tests/test_gpu_configs.py)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_gpu_configs as tc

    def full_grid(coords, numbers):
        return tc._grid(coords, numbers)

    if (GOLD / "config2_organic20.npz").exists():
        print("== config 2: 20 atoms, 582,000 points (full size)")
        z = np.load(GOLD / "config2_organic20.npz")
        coords, numbers = z["coordinates"], z["numbers"]
        grid = full_grid(coords, numbers)
        case = dict(coords=coords, numbers=numbers, pseudo=numbers.astype(float), grid=grid,
                    rho=synthetic.slater_promolecule_host(grid.points, coords, numbers))
        report("config2 mbis", run("MBISWPart", case), gold(z, "mbis"))
        report("config2 isa", run("ISAWPart", case), gold(z, "isa"))
    z = np.load(GOLD / "config3_water24.npz")
    print("== config 3 reduced: 24 atoms, 698,400 points")
    coords, numbers = z["coordinates"], z["numbers"]
    grid = full_grid(coords, numbers)
    case = dict(coords=coords, numbers=numbers, pseudo=numbers.astype(float), grid=grid,
                rho=synthetic.expbasis_promolecule_host(grid.points, coords, numbers, helper, scale={8: 8.6, 1: 0.7}))
    for basis in ("gauss", "slater"):
        report(f"config3 lisa sc {basis}", run("LinearISAWPart", case, solver="sc", basis_func=basis), gold(z, f"lisa_sc_{basis}"))
    z = np.load(GOLD / "config4_peptide12.npz")
    print("== config 4 reduced: 12 atoms, 349,200 points")
    coords, numbers = z["coordinates"], z["numbers"]
    grid = full_grid(coords, numbers)
    case = dict(coords=coords, numbers=numbers, pseudo=numbers.astype(float), grid=grid,
                rho=synthetic.expbasis_promolecule_host(grid.points, coords, numbers, helper,
                                                        scale={1: 0.75, 6: 6.2, 7: 7.3, 8: 8.4}))
    report("config4 glisa newton", run("GlobalLinearISAWPart", case, solver="newton"), gold(z, "glisa_newton"))
    report("config4 glisa sc", run("GlobalLinearISAWPart", case, solver="sc", maxiter=60), gold(z, "glisa_sc"))


if __name__ == "__main__":
    main()
