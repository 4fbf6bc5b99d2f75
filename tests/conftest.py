"""Test configuration: markers, import paths and shared fixtures.

``-m "not gpu"`` covers the oracle against the golden vectors, host logic and the C-ABI symbol
table; ``-m gpu`` are the parity tests proper (CUDA path vs oracle / golden vectors).
Only tests may import from oracle/.
"""

import logging
import pathlib
import sys

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    logging.disable(logging.INFO)


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree C-ABI library (built on demand; nvcc cross-compiles without a GPU)."""
    from horton_part_b200 import build

    return build.build()


def _h2o_case(gridmod):
    z = np.load(GOLDEN / "h2o_hf_sto3g.npz")
    coords, numbers, pseudo = z["coordinates"], z["numbers"], z["pseudo_numbers"]
    rgrid = gridmod.ExpRTransform(5e-4, 2e1, 119).transform_1d_grid(gridmod.UniformInteger(120))
    grid = gridmod.MolGrid.from_size(numbers, coords, 110, rgrid, z["aim_weights"], store=True)
    return dict(coords=coords, numbers=numbers, pseudo=pseudo, grid=grid, rho=z["dens"], gold=z)


@pytest.fixture(scope="session")
def h2o():
    """Config 1: water HF/STO-3G on the reference's test grid, built with the product's gridlite."""
    from horton_part_b200 import gridlite

    return _h2o_case(gridlite)


def _water_case(gridmod, natom, nrad, nang, seed=0, gold=None):
    from horton_part_b200 import synthetic

    coords, numbers = synthetic.water_cluster(natom, seed)
    rgrid = gridmod.BeckeRTransform(1e-4, 1.5).transform_1d_grid(gridmod.GaussChebyshev(nrad))
    # the host Becke weights are O(natom^2 Npts) NumPy: beyond a few dozen atoms use the device kernel
    becke = gridmod.DeviceBeckeWeights() if natom > 24 and hasattr(gridmod, "DeviceBeckeWeights") else gridmod.BeckeWeights()
    grid = gridmod.MolGrid.from_size(numbers, coords, nang, rgrid, becke, store=True)
    rho = synthetic.slater_promolecule_host(grid.points, coords, numbers)
    return dict(coords=coords, numbers=numbers, pseudo=numbers.astype(float), grid=grid, rho=rho, gold=gold)


@pytest.fixture(scope="session")
def water6():
    from horton_part_b200 import gridlite

    return _water_case(gridlite, 6, 40, 50, gold=np.load(GOLDEN / "water6_slater.npz"))


@pytest.fixture(scope="session")
def make_water():
    from horton_part_b200 import gridlite

    return lambda natom, nrad=40, nang=50, seed=0: _water_case(gridlite, natom, nrad, nang, seed)


@pytest.fixture(scope="session")
def water6g():
    """Synthetic Gaussian promolecule (gauss table initials scaled to 8.6 / 0.7 electrons)."""
    from horton_part_b200 import gridlite, synthetic
    from horton_part_b200.core.basis import ExpBasisFuncHelper

    case = _water_case(gridlite, 6, 40, 50, gold=np.load(GOLDEN / "water6_gauss.npz"))
    helper = ExpBasisFuncHelper.from_function_type("gauss")
    case["rho"] = synthetic.expbasis_promolecule_host(case["grid"].points, case["coords"], case["numbers"],
                                                      helper, scale={8: 8.6, 1: 0.7})
    return case


@pytest.fixture(scope="session")
def h2o_proatomdb():
    """The reference's HF/STO-3G isolated-atom records (packed by oracle/gen_golden.py) as a
    product-side ProAtomDB on PowerRTransform radial grids (tests/common.py:64-111)."""
    from horton_part_b200 import gridlite
    from horton_part_b200.core.proatomdb import ProAtomDB, ProAtomRecord

    z = np.load(GOLDEN / "h2o_hirshfeld.npz")
    records = []
    for key in z.files:
        if not key.startswith("record/"):
            continue
        v = z[key]
        number, charge, energy, rmin, rmax, npoint = int(v[0]), int(v[1]), float(v[2]), v[3], v[4], int(v[5])
        rgrid = gridlite.PowerRTransform(rmin, rmax, npoint - 1).transform_1d_grid(gridlite.UniformInteger(npoint))
        records.append(ProAtomRecord(number, charge, energy, rgrid, v[6 : 6 + npoint].copy(), v[6 + npoint :].copy()))
    return ProAtomDB(records), z
