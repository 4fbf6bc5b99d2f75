"""Two more molecules of the reference's own test suite, against runs of the unmodified reference
(tests/golden/ref_molecules.npz, oracle/gen_golden.py::case_molecules) and against the expected
values the reference's tests carry: N2 (tests/test_becke.py:31-57) and monosilicic acid with LANL
effective core potentials, where pseudo_numbers != numbers (tests/test_wpart.py:104-217)."""

import contextlib
import io

import numpy as np
import pytest
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

GOLD = np.load(GOLDEN / "ref_molecules.npz")


def _case(name, rgrid_args, npoint):
    from horton_part_b200 import gridlite

    coords, numbers, pseudo = GOLD[f"{name}/coordinates"], GOLD[f"{name}/numbers"], GOLD[f"{name}/pseudo_numbers"]
    rgrid = gridlite.ExpRTransform(*rgrid_args).transform_1d_grid(gridlite.UniformInteger(npoint))
    grid = gridlite.MolGrid.from_size(numbers, coords, 110, rgrid, gridlite.BeckeWeights(), store=True)
    np.testing.assert_allclose(grid.aim_weights[::211], GOLD[f"{name}/aim_weights_sample"], rtol=1e-10, atol=1e-14)
    return coords, numbers, pseudo, grid, GOLD[f"{name}/dens"]


@pytest.fixture(scope="module")
def n2():
    return _case("n2", (1e-3, 1e1, 99), 100)


@pytest.fixture(scope="module")
def msa():
    return _case("msa", (5e-4, 2e1, 119), 120)


def _run(case, scheme, **kw):
    from horton_part_b200 import wpart_schemes

    coords, numbers, pseudo, grid, rho = case
    with contextlib.redirect_stdout(io.StringIO()):
        part = wpart_schemes(scheme)(coords, numbers, pseudo, grid, rho, **kw)
        part.do_charges()
    return part


def _compare(part, tag, rtol=1e-8, atol=1e-9):
    if f"{tag}/niter" in GOLD.files:
        assert part["niter"] == int(GOLD[f"{tag}/niter"])
    # north_star tolerance: 1e-8 relative
    np.testing.assert_allclose(part["charges"], GOLD[f"{tag}/charges"], rtol=rtol, atol=atol)
    np.testing.assert_allclose(part["populations"], GOLD[f"{tag}/populations"], rtol=max(rtol, 1e-9), atol=atol)


def test_n2_becke_is_the_references_own_test(n2):
    part = _run(n2, "b")
    assert abs(part["populations"] - 7).max() < 1e-4  # tests/test_becke.py:49
    assert abs(part["charges"]).max() < 1e-4
    _compare(part, "n2/becke")


@pytest.mark.parametrize("tag,scheme", [("mbis", "mbis"), ("isa", "is")])
def test_n2_iterative_schemes(n2, tag, scheme):
    _compare(_run(n2, scheme), f"n2/{tag}")


def _lan_database():
    from horton_part_b200 import gridlite
    from horton_part_b200.core.proatomdb import ProAtomDB, ProAtomRecord

    records = []
    for key in GOLD.files:
        if not key.startswith("msa/record/"):
            continue
        v = GOLD[key]
        number, charge, energy, rmin, rmax, npoint, pn = int(v[0]), int(v[1]), float(v[2]), v[3], v[4], int(v[5]), float(v[6])
        rgrid = gridlite.PowerRTransform(rmin, rmax, npoint - 1).transform_1d_grid(gridlite.UniformInteger(npoint))
        records.append(ProAtomRecord(number, charge, energy, rgrid, v[7 : 7 + npoint].copy(), v[7 + npoint :].copy(),
                                     pseudo_number=pn))  # fmt: skip
    return ProAtomDB(records)


# expected charges quoted by the reference's tests (from HiPart), tolerance 4e-3 there
MSA_EXPECTED = {
    "h": [0.56175431, -0.30002709, -0.28602105, -0.28335086, -0.26832878, 0.13681904, 0.14535691, 0.14206876, 0.15097682],
    "hi": [1.14305602, -0.52958298, -0.51787452, -0.51302759, -0.50033981, 0.21958586, 0.23189187, 0.22657354, 0.23938904],
    "isa": [1.1721364, -0.5799622, -0.5654549, -0.5599638, -0.5444145, 0.2606699, 0.2721848, 0.2664377, 0.2783666],
}


@pytest.mark.parametrize("tag,scheme", [("h", "h"), ("hi", "hi"), ("isa", "is"), ("mbis", "mbis")])
def test_monosilicic_acid_with_effective_core_potentials(msa, tag, scheme):
    kw = dict(proatomdb=_lan_database()) if scheme in ("h", "hi") else {}
    part = _run(msa, scheme, **kw)
    # ISA needs 180 iterations here and is ill-conditioned: the reference's own answer moves by
    # 4.7e-7 when its input density is perturbed by 1e-15 relative (NumPy restatement, which otherwise
    # reproduces the reference run to 2e-15; measured with oracle/stockholder_oracle.isa).  The
    # iteration count is still identical; the charges are compared at that sensitivity.
    tol = dict(rtol=5e-6, atol=5e-6) if tag == "isa" else {}
    _compare(part, f"msa/{tag}", **tol)
    np.testing.assert_allclose(part["pseudo_populations"], GOLD[f"msa/{tag}/pseudo_populations"],
                               rtol=tol.get("rtol", 1e-9), atol=tol.get("atol", 0))  # fmt: skip
    if tag in MSA_EXPECTED:
        assert abs(part["charges"] - np.array(MSA_EXPECTED[tag])).max() < 4e-3  # tests/test_wpart.py:127
    assert (msa[2] != msa[1]).any()  # pseudo numbers really differ from the atomic numbers
