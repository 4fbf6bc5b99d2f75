"""Becke weights on the device against the host implementation and against
the oracle's restatement of qc-grid's BeckeWeights."""

import sys

import numpy as np
import pytest
from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_device_becke_matches_host_and_oracle(make_water):
    from horton_part_b200 import gridlite, synthetic

    sys.path.insert(0, str(ROOT / "oracle" / "qcgrid_shim"))
    try:
        import grid as qcgrid  # the oracle's qc-grid restatement
    finally:
        sys.path.pop(0)

    for coords, numbers in (synthetic.water_cluster(12, seed=5), synthetic.organic_like(20, seed=1)):
        rgrid = gridlite.BeckeRTransform(1e-4, 1.5).transform_1d_grid(gridlite.GaussChebyshev(30))
        host = gridlite.MolGrid.from_size(numbers, coords, 38, rgrid, gridlite.BeckeWeights(), store=True)
        dev = gridlite.MolGrid.from_size(numbers, coords, 38, rgrid, gridlite.DeviceBeckeWeights(), store=True)
        np.testing.assert_allclose(dev.aim_weights, host.aim_weights, rtol=1e-11, atol=1e-15)
        ref = qcgrid.BeckeWeights()(host.points, coords, numbers, host.indices)
        np.testing.assert_allclose(dev.aim_weights, ref, rtol=1e-11, atol=1e-15)
        assert dev.aim_weights.min() >= 0.0 and dev.aim_weights.max() <= 1.0


def test_partition_of_unity(make_water):
    """Sum over atoms of the cell functions is 1: check through per-atom owner permutations."""
    import torch

    from horton_part_b200 import gridlite, synthetic

    coords, numbers = synthetic.water_cluster(6, seed=2)
    rng = np.random.default_rng(0)
    pts = coords[rng.integers(0, 6, 4000)] + rng.normal(scale=2.0, size=(4000, 3))
    total = np.zeros(len(pts))
    bw = gridlite.DeviceBeckeWeights()
    for a in range(6):
        ind = np.zeros(7, dtype=np.int64)
        ind[a + 1 :] = len(pts)  # every point owned by atom a
        total += bw(pts, coords, numbers, ind)
    np.testing.assert_allclose(total, 1.0, rtol=0, atol=1e-13)
