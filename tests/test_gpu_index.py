"""Row L: the cut-off local index must be bit-exact against cKDTree.query_ball_point (integer
result), including adversarial tie radii that sit exactly on Lebedev shells."""

import numpy as np
import pytest
import stockholder_oracle as oracle

pytestmark = pytest.mark.gpu


def _gpu_index(points, center, radius, begin, end):
    import torch

    from horton_part_b200 import _lib

    dev = torch.device("cuda:0")
    n = len(points)
    pts = torch.from_numpy(np.ascontiguousarray(points)).to(dev)
    idx = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    ovl = torch.empty(max(n, 1), dtype=torch.uint8, device=dev)
    dist = torch.empty(max(n, 1), dtype=torch.float64, device=dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    nbytes = _lib.call("hp_local_index_scratch_bytes", n)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    _lib.call("hp_build_local_index", pts, n, np.ascontiguousarray(center, dtype=np.float64), float(radius),
              int(begin), int(end), idx, ovl, dist, cnt, scratch, nbytes, torch.cuda.current_stream(dev).cuda_stream)
    c = int(cnt.item())
    return idx[:c].cpu().numpy(), ovl[:c].cpu().numpy().astype(bool), dist[:c].cpu().numpy()


def test_bit_exact_against_kdtree(h2o):
    grid = h2o["grid"]
    pts = grid.points
    rng = np.random.default_rng(7)
    for a in range(3):
        center = h2o["coords"][a]
        begin, end = grid.indices[a], grid.indices[a + 1]
        shell_r = grid.atgrids[a].rgrid.points
        radii = list(shell_r[[5, 40, 80, 100, 119]]) + list(rng.uniform(0.01, 12.0, 6)) + [0.0, 1e3]
        # tie radii: exact distances of existing points
        d = np.linalg.norm(pts - center, axis=1)
        radii += list(d[rng.integers(0, len(d), 6)])
        for radius in radii:
            ridx, rovl, rdist = oracle.local_index(pts, center, radius, begin, end)
            gidx, govl, gdist = _gpu_index(pts, center, radius, begin, end)
            assert np.array_equal(gidx, ridx), (a, radius, len(gidx), len(ridx))
            assert np.array_equal(govl, rovl)
            assert np.array_equal(gdist, rdist)  # bit-exact distances


def test_empty_and_ragged():
    pts = np.zeros((0, 3))
    gidx, _, _ = _gpu_index(pts, np.zeros(3), 1.0, 0, 0)
    assert gidx.size == 0
    rng = np.random.default_rng(0)
    for n in (1, 31, 2047, 2049, 100003):
        pts = rng.normal(size=(n, 3))
        ridx, rovl, rdist = oracle.local_index(pts, np.array([0.1, -0.2, 0.3]), 1.1, n // 3, n // 2)
        gidx, govl, gdist = _gpu_index(pts, np.array([0.1, -0.2, 0.3]), 1.1, n // 3, n // 2)
        assert np.array_equal(gidx, ridx) and np.array_equal(govl, rovl) and np.array_equal(gdist, rdist)
