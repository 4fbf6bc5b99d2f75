"""Host <-> device transfer helpers: pipelined uploads from pageable NumPy arrays, direct uploads
from page-locked ones, pooled page-locked result arrays that are never re-issued while in use."""

import gc

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_upload_pageable_and_pinned_are_exact():
    import torch

    from horton_part_b200.core import hostmem

    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(0)
    n = (40 << 20) // 8 + 12345  # below one staging half (64 MB); the last size spans several chunks
    for size in (1000, n, 3 * (64 << 20) // 8 + 7):
        a = rng.random(size)
        t = hostmem.upload(a, dev)
        torch.cuda.synchronize()
        assert t.dtype == torch.float64 and t.shape == (size,)
        assert np.array_equal(t.cpu().numpy(), a)
        p = hostmem.pinned_empty(size)
        assert hostmem.is_pinned(p) and not hostmem.is_pinned(a)
        p[:] = a
        t2 = hostmem.upload(p, dev)
        torch.cuda.synchronize()
        assert np.array_equal(t2.cpu().numpy(), a)
    # dtype conversion and non-contiguous input
    m = rng.integers(0, 100, size=(2000, 3000)).astype(np.int32)
    t = hostmem.upload(m[:, ::2], dev, np.float64)
    assert np.array_equal(t.cpu().numpy(), m[:, ::2].astype(np.float64))
    stats = hostmem.pool_stats()
    assert stats["staged_uploads"] >= 2 and stats["direct_uploads"] >= 2


def test_download_pool_never_reissues_a_buffer_in_use():
    import torch

    from horton_part_b200.core import hostmem

    dev = torch.device("cuda", 0)
    n = (16 << 20) // 8
    a = torch.arange(n, dtype=torch.float64, device=dev)
    b = -torch.arange(n, dtype=torch.float64, device=dev)
    ha = hostmem.download(a)
    view = ha[100:200]  # a view keeps the whole buffer busy
    hb = hostmem.download(b)
    assert ha.ctypes.data != hb.ctypes.data
    assert ha[12345] == 12345.0 and hb[12345] == -12345.0
    addr_a = ha.ctypes.data
    del ha
    gc.collect()
    hc = hostmem.download(b)  # the view is still alive: buffer a must not be reused
    assert hc.ctypes.data != addr_a
    assert view[5] == 105.0
    del view
    gc.collect()
    before = hostmem.pool_stats()["pinned_reuses"]
    hd = hostmem.download(a)  # now it may be
    assert hostmem.pool_stats()["pinned_reuses"] == before + 1
    assert hd.ctypes.data == addr_a and hd[7] == 7.0
    assert hb[3] == -3.0 and hc[3] == -3.0  # untouched
    assert hostmem.is_pinned(hd)
    # small tensors take the plain path
    small = hostmem.download(a[:10])
    assert np.array_equal(small, np.arange(10.0))


def test_results_survive_a_second_job(make_water):
    """Arrays published by one partitioning job stay intact when another job runs afterwards."""
    from horton_part_b200 import MBISWPart

    case = make_water(12, nrad=30, nang=38, seed=3)
    args = (case["coords"], case["numbers"], case["pseudo"], case["grid"], case["rho"])
    first = MBISWPart(*args, maxiter=3)
    first.do_partitioning()
    promol = first["promoldens"]
    w0 = first["at_weights_0"]
    keep_p, keep_w = promol.copy(), w0.copy()
    second = MBISWPart(*args, maxiter=5)
    second.do_partitioning()
    assert np.array_equal(promol, keep_p) and np.array_equal(w0, keep_w)
    assert not np.array_equal(second["promoldens"], promol)


@pytest.mark.parametrize("pinned", [False, True])
def test_split_upload_gives_bit_identical_results(pinned, monkeypatch):
    """A slab uploaded in two parts (second part under the first pass's kernel) gives the same bits as the
    one-piece upload: charges, parameters, entropies, pair counters -- from pageable and page-locked arrays."""
    from horton_part_b200 import MBISWPart, gridlite, synthetic
    from horton_part_b200.core import hostmem
    from horton_part_b200.core.device import GridSlab

    coords, numbers = synthetic.water_cluster(30, seed=3)
    rgrid = gridlite.BeckeRTransform(1e-4, 1.5).transform_1d_grid(gridlite.GaussChebyshev(40))
    grid = gridlite.MolGrid.from_size(numbers, coords, 50, rgrid, np.ones(len(numbers) * 40 * 50), store=True)
    rho, w, _, _ = synthetic.slater_promolecule_device(grid, coords, numbers, device="cuda:0")
    grid.aim_weights[:] = w
    grid.weights[:] = grid.atweights * w
    if pinned:
        for name in ("points", "weights"):
            pin = hostmem.pinned_empty(getattr(grid, name).shape)
            pin[...] = getattr(grid, name)
            setattr(grid, name, pin)
        pin = hostmem.pinned_empty(rho.shape)
        pin[...] = rho
        rho = pin
    monkeypatch.setattr(GridSlab, "split_upload_min_bytes", 1 << 20)

    def run(split):
        monkeypatch.setenv("HP_B200_SPLIT_UPLOAD", "1" if split else "0")
        part = MBISWPart(coords, numbers, numbers.astype(float), grid, rho, maxiter=6)
        part.do_partitioning()
        assert bool(part.slab._split_atoms) == split and part.slab._pending_upload is None
        return part

    one, two = run(False), run(True)
    assert one["niter"] == two["niter"] == 6
    for key in ("charges", "propars", "history_entropies", "history_changes", "promoldens"):
        assert np.array_equal(np.asarray(one[key]), np.asarray(two[key])), key
    for a in range(len(numbers)):
        assert np.array_equal(one[f"at_weights_{a}"], two[f"at_weights_{a}"])


def test_split_upload_is_completed_by_unaware_consumers(monkeypatch):
    """Reading a point array of a half-uploaded slab completes the upload first."""
    import torch

    from horton_part_b200 import gridlite, synthetic
    from horton_part_b200.core.device import GridSlab

    coords, numbers = synthetic.water_cluster(30, seed=3)
    rgrid = gridlite.BeckeRTransform(1e-4, 1.5).transform_1d_grid(gridlite.GaussChebyshev(40))
    grid = gridlite.MolGrid.from_size(numbers, coords, 50, rgrid, np.ones(len(numbers) * 40 * 50), store=True)
    rho = np.random.default_rng(1).random(grid.size)
    monkeypatch.setattr(GridSlab, "split_upload_min_bytes", 1 << 20)
    slab = GridSlab(grid, rho, coords, device="cuda:0")
    assert slab._pending_upload is not None and slab._split_atoms > 0
    assert np.array_equal(slab.rho.cpu().numpy(), rho)
    assert slab._pending_upload is None
    torch.cuda.synchronize()
    assert np.array_equal(slab.px.cpu().numpy(), grid.points[:, 0])
    assert np.array_equal(slab.pz.cpu().numpy(), grid.points[:, 2])
    assert np.array_equal(slab.molw.cpu().numpy(), grid.weights)
    assert np.array_equal(slab.atw.cpu().numpy(), np.concatenate([g.weights for g in grid.atgrids]))


def test_pool_keeps_idle_buffers_of_other_sizes_within_its_budget(monkeypatch):
    """Alternating result sizes re-use their own buffers (page-locked memory is expensive to free and to
    allocate); idle buffers are dropped only when the pool would exceed HP_B200_PINNED_POOL_BYTES."""
    import torch

    from horton_part_b200.core import hostmem

    dev = torch.device("cuda", 0)
    small = torch.arange((16 << 20) // 8, dtype=torch.float64, device=dev)
    large = torch.arange((40 << 20) // 8, dtype=torch.float64, device=dev)

    def cycle():
        for t in (small, large, small, large):
            a = hostmem.download(t)
            assert a[-1] == t.numel() - 1
            del a
            gc.collect()

    cycle()  # both sizes are in the pool now
    before = hostmem.pool_stats()
    cycle()
    after = hostmem.pool_stats()
    assert after["pinned_allocs"] == before["pinned_allocs"] and after["pinned_reuses"] == before["pinned_reuses"] + 4
    monkeypatch.setenv("HP_B200_PINNED_POOL_BYTES", "1")  # nothing idle may stay when a new buffer is needed
    odd = torch.zeros((100 << 20) // 8 + 1, dtype=torch.float64, device=dev)  # fits neither of the two
    a = hostmem.download(odd)
    stats = hostmem.pool_stats()
    assert stats["pool_buffers"] == 1 and stats["pinned_allocs"] == after["pinned_allocs"] + 1
    del a
