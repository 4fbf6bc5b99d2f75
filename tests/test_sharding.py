"""Host-side logic of the multi-GPU path on CPU: atom-block sharding and the state-vector exchange
(world_size 2 over gloo).  Per-rank arithmetic is done by the ORACLE here (tests may use it); what is
under test is the product's Shard partition and IterationState gather protocol."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_partition_covers_all_atoms():
    from horton_part_b200.core.device import Shard

    rng = np.random.default_rng(0)
    for natom in (1, 2, 3, 7, 96, 2000):
        sizes = rng.integers(1, 5, natom) * 1000 if natom < 2000 else np.full(natom, 29100)
        off = np.concatenate([[0], np.cumsum(sizes)])
        for world in (1, 2, 4, 8):
            shards = [Shard(natom, off, r, world) for r in range(world)]
            assert shards[0].atom_lo == 0 and shards[-1].atom_hi == natom
            for a, b in zip(shards[:-1], shards[1:]):
                assert a.atom_hi == b.atom_lo and a.point_hi == b.point_lo
            assert sum(s.nlocal for s in shards) == natom
            if natom == 2000:  # dense mode balance: equal atom blocks
                assert max(s.nlocal for s in shards) - min(s.nlocal for s in shards) <= 1


def test_work_balanced_shard_boundaries():
    """Shard(work=...) cuts the atom list at equal cumulative work; without weights at equal points."""
    from horton_part_b200.core.device import Shard

    natom = 1000
    off = np.arange(natom + 1) * 29100
    rng = np.random.default_rng(2)
    work = 1.0 + np.sin(np.linspace(0, np.pi, natom)) + 0.05 * rng.random(natom)  # light ends, heavy middle
    for world in (2, 4, 8):
        by_work = [Shard(natom, off, r, world, work=work) for r in range(world)]
        by_points = [Shard(natom, off, r, world) for r in range(world)]
        assert by_work[0].atom_lo == 0 and by_work[-1].atom_hi == natom
        for a, b in zip(by_work[:-1], by_work[1:]):
            assert a.atom_hi == b.atom_lo and a.point_hi == b.point_lo
        load = lambda shards: np.array([work[s.atom_lo : s.atom_hi].sum() for s in shards])  # noqa: E731
        lw, lp = load(by_work), load(by_points)
        assert lw.max() / lw.mean() < 1.02
        if world > 2:  # (two ranks: the profile is symmetric, both splits are even)
            assert lw.max() / lw.mean() < lp.max() / lp.mean()
            assert by_work[0].nlocal > by_points[0].nlocal  # the light end gets more atoms
    assert Shard(natom, off, 0, 1, work=work).nlocal == natom  # one rank: weights ignored


def test_proatom_reach_solver():
    """Radius at which a pro-atom bound has decayed to a target (used by the work estimate)."""
    from horton_part_b200.core.device import _reach

    A, alpha = np.array([326.0, 1.9]), np.array([16.0, 2.0])
    targets = np.array([1e3, 1.0, 1e-10, 1e-25])
    for gaussian in (False, True):
        d = _reach(A, alpha, targets, gaussian)
        x = d * d if gaussian else d
        val = (A[None, :] * np.exp(-alpha[None, :] * x[:, None])).sum(1)
        assert d[0] == 0.0  # already below the target at the nucleus
        np.testing.assert_allclose(val[1:], targets[1:], rtol=1e-9)
        assert (np.diff(d) > 0).all()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, out):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "oracle")]
    import stockholder_oracle as oracle

    from horton_part_b200.core.device import Shard
    from horton_part_b200.core.iterstock import IterationState

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    coords, numbers, pseudo, grid, rho = case
    natom = len(numbers)
    ranges = [0]
    for z in numbers:
        ranges.append(ranges[-1] + 2 * oracle.mbis_nshell(int(z)))
    shard = Shard(natom, grid.indices, rank, world)
    st = IterationState(natom, ranges[-1], "cpu")
    st.propars[:] = torch.from_numpy(np.concatenate([oracle.mbis_initial(int(z)) for z in numbers]))
    dist_a = [oracle.distances(grid.points, c) for c in coords]
    changes = []
    for _ in range(4):
        propars = st.propars.numpy().copy()
        # every rank evaluates the promolecule on ITS OWN points only (all atoms contribute)
        lo, hi = shard.point_lo, shard.point_hi
        promol = np.zeros(hi - lo)
        own = {}
        for a in range(natom):
            work = oracle.mbis_proatom(propars[ranges[a] : ranges[a + 1]], dist_a[a][lo:hi])
            promol += work
            promol += 1e-100
            if shard.atom_lo <= a < shard.atom_hi:
                own[a] = work[grid.indices[a] - lo : grid.indices[a + 1] - lo]
        st.begin_sharded_update(ranges[shard.atom_lo], ranges[shard.atom_hi])
        st.entropy[0] = oracle.entropy(grid.weights[lo:hi], rho[lo:hi], promol)
        for a in range(shard.atom_lo, shard.atom_hi):
            g = grid.atgrids[a]
            w = np.clip(own[a] / promol[grid.indices[a] - lo : grid.indices[a + 1] - lo], 0, 1)
            sph = oracle.shell_average(g, w * rho[grid.indices[a] : grid.indices[a + 1]])
            r = g.rgrid.points
            w4 = 4 * np.pi * r**2 * g.rgrid.weights
            old = propars[ranges[a] : ranges[a + 1]]
            new, _ = oracle.mbis_inner(sph, old.copy(), w4, r, 1e-8)
            st.propars[ranges[a] : ranges[a + 1]] = torch.from_numpy(new)
            st.charges[a] = pseudo[a] - np.einsum("p,p->", w4, sph)
            d = oracle.mbis_proatom(new, r) - oracle.mbis_proatom(old, r)
            st.msd[a] = np.einsum("i,i,i", w4, d, d)
        st.gather(dist.group.WORLD)
        changes.append(float(np.sqrt(st.msd.numpy().sum())))
    out[rank] = (st.vec.numpy().copy(), changes)
    dist.destroy_process_group()


def test_state_exchange_world2_matches_single_process(make_water):
    import stockholder_oracle as oracle

    case = make_water(6, nrad=24, nang=26, seed=1)
    ref = oracle.mbis(case["coords"], case["numbers"], case["pseudo"], case["grid"], case["rho"], maxiter=4)
    payload = (case["coords"], case["numbers"], case["pseudo"], case["grid"], case["rho"])
    port = _free_port()
    manager = mp.Manager()
    out = manager.dict()
    mp.spawn(_worker, args=(2, port, payload, out), nprocs=2, join=True)
    natom = 6
    for rank in (0, 1):
        vec, changes = out[rank]
        np.testing.assert_allclose(vec[1 + natom : 1 + 2 * natom], ref["charges"], rtol=0, atol=1e-13)
        np.testing.assert_allclose(vec[1 + 2 * natom :], ref["propars"], rtol=1e-12)
        np.testing.assert_allclose(changes, ref["history_changes"], rtol=1e-10)
        np.testing.assert_allclose(vec[0], ref["history_entropies"][-1], rtol=1e-12)
    assert np.array_equal(out[0][0], out[1][0])  # bit-identical state on both ranks
