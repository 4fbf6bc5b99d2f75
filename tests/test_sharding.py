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


def test_work_balanced_sharding_of_the_screened_dense_pass():
    """The screened dense pass does less work for atoms at the surface of a cluster: balancing atom
    blocks by the geometric work estimate evens out the ranks (by points they differ by ~10 %)."""
    from horton_part_b200 import gridlite, synthetic
    from horton_part_b200.core.device import Shard
    from horton_part_b200.mbis import mbis_atom_work

    natom = 1200
    coords, numbers = synthetic.water_cluster(natom, 0)
    rgrid = gridlite.BeckeRTransform(1e-4, 1.5).transform_1d_grid(gridlite.GaussChebyshev(150))
    grid = gridlite.MolGrid.from_size(numbers, coords, 194, rgrid, np.ones(natom * 150 * 194), store=True)
    work = mbis_atom_work(coords, numbers, grid)
    npts = np.diff(grid.indices)
    assert work.shape == (natom,) and (work > 0).all() and (work <= natom * npts).all()
    assert work.min() < 0.8 * work.max()  # surface vs interior
    for world in (2, 4, 8):
        by_work = [Shard(natom, grid.indices, r, world, work=work) for r in range(world)]
        by_points = [Shard(natom, grid.indices, r, world) for r in range(world)]
        assert by_work[0].atom_lo == 0 and by_work[-1].atom_hi == natom
        for a, b in zip(by_work[:-1], by_work[1:]):
            assert a.atom_hi == b.atom_lo and a.point_hi == b.point_lo
        load = lambda shards: np.array([work[s.atom_lo : s.atom_hi].sum() for s in shards])  # noqa: E731
        lw, lp = load(by_work), load(by_points)
        assert lw.max() / lw.mean() < 1.02
        assert lw.max() / lw.mean() <= lp.max() / lp.mean() + 1e-12
    # one rank: the estimate is not needed and ignored
    assert Shard(natom, grid.indices, 0, 1, work=work).nlocal == natom


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
