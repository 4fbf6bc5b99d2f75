"""Sharded (multi-GPU) MBIS equals the single-GPU run: identical iteration count, charges to 1e-12.
Needs >= 2 GPUs (run with `gpurun --gpus 2`); skipped on a single-GPU box."""

import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, payload, out):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import logging

    import torch
    import torch.distributed as dist

    logging.disable(logging.INFO)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from horton_part_b200 import ISAWPart, LinearISAWPart, MBISWPart

    coords, numbers, pseudo, grid, rho = payload
    res = {}
    for name, cls, kw in (("mbis", MBISWPart, {}), ("lisa", LinearISAWPart, dict(solver="sc", maxiter=15)),
                          ("isa", ISAWPart, dict(maxiter=10))):
        part = cls(coords, numbers, pseudo, grid, rho, device=torch.device("cuda", rank), comm=dist.group.WORLD, **kw)
        part.do_charges()
        res[name] = (int(part["niter"]), part["charges"].copy(), part["propars"].copy(),
                     np.array(part["history_entropies"]), np.array(part["history_changes"]))
    out[rank] = res
    dist.destroy_process_group()


def test_sharded_equals_single_gpu(make_water):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from horton_part_b200 import ISAWPart, LinearISAWPart, MBISWPart

    case = make_water(9, nrad=30, nang=38, seed=2)
    payload = (case["coords"], case["numbers"], case["pseudo"], case["grid"], case["rho"])
    manager = mp.Manager()
    out = manager.dict()
    mp.spawn(_worker, args=(2, _free_port(), payload, out), nprocs=2, join=True)
    for name, cls, kw in (("mbis", MBISWPart, {}), ("lisa", LinearISAWPart, dict(solver="sc", maxiter=15)),
                          ("isa", ISAWPart, dict(maxiter=10))):
        single = cls(*payload, **kw)
        single.do_charges()
        for rank in (0, 1):
            niter, charges, propars, ent, chg = out[rank][name]
            assert niter == single["niter"], name
            np.testing.assert_allclose(charges, single["charges"], rtol=0, atol=1e-12, err_msg=name)
            np.testing.assert_allclose(propars, single["propars"], rtol=1e-11, atol=1e-14, err_msg=name)
            np.testing.assert_allclose(ent, single["history_entropies"], rtol=1e-12, atol=1e-14, err_msg=name)
            np.testing.assert_allclose(chg, single["history_changes"], rtol=1e-9, err_msg=name)
        assert np.array_equal(out[0][name][1], out[1][name][1])  # ranks agree bit for bit


def _worker_hpcomm(rank, world, idfile, payload, out):
    """Same run with the communicator of the C ABI (hp_comm_*): no torch.distributed anywhere; the NCCL id
    travels through a file, as a launcher without any Python messaging layer would do it."""
    import sys
    import time

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import logging

    import torch

    logging.disable(logging.INFO)
    torch.cuda.set_device(rank)
    from horton_part_b200 import MBISWPart
    from horton_part_b200.core.comm import HpComm

    if rank == 0:
        with open(idfile + ".tmp", "wb") as fh:
            fh.write(HpComm.unique_id())
        os.replace(idfile + ".tmp", idfile)
    else:
        for _ in range(600):
            if os.path.exists(idfile):
                break
            time.sleep(0.05)
    uid = open(idfile, "rb").read()
    comm = HpComm(world, rank, uid, device=torch.device("cuda", rank))
    coords, numbers, pseudo, grid, rho = payload
    part = MBISWPart(coords, numbers, pseudo, grid, rho, device=torch.device("cuda", rank), comm=comm)
    part.do_charges()
    part.do_moments()
    out[rank] = (int(part["niter"]), part["charges"].copy(), part["propars"].copy(), part["cartesian_multipoles"].copy())
    comm.close()


def test_sharded_through_the_c_abi_communicator(make_water, tmp_path):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from horton_part_b200 import MBISWPart

    case = make_water(9, nrad=30, nang=38, seed=2)
    payload = (case["coords"], case["numbers"], case["pseudo"], case["grid"], case["rho"])
    out = mp.Manager().dict()
    mp.spawn(_worker_hpcomm, args=(2, str(tmp_path / "nccl_id"), payload, out), nprocs=2, join=True)
    single = MBISWPart(*payload)
    single.do_charges()
    single.do_moments()
    for rank in (0, 1):
        niter, charges, propars, moments = out[rank]
        assert niter == single["niter"]
        np.testing.assert_allclose(charges, single["charges"], rtol=0, atol=1e-12)
        np.testing.assert_allclose(propars, single["propars"], rtol=1e-11, atol=1e-14)
        np.testing.assert_allclose(moments, single["cartesian_multipoles"], rtol=1e-10, atol=1e-12)
    assert np.array_equal(out[0][1], out[1][1])


def test_c_abi_communicator_symbols_load():
    """hp_comm_* bind NCCL at run time; on any GPU box the library must be found."""
    from horton_part_b200 import _lib

    assert int(_lib.call("hp_comm_nccl_version")) >= 20000


def test_work_estimate_matches_the_kernel_counters(make_water):
    """estimate_dense_work (geometry only) against the pairs the screened dense pass really evaluates:
    same total to ~15 %, and per-rank loads of a work-balanced split within a few per cent."""
    import torch

    from horton_part_b200 import MBISWPart
    from horton_part_b200.core.device import Shard
    from horton_part_b200.mbis import mbis_atom_work

    case = make_water(384, nrad=40, nang=50)
    natom, grid = len(case["numbers"]), case["grid"]
    work = mbis_atom_work(case["coords"], case["numbers"], grid)
    assert work.shape == (natom,) and (work > 0).all()
    part = MBISWPart(case["coords"], case["numbers"], case["pseudo"], grid, case["rho"])
    part._init_propars()
    part._launch_promol_weights()
    torch.cuda.synchronize()
    pairs = part._table.pairs_evaluated()
    setup = 55.0 * natom * sum(-(-int(n) // 1024) for n in np.diff(grid.indices))
    assert abs((work.sum() - setup) - pairs) < 0.15 * pairs, (work.sum() - setup, pairs)
    assert work.min() < 0.9 * work.max()  # surface atoms do less work than interior ones
    for world in (2, 4):
        shards = [Shard(natom, grid.indices, r, world, work=work) for r in range(world)]
        loads = np.array([work[s.atom_lo : s.atom_hi].sum() for s in shards])
        assert loads.max() / loads.mean() < 1.05
