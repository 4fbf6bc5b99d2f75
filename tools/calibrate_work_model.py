"""Calibration of the sharding cost model (core/device.py estimate_dense_work): time of the screened
dense pass on each of the 8 atom-block shards of config 5, run one after the other on ONE GPU, against
the pairs it evaluates and the chunks it processes.  Least squares  t = a * pairs + b * chunks * natom
gives the per-chunk screening cost in units of point evaluations (setup_points = b / a)."""
import logging
import sys

sys.path.insert(0, ".")
import numpy as np
import torch

logging.disable(logging.INFO)
import bench
from horton_part_b200 import _lib, synthetic
from horton_part_b200.core.device import GridSlab, ShellTable, Shard, stream_ptr, to_device
from horton_part_b200.mbis import get_initial_mbis_propars, mbis_atom_work

natom, world = 2000, 8
dev = torch.device("cuda:0")
coords, numbers, grid = bench.build_system(natom)
work = mbis_atom_work(coords, numbers, grid, dev)
propars = np.concatenate([get_initial_mbis_propars(int(z)) for z in numbers])
rows = []
for mode, w in (("points", None), ("work", work)):
    for r in range(world):
        shard = Shard(natom, grid.indices, r, world, work=w)
        slab = GridSlab(grid, np.zeros(grid.size), coords, dev, shard, need_atgrids=False)
        table = ShellTable(slab, 1, [len(get_initial_mbis_propars(int(z))) // 2 for z in numbers])
        _lib.call("hp_table_mbis", table.nshell, to_device(propars, dev), table.A, table.alpha, stream_ptr(dev))
        table.promol_weights(1e-15, True, True, False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            table.promol_weights(1e-15, True, True, False)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        pairs = table.pairs_evaluated()
        rows.append((mode, r, shard.nlocal, table._loc_nchunk, pairs, ms, float(work[shard.atom_lo:shard.atom_hi].sum())))
        print(mode, r, "atoms", shard.nlocal, "chunks", table._loc_nchunk, "pairs", pairs, "ms", round(ms, 3), flush=True)
        del slab, table
        torch.cuda.empty_cache()
A = np.array([[row[4], row[3] * natom] for row in rows], float)
t = np.array([row[5] for row in rows])
(a, b), *_ = np.linalg.lstsq(A, t, rcond=None)
print("fit: ms = %.4e * pairs + %.4e * chunks*natom  ->  setup_points = %.1f" % (a, b, b / a))
print("residuals (ms):", np.round(A @ np.array([a, b]) - t, 3))
for mode in ("points", "work"):
    ts = np.array([row[5] for row in rows if row[0] == mode])
    print(mode, "max/mean of the shard times:", round(ts.max() / ts.mean(), 4))
