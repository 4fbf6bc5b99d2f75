"""The device-resident outer loop (csrc/hp_loop.cu: one CUDA-graph launch with a conditional WHILE node,
convergence test on the device) must reproduce the host-driven loop bit for bit: same kernels, same order,
same stopping rule (core/iterstock.py:171-188)."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _pair(cls, case, **kw):
    parts = []
    for flag in (True, False):
        part = cls(case["coords"], case["numbers"], case["pseudo"], case["grid"], case["rho"], device_loop=flag, **kw)
        part.do_partitioning()
        parts.append(part)
    return parts


def _same(dev, host):
    assert "device_loop" in dev.time_usage and "device_loop" not in host.time_usage
    assert dev["niter"] == host["niter"]
    for key in ("charges", "propars", "history_changes", "history_entropies", "history_charges", "history_propars",
                "promoldens", "at_weights_0"):
        assert np.array_equal(dev[key], host[key]), key
    assert dev["change"] == host["change"]
    assert len(dev.history_time_update_at_weights) == dev["niter"] == len(dev.history_time_update_propars)
    assert min(dev.history_time_update_at_weights) > 0 and min(dev.history_time_update_propars) > 0


def test_mbis_h2o(h2o):
    from horton_part_b200 import MBISWPart

    dev, host = _pair(MBISWPart, h2o)
    _same(dev, host)
    assert dev["niter"] == 27


@pytest.mark.parametrize("maxiter", [1, 2, 3])
def test_maxiter_is_honoured(h2o, maxiter):
    from horton_part_b200 import MBISWPart

    dev, host = _pair(MBISWPart, h2o, maxiter=maxiter)
    assert dev["niter"] == host["niter"] == maxiter
    assert np.array_equal(dev["history_changes"], host["history_changes"])


def test_isa_water6(water6):
    from horton_part_b200 import ISAWPart

    _same(*_pair(ISAWPart, water6, maxiter=60))


def test_nlis_h2o(h2o):
    from horton_part_b200 import NLISWPart

    _same(*_pair(NLISWPart, h2o, exp_n_dict={}))


@pytest.mark.parametrize("basis", ["gauss", "slater"])
def test_alisa_sc_water6(water6, basis):
    from horton_part_b200 import LinearISAWPart

    _same(*_pair(LinearISAWPart, water6, solver="sc", basis_func=basis))


def test_host_plugins_keep_the_host_loop(water6):
    from horton_part_b200 import LinearISAWPart

    part = LinearISAWPart(water6["coords"], water6["numbers"], water6["pseudo"], water6["grid"], water6["rho"],
                          solver="diis", maxiter=3)
    part.do_partitioning()
    assert "device_loop" not in part.time_usage


def test_two_partitionings_interleaved_on_one_device(h2o, water6):
    """Per-launch work counters (chunk_scratch) and per-object graphs: two objects of one process do not
    disturb each other."""
    from horton_part_b200 import MBISWPart

    a = MBISWPart(h2o["coords"], h2o["numbers"], h2o["pseudo"], h2o["grid"], h2o["rho"])
    b = MBISWPart(water6["coords"], water6["numbers"], water6["pseudo"], water6["grid"], water6["rho"])
    a._init_propars()
    b._init_propars()
    for _ in range(3):
        a._run_iteration()
        b._run_iteration()
    ref = MBISWPart(h2o["coords"], h2o["numbers"], h2o["pseudo"], h2o["grid"], h2o["rho"], maxiter=3, device_loop=False)
    ref.do_partitioning()
    assert np.array_equal(a.cache.load("charges"), ref["charges"])


def test_cutoff_mode_in_the_device_loop(water6):
    """local_radius (the reference's removed local-grid design) uses the same chunk kernel: the graph loop
    reproduces the host loop there too."""
    from horton_part_b200 import MBISWPart

    dev, host = _pair(MBISWPart, water6, local_radius=9.0, maxiter=25)
    _same(dev, host)
