"""The WPart constructor contract of SURVEY section 8b, checked on a machine WITHOUT a GPU: the
classes construct (they only validate and log), mirror the reference's properties and error
behaviour (tests/core/test_base.py:29-56, core/base.py:416-470, core/iterstock.py:101), and refuse
to partition -- loudly -- when no CUDA device is there.  CPU only."""

import logging

import numpy as np
import pytest
import torch

import horton_part_b200 as hp
from horton_part_b200 import gridlite, synthetic
from horton_part_b200.core.base import WPart, get_ncart_cumul, get_npure_cumul


@pytest.fixture(scope="module")
def water():
    logging.disable(logging.CRITICAL)
    coords, numbers = synthetic.water_cluster(3, 0)
    rgrid = gridlite.BeckeRTransform(1e-4, 1.5).transform_1d_grid(gridlite.GaussChebyshev(10))
    grids = {store: gridlite.MolGrid.from_size(numbers, coords, 26, rgrid, gridlite.BeckeWeights(), store=store)
             for store in (False, True)}  # fmt: skip
    rho = synthetic.slater_promolecule_host(grids[True].points, coords, numbers)
    yield coords, numbers, numbers.astype(float), grids, rho
    logging.disable(logging.INFO)  # the level tests/conftest.py sets


def test_base_exceptions(water):
    coords, numbers, pseudo, grids, rho = water
    with pytest.raises(ValueError, match="Atomic grids are discarded"):
        WPart(coords, numbers, pseudo, grids[False], rho)  # local integrations need the atomic grids
    with pytest.raises(ValueError, match="Atomic grids are discarded"):
        hp.MBISWPart(coords, numbers, pseudo, grids[False], rho)
    with pytest.raises(NotImplementedError):
        WPart(coords, numbers, pseudo, grids[True], rho)  # the base class is abstract
    hp.MBISWPart(coords, numbers, pseudo, grids[False], rho, grid_type=3)  # molecular grid only: fine
    with pytest.raises(TypeError):
        hp.MBISWPart(coords, numbers.astype(np.int32), pseudo, grids[True], rho)
    with pytest.raises(TypeError):
        hp.MBISWPart(coords.astype(np.float32), numbers, pseudo, grids[True], rho)
    with pytest.raises(AssertionError):
        hp.MBISWPart(coords, numbers, pseudo, grids[True], rho, grid_type=4)


def test_properties_and_ignored_keywords(water):
    coords, numbers, pseudo, grids, rho = water
    part = hp.MBISWPart(coords, numbers, None, grids[True], rho, lmax=2, threshold=1e-7, inner_threshold=1e-5,
                        maxiter=17, density_cutoff=1e-14, some_unknown_keyword=1)  # **ignored, as in the reference
    assert part.natom == 3 and part.local and part.grid_type == 1 and not part.on_molgrid and not part.only_use_molgrid
    assert part.lmax == 2 and part.density_cutoff == 1e-14 and part.negative_cutoff == -1e-12 and part.population_cutoff == 1e-4
    assert part.coordinates is coords and part.numbers is numbers and (part.pseudo_numbers == numbers).all()
    assert part.pseudo_numbers.dtype == float and part.grid is grids[True]
    assert abs(part.nelec - grids[True].integrate(rho)) < 1e-12
    assert part._inner_threshold == 1e-7  # clamped to the outer threshold (core/iterstock.py:101)
    assert part.get_grid(1) is grids[True].atgrids[1] and part.get_grid() is grids[True]
    lo, hi = grids[True].indices[1], grids[True].indices[2]
    assert np.array_equal(part.get_moldens(1), rho[lo:hi]) and part.get_moldens() is rho
    assert np.array_equal(part.to_atomic_grid(2, rho), rho[grids[True].indices[2] :])
    assert part.variables_stored_in_cache() is not None and "charges" not in part.cache
    for gt, (on, only) in {2: (True, False), 3: (True, True)}.items():
        p = hp.MBISWPart(coords, numbers, pseudo, grids[True], rho, grid_type=gt)
        assert (p.on_molgrid, p.only_use_molgrid, p.local) == (on, only, not only)


@pytest.mark.parametrize("scheme,kw", [
    ("mbis", {}), ("is", {}), ("lisa", {}), ("lisa", {"solver": "sc", "basis_func": "slater"}), ("gisa", {}),
    ("glisa", {}), ("nlis", {"exp_n_dict": {}}), ("gmbis", {"exp_n_dict": {}}), ("b", {}),
])  # fmt: skip
def test_every_scheme_constructs_and_refuses_to_run_without_cuda(water, scheme, kw):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    coords, numbers, pseudo, grids, rho = water
    part = hp.wpart_schemes(scheme)(coords, numbers, pseudo, grids[True], rho, **kw)
    assert part.name == scheme
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        part.do_partitioning()


def test_multipole_counts():
    assert [get_ncart_cumul(l) for l in range(5)] == [1, 4, 10, 20, 35]  # core/base.py:685-692
    assert [get_npure_cumul(l) for l in range(5)] == [1, 4, 9, 16, 25]
