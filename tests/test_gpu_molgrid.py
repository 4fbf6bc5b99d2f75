"""grid_type 2/3: per-atom fixed points on the molecular grid (row a9) against the reference's own
runs (tests/golden: mbis_gt2, lisa_sc_gt2)."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _gold(gold, tag):
    return {k.split("/", 1)[1]: gold[k] for k in gold.files if k.startswith(tag + "/")}


def _compare(part, ref, ptol=1e-8):
    assert part["niter"] == int(ref["niter"])
    np.testing.assert_allclose(part["charges"], ref["charges"], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(part["propars"], ref["propars"], rtol=ptol, atol=1e-9)
    np.testing.assert_allclose(part["history_changes"], ref["history_changes"], rtol=1e-5)
    np.testing.assert_allclose(part["history_entropies"], ref["history_entropies"], rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(part["promoldens"][::97], ref["promoldens_sample"], rtol=1e-8)


def test_mbis_grid_type_2_h2o(h2o):
    from horton_part_b200 import MBISWPart

    part = MBISWPart(h2o["coords"], h2o["numbers"], h2o["pseudo"], h2o["grid"], h2o["rho"], grid_type=2)
    part.do_partitioning()
    _compare(part, _gold(h2o["gold"], "mbis_gt2"))
    assert part["niter"] == 27


def test_alisa_sc_grid_type_2_h2o(h2o):
    from horton_part_b200 import LinearISAWPart

    part = LinearISAWPart(h2o["coords"], h2o["numbers"], h2o["pseudo"], h2o["grid"], h2o["rho"], solver="sc",
                          grid_type=2)
    part.do_partitioning()
    _compare(part, _gold(h2o["gold"], "lisa_sc_gt2"), ptol=1e-5)
    assert part["niter"] == 22


def test_mbis_grid_type_3_equals_2(water6):
    """grid_type 3 (molecular grid only, no atomic grids needed) gives the same partitioning."""
    from horton_part_b200 import MBISWPart

    args = (water6["coords"], water6["numbers"], water6["pseudo"], water6["grid"], water6["rho"])
    p2 = MBISWPart(*args, grid_type=2, maxiter=8)
    p3 = MBISWPart(*args, grid_type=3, maxiter=8)
    p2.do_partitioning()
    p3.do_partitioning()
    np.testing.assert_allclose(p3["charges"], p2["charges"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(p3["history_changes"], p2["history_changes"], rtol=1e-12)


def test_nlis_grid_type_2_h2o(h2o):
    from horton_part_b200 import NLISWPart

    part = NLISWPart(h2o["coords"], h2o["numbers"], h2o["pseudo"], h2o["grid"], h2o["rho"], exp_n_dict={},
                     grid_type=2)
    part.do_partitioning()
    _compare(part, _gold(h2o["gold"], "nlis_gt2"))


def test_glisa_sc_grid_type_2_h2o(h2o):
    from horton_part_b200 import GlobalLinearISAWPart

    part = GlobalLinearISAWPart(h2o["coords"], h2o["numbers"], h2o["pseudo"], h2o["grid"], h2o["rho"],
                                solver="sc", grid_type=2)
    part.do_partitioning()
    _compare(part, _gold(h2o["gold"], "glisa_sc_gt2"), ptol=1e-5)


def test_alisa_default_solver_grid_type_2(water6g):
    """grid_type 2 with a HOST plug-in (the reference's default aLISA solver, the convex programme):
    promolecule and entropy from the device, the K_a x Npts per-atom problems on the host exactly
    as the reference does them (gisa.py:257-279).  Golden: reference run through the oracle's
    cvxopt stand-in (tests/golden/water6_convex.npz)."""
    from conftest import GOLDEN

    from horton_part_b200 import LinearISAWPart

    gold = np.load(GOLDEN / "water6_convex.npz")
    tag = "g/lisa_cvxopt_gt2/"
    ref = {k[len(tag) :]: gold[k] for k in gold.files if k.startswith(tag)}
    c = water6g
    part = LinearISAWPart(c["coords"], c["numbers"], c["pseudo"], c["grid"], c["rho"], grid_type=2)
    part.do_partitioning()
    assert part["niter"] == int(ref["niter"])
    np.testing.assert_allclose(part["charges"], ref["charges"], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(part["propars"], ref["propars"], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(part["history_changes"], ref["history_changes"], rtol=1e-5, atol=1e-10)
    np.testing.assert_allclose(part["history_entropies"], ref["history_entropies"], rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(part["promoldens"][::97], ref["promoldens_sample"], rtol=1e-8)


def test_alisa_callable_solver_grid_type_3(water6g):
    """A user callable with the reference's plug-in signature on grid_type 3 gets the molecular-grid
    arrays (K_a x Npts basis table, w_a*rho, grid points and weights) and reproduces the device
    fixed point when it performs the same update."""
    from horton_part_b200 import LinearISAWPart

    c = water6g
    seen = []

    def one_sc_step(bs_funcs, rho, propars, points, weights, threshold, logger, density_cutoff, **kw):
        seen.append((bs_funcs.shape, rho.shape, points.shape, weights.shape))
        pro = propars @ bs_funcs
        ok = (rho >= density_cutoff) & (pro >= density_cutoff)
        ratio = np.divide(rho, pro, out=np.zeros_like(rho), where=ok)
        return np.einsum("kp,p,p->k", bs_funcs * propars[:, None], ratio, weights)

    args = (c["coords"], c["numbers"], c["pseudo"], c["grid"], c["rho"])
    host = LinearISAWPart(*args, grid_type=3, solver=one_sc_step, maxiter=5)
    host.do_partitioning()
    dev = LinearISAWPart(*args, grid_type=3, solver="sc-1-iter", maxiter=5)
    dev.do_partitioning()
    npts = c["grid"].size
    assert seen[0][1:] == ((npts,), (npts, 3), (npts,)) and seen[0][0][1] == npts
    np.testing.assert_allclose(host["charges"], dev["charges"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(host["history_changes"], dev["history_changes"], rtol=1e-6)


@pytest.mark.parametrize("tag,scheme,case_name,kw,rtol", [
    ("g/gisa_gt2", "GaussianISAWPart", "water6g", dict(), 1e-8),
    ("s/lisa_diis_gt2", "LinearISAWPart", "water6", dict(solver="diis", maxiter=8, solver_options=dict(check_mono=False)), 2e-6),
])  # fmt: skip
def test_host_plugins_on_the_molecular_grid(request, tag, scheme, case_name, kw, rtol):
    """GISA's quadratic programme and an aLISA host plug-in (DIIS) with grid_type 2 against reference
    runs (tests/golden/water6_convex.npz); DIIS tolerance as in tests/test_gpu_solvers.py."""
    import contextlib
    import io
    import warnings

    from conftest import GOLDEN

    import horton_part_b200 as hp

    gold = np.load(GOLDEN / "water6_convex.npz")
    ref = {k[len(tag) + 1 :]: gold[k] for k in gold.files if k.startswith(tag + "/")}
    c = request.getfixturevalue(case_name)
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        part = getattr(hp, scheme)(c["coords"], c["numbers"], c["pseudo"], c["grid"], c["rho"], grid_type=2, **kw)
        part.do_partitioning()
    assert part["niter"] == int(ref["niter"])
    np.testing.assert_allclose(part["charges"], ref["charges"], rtol=rtol, atol=1e-9)
    np.testing.assert_allclose(part["history_changes"], ref["history_changes"], rtol=max(1e-5, 5e3 * rtol), atol=1e-11)
    # (the promolecule of the PENULTIMATE parameters is cached; with DIIS those carry the solver's
    # restart noise: 1e-5 relative at single points, measured)
    np.testing.assert_allclose(part["promoldens"][::97], ref["promoldens_sample"], rtol=1e-8 if rtol <= 1e-8 else 1e-4)


def test_spin_charges_without_atomic_grids(water6):
    """grid_type 3 has no atomic grids: populations of a second density are integrated over the whole
    molecular grid with the weight functions regenerated in hp_atom_weight_integrals (the reference does
    grid.integrate(at_weights, spindens), core/base.py:287-298, 313-327)."""
    from horton_part_b200 import MBISWPart

    spin = 0.125 * water6["rho"]
    part = MBISWPart(water6["coords"], water6["numbers"], water6["pseudo"], water6["grid"], water6["rho"],
                     spindens=spin, grid_type=3)
    part.do_spin_charges()
    pops = water6["pseudo"] - part["charges"]
    # the weights are those of the last iteration (penultimate parameters), the charges come from the same
    # weights: spin populations = 0.125 x populations up to the quadrature of the two code paths
    np.testing.assert_allclose(part["spin_charges"], 0.125 * pops, rtol=1e-9, atol=1e-12)
    with pytest.raises(NotImplementedError, match="do_moments"):
        part.do_moments()
