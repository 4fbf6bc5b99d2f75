"""Exactness of the two work-skipping devices of the dense pass: shell screening (2^-100 of the
atom's most diffuse shell) and atom screening (upper bound of the pro-atom below 2^-(55+log2 natom)
of a lower bound of the promolecule).  Neither may change the promolecule beyond the rounding noise
of the sequential FP64 sum itself (a different set of additions rounds differently: tens of ulp at
worst, measured 50 ulp on config 5), nor the charges beyond 1e-13."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(case, monkeypatch, atom_screen, shell_bits=None, niter=3, cls=None, **kw):
    from horton_part_b200 import MBISWPart

    monkeypatch.setenv("HP_B200_ATOM_SCREEN", "1" if atom_screen else "0")
    if shell_bits is not None:
        monkeypatch.setenv("HP_B200_SCREEN_BITS", str(shell_bits))
    else:
        monkeypatch.delenv("HP_B200_SCREEN_BITS", raising=False)
    part = (cls or MBISWPart)(case["coords"], case["numbers"], case["pseudo"], case["grid"], case["rho"],
                              maxiter=niter, **kw)  # fmt: skip
    part.do_partitioning()
    pairs = part._table.pairs_evaluated()
    return part, pairs


def test_atom_screening_is_exact(make_water, monkeypatch):
    case = make_water(384, nrad=40, nang=50)  # 2,000 points per atom: chunks have one owner
    natom, npts = len(case["numbers"]), case["grid"].size
    on, pairs_on = _run(case, monkeypatch, True)
    off, pairs_off = _run(case, monkeypatch, False)
    plain, _ = _run(case, monkeypatch, False, shell_bits=0)  # the plain dense kernel, nothing skipped
    assert pairs_off == natom * npts * 1 or pairs_off == natom * npts  # last launch, all pairs
    assert pairs_on < 0.99 * pairs_off, (pairs_on, pairs_off)  # skips work even in a 30-bohr cluster
    # `off` is the same kernel with every pair evaluated: same operation sequence except for the
    # dropped terms.  `plain` is the reference-ordered kernel (each pro-atom summed over its shells
    # first, then added): the chunk kernel feeds every shell straight into the running sum with one
    # FMA, an equally valid but different rounding sequence of ~natom*K additions (random walk of
    # half-ulp steps: a few hundred ulp at worst over 768,000 points).
    # The chunk kernel's exponential reduces its argument with ONE constant (r = x - k fl(ln2), hp_math.cuh):
    # a systematic relative error of k * 2.3e-17 = 0.1-0.2 ulp per unit of k = -x / ln2.  Where the
    # promolecule is above 1e-30 (k <= 100) that is below the rounding-sequence noise; in the far field
    # (values down to 1e-98 here, k up to ~1000) it reaches a few hundred ulp, i.e. 1e-13 relative.
    for other, max_ulp, same, wtol in ((off, 64, 0.9, 1e-15), (plain, 512, 0.0, 1e-12)):
        a, b = on["promoldens"], other["promoldens"]
        ulp = np.spacing(np.abs(b))
        tol = np.where(b >= 1e-30, max_ulp, max(max_ulp, 2048)) if other is plain else max_ulp
        assert (np.abs(a - b) <= tol * ulp).all(), float((np.abs(a - b) / ulp).max())
        assert (a == b).mean() >= same
        np.testing.assert_allclose(on["charges"], other["charges"], rtol=0, atol=5e-14)
        np.testing.assert_allclose(on["propars"], other["propars"], rtol=1e-13)
        np.testing.assert_allclose(on["history_entropies"], other["history_entropies"], rtol=1e-13)
        for atom in (0, 17, natom - 1):
            np.testing.assert_allclose(on[f"at_weights_{atom}"], other[f"at_weights_{atom}"], rtol=wtol, atol=1e-300)
    # far outside the cluster the promolecule is tiny: the test must have stayed off there
    assert on["promoldens"].min() < 1e-60


def test_atom_screening_gaussian_functor(make_water, monkeypatch):
    from horton_part_b200 import LinearISAWPart

    case = make_water(192, nrad=40, nang=50)
    on, pairs_on = _run(case, monkeypatch, True, cls=LinearISAWPart, solver="sc", niter=2)
    off, pairs_off = _run(case, monkeypatch, False, cls=LinearISAWPart, solver="sc", niter=2)
    assert pairs_on < pairs_off
    a, b = on["promoldens"], off["promoldens"]
    assert (np.abs(a - b) <= 64 * np.spacing(np.abs(b))).all()
    np.testing.assert_allclose(on["charges"], off["charges"], rtol=0, atol=5e-14)


def test_negative_amplitudes_switch_atom_screening_off(make_water, monkeypatch):
    """The lower bounds need non-negative pro-atoms: hp_shell_screen flags a negative amplitude and
    the kernel then evaluates every pair."""
    import torch

    from horton_part_b200 import MBISWPart

    case = make_water(96, nrad=40, nang=50)
    monkeypatch.setenv("HP_B200_ATOM_SCREEN", "1")
    part = MBISWPart(case["coords"], case["numbers"], case["pseudo"], case["grid"], case["rho"])
    part._init_propars()
    part._launch_promol_weights()
    torch.cuda.synchronize()
    natom, npts = len(case["numbers"]), case["grid"].size
    assert float(part._table.skip[-1].item()) == 0.0
    part._state.propars[0] = -part._state.propars[0]  # a negative population
    part._launch_promol_weights()
    torch.cuda.synchronize()
    assert float(part._table.skip[-1].item()) == 1.0
    assert part._table.pairs_evaluated() == natom * npts
