"""Hirshfeld-I host algebra: the vectorised mixing of the database states (all atoms of an element at once)
against the per-atom route that follows the reference line by line (hirshfeld_i.py:116-158, core/iterstock.py:32-45)."""

import numpy as np
import pytest
from conftest import GOLDEN


def _database(z):
    from horton_part_b200 import gridlite
    from horton_part_b200.core.proatomdb import ProAtomDB, ProAtomRecord

    records = []
    for key in z.files:
        if key.startswith("record/"):
            v = z[key]
            n = int(v[5])
            rgrid = gridlite.PowerRTransform(v[3], v[4], n - 1).transform_1d_grid(gridlite.UniformInteger(n))
            records.append(ProAtomRecord(int(v[0]), int(v[1]), float(v[2]), rgrid, v[6 : 6 + n].copy(), v[6 + n :].copy()))
    return ProAtomDB(records)


def _stub(numbers, db):
    from horton_part_b200.hirshfeld_i import HirshfeldIWPart

    class Stub(HirshfeldIWPart):  # the algebra only needs numbers, pseudo numbers and the database
        def __init__(self):
            self._numbers, self._pseudo_numbers = numbers, numbers.astype(float)
            self._proatomdb, self._coef_cache = db, {}

        numbers = property(lambda s: s._numbers)
        pseudo_numbers = property(lambda s: s._pseudo_numbers)
        natom = property(lambda s: len(s._numbers))

    return Stub()


def test_vectorised_mixing_equals_the_per_atom_route():
    from horton_part_b200.core.iterstock import AbstractISAWPart

    if not (GOLDEN / "config2_hi.npz").exists():
        pytest.skip("tests/golden/config2_hi.npz missing")
    z = np.load(GOLDEN / "config2_hi.npz")
    numbers = z["numbers"]
    part = _stub(numbers, _database(z))
    rng = np.random.default_rng(0)
    c1, c2 = rng.uniform(-0.6, 0.6, len(numbers)), rng.uniform(-0.6, 0.6, len(numbers))
    c1[3] = 0.0  # integer charge: one state only
    c1[numbers == 1] = np.abs(c1[numbers == 1]) * 0.0 + 0.25  # H at +0.25: floor 0, one-electron rule does not apply
    assert part.compute_change(c1, c2) == AbstractISAWPart.compute_change(part, c1, c2)
    for zz, tab in part._element_tables().items():
        mixed = part._mix(tab, c1, tab["coef"])
        for k, a in enumerate(tab["atoms"]):
            ic = int(np.floor(c1[a]))
            x = c1[a] - ic
            ref = part._state_coefficients(zz, ic) * (1 - x)
            if part.pseudo_numbers[a] - ic > 1 and x != 0.0:
                ref = ref + part._state_coefficients(zz, ic + 1) * x
            assert np.array_equal(mixed[k], ref.ravel())
    # a charge outside the database: the vectorised route steps aside, the per-atom route raises as before
    c3 = c1.copy()
    c3[numbers == 8] = -2.5  # the database ends at charge -1
    assert any(part._mix(tab, c3, tab["rho"]) is None for tab in part._element_tables().values())
    with pytest.raises(Exception):
        part.compute_change(c3, c2)
