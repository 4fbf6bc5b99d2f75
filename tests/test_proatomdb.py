"""ProAtomDB / ProAtomRecord host API (core/proatomdb.py of the reference): the assertions of the
reference's own tests/core/test_proatomdb.py on its HF/STO-3G records (packed into
tests/golden/h2o_hirshfeld.npz), plus values the reference produced for compute_radii, compact and
normalize (tests/golden/proatomdb.npz, oracle/gen_golden.py::case_proatomdb).  CPU only."""

import copy

import numpy as np
import pytest
from conftest import GOLDEN

GOLD = np.load(GOLDEN / "proatomdb.npz")


def test_db_basics(h2o_proatomdb):
    db, _ = h2o_proatomdb
    assert db.get_numbers() == [1, 6, 8]
    for z in db.get_numbers():
        assert db.get_charges(z) == list(GOLD[f"charges/{z}"])
        assert db.get_charges(z, safe=True) == list(GOLD[f"safe/{z}"])
    assert db.size == sum(len(db.get_charges(z)) for z in db.get_numbers())
    r1 = db.get_record(8, -1)
    assert (r1.number, r1.charge, r1.population, r1.pseudo_number, r1.pseudo_population) == (8, -1, 9, 8, 9)
    assert r1.energy < -70 and r1.rgrid.size == 59  # (HF/STO-3G records; the reference test uses another level)
    assert r1.ipot_energy == db.get_record(8, 0).energy - r1.energy
    assert db.get_record(1, 0).ipot_energy == -db.get_record(1, 0).energy
    highest = db.get_charges(8)[0]
    assert db.get_record(8, highest).ipot_energy is None
    # equality is by value (tests/core/test_proatomdb.py:47-50)
    assert r1 == db.get_record(8, -1) and r1 == copy.deepcopy(r1)
    assert r1 != db.get_record(8, 0)
    assert (db.get_rho(8, {}) == 0.0).all()  # test_empty_proatom


def test_compute_radii_matches_reference(h2o_proatomdb):
    db, _ = h2o_proatomdb
    for z in db.get_numbers():
        for q in db.get_charges(z):
            rec = db.get_record(z, q)
            idx, radii = rec.compute_radii([0.5 * rec.pseudo_population, rec.pseudo_population - 0.1, 1e3])
            ref = GOLD[f"radii/{z}/{q}"]
            assert list(idx) == [int(v) for v in ref[:3]]
            np.testing.assert_allclose(radii, ref[3:], rtol=1e-12)
            assert radii[2] == rec.rgrid.points[-1]  # more electrons than the atom has: last point


def test_compact_and_normalize_match_reference(h2o_proatomdb):
    db = copy.deepcopy(h2o_proatomdb[0])
    before = {z: db.get_rgrid(z).size for z in db.get_numbers()}
    db.compact(0.1)
    for z in db.get_numbers():
        assert db.get_rgrid(z).size == int(GOLD[f"compact_size/{z}"]) < before[z]
        for q in db.get_charges(z):
            rec = db.get_record(z, q)
            assert rec.rgrid.size == rec.rho.size == rec.deriv.size == db.get_rgrid(z).size
    db.normalize()
    for z in db.get_numbers():
        rgrid = db.get_rgrid(z)
        for q in db.get_charges(z):
            rec = db.get_record(z, q)
            assert abs(rgrid.integrate(rec.rho) - (rec.pseudo_number - q)) < 1e-10  # test_normalize
            np.testing.assert_allclose(rec.rho, GOLD[f"normalized/{z}/{q}"], rtol=1e-12, atol=1e-300)
    # splines still come out of the compacted database
    spline = db.get_spline(8, {0: 0.5, -1: 0.5})
    assert spline(db.get_rgrid(8).points).shape == (db.get_rgrid(8).size,)


def test_records_are_not_hashable_and_compare_only_with_records(h2o_proatomdb):
    rec = h2o_proatomdb[0].get_record(1, 0)
    with pytest.raises(TypeError):
        hash(rec)
    assert rec != "H"
