"""horton_part_b200/mulliken.py: sum rule and structure on a water-like minimal basis, and -- where
the reference tree is present -- equality with the reference's own functions on random overlap
matrices and shell layouts (the reference module needs NumPy only).  CPU only."""

import importlib.util
import pathlib

import numpy as np
import pytest

from horton_part_b200 import mulliken

REF = pathlib.Path("/root/reference/src/horton_part/mulliken.py")


def _random_case(rng):
    ncenter = int(rng.integers(1, 5))
    nshell = int(rng.integers(ncenter, 3 * ncenter + 2))
    shell_types = [int(t) for t in rng.choice([0, 1, 2, -2, -3, 3], size=nshell)]
    shell_maps = [int(c) for c in rng.integers(0, ncenter, size=nshell)]
    nbasis = sum(mulliken.get_shell_nbasis(t) for t in shell_types)
    a = rng.normal(size=(nbasis, nbasis))
    return a @ a.T + np.identity(nbasis), ncenter, shell_types, shell_maps


def test_shell_sizes():
    assert [mulliken.get_shell_nbasis(t) for t in (0, 1, 2, 3, -2, -3, -4)] == [1, 3, 6, 10, 5, 7, 9]
    assert mulliken.get_shell_nbasis(-1) == -1


def test_operators_sum_to_the_overlap_and_count_the_electrons():
    # minimal basis of water: O 1s 2s 2p, H 1s, H 1s (shell types / maps of tests/test_mulliken.py:27-30)
    shell_types, shell_maps = [0, 0, 1, 0, 0], [0, 0, 0, 1, 2]
    rng = np.random.default_rng(0)
    a = rng.normal(size=(7, 7))
    overlap = a @ a.T + np.identity(7)
    c = rng.normal(size=(7, 5))
    dm = 2 * c @ np.linalg.solve(c.T @ overlap @ c, c.T)  # idempotent-like: tr(dm S) = 10 electrons
    ops = mulliken.get_mulliken_operators(overlap, 3, shell_types, shell_maps)
    assert len(ops) == 3 and all(np.allclose(p, p.T) for p in ops)
    np.testing.assert_allclose(sum(ops), overlap, atol=1e-14)
    pops = [np.einsum("ab,ba", p, dm) for p in ops]
    assert abs(sum(pops) - 10.0) < 1e-10
    # the oxygen operator touches only rows/columns of its five functions
    assert not ops[0][5:, 5:].any() and ops[1][5, 5] == overlap[5, 5] and ops[2][6, 6] == overlap[6, 6]


def test_equal_to_the_reference_functions():
    if not REF.is_file():
        pytest.skip("reference tree not present on this machine")
    spec = importlib.util.spec_from_file_location("ref_mulliken", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = np.random.default_rng(5)
    for _ in range(50):
        overlap, ncenter, shell_types, shell_maps = _random_case(rng)
        mine = mulliken.get_mulliken_operators(overlap, ncenter, shell_types, shell_maps)
        theirs = ref.get_mulliken_operators(overlap, ncenter, shell_types, shell_maps)
        for a, b in zip(mine, theirs):
            np.testing.assert_array_equal(a, b)
        op1, op2 = overlap.copy(), overlap.copy()
        mulliken.partition_mulliken(op1, len(overlap), shell_types, shell_maps, 0)
        ref.partition_mulliken(op2, len(overlap), shell_types, shell_maps, 0)
        np.testing.assert_array_equal(op1, op2)
    for t in range(-6, 7):
        assert mulliken.get_shell_nbasis(t) == ref.get_shell_nbasis(t)
