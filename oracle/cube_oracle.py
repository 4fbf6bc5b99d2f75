"""NumPy restatement of the AIM arrays of the reference's `part-cube` program (test infrastructure: only
tests/ import this).  Follows /root/reference/src/horton_part/scripts/generate_cube.py:140-157 (`_compute_rho0`)
and :213-227 (distances, promolecule + 1e-100, weight functions, AIM densities) line by line; the basis
evaluation is the helper's `compute_proatom_dens` (core/basis.py:214-232).  Pinned: the helper is checked against
the reference's known answers in tests/test_basis_host.py, and for a promolecular density the result has the
closed form checked in tests/test_cube_host.py."""

import numpy as np


def aim_on_points(helper, atnums, atcoords, points, density, propars):
    points = np.asarray(points, dtype=float)
    dis_array = np.linalg.norm(points[None, :, :] - np.asarray(atcoords, dtype=float)[:, None, :], axis=2)  # :213-215
    rho0 = np.zeros_like(dis_array)
    begin = 0
    for i, number in enumerate(atnums):  # :149-156
        nshell = helper.get_nshell(int(number))
        rho0[i, :] = helper.compute_proatom_dens(int(number), propars[begin : begin + nshell], dis_array[i, :], 0)
        begin += nshell
    promol = np.sum(rho0, axis=0)  # :219
    promol += 1e-100  # :220
    weights_funcs = rho0 / promol  # :222
    aim_rho = weights_funcs * np.asarray(density, dtype=float)[None, :]  # :225
    return rho0, promol, aim_rho
