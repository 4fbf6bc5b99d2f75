"""AIM quantities on a uniform grid and Gaussian cube files (post-processing, host NumPy).

Counterpart of the AIM half of the reference's ``part-cube`` program
(scripts/generate_cube.py:100-157, 206-272): from the converged pro-atom coefficients of an aLISA /
gLISA run it re-evaluates every pro-atom on the points of a uniform grid, forms the promolecule
(+1e-100), the weight functions w_a = rho0_a / rho0 and the atoms-in-molecules densities
w_a * rho, and writes them as cube files.  The other half of the reference program (the molecular
density on the uniform grid from a wavefunction, through iodata / gbasis) is outside this
repository: the density comes in as an array.

This is output formatting, O(natom * Npts * K) once per job; nothing here is on the GPU path.
"""

from __future__ import annotations

import os

import numpy as np

from ..core.basis import ExpBasisFuncHelper

__all__ = ["UniformGrid", "to_cube", "read_cube", "compute_rho0", "aim_on_points", "write_aim_cubes"]


class UniformGrid:
    """Origin, three axis vectors (rows of ``axes``) and point counts; points run with x as the
    outer and z as the inner loop, the cube-file order."""

    def __init__(self, origin, axes, shape):
        self.origin = np.asarray(origin, dtype=float).reshape(3)
        self.axes = np.asarray(axes, dtype=float).reshape(3, 3)
        self.shape = tuple(int(n) for n in shape)
        if len(self.shape) != 3 or min(self.shape) < 1:
            raise ValueError("shape must hold three positive point counts")

    @classmethod
    def from_molecule(cls, atnums, atcoords, spacing=0.2, extension=5.0):
        """Axis-aligned box around the molecule (the ``rotate=False`` set-up of the reference,
        scripts/generate_cube.py:70-97): ``extension`` bohr on every side, ``spacing`` between points."""
        atcoords = np.asarray(atcoords, dtype=float)
        lo = atcoords.min(axis=0) - extension
        hi = atcoords.max(axis=0) + extension
        shape = np.ceil((hi - lo) / spacing).astype(int) + 1
        return cls(lo, np.identity(3) * spacing, shape)

    size = property(lambda self: int(np.prod(self.shape)))

    @property
    def points(self):
        i, j, k = np.meshgrid(*(np.arange(n) for n in self.shape), indexing="ij")
        steps = np.stack([i.ravel(), j.ravel(), k.ravel()], axis=1).astype(float)
        return self.origin + steps @ self.axes

    @property
    def weights(self):
        return np.full(self.size, abs(np.linalg.det(self.axes)))

    def integrate(self, *arrays):
        return float(np.einsum(",".join("i" * (len(arrays) + 1)), self.weights, *arrays))


def to_cube(fname, atnums, atcorenums, atcoords, grid, data):
    """Write ``data`` (one value per grid point) as a cube file (scripts/generate_cube.py:100-137:
    two comment lines, atom count + origin, one line per axis, one per atom, six values per line)."""
    fname = os.fspath(fname)
    if not fname.endswith(".cube"):
        raise ValueError("Argument fname should be a cube file with `*.cube` extension!")
    data = np.asarray(data, dtype=float).ravel()
    if data.size != grid.size:
        raise ValueError(f"Argument data should have the same size as the grid. {data.size}!={grid.size}")
    lines = ["Cubefile created with HORTON-PART", "OUTER LOOP: X, MIDDLE LOOP: Y, INNER LOOP: Z"]
    lines.append("%5d %11.6f %11.6f %11.6f" % (len(atnums), *grid.origin))
    for n, axis in zip(grid.shape, grid.axes):
        lines.append("%5d %11.6f %11.6f %11.6f" % (n, *axis))
    for z, q, xyz in zip(atnums, atcorenums, atcoords):
        lines.append("%5d %11.6f %11.6f %11.6f %11.6f" % (z, q, *xyz))
    for lo in range(0, data.size, 6):
        lines.append("".join(" %12.5E" % v for v in data[lo : lo + 6]))
    with open(fname, "w") as fh:
        fh.write("\n".join(lines) + "\n")


def read_cube(fname):
    """Inverse of :func:`to_cube` (values to the 6 digits the format keeps)."""
    with open(fname) as fh:
        rows = fh.read().splitlines()
    head = rows[2].split()
    natom, origin = int(head[0]), np.array(head[1:4], dtype=float)
    shape, axes = [], []
    for row in rows[3:6]:
        parts = row.split()
        shape.append(int(parts[0]))
        axes.append([float(v) for v in parts[1:4]])
    atoms = np.array([[float(v) for v in row.split()] for row in rows[6 : 6 + natom]]).reshape(natom, 5)
    data = np.array(" ".join(rows[6 + natom :]).split(), dtype=float)
    return {"grid": UniformGrid(origin, axes, shape), "atnums": atoms[:, 0].astype(int), "atcorenums": atoms[:, 1],
            "atcoords": atoms[:, 2:5], "data": data}  # fmt: skip


def _helper(basis_func):
    if isinstance(basis_func, ExpBasisFuncHelper):
        return basis_func
    if basis_func in ("gauss", "slater"):
        return ExpBasisFuncHelper.from_function_type(basis_func)
    if isinstance(basis_func, (str, os.PathLike)) and os.path.exists(basis_func):
        return ExpBasisFuncHelper.from_file(basis_func)
    raise RuntimeError(f"Invalid func_type: {basis_func}!")


def compute_rho0(atnums, distances, pops, func_type="gauss", nderiv=0):
    """Pro-atom densities, one row per atom, at the given atom-to-point distances (natom, Npts)
    from the concatenated coefficients ``pops`` (scripts/generate_cube.py:140-157)."""
    helper = _helper(func_type)
    distances = np.asarray(distances, dtype=float)
    counts = [helper.get_nshell(int(z)) for z in atnums]
    if sum(counts) != len(pops):
        raise ValueError("the number of coefficients does not match the basis functions of the atoms")
    rho0 = np.zeros_like(distances)
    begin = 0
    for a, (z, k) in enumerate(zip(atnums, counts)):
        rho0[a] = helper.compute_proatom_dens(int(z), pops[begin : begin + k], distances[a], nderiv)
        begin += k
    return rho0


def aim_on_points(atnums, atcoords, points, density, propars, basis_func="gauss"):
    """(rho0 (natom, Npts), promolecule, w_a * density) on arbitrary points
    (scripts/generate_cube.py:213-227)."""
    points = np.asarray(points, dtype=float)
    distances = np.linalg.norm(points[None, :, :] - np.asarray(atcoords, dtype=float)[:, None, :], axis=2)
    rho0 = compute_rho0(atnums, distances, np.asarray(propars, dtype=float), basis_func)
    promol = rho0.sum(axis=0)
    promol += 1e-100
    return rho0, promol, rho0 / promol * np.asarray(density, dtype=float)[None, :]


def write_aim_cubes(prefix, atnums, atcorenums, atcoords, grid, density, propars, basis_func="gauss"):
    """The cube files of the reference program (scripts/generate_cube.py:239-271):
    ``<prefix>_rho_mol.cube``, per atom ``_rho_<a>.cube`` (AIM density) and ``_rho0_<a>.cube``
    (pro-atom), and ``_rho0_mol.cube`` (promolecule).  Returns the arrays."""
    rho0, promol, aim = aim_on_points(atnums, atcoords, grid.points, density, propars, basis_func)
    to_cube(f"{prefix}_rho_mol.cube", atnums, atcorenums, atcoords, grid, density)
    for a in range(len(atnums)):
        to_cube(f"{prefix}_rho_{a}.cube", atnums, atcorenums, atcoords, grid, aim[a])
        to_cube(f"{prefix}_rho0_{a}.cube", atnums, atcorenums, atcoords, grid, rho0[a])
    to_cube(f"{prefix}_rho0_mol.cube", atnums, atcorenums, atcoords, grid, promol)
    return {"rho0": rho0, "promol": promol, "aim_rho": aim}
