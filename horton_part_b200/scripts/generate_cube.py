"""AIM quantities on a uniform grid (device kernel ``hp_aim_on_points``) and Gaussian cube files.

Counterpart of the AIM half of the reference's ``part-cube`` program
(scripts/generate_cube.py:100-157, 206-272): from the converged pro-atom coefficients of an aLISA /
gLISA run it re-evaluates every pro-atom on the points of a uniform grid, forms the promolecule
(+1e-100), the weight functions w_a = rho0_a / rho0 and the atoms-in-molecules densities
w_a * rho, and writes them as cube files.  The other half of the reference program (the molecular
density on the uniform grid from a wavefunction, through iodata / gbasis) is outside this
repository: the density comes in as an array.

The O(natom * Npts * K) evaluation of the pro-atoms, the promolecule and the AIM densities runs on the
GPU with the same shell table and functors as the partitioning kernels (``aim_on_points``); writing the
text files is host work.  ``compute_rho0`` mirrors the reference's small ``_compute_rho0`` helper
(distances in, pro-atoms out) on the host for API users.
"""

from __future__ import annotations

import os

import numpy as np

from ..core.basis import ExpBasisFuncHelper

__all__ = ["UniformGrid", "to_cube", "read_cube", "compute_rho0", "aim_on_points", "write_aim_cubes", "main"]


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


def aim_on_points(atnums, atcoords, points, density, propars, basis_func="gauss", device=None, chunk=1 << 22):
    """(rho0 (natom, Npts), promolecule, w_a * density) on arbitrary points, evaluated on the GPU
    (scripts/generate_cube.py:213-227 of the reference: basis evaluation per atom, sum over atoms + 1e-100,
    rho0 / promol * density).  Points are processed in chunks of ``chunk`` so that the natom x chunk
    device buffers stay small; the results come back as NumPy arrays."""
    import torch

    from .. import _lib
    from ..core.basis import shell_norm
    from ..core.device import require_cuda, stream_ptr, to_device

    dev = require_cuda(device)
    helper = _helper(basis_func)
    atnums = np.asarray(atnums)
    points = np.ascontiguousarray(points, dtype=np.float64)
    density = np.ascontiguousarray(density, dtype=np.float64)
    propars = np.asarray(propars, dtype=np.float64)
    counts = np.array([helper.get_nshell(int(z)) for z in atnums], dtype=np.int64)
    if counts.sum() != len(propars):
        raise ValueError("the number of coefficients does not match the basis functions of the atoms")
    orders = np.concatenate([np.asarray(helper.get_order(int(z)), float) for z in atnums])
    alphas = np.concatenate([np.asarray(helper.get_exponent(int(z)), float) for z in atnums])
    functor = 2 if np.all(orders == 2.0) else (1 if np.all(orders == 1.0) else 3)
    natom, npts = len(atnums), len(points)
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    max_atoms, max_shells = np.zeros(1, np.int32), np.zeros(1, np.int32)
    _lib.call("hp_tile_limits", max_atoms, max_shells)
    tiles, a = [0], 0
    while a < natom:
        b, nsh = a, 0
        while b < natom and b - a < int(max_atoms[0]) and nsh + counts[b] <= int(max_shells[0]):
            nsh += counts[b]
            b += 1
        if b == a:
            raise ValueError("an atom has more shells than one shared-memory tile can hold")
        tiles.append(b)
        a = b
    d_xyz = to_device(np.asarray(atcoords, dtype=np.float64), dev)
    d_off = to_device(offsets, dev, np.int32)
    d_A = to_device(propars * shell_norm(orders, alphas), dev)  # c_k N(alpha_k, n_k), core/basis.py:161
    d_alpha = to_device(alphas, dev)
    d_order = to_device(orders, dev) if functor == 3 else None
    d_tiles = to_device(np.asarray(tiles, dtype=np.int32), dev, np.int32)
    rho0 = np.empty((natom, npts))
    aim = np.empty((natom, npts))
    promol = np.empty(npts)
    for lo in range(0, npts, chunk):
        hi = min(lo + chunk, npts)
        n = hi - lo
        pts = to_device(points[lo:hi].T.copy(), dev)  # structure of arrays
        dens = to_device(density[lo:hi], dev)
        d_rho0 = torch.empty((natom, n), dtype=torch.float64, device=dev)
        d_aim = torch.empty_like(d_rho0)
        d_pro = torch.empty(n, dtype=torch.float64, device=dev)
        _lib.call("hp_aim_on_points", functor, n, pts[0], pts[1], pts[2], natom, d_xyz, d_off, d_A, d_alpha, d_order,
                  len(tiles) - 1, d_tiles, dens, 1e-100, d_rho0, d_pro, d_aim, stream_ptr(dev))  # fmt: skip
        rho0[:, lo:hi] = d_rho0.cpu().numpy()
        aim[:, lo:hi] = d_aim.cpu().numpy()
        promol[lo:hi] = d_pro.cpu().numpy()
    return rho0, promol, aim


def write_aim_cubes(prefix, atnums, atcorenums, atcoords, grid, density, propars=None, basis_func="gauss",
                    arrays=None, device=None):
    """The cube files of the reference program (scripts/generate_cube.py:239-271):
    ``<prefix>_rho_mol.cube``, per atom ``_rho_<a>.cube`` (AIM density) and ``_rho0_<a>.cube``
    (pro-atom), and ``_rho0_mol.cube`` (promolecule).  The arrays come from ``aim_on_points`` (GPU) unless
    ``arrays = (rho0, promol, aim_rho)`` is given.  Returns the arrays."""
    if arrays is None:
        arrays = aim_on_points(atnums, atcoords, grid.points, density, propars, basis_func, device=device)
    rho0, promol, aim = arrays
    to_cube(f"{prefix}_rho_mol.cube", atnums, atcorenums, atcoords, grid, density)
    for a in range(len(atnums)):
        to_cube(f"{prefix}_rho_{a}.cube", atnums, atcorenums, atcoords, grid, aim[a])
        to_cube(f"{prefix}_rho0_{a}.cube", atnums, atcorenums, atcoords, grid, rho0[a])
    to_cube(f"{prefix}_rho0_mol.cube", atnums, atcorenums, atcoords, grid, promol)
    return {"rho0": rho0, "promol": promol, "aim_rho": aim}


def main(args=None) -> int:
    """``part-cube``, AIM half: YAML settings ``inputs`` (NPZ files with ``atnums, atcorenums, atcoords,
    origin, axes, shape, density`` on the uniform grid -- what the reference program stores after its
    wavefunction step), ``partdens`` (the matching ``part-dens`` outputs: ``history_propars``), ``outputs``,
    ``basis_func``, ``with_cube_files``, ``with_aim_cache`` (scripts/generate_cube.py:166-272)."""
    import argparse
    import sys

    import yaml

    parser = argparse.ArgumentParser(prog="part-cube", description="AIM densities on a uniform grid (B200 path)")
    parser.add_argument("config_file", type=str, help="Use configure file.")
    ns = parser.parse_args(args)
    with open(ns.config_file) as fh:
        settings = yaml.safe_load(fh)
    settings = settings.get("part-cube", settings)
    inputs, partdens = settings["inputs"], settings["partdens"]
    outputs = settings.get("outputs") or [f"cube_{i + 1}.npz" for i in range(len(inputs))]
    if not (len(inputs) == len(partdens) == len(outputs)):
        print("The settings for part-cube is not fully correct.", file=sys.stderr)
        return 1
    for fn_in, fn_part, fn_out in zip(inputs, partdens, outputs):
        data = dict(np.load(fn_in))
        grid = UniformGrid(data["origin"], data["axes"], data["shape"])
        propars = np.load(fn_part)["history_propars"][-1, :]
        rho0, promol, aim = aim_on_points(data["atnums"], data["atcoords"], grid.points, data["density"], propars,
                                          settings.get("basis_func", "gauss"))  # fmt: skip
        if settings.get("with_aim_cache", True):
            data.update({"rho0": rho0, "aim_rho": aim})
        os.makedirs(os.path.dirname(os.path.abspath(fn_out)), exist_ok=True)
        np.savez_compressed(fn_out, **data)
        if settings.get("with_cube_files", True):
            prefix = ".".join(str(fn_out).split(".")[:-1])
            write_aim_cubes(prefix, data["atnums"], data["atcorenums"], data["atcoords"], grid, data["density"],
                            arrays=(rho0, promol, aim))  # fmt: skip
    return 0


if __name__ == "__main__":
    import sys

    sys.exit(main())
