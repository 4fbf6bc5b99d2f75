"""Seeded synthetic systems for tests and the benchmark (BASELINE.json configs 2-5).

Densities are exact promolecules (sums of normalised Slater or Gaussian shells on the atoms), so
the partitioning has a known answer: MBIS recovers the generating populations (SURVEY.md
section 8c, "known-answer test for any size").  Geometry generators are plain NumPy and
deterministic in ``seed``.
"""

from __future__ import annotations

import numpy as np

__all__ = [
    "water_cluster",
    "organic_like",
    "peptide_like",
    "SLATER_SHELLS",
    "slater_promolecule_host",
    "expbasis_promolecule_host",
]

# (population N, Slater exponent S) per shell, rho(r) = N S^3 exp(-S r) / (8 pi)
SLATER_SHELLS = {
    1: ((0.70, 2.0),),
    6: ((1.70, 11.3), (4.20, 1.9)),
    7: ((1.70, 13.2), (5.40, 2.2)),
    8: ((1.65, 15.0), (6.95, 2.1)),
}

_R_OH = 1.8088  # bohr
_HOH = np.deg2rad(104.52)
_LATTICE = 5.86  # bohr


def _random_rotation(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array(
        [
            [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
            [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
            [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
        ]
    )


def water_cluster(natom: int, seed: int = 0):
    """``natom`` atoms (O,H,H,O,H,H,... truncated) of waters on a cubic 5.86-bohr lattice with
    random rigid orientations.  Returns (coordinates (natom,3), numbers int64 (natom,))."""
    nmol = -(-natom // 3)
    side = int(np.ceil(nmol ** (1.0 / 3.0)))
    rng = np.random.default_rng(seed)
    mono = np.array(
        [
            [0.0, 0.0, 0.0],
            [_R_OH * np.sin(_HOH / 2), 0.0, _R_OH * np.cos(_HOH / 2)],
            [-_R_OH * np.sin(_HOH / 2), 0.0, _R_OH * np.cos(_HOH / 2)],
        ]
    )
    coords = np.empty((nmol * 3, 3))
    m = 0
    for ix in range(side):
        for iy in range(side):
            for iz in range(side):
                if m == nmol:
                    break
                rot = _random_rotation(rng)
                coords[3 * m : 3 * m + 3] = mono @ rot.T + _LATTICE * np.array([ix, iy, iz], float)
                m += 1
    numbers = np.tile(np.array([8, 1, 1], dtype=np.int64), nmol)
    return coords[:natom].copy(), numbers[:natom].copy()


def _chain(natom, pattern, seed, bond=(2.0, 2.9), min_sep=1.9):
    """Self-avoiding random chain with bonded distances in ``bond`` (bohr)."""
    rng = np.random.default_rng(seed)
    coords = np.zeros((natom, 3))
    for i in range(1, natom):
        for _ in range(10000):
            step = rng.normal(size=3)
            step *= rng.uniform(*bond) / np.linalg.norm(step)
            trial = coords[rng.integers(max(0, i - 3), i)] + step
            if np.min(np.linalg.norm(coords[:i] - trial, axis=1)) >= min_sep:
                coords[i] = trial
                break
        else:  # pragma: no cover
            raise RuntimeError("could not place atom")
    numbers = np.array([pattern[i % len(pattern)] for i in range(natom)], dtype=np.int64)
    return coords, numbers


def organic_like(natom: int = 20, seed: int = 0):
    """~20-atom C6N2O2H10-like blob (config 2)."""
    pattern = [6, 1, 6, 1, 7, 1, 6, 1, 8, 1, 6, 1, 6, 1, 7, 1, 6, 1, 8, 1]
    return _chain(natom, pattern, seed)


def peptide_like(natom: int = 300, seed: int = 0):
    """Random-coil chain with H:C:N:O = 10:6:2:2 per 20 atoms (config 4)."""
    pattern = [6, 1, 7, 1, 6, 1, 8, 1, 6, 1, 6, 1, 7, 1, 6, 1, 8, 1, 6, 1]
    return _chain(natom, pattern, seed)


def slater_promolecule_host(points, coordinates, numbers, shells=SLATER_SHELLS, chunk=65536):
    """Exact Slater promolecule on ``points`` (host NumPy; small systems / tests only)."""
    out = np.zeros(len(points))
    for lo in range(0, len(points), chunk):
        p = points[lo : lo + chunk]
        acc = np.zeros(len(p))
        for R, z in zip(coordinates, numbers):
            r = np.sqrt(((p - R) ** 2).sum(axis=1))
            for N, S in shells[int(z)]:
                acc += N * S**3 * np.exp(-S * r) / (8 * np.pi)
        out[lo : lo + chunk] = acc
    return out


def expbasis_promolecule_host(points, coordinates, numbers, helper, scale=None, chunk=65536):
    """Promolecule sum_a sum_k c_ak g_ak with c proportional to the basis table's initials,
    scaled per element to ``scale[Z]`` electrons (default: Z electrons)."""
    out = np.zeros(len(points))
    for lo in range(0, len(points), chunk):
        p = points[lo : lo + chunk]
        acc = np.zeros(len(p))
        for R, z in zip(coordinates, numbers):
            z = int(z)
            r = np.sqrt(((p - R) ** 2).sum(axis=1))
            c = np.asarray(helper.get_initial(z), dtype=float)
            c = c / c.sum() * (float(z) if scale is None else scale[z])
            acc += helper.compute_proatom_dens(z, c, r)
        out[lo : lo + chunk] = acc
    return out


def slater_promolecule_device(grid, coordinates, numbers, shells=SLATER_SHELLS, device=None, shard=None):
    """Exact Slater promolecule and its Hirshfeld partition of unity on this rank's slab, computed
    with the fused promolecule kernel itself (large synthetic systems: 2,000 atoms x 58 M points
    is out of reach for host NumPy).  Returns (rho_local, owner_weight_local, point_lo, point_hi).
    """
    import numpy as _np

    from . import _lib
    from .core.device import GridSlab, ShellTable, stream_ptr, to_device

    zeros = _np.zeros(grid.size)
    slab = GridSlab(grid, zeros, _np.asarray(coordinates, float), device, shard, need_atgrids=False)
    counts = [len(shells[int(z)]) for z in numbers]
    table = ShellTable(slab, 1, counts)
    flat = _np.array([v for z in numbers for ns in shells[int(z)] for v in ns], dtype=float)
    _lib.call("hp_table_mbis", table.nshell, to_device(flat, slab.device), table.A, table.alpha,
              stream_ptr(slab.device))  # fmt: skip
    table.promol_weights(1e-15, True, True, False)
    rho = slab.promol.cpu().numpy()
    w = slab.at_w.cpu().numpy()
    return rho, w, slab.point_base, slab.point_base + slab.npts


def shell_promolecule_device(grid, coordinates, functor, counts, A, alpha, order=None, device=None, shard=None):
    """Promolecule sum_shells A exp(-alpha r^n) and its Hirshfeld owner weights on this rank's slab for
    an arbitrary shell table (used to synthesise Gaussian / Slater-basis test densities on the GPU)."""
    import numpy as _np

    from .core.device import GridSlab, ShellTable, to_device

    slab = GridSlab(grid, _np.zeros(grid.size), _np.asarray(coordinates, float), device, shard, need_atgrids=False)
    table = ShellTable(slab, functor, counts)
    table.A.copy_(to_device(_np.asarray(A, float), slab.device))
    table.alpha.copy_(to_device(_np.asarray(alpha, float), slab.device))
    if functor == 3:
        table.order.copy_(to_device(_np.asarray(order, float), slab.device))
    table.promol_weights(1e-15, True, True, False)
    return slab.promol.cpu().numpy(), slab.at_w.cpu().numpy(), slab.point_base, slab.point_base + slab.npts


def expbasis_promolecule_device(grid, coordinates, numbers, helper, scale=None, device=None, shard=None):
    """Device version of ``expbasis_promolecule_host``; returns (rho, owner weights) on the slab."""
    from .core.basis import shell_norm

    counts, A, alpha, order = [], [], [], []
    for z in numbers:
        z = int(z)
        c = np.asarray(helper.get_initial(z), dtype=float)
        c = c / c.sum() * (float(z) if scale is None else scale[z])
        n = np.asarray(helper.get_order(z), dtype=float)
        al = np.asarray(helper.get_exponent(z), dtype=float)
        counts.append(len(al))
        A.append(c * shell_norm(n, al))
        alpha.append(al)
        order.append(n)
    order = np.concatenate(order)
    functor = 2 if np.all(order == 2.0) else (1 if np.all(order == 1.0) else 3)
    rho, w, lo, hi = shell_promolecule_device(grid, coordinates, functor, counts, np.concatenate(A),
                                              np.concatenate(alpha), order, device, shard)
    return rho, w


def contraction_map(n=12, seed=1):
    """Seeded linear part (A, b) of a contraction x -> A x + b + 0.05 sin(x) with ||A||_2 = 0.4:
    the toy fixed-point problem of the DIIS / CDIIS known-answer tests."""
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(n, n))
    A = 0.4 * A / np.linalg.norm(A, 2)
    return A, rng.normal(size=n)


def radial_problem(helper, number, population, nrad=120):
    """A seeded 1-D aLISA problem on an exponential radial grid: tabulated unit-population basis
    functions (K, nrad), a target density (random non-negative coefficients plus a Slater tail that
    is outside the basis), start coefficients, radial points and 4 pi r^2 weights."""
    r = 5e-4 * np.exp(np.arange(nrad) * np.log(2e1 / 5e-4) / (nrad - 1))
    w = 4 * np.pi * r**2 * np.gradient(r)
    K = helper.get_nshell(number)
    bs = np.array([helper.compute_proshell_dens(number, k, 1.0, r) for k in range(K)])
    rng = np.random.default_rng(int(number))
    c = rng.random(K)
    c *= population / c.sum()
    rho = c @ bs + 1e-3 * np.exp(-1.3 * r)
    return bs, rho, np.full(K, population / K), r, w
