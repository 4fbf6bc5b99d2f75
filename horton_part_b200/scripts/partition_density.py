"""``part-dens`` on the B200: NPZ in (the reference's ``part-gen`` wire format) -> partitioning on the
GPU -> NPZ out (the reference's ``part-dens`` output keys).

Row (f)-1 of SURVEY.md section 8.  Input keys consumed (written by the reference's
``scripts/generate_density.py:232-257``): ``atcoords, atnums, atcorenums, density, aim_weights,
points, weights, atom_idxs, atom{i}/points, atom{i}/weights, atom{i}/shell_idxs,
atom{i}/rgrid/points, atom{i}/rgrid/weights``.  Unlike the reference's consumer
(``scripts/partition_density.py:65-85``), which rebuilds every AtomGrid with qc-grid, the stored
points and weights are used as they are, so no Lebedev tables are needed and the density stays
aligned with the points whatever produced them.  Output keys as in
``scripts/partition_density.py:242-263`` (checked by the reference's tests/scripts/test_main.py:101-119).

    python -m horton_part_b200.scripts.partition_density config.yaml [--skip_exist_files]

``config.yaml`` has the reference's layout (optional top-level ``part-dens:`` section; keys
``inputs, outputs, log_files, type, basis_func, func_file, maxiter, threshold, inner_threshold, lmax,
solver, solver_options, exp_n_dict, nshell_dict, part_job_type, grid_type, log_level``).
"""

from __future__ import annotations

import argparse
import logging
import os
import sys

import numpy as np

from ..core.logging import setup_logger
from ..utils import wpart_schemes

__all__ = ["construct_molgrid_from_dict", "write_part_gen_npz", "main", "DEFAULTS", "CLASS_ARGS"]

# data/part-dens.yaml of the reference
DEFAULTS = {
    "type": "lisa", "basis_func": "gauss", "func_file": None, "maxiter": 1000, "threshold": 1.0e-6,
    "inner_threshold": 1.0e-8, "lmax": 3, "log_level": "INFO", "solver": "sc", "exp_n_dict": None,
    "nshell_dict": None, "part_job_type": "do_partitioning", "grid_type": 1,
}  # fmt: skip

_COMMON = ["maxiter", "threshold", "lmax", "density_cutoff", "population_cutoff", "negative_cutoff"]
# data/keywords.yaml of the reference: constructor arguments each scheme accepts from the settings
CLASS_ARGS = {
    "is": _COMMON + ["inner_threshold"],
    "mbis": _COMMON + ["inner_threshold", "grid_type"],
    "gisa": _COMMON + ["inner_threshold", "solver", "solver_options", "grid_type"],
    "lisa": _COMMON + ["inner_threshold", "solver", "solver_options", "grid_type", "basis_func", "basis_type"],
    "nlis": _COMMON + ["inner_threshold", "exp_n_dict", "nshell_dict", "grid_type"],
    "gmbis": _COMMON + ["inner_threshold", "exp_n_dict", "grid_type"],
    "glisa": _COMMON + ["solver", "solver_options", "grid_type", "basis_func", "basis_type"],
}


class _Stored:
    """Points + weights container with the attributes the partitioning classes read."""

    def __init__(self, points, weights):
        self.points = np.ascontiguousarray(points, dtype=np.float64)
        self.weights = np.ascontiguousarray(weights, dtype=np.float64)

    @property
    def size(self):
        return self.weights.shape[0]

    def integrate(self, *arrays):
        return np.einsum("i" + ",i" * len(arrays), self.weights, *arrays)


class _StoredAtomGrid(_Stored):
    def __init__(self, points, weights, rgrid, indices, center):
        super().__init__(points, weights)
        self.rgrid, self.center = rgrid, np.asarray(center, dtype=float)
        self.indices = np.asarray(indices, dtype=np.int64)
        # Lebedev degree of every radial shell from its point count (qc-grid AtomGrid.degrees / l_max,
        # read by do_density_decomposition, core/base.py:646-657)
        from ..gridlite import _DEGREE_OF_SIZE

        sizes = np.diff(self.indices)
        unknown = sorted(set(int(n) for n in sizes) - set(_DEGREE_OF_SIZE))
        if unknown:
            raise ValueError(f"shell sizes {unknown} are not Lebedev-Laikov grid sizes")
        self.degrees = [_DEGREE_OF_SIZE[int(n)] for n in sizes]

    @property
    def l_max(self):
        return max(self.degrees)

    @property
    def n_shells(self):
        return len(self.degrees)


class _StoredMolGrid(_Stored):
    def __init__(self, points, weights, aim_weights, indices, atgrids, atweights):
        super().__init__(points, weights)
        self.aim_weights = np.asarray(aim_weights, dtype=float)
        self.indices = np.asarray(indices, dtype=np.int64)
        self.atgrids = atgrids
        self.atweights = atweights


def construct_molgrid_from_dict(data):
    """Molecular grid from a ``part-gen`` NPZ (or dict with the same keys)."""
    natom = len(data["atnums"])
    atgrids, atw = [], []
    for a in range(natom):
        rgrid = _Stored(data[f"atom{a}/rgrid/points"], data[f"atom{a}/rgrid/weights"])
        shell_idxs = np.asarray(data[f"atom{a}/shell_idxs"], dtype=np.int64)
        if f"atom{a}/points" in data:
            pts, wts = data[f"atom{a}/points"], data[f"atom{a}/weights"]
        else:  # older files without the per-atom arrays: rebuild with the Lebedev tables
            from .. import gridlite

            g = gridlite.AtomGrid(gridlite.OneDGrid(rgrid.points, rgrid.weights), sizes=np.diff(shell_idxs),
                                  center=data["atcoords"][a])  # fmt: skip
            pts, wts = g.points, g.weights
        atgrids.append(_StoredAtomGrid(pts, wts, rgrid, shell_idxs, data["atcoords"][a]))
        atw.append(np.asarray(wts, dtype=float))
    atweights = np.concatenate(atw)
    aim = np.asarray(data["aim_weights"], dtype=float)
    if "points" in data and "atom_idxs" in data:
        points, indices = data["points"], data["atom_idxs"]
        weights = data["weights"] if "weights" in data else atweights * aim
    else:
        points = np.concatenate([g.points for g in atgrids])
        indices = np.concatenate([[0], np.cumsum([g.size for g in atgrids])])
        weights = atweights * aim
    return _StoredMolGrid(points, weights, aim, indices, atgrids, atweights)


def write_part_gen_npz(filename, coordinates, numbers, pseudo_numbers, grid, density, nelec=None):
    """Write the ``part-gen`` wire format from arrays (the reference needs iodata + gbasis for this)."""
    data = {
        "atcoords": np.asarray(coordinates), "atnums": np.asarray(numbers),
        "atcorenums": np.asarray(pseudo_numbers, dtype=float), "density": np.asarray(density),
        "points": grid.points, "weights": grid.weights, "aim_weights": grid.aim_weights,
        "cellvecs": np.zeros((0, 3)), "atom_idxs": np.asarray(grid.indices),
        "nelec": np.float64(grid.integrate(np.asarray(density)) if nelec is None else nelec),
    }  # fmt: skip
    for a, g in enumerate(grid.atgrids):
        data[f"atom{a}/points"] = g.points
        data[f"atom{a}/weights"] = g.weights
        data[f"atom{a}/shell_idxs"] = np.asarray(g.indices)
        data[f"atom{a}/rgrid/points"] = g.rgrid.points
        data[f"atom{a}/rgrid/weights"] = g.rgrid.weights
    folder = os.path.dirname(os.path.abspath(filename))
    os.makedirs(folder, exist_ok=True)
    np.savez_compressed(filename, **data)


def _prepare_exp_n_dict(spec):
    out = {}
    for z, values in (spec or {}).items():
        for k, v in enumerate(values):
            out[(int(z), k)] = float(v)
    return out


def single_launch(settings, fn_in, fn_out, fn_log, logger):
    setup_logger(logger, getattr(logging, settings.get("log_level", "INFO")), fn_log, overwrite=False)
    kind = settings["type"]
    if kind not in CLASS_ARGS:
        raise NotImplementedError(f"scheme {kind!r} is not available through part-dens on the B200 path")
    logger.info(f"Load grid and density data from {fn_in} ...")
    data = np.load(fn_in)
    grid = construct_molgrid_from_dict(data)
    kwargs = {
        "coordinates": np.asarray(data["atcoords"], dtype=float),
        "numbers": np.asarray(data["atnums"]).astype(np.int64),
        "pseudo_numbers": np.asarray(data["atcorenums"], dtype=float),
        "grid": grid, "moldens": np.asarray(data["density"], dtype=float), "logger": logger,
    }  # fmt: skip
    for key in CLASS_ARGS[kind]:
        if key in settings:
            kwargs[key] = settings[key]
    if "basis_func" in CLASS_ARGS[kind]:
        kwargs["basis_func"] = settings.get("func_file") or settings.get("basis_func")
    if "exp_n_dict" in CLASS_ARGS[kind]:
        kwargs["exp_n_dict"] = _prepare_exp_n_dict(settings.get("exp_n_dict"))
    if "nshell_dict" in CLASS_ARGS[kind]:
        kwargs["nshell_dict"] = settings.get("nshell_dict")
    part = wpart_schemes(kind)(**kwargs)
    try:
        getattr(part, settings["part_job_type"])()
    except RuntimeError as exc:  # as the reference: report and return 1
        logger.info(exc)
        return 1
    cache = part.cache
    out = {
        "natom": len(data["atnums"]), "atnums": data["atnums"], "atcorenums": data["atcorenums"],
        "type": kind, "lmax": settings["lmax"], "maxiter": settings["maxiter"],
        "threshold": settings["threshold"],
        "inner_threshold": settings["inner_threshold"] if kind != "glisa" else np.nan,
        "solver": settings["solver"], "charges": cache["charges"],
        "time": part.time_usage["do_partitioning"],
        "time_update_at_weights": cache["time_update_at_weights"],
        "time_update_propars": cache["time_update_propars"], "niter": cache["niter"],
        "history_charges": cache["history_charges"], "history_propars": cache["history_propars"],
        "history_entropies": cache["history_entropies"], "history_changes": cache["history_changes"],
    }  # fmt: skip
    if settings["part_job_type"] == "do_density_decomposition":
        # scripts/partition_density.py:290-322 of the reference: radial projections and the final basis
        # table (order, exponent, population) of every atom
        propars = np.asarray(cache["history_propars"])[-1, :]
        for a in range(part.natom):
            for key in (f"radial_points_{a}", f"spherical_average_{a}", f"radial_weights_{a}"):
                out[key] = cache[key]
            mine = propars[part._ranges[a] : part._ranges[a + 1]]
            if kind in ("gisa", "lisa", "glisa"):
                helper, z = part.bs_helper, int(part.numbers[a])
                info = np.asarray([helper.orders[z], helper.exponents[z], helper.initials[z]], dtype=float).T
                info[:, -1] = mine
            elif kind == "mbis":
                mine = mine.reshape((-1, 2))
                info = np.ones((mine.shape[0], 3))
                info[:, 1], info[:, 2] = mine[:, 1], mine[:, 0]
            elif kind in ("gmbis", "nlis"):
                mine = mine.reshape((-1, 3))
                info = np.stack([mine[:, 2], mine[:, 1], mine[:, 0]], axis=1)
            elif kind == "is":
                info = mine
            else:
                raise NotImplementedError
            out[f"bs_info_{a}"] = info
    for key in settings.get("save") or []:
        # nested attributes of the partitioning object in dot notation first, cache keys second
        # (scripts/partition_density.py:324-340)
        if isinstance(key, list):
            key, value = tuple(key), None
        else:
            value = part
            for attr in key.split("."):
                value = getattr(value, attr, None)
                if value is None:
                    break
        if isinstance(value, np.ndarray):
            out[f"save/part.{key}"] = value
        if value is None and key in cache:
            out[f"save/part.cache/{key}"] = cache[key]
    os.makedirs(os.path.dirname(os.path.abspath(fn_out)), exist_ok=True)
    np.savez_compressed(fn_out, **out)
    return 0


def main(args=None) -> int:
    import yaml

    parser = argparse.ArgumentParser(prog="part-dens", description="Partition densities on the B200 path")
    parser.add_argument("config_file", type=str, help="Use configure file.")
    parser.add_argument("--skip_exist_files", action="store_true",
                        help="Skip the calculation if the output files and log files exist")  # fmt: skip
    ns = parser.parse_args(args)
    with open(ns.config_file) as fh:
        settings = yaml.safe_load(fh)
    settings = settings["part-dens"] if "part-dens" in settings else settings
    for key, value in DEFAULTS.items():
        settings.setdefault(key, value)
    inputs = settings["inputs"]
    assert isinstance(inputs, list)
    settings.setdefault("outputs", [f"output_{i+1}.npz" for i in range(len(inputs))])
    settings.setdefault("log_files", [None] * len(inputs))
    if not (len(inputs) == len(settings["outputs"]) == len(settings["log_files"])):
        raise RuntimeError("The settings for part-dens is not fully correct.")
    logger = logging.getLogger("part-dens")
    status = 0
    for fn_in, fn_out, fn_log in zip(inputs, settings["outputs"], settings["log_files"]):
        if ns.skip_exist_files and os.path.exists(fn_out) and (fn_log is None or os.path.exists(fn_log)):
            print(f"Skip the calculations with input: {fn_in}")
            continue
        status |= single_launch(settings, fn_in, fn_out, fn_log, logger)
    return status


if __name__ == "__main__":
    sys.exit(main())
