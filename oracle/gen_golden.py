"""Generate tests/golden/* by running the UNMODIFIED reference (/root/reference/src/horton_part)
through oracle/qcgrid_shim in the build container.  The reference tree does not exist on the GPU
box, so the vectors are committed; this script is the record of how they were made.

    python oracle/gen_golden.py            # all cases
    python oracle/gen_golden.py h2o        # one case

Inputs that cannot be regenerated on the GPU box (the water HF/STO-3G density from the
reference's tests/cached fixture, re-ordered onto the rebuilt grid) are stored alongside.
"""

from __future__ import annotations

import logging
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "oracle" / "qcgrid_shim"), "/root/reference/src", str(ROOT)]

import grid as qcgrid  # noqa: E402  (the shim)
from horton_part.utils import wpart_schemes  # noqa: E402  (the reference itself)

from horton_part_b200 import synthetic  # noqa: E402  (pure-NumPy generators only)

GOLD = ROOT / "tests" / "golden"
REF_CACHED = pathlib.Path("/root/reference/tests/cached")
logging.disable(logging.CRITICAL)


def run_reference(scheme, coords, numbers, pseudo, grid, rho, **kwargs):
    t0 = time.time()
    part = wpart_schemes(scheme)(coords, numbers, pseudo, grid, rho, **kwargs)
    part.do_charges()
    out = {
        "niter": np.int64(part["niter"]),
        "charges": part["charges"],
        "propars": part["propars"],
        "history_changes": part["history_changes"],
        "history_entropies": part["history_entropies"],
        "history_charges": part["history_charges"],
        "promoldens_sample": part["promoldens"][::97].copy(),
        "seconds": np.float64(time.time() - t0),
    }
    for a in range(min(len(numbers), 3)):
        w = part[f"at_weights_{a}"]
        if w.shape == grid.weights.shape:
            w = w[grid.indices[a] : grid.indices[a + 1]]
        out[f"at_weights_{a}_sample"] = w[::53].copy()
    for key in ("core_charges", "valence_charges", "valence_widths"):
        if key in part.cache:
            out[key] = part[key]
    if "spherical_average_0" in part.cache:
        out["spherical_average_0"] = part["spherical_average_0"]
    if scheme == "mbis" and kwargs.get("grid_type", 1) == 1:
        import contextlib
        import io

        with contextlib.redirect_stdout(io.StringIO()):
            part.do_moments()  # core/base.py:329-402 (prints a progress line)
        for key in ("cartesian_multipoles", "pure_multipoles", "radial_moments"):
            out[key] = part[key]
    return out


def save(name, prefix_results, **extra):
    flat = dict(extra)
    for prefix, res in prefix_results.items():
        for k, v in res.items():
            flat[f"{prefix}/{k}"] = v
    GOLD.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(GOLD / name, **flat)
    print("wrote", GOLD / name, f"{(GOLD / name).stat().st_size / 1024:.0f} KiB")


H2O_SCHEMES = {
    "mbis": ("mbis", {}),
    "isa": ("is", {}),
    "lisa_sc_gauss": ("lisa", dict(solver="sc")),
    "lisa_sc_slater": ("lisa", dict(solver="sc", basis_func="slater")),
    "nlis": ("nlis", dict(exp_n_dict={})),
    "gmbis": ("gmbis", dict(exp_n_dict={})),
    "glisa_sc": ("glisa", dict(solver="sc")),
    "mbis_gt2": ("mbis", dict(grid_type=2)),
    "lisa_sc_gt2": ("lisa", dict(solver="sc", grid_type=2)),
    "nlis_gt2": ("nlis", dict(exp_n_dict={}, grid_type=2)),
    "glisa_sc_gt2": ("glisa", dict(solver="sc", grid_type=2)),
}


def case_h2o():
    """Config 1: water HF/STO-3G, ExpRTransform(5e-4, 2e1, 119) x 110 Lebedev (tests/test_wpart.py:38-56)."""
    from scipy.spatial import cKDTree

    npz = np.load(REF_CACHED / "water_sto3g_hf_g03_fchk_exp:5e-4:2e1:120:110.npz")
    coords, numbers, pseudo = npz["coordinates"], npz["numbers"], npz["pseudo_numbers"]
    rgrid = qcgrid.ExpRTransform(5e-4, 2e1, 119).transform_1d_grid(qcgrid.UniformInteger(120))
    grid = qcgrid.MolGrid.from_size(numbers, coords, 110, rgrid, qcgrid.BeckeWeights(), rotate=False, store=True)
    dist, order = cKDTree(npz["points"]).query(grid.points)
    assert dist.max() < 1e-12 and len(set(order)) == grid.size
    rho = npz["dens"][order]
    results = {}
    for tag, (scheme, kw) in H2O_SCHEMES.items():
        try:
            results[tag] = run_reference(scheme, coords, numbers, pseudo, grid, rho, **kw)
            print(f"  h2o {tag}: niter={results[tag]['niter']} q={results[tag]['charges']}")
        except Exception as exc:  # reference behaviour on this density (e.g. singular Newton)
            print(f"  h2o {tag}: reference raised {type(exc).__name__}: {exc}")
    save(
        "h2o_hf_sto3g.npz", results, coordinates=coords, numbers=numbers, pseudo_numbers=pseudo,
        dens=rho, aim_weights=grid.aim_weights, nelec=np.float64(grid.integrate(rho)),
        grid_spec=np.array("ExpRTransform(5e-4,2e1,119) o UniformInteger(120) x Lebedev110, BeckeWeights()"),
    )  # fmt: skip


def load_proatom_records(level="hf_sto3g", numbers=(1, 6, 8)):
    """The reference's cached isolated-atom records (tests/common.py:64-111, PowerRTransform grids)."""
    from horton_part.core.proatomdb import ProAtomRecord

    records, raw = [], {}
    for z in numbers:
        for path in sorted(REF_CACHED.glob(f"atom_{level}_Z{z:02d}_N*_pow.npz")):
            with np.load(path) as f:
                number, charge, energy = int(f["number"]), int(f["charge"]), float(f["energy"])
                rmin, rmax, npoint = f["rgrid"]
                rgrid = qcgrid.PowerRTransform(rmin, rmax, int(npoint) - 1).transform_1d_grid(qcgrid.UniformInteger(int(npoint)))
                records.append(ProAtomRecord(number, charge, energy, rgrid, f["dens"], f["deriv"]))
                raw[f"Z{number}_q{charge}"] = np.concatenate([[number, charge, energy, rmin, rmax, npoint], f["dens"], f["deriv"]])
    return records, raw


def case_hirshfeld():
    """Hirshfeld and Hirshfeld-I on water (tests/test_wpart.py:70-87) with the reference's database."""
    from horton_part.core.proatomdb import ProAtomDB
    from scipy.spatial import cKDTree

    npz = np.load(REF_CACHED / "water_sto3g_hf_g03_fchk_exp:5e-4:2e1:120:110.npz")
    coords, numbers, pseudo = npz["coordinates"], npz["numbers"], npz["pseudo_numbers"]
    rgrid = qcgrid.ExpRTransform(5e-4, 2e1, 119).transform_1d_grid(qcgrid.UniformInteger(120))
    grid = qcgrid.MolGrid.from_size(numbers, coords, 110, rgrid, qcgrid.BeckeWeights(), rotate=False, store=True)
    _, order = cKDTree(npz["points"]).query(grid.points)
    rho = npz["dens"][order]
    records, raw = load_proatom_records()
    results = {}
    for tag, scheme in (("h", "h"), ("hi", "hi")):
        part = wpart_schemes(scheme)(coords, numbers, pseudo, grid, rho, proatomdb=ProAtomDB(records))
        part.do_charges()
        out = {"charges": part["charges"], "promoldens_sample": part["promoldens"][::97].copy(),
               "at_weights_0_sample": part["at_weights_0"][::53].copy()}
        if scheme == "hi":
            out.update(niter=np.int64(part["niter"]), history_changes=part["history_changes"],
                       history_charges=part["history_charges"], history_entropies=part["history_entropies"])
        results[tag] = out
        print(f"  h2o {tag}: q={part['charges']} niter={out.get('niter')}")
    for gt in (2, 3):  # Hirshfeld on the molecular grid (Hirshfeld-I raises ValueError there in the reference)
        part = wpart_schemes("h")(coords, numbers, pseudo, grid, rho, proatomdb=ProAtomDB(records), grid_type=gt)
        part.do_charges()
        results[f"h_gt{gt}"] = {"charges": part["charges"], "promoldens_sample": part["promoldens"][::97].copy()}
        print(f"  h2o h grid_type {gt}: q={part['charges']}")
    save("h2o_hirshfeld.npz", results, **{f"record/{k}": v for k, v in raw.items()})


def synthetic_grid(coords, numbers, nrad, nang):
    rgrid = qcgrid.BeckeRTransform(1e-4, 1.5).transform_1d_grid(qcgrid.GaussChebyshev(nrad))
    return qcgrid.MolGrid.from_size(numbers, coords, nang, rgrid, qcgrid.BeckeWeights(), rotate=0, store=True)


def case_water_cluster(natom=6, nrad=40, nang=50, seed=0):
    """Synthetic Slater promolecule water cluster (BASELINE.md section 2 generator), small grid."""
    coords, numbers = synthetic.water_cluster(natom, seed)
    grid = synthetic_grid(coords, numbers, nrad, nang)
    rho = synthetic.slater_promolecule_host(grid.points, coords, numbers)
    pseudo = numbers.astype(float)
    results = {}
    for tag, (scheme, kw) in {
        "mbis": ("mbis", {}),
        "isa": ("is", dict(maxiter=60)),
        "lisa_sc_gauss": ("lisa", dict(solver="sc")),
        "nlis": ("nlis", dict(exp_n_dict={})),
        "glisa_sc": ("glisa", dict(solver="sc", maxiter=500)),
    }.items():
        try:
            results[tag] = run_reference(scheme, coords, numbers, pseudo, grid, rho, **kw)
            print(f"  water{natom} {tag}: niter={results[tag]['niter']} q={results[tag]['charges'][:3]}")
        except Exception as exc:
            print(f"  water{natom} {tag}: reference raised {type(exc).__name__}: {exc}")
    save(
        f"water{natom}_slater.npz", results, coordinates=coords, numbers=numbers, pseudo_numbers=pseudo,
        dens_sample=rho[::101].copy(), nelec=np.float64(grid.integrate(rho)),
        grid_spec=np.array(f"BeckeRTransform(1e-4,1.5) o GaussChebyshev({nrad}) x Lebedev{nang}, BeckeWeights(); seed={seed}"),
    )  # fmt: skip


def case_water_gauss(natom=6, nrad=40, nang=50, seed=0):
    """Synthetic Gaussian promolecule (sum_a sum_k c_ak g_ak, c from the gauss table's initials scaled
    to 8.6 / 0.7 electrons): the exact-Newton gLISA solver converges on it."""
    from horton_part.core.basis import ExpBasisFuncHelper

    coords, numbers = synthetic.water_cluster(natom, seed)
    grid = synthetic_grid(coords, numbers, nrad, nang)
    helper = ExpBasisFuncHelper.from_function_type("gauss")
    rho = synthetic.expbasis_promolecule_host(grid.points, coords, numbers, helper, scale={8: 8.6, 1: 0.7})
    pseudo = numbers.astype(float)
    results = {}
    for tag, (scheme, kw) in {
        "glisa_newton": ("glisa", dict(solver="newton")),
        "glisa_sc": ("glisa", dict(solver="sc")),
        "lisa_sc_gauss": ("lisa", dict(solver="sc")),
    }.items():
        try:
            results[tag] = run_reference(scheme, coords, numbers, pseudo, grid, rho, **kw)
            print(f"  water{natom}g {tag}: niter={results[tag]['niter']} q={results[tag]['charges'][:3]}")
        except Exception as exc:
            print(f"  water{natom}g {tag}: reference raised {type(exc).__name__}: {exc}")
    save(
        f"water{natom}_gauss.npz", results, coordinates=coords, numbers=numbers, pseudo_numbers=pseudo,
        dens_sample=rho[::101].copy(), nelec=np.float64(grid.integrate(rho)),
        grid_spec=np.array(f"BeckeRTransform(1e-4,1.5) o GaussChebyshev({nrad}) x Lebedev{nang}, BeckeWeights(); seed={seed}"),
    )  # fmt: skip


def run_reference_light(scheme, coords, numbers, pseudo, grid, rho, **kwargs):
    """Reference run keeping only what the solver parity tests compare."""
    import contextlib
    import io
    import warnings

    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        part = wpart_schemes(scheme)(coords, numbers, pseudo, grid, rho, **kwargs)
        part.do_charges()
    out = {"charges": part["charges"], "propars": part["propars"],
           "promoldens_sample": part["promoldens"][::97].copy()}
    for key in ("niter", "history_changes", "history_entropies"):
        if key in part.cache:
            out[key] = np.asarray(part[key])
    return out


SOLVER_CASES = {
    # aLISA host plug-ins (alisa.py:460-1127) on the Slater promolecule
    "s/lisa_diis": ("lisa", "s", dict(solver="diis")),
    "s/lisa_diis_A": ("lisa", "s", dict(solver="diis", solver_options=dict(version="A"))),
    "s/lisa_cdiis": ("lisa", "s", dict(solver="cdiis")),
    "s/lisa_cdiis_ad": ("lisa", "s", dict(solver="cdiis", solver_options=dict(mode="AD-CDIIS"))),
    "s/lisa_newton": ("lisa", "s", dict(solver="newton")),
    "s/lisa_m_newton": ("lisa", "s", dict(solver="m-newton")),
    "s/lisa_quasi_newton": ("lisa", "s", dict(solver="quasi-newton")),
    "s/lisa_trust_region": ("lisa", "s", dict(solver="trust-region", maxiter=6)),
    "s/lisa_sc_1_iter": ("lisa", "s", dict(solver="sc-1-iter")),
    # GISA: the reference's qpsolvers call answered by the shim's brute-force KKT enumeration
    "s/gisa": ("gisa", "s", dict()),
    "g/gisa": ("gisa", "g", dict()),
    # gLISA solver family (glisa.py:572-1028) on both promolecules
    "g/glisa_diis": ("glisa", "g", dict(solver="diis")),
    "g/glisa_diis_A": ("glisa", "g", dict(solver="diis", solver_options=dict(version="A"))),
    "g/glisa_diis_dmrs": ("glisa", "g", dict(solver="diis", solver_options=dict(use_dmrs=True))),
    "g/glisa_cdiis": ("glisa", "g", dict(solver="cdiis")),
    "g/glisa_cdiis_ad": ("glisa", "g", dict(solver="cdiis", solver_options=dict(mode="AD-CDIIS"))),
    "g/glisa_cdiis_fd": ("glisa", "g", dict(solver="cdiis", solver_options=dict(mode="FD-CDIIS"))),
    "g/glisa_m_newton": ("glisa", "g", dict(solver="m-newton")),
    "g/glisa_m_newton_kl": ("glisa", "g", dict(solver="m-newton", solver_options=dict(linesearch_mode="with-extended-kl"))),
    "g/glisa_quasi_newton": ("glisa", "g", dict(solver="quasi-newton")),
    "g/glisa_quasi_newton_2": ("glisa", "g", dict(solver="quasi-newton", solver_options=dict(niter_exact_newton=2))),
    "g/glisa_trust_region": ("glisa", "g", dict(solver="trust-region")),
    "s/glisa_diis": ("glisa", "s", dict(solver="diis")),
    "s/glisa_cdiis": ("glisa", "s", dict(solver="cdiis")),
    "s/glisa_newton": ("glisa", "s", dict(solver="newton")),
    "s/glisa_m_newton": ("glisa", "s", dict(solver="m-newton")),
    "s/glisa_quasi_newton": ("glisa", "s", dict(solver="quasi-newton")),
}


def case_water6_solvers(natom=6, nrad=40, nang=50, seed=0):
    """Every built-in solver of aLISA / gLISA that runs without third-party packages, on the two
    synthetic 6-atom promolecules of the cases above ("s" Slater, "g" Gaussian)."""
    from horton_part.core.basis import ExpBasisFuncHelper

    coords, numbers = synthetic.water_cluster(natom, seed)
    grid = synthetic_grid(coords, numbers, nrad, nang)
    helper = ExpBasisFuncHelper.from_function_type("gauss")
    rhos = {"s": synthetic.slater_promolecule_host(grid.points, coords, numbers),
            "g": synthetic.expbasis_promolecule_host(grid.points, coords, numbers, helper, scale={8: 8.6, 1: 0.7})}
    pseudo = numbers.astype(float)
    results = {}
    for tag, (scheme, dens, kw) in SOLVER_CASES.items():
        try:
            t0 = time.time()
            results[tag] = run_reference_light(scheme, coords, numbers, pseudo, grid, rhos[dens], **kw)
            print(f"  {tag}: niter={results[tag].get('niter')} q={results[tag]['charges'][:3]} {time.time()-t0:.1f}s")
        except Exception as exc:
            results[tag] = {"raised": np.array(f"{type(exc).__name__}: {exc}")}
            print(f"  {tag}: reference raised {type(exc).__name__}: {exc}")
    save("water6_solvers.npz", results, coordinates=coords, numbers=numbers)


CONVEX_CASES = {
    # the reference's DEFAULT aLISA / gLISA solver (cvxopt.solvers.cp, alisa.py:67-190,
    # glisa.py:488-570), answered by the shim's SciPy trust-constr + active-set polish
    "g/lisa_cvxopt": ("lisa", "g", dict()),
    "s/lisa_cvxopt": ("lisa", "s", dict()),
    "s/lisa_cvxopt_slater": ("lisa", "s", dict(basis_func="slater")),
    # ("sc-plus-convex" cannot be generated: its fall-back call omits `population_cutoff`,
    # alisa.py:446-457, and raises TypeError in the reference itself)
    "g/lisa_cvxopt_gt2": ("lisa", "g", dict(grid_type=2)),
    # GISA's quadratic programme on the molecular grid (gisa.py:257-279; QP answered by the qpsolvers stand-in)
    "g/gisa_gt2": ("gisa", "g", dict(grid_type=2)),
    "s/lisa_diis_gt2": ("lisa", "s", dict(solver="diis", grid_type=2, maxiter=8, solver_options=dict(check_mono=False))),
    "g/glisa_cvxopt": ("glisa", "g", dict()),
    "s/glisa_cvxopt": ("glisa", "s", dict()),
}


def case_water6_convex(natom=6, nrad=40, nang=50, seed=0):
    """The convex-programme solvers on the two synthetic 6-atom promolecules.  cvxopt itself is
    absent: these runs pin the product up to the choice of convex solver (unique minimiser)."""
    from horton_part.core.basis import ExpBasisFuncHelper

    coords, numbers = synthetic.water_cluster(natom, seed)
    grid = synthetic_grid(coords, numbers, nrad, nang)
    helper = ExpBasisFuncHelper.from_function_type("gauss")
    rhos = {"s": synthetic.slater_promolecule_host(grid.points, coords, numbers),
            "g": synthetic.expbasis_promolecule_host(grid.points, coords, numbers, helper, scale={8: 8.6, 1: 0.7})}
    pseudo = numbers.astype(float)
    results = {}
    for tag, (scheme, dens, kw) in CONVEX_CASES.items():
        t0 = time.time()
        results[tag] = run_reference_light(scheme, coords, numbers, pseudo, grid, rhos[dens], **kw)
        print(f"  {tag}: niter={results[tag].get('niter')} q={results[tag]['charges'][:3]} {time.time()-t0:.1f}s")
    save("water6_convex.npz", results, coordinates=coords, numbers=numbers)


def case_algo():
    """Known-answer vectors for the host algebra modules (algo/diis.py, algo/cdiis.py,
    algo/quasi_newton.py) and the aLISA radial plug-in solvers (alisa.py), from the reference
    itself on a seeded contraction map / seeded radial problems.  CPU-only tests use these."""
    import warnings

    import horton_part.alisa as ra
    from horton_part.algo.cdiis import cdiis
    from horton_part.algo.diis import diis, lstsq_solver_dyn
    from horton_part.algo.quasi_newton import bfgs
    from horton_part.core.basis import ExpBasisFuncHelper
    import contextlib
    import io

    from horton_part_b200 import synthetic as syn

    out = {}
    A, b = syn.contraction_map(12, seed=1)
    f = lambda x: A @ x + b + 0.05 * np.sin(x)  # noqa: E731
    x0 = np.zeros(12)
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for mode in ("R-CDIIS", "AD-CDIIS", "FD-CDIIS", "Roothaan"):
            for qr in ("full", "economic"):
                conv, n, rn, mk, cn, xl, hist = cdiis(x0.copy(), f, 1e-10, 200, modeQR=qr, mode=mode)
                out[f"cdiis/{mode}/{qr}/x"] = xl
                out[f"cdiis/{mode}/{qr}/niter"] = np.int64(n)
                out[f"cdiis/{mode}/{qr}/rnorm"] = np.asarray(rn)
                out[f"cdiis/{mode}/{qr}/mk"] = np.asarray(mk)
        for ver in "PA":
            for name, ls in (("sp", None), ("dyn", lstsq_solver_dyn)):
                x, n, hist = diis(x0.copy(), f, 1e-10, version=ver, lstsq_solver=ls)
                out[f"diis/{ver}/{name}/x"] = x
                out[f"diis/{ver}/{name}/niter"] = np.int64(n)
                out[f"diis/{ver}/{name}/history"] = np.asarray(hist)
        rng = np.random.default_rng(5)
        s, d0, d1 = rng.normal(size=5), rng.normal(size=5), rng.normal(size=5)
        H0 = np.eye(5) + 0.1 * np.outer(s, s)
        out["bfgs/s"], out["bfgs/d0"], out["bfgs/d1"], out["bfgs/H0"] = s, d0, d1, H0
        out["bfgs/H1"] = bfgs(d1, s, d0, H0)
        lg = logging.getLogger("golden")
        for func_type in ("gauss", "slater"):
            h = ExpBasisFuncHelper.from_function_type(func_type)
            for Z, pop in ((8, 8.5), (1, 0.7), (6, 6.1)):
                bs, rho, c0, r, w = syn.radial_problem(h, Z, pop)
                for name in ("solver_sc", "solver_sc_1_iter", "solver_diis", "solver_cdiis", "solver_m_newton",
                             "solver_quasi_newton", "solver_newton", "solver_trust_region"):
                    key = f"radial/{func_type}/{Z}/{name}"
                    try:
                        res = getattr(ra, name)(bs, rho, c0.copy(), r, w, 1e-8, lg, 1e-15, -1e-12, 1e-4)
                        out[key] = np.asarray(res)
                    except Exception as exc:
                        out[key + "/raised"] = np.array(type(exc).__name__)
    GOLD.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(GOLD / "algo_host.npz", **out)
    print("wrote", GOLD / "algo_host.npz", len(out), "entries")


def case_convex_radial():
    """The reference's ``solver_cvxopt`` (alisa.py:67-190) on the seeded radial problems of
    case_algo, its ``cvxopt.solvers.cp`` call answered by the shim.  CPU-only tests use these."""
    import horton_part.alisa as ra
    from horton_part.core.basis import ExpBasisFuncHelper

    from horton_part_b200 import synthetic as syn

    out = {}
    lg = logging.getLogger("golden")
    lg.setLevel(logging.INFO)
    for func_type in ("gauss", "slater"):
        h = ExpBasisFuncHelper.from_function_type(func_type)
        for Z, pop in ((8, 8.5), (1, 0.7), (6, 6.1)):
            bs, rho, c0, r, w = syn.radial_problem(h, Z, pop)
            out[f"{func_type}/{Z}/nonneg"] = ra.solver_cvxopt(bs, rho, c0.copy(), r, w, 1e-8, lg, 1e-15, -1e-12, 1e-4)
            if func_type == "gauss":  # (sign-free Slater fits leave the region where f is convex)
                out[f"{func_type}/{Z}/free"] = ra.solver_cvxopt(bs, rho, c0.copy(), r, w, 1e-8, lg, 1e-15, -1e-12,
                                                                1e-4, allow_neg_params=True)  # fmt: skip
    np.savez_compressed(GOLD / "convex_radial.npz", **out)
    print("wrote", GOLD / "convex_radial.npz", len(out), "entries")


def case_numeric(natom=6, nrad=40, nang=50, seed=0):
    """basis_type="numeric" (tabulated basis functions, core/basis.py:330-390): helper values for the
    CPU test and aLISA runs on the 6-atom Slater promolecule."""
    from horton_part.core.basis import NumericBasisFuncHelper

    coords, numbers = synthetic.water_cluster(natom, seed)
    grid = synthetic_grid(coords, numbers, nrad, nang)
    rho = synthetic.slater_promolecule_host(grid.points, coords, numbers)
    pseudo = numbers.astype(float)
    results = {}
    for tag, kw in {
        "lisa_sc": dict(solver="sc"),
        "lisa_sc_slater": dict(solver="sc", basis_func="slater"),
        "lisa_cvxopt": dict(),
        # (grid_type 2/3 raise AttributeError in the reference: gisa.py:265 asks the helper for exponents)
    }.items():
        t0 = time.time()
        results[tag] = run_reference_light("lisa", coords, numbers, pseudo, grid, rho, basis_type="numeric", **kw)
        print(f"  {tag}: niter={results[tag].get('niter')} q={results[tag]['charges'][:3]} {time.time()-t0:.1f}s")
    r = np.concatenate([[0.0, 1e-6], np.geomspace(1e-5, 60.0, 300)])  # incl. both extrapolation sides
    extra = {"helper/r": r}
    for ft in ("gauss", "slater"):
        h = NumericBasisFuncHelper.from_function_type(ft)
        for z in (1, 6, 8):
            extra[f"helper/{ft}/{z}"] = np.array([h.compute_proshell_dens(z, k, 1.0, r) for k in range(h.get_nshell(z))])
            pops = np.linspace(0.3, 1.1, h.get_nshell(z))
            extra[f"helper/{ft}/{z}/proatom"] = h.compute_proatom_dens(z, pops, r, 0)
    save("water6_numeric.npz", results, coordinates=coords, numbers=numbers, **extra)


def case_numeric_glisa(natom=6, nrad=40, nang=50, seed=0):
    """gLISA with basis_type="numeric" (core/basis.py:330-387 + glisa.py:226-246) on the 6-atom Slater
    promolecule: fixed point, DIIS, exact Newton and the default convex programme."""
    coords, numbers = synthetic.water_cluster(natom, seed)
    grid = synthetic_grid(coords, numbers, nrad, nang)
    rho = synthetic.slater_promolecule_host(grid.points, coords, numbers)
    pseudo = numbers.astype(float)
    results = {}
    for tag, kw in {"glisa_sc": dict(solver="sc"), "glisa_diis": dict(solver="diis"), "glisa_newton": dict(solver="newton"),
                    "glisa_m_newton": dict(solver="m-newton"), "glisa_cvxopt": dict(),
                    "glisa_sc_slater": dict(solver="sc", basis_func="slater")}.items():  # fmt: skip
        t0 = time.time()
        try:
            results[tag] = run_reference_light("glisa", coords, numbers, pseudo, grid, rho, basis_type="numeric", **kw)
            print(f"  {tag}: niter={results[tag].get('niter')} q={results[tag]['charges'][:3]} {time.time()-t0:.1f}s")
        except Exception as exc:
            results[tag] = {"raised": np.array(f"{type(exc).__name__}: {exc}")}
            print(f"  {tag}: reference raised {type(exc).__name__}: {exc}")
    save("water6_numeric_glisa.npz", results, coordinates=coords, numbers=numbers)


def case_proatomdb():
    """ProAtomDB.compact / normalize / compute_radii of the reference (core/proatomdb.py:154-190,
    390-447) on its cached HF/STO-3G records (the ones packed into h2o_hirshfeld.npz)."""
    from horton_part.core.proatomdb import ProAtomDB

    records, _ = load_proatom_records()
    db = ProAtomDB(records)
    out = {}
    for z in db.get_numbers():
        out[f"charges/{z}"] = np.array(db.get_charges(z))
        out[f"safe/{z}"] = np.array(db.get_charges(z, safe=True))
        for q in db.get_charges(z):
            rec = db.get_record(z, q)
            idx, radii = rec.compute_radii([0.5 * rec.pseudo_population, rec.pseudo_population - 0.1, 1e3])
            out[f"radii/{z}/{q}"] = np.concatenate([idx, radii])
    db.compact(0.1)
    for z in db.get_numbers():
        out[f"compact_size/{z}"] = np.int64(db.get_rgrid(z).size)
    db.normalize()
    for z in db.get_numbers():
        for q in db.get_charges(z):
            out[f"normalized/{z}/{q}"] = db.get_record(z, q).rho
    np.savez_compressed(GOLD / "proatomdb.npz", **out)
    print("wrote", GOLD / "proatomdb.npz", len(out), "entries")


def case_postproc():
    """Post-processing beyond charges on water HF/STO-3G: Becke scheme
    (becke.py), density decomposition splines (core/base.py:637-659), pro-atom splines
    (core/stockholder.py:386-393), Tkatchenko-Scheffler dispersion (hirshfeld.py:48-124)."""
    import contextlib
    import io

    from horton_part.core.proatomdb import ProAtomDB
    from scipy.spatial import cKDTree

    npz = np.load(REF_CACHED / "water_sto3g_hf_g03_fchk_exp:5e-4:2e1:120:110.npz")
    coords, numbers, pseudo = npz["coordinates"], npz["numbers"], npz["pseudo_numbers"]
    rgrid = qcgrid.ExpRTransform(5e-4, 2e1, 119).transform_1d_grid(qcgrid.UniformInteger(120))
    grid = qcgrid.MolGrid.from_size(numbers, coords, 110, rgrid, qcgrid.BeckeWeights(), rotate=False, store=True)
    _, order = cKDTree(npz["points"]).query(grid.points)
    rho = npz["dens"][order]
    out = {}
    r_mid = np.sqrt(rgrid.points[:-1] * rgrid.points[1:])
    with contextlib.redirect_stdout(io.StringIO()):
        part = wpart_schemes("b")(coords, numbers, pseudo, grid, rho)
        part.do_charges()
        out["becke/charges"] = part["charges"]
        out["becke/populations"] = part["populations"]
        out["becke/at_weights_0_sample"] = part["at_weights_0"][::53].copy()
        part.do_moments()
        out["becke/cartesian_multipoles"] = part["cartesian_multipoles"]

        part = wpart_schemes("mbis")(coords, numbers, pseudo, grid, rho)
        part.do_density_decomposition()
        for a in range(3):
            dec = part.cache.load("density_decomposition", a)
            keys = sorted(dec)
            out[f"mbis/decomp_{a}_knots"] = np.array([dec[k](rgrid.points) for k in keys])
            out[f"mbis/decomp_{a}_mid"] = np.array([dec[k](r_mid) for k in keys])
        part.do_prosplines()
        for a in range(3):
            out[f"mbis/prospline_{a}_mid"] = part.cache.load("spline_prodensity", a)(r_mid)

        records, _ = load_proatom_records()
        part = wpart_schemes("hi")(coords, numbers, pseudo, grid, rho, proatomdb=ProAtomDB(records))
        part.do_dispersion()
        for key in ("volumes", "volume_ratios", "c6s", "radial_moments"):
            out[f"hi/{key}"] = part[key]
        part.do_prosplines()
        pts = ProAtomDB(records).get_rgrid(8).points
        out["hi/prospline_0_mid"] = part.cache.load("spline_prodensity", 0)(np.sqrt(pts[:-1] * pts[1:]))
    GOLD.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(GOLD / "h2o_postproc.npz", **out)
    print("wrote", GOLD / "h2o_postproc.npz", {k: v.shape for k, v in out.items()})
    print("  becke q", out["becke/charges"], " c6", out["hi/c6s"])


def _load_molecule(name, rgrid):
    """A molecule of the reference's test suite on its own grid (tests/common.py:38-61): density
    re-ordered onto the rebuilt Becke-Lebedev grid."""
    from scipy.spatial import cKDTree

    npz = np.load(REF_CACHED / name)
    coords, numbers, pseudo = npz["coordinates"], npz["numbers"], npz["pseudo_numbers"]
    grid = qcgrid.MolGrid.from_size(numbers, coords, 110, rgrid, qcgrid.BeckeWeights(), rotate=False, store=True)
    dist, order = cKDTree(npz["points"]).query(grid.points)
    assert dist.max() < 1e-6 and len(set(order)) == grid.size
    return coords, numbers, pseudo, grid, npz["dens"][order]


def case_molecules():
    """Two more molecules of the reference's own tests: N2 (tests/test_becke.py:31-57) and
    monosilicic acid with LANL effective core potentials, i.e. pseudo_numbers != numbers
    (tests/test_wpart.py:104-217), with the hf_lan pro-atom database."""
    import contextlib
    import io

    from horton_part.core.proatomdb import ProAtomDB, ProAtomRecord

    out = {}
    rg = qcgrid.ExpRTransform(1e-3, 1e1, 99).transform_1d_grid(qcgrid.UniformInteger(100))
    coords, numbers, pseudo, grid, rho = _load_molecule("n2_hfs_sto3g_fchk_exp:1e-3:1e1:100:110.npz", rg)
    out.update({"n2/coordinates": coords, "n2/numbers": numbers, "n2/pseudo_numbers": pseudo, "n2/dens": rho,
                "n2/aim_weights_sample": grid.aim_weights[::211].copy()})
    for tag, scheme in (("becke", "b"), ("mbis", "mbis"), ("isa", "is")):
        with contextlib.redirect_stdout(io.StringIO()):
            part = wpart_schemes(scheme)(coords, numbers, pseudo, grid, rho)
            part.do_charges()
        out[f"n2/{tag}/charges"] = part["charges"]
        out[f"n2/{tag}/populations"] = part["populations"]
        if "niter" in part.cache:
            out[f"n2/{tag}/niter"] = np.int64(part["niter"])
        print(f"  n2 {tag}: q={part['charges']} niter={out.get(f'n2/{tag}/niter')}")

    rg = qcgrid.ExpRTransform(5e-4, 2e1, 119).transform_1d_grid(qcgrid.UniformInteger(120))
    coords, numbers, pseudo, grid, rho = _load_molecule("monosilicic_acid_hf_lan_fchk_exp:5e-4:2e1:120:110.npz", rg)
    out.update({"msa/coordinates": coords, "msa/numbers": numbers, "msa/pseudo_numbers": pseudo, "msa/dens": rho,
                "msa/aim_weights_sample": grid.aim_weights[::211].copy()})
    records = []
    for z in (14, 8, 1):
        for path in sorted(REF_CACHED.glob(f"atom_hf_lan_Z{z:02d}_N*_pow.npz")):
            with np.load(path) as f:
                number, charge, energy = int(f["number"]), int(f["charge"]), float(f["energy"])
                rmin, rmax, npoint = f["rgrid"]
                pn = float(f["pseudo_number"]) if "pseudo_number" in f.files else float(number)
                r = qcgrid.PowerRTransform(rmin, rmax, int(npoint) - 1).transform_1d_grid(qcgrid.UniformInteger(int(npoint)))
                records.append(ProAtomRecord(number, charge, energy, r, f["dens"], f["deriv"], pseudo_number=pn))
                out[f"msa/record/Z{number}_q{charge}"] = np.concatenate(
                    [[number, charge, energy, rmin, rmax, npoint, pn], f["dens"], f["deriv"]])
    for tag, scheme, kw in (("h", "h", dict(proatomdb=ProAtomDB(records))), ("hi", "hi", dict(proatomdb=ProAtomDB(records))),
                            ("isa", "is", {}), ("mbis", "mbis", {})):
        with contextlib.redirect_stdout(io.StringIO()):
            part = wpart_schemes(scheme)(coords, numbers, pseudo, grid, rho, **kw)
            part.do_charges()
        out[f"msa/{tag}/charges"] = part["charges"]
        out[f"msa/{tag}/populations"] = part["populations"]
        out[f"msa/{tag}/pseudo_populations"] = part["pseudo_populations"]
        if "niter" in part.cache:
            out[f"msa/{tag}/niter"] = np.int64(part["niter"])
        print(f"  msa {tag}: q={np.round(part['charges'], 5)} niter={out.get(f'msa/{tag}/niter')}")
    GOLD.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(GOLD / "ref_molecules.npz", **out)
    print("wrote", GOLD / "ref_molecules.npz", f"{(GOLD / 'ref_molecules.npz').stat().st_size / 1024:.0f} KiB")


def _pack_records(numbers, level=None):
    """ProAtomDB records for `numbers` from the reference's cached isolated-atom files
    (tests/common.py:64-111) + the packed raw arrays the GPU-side test rebuilds them from."""
    from horton_part.core.proatomdb import ProAtomRecord

    records, raw = [], {}
    for z in numbers:
        pat = f"atom_Z{z:02d}_N*_pow.npz" if level is None else f"atom_{level}_Z{z:02d}_N*_pow.npz"
        for path in sorted(REF_CACHED.glob(pat)):
            with np.load(path) as f:
                number, charge, energy = int(f["number"]), int(f["charge"]), float(f["energy"])
                rmin, rmax, npoint = f["rgrid"]
                rgrid = qcgrid.PowerRTransform(rmin, rmax, int(npoint) - 1).transform_1d_grid(qcgrid.UniformInteger(int(npoint)))
                records.append(ProAtomRecord(number, charge, energy, rgrid, f["dens"], f["deriv"]))
                raw[f"Z{number}_q{charge}"] = np.concatenate([[number, charge, energy, rmin, rmax, npoint], f["dens"], f["deriv"]])
    return records, raw


def _full_out(part, grid, natom_samples=3):
    out = {"charges": part["charges"], "promoldens_sample": part["promoldens"][::997].copy()}
    for key in ("niter", "propars", "history_changes", "history_entropies"):
        if key in part.cache:
            out[key] = np.asarray(part[key])
    if "history_charges" in part.cache:
        out["history_charges_last"] = np.asarray(part["history_charges"])[-1]
    for a in range(natom_samples):
        w = part[f"at_weights_{a}"]
        if w.shape == grid.weights.shape:
            w = w[grid.indices[a] : grid.indices[a + 1]]
        out[f"at_weights_{a}_sample"] = w[::53].copy()
    return out


def case_config2(natom=20, nrad=150, nang=194, seed=0):
    """BASELINE.json config 2 at FULL size: 20-atom organic-like chain, 150 x 194 grid per atom
    (582,000 points), exact Slater promolecule; ISA, Hirshfeld, Hirshfeld-I (database = the reference's
    cached H/C/N/O records) and MBIS, all through the unmodified reference."""
    import contextlib
    import io

    from horton_part.core.proatomdb import ProAtomDB

    coords, numbers = synthetic.organic_like(natom, seed)
    t0 = time.time()
    grid = synthetic_grid(coords, numbers, nrad, nang)
    print(f"  grid {grid.size} points, {time.time() - t0:.0f} s", flush=True)
    rho = synthetic.slater_promolecule_host(grid.points, coords, numbers)
    pseudo = numbers.astype(float)
    records, raw = _pack_records((1, 6, 7, 8))
    results = {}
    for tag, scheme, kw in (("mbis", "mbis", {}), ("h", "h", dict(proatomdb=ProAtomDB(records))), ("isa", "is", {})):
        t0 = time.time()
        with contextlib.redirect_stdout(io.StringIO()):
            part = wpart_schemes(scheme)(coords, numbers, pseudo, grid, rho, **kw)
            part.do_charges()
        results[tag] = _full_out(part, grid)
        results[tag]["seconds"] = np.float64(time.time() - t0)
        print(f"  config2 {tag}: niter={results[tag].get('niter')} q={np.round(part['charges'][:4], 6)} "
              f"{time.time() - t0:.0f} s", flush=True)
    # Hirshfeld-I: case_config2_hi (its own file; it needs a database-compatible element set and density)
    save("config2_organic20.npz", results, coordinates=coords, numbers=numbers, pseudo_numbers=pseudo,
         dens_sample=rho[::997].copy(), aim_weights_sample=grid.aim_weights[::997].copy(),
         nelec=np.float64(grid.integrate(rho)),
         grid_spec=np.array(f"BeckeRTransform(1e-4,1.5) o GaussChebyshev({nrad}) x Lebedev{nang}, BeckeWeights(); "
                            f"synthetic.organic_like({natom}, {seed})"),
         **{f"record/{k}": v for k, v in raw.items()})  # fmt: skip


#: target charges of the Hirshfeld-I variant of config 2 (per element)
CONFIG2_HI_CHARGES = {1: 0.12, 6: 0.08, 8: -0.42}


def database_promolecule(db, points, coords, numbers, charges):
    """sum_a rho_db(Z_a, q_a)(|r - R_a|) with linearly interpolated charge states (hirshfeld_i.py:116-134):
    a density for which Hirshfeld-I has the exact fixed point q_a.  `db` is a ProAtomDB (reference or
    product: same get_spline API)."""
    rho = np.zeros(len(points))
    for R, z, q in zip(coords, numbers, charges):
        ic = int(np.floor(q))
        x = float(q - ic)
        # hirshfeld_i.py:125-132: a one-electron atom (or an integer charge) uses the scaled lower state only
        one = (int(z) - ic) == 1 or x == 0.0
        spline = db.get_spline(int(z), {ic: 1 - x} if one else {ic: 1 - x, ic + 1: x})
        r = np.linalg.norm(points - R, axis=1)
        # inside the record's radial grid only: beyond it the extrapolating cubic is meaningless (the
        # outermost points of the Becke-transformed grid sit at 1e4 bohr and more)
        rho += np.where(r <= db.get_rgrid(int(z)).points[-1], np.clip(spline(np.minimum(r, 1e3)), 0.0, None), 0.0)
    return rho


def case_config2_hi(natom=20, nrad=150, nang=194, seed=0):
    """Hirshfeld-I at config-2 size: the same 20-atom chain with N replaced by O (the reference's cached
    database has no N anion: atom_Z07_N08 is absent), 582,000 points.  The density is the promolecule of
    the DATABASE pro-atoms at the charges CONFIG2_HI_CHARGES: the Slater promolecule of the other cases
    is too diffuse for the compact HF/STO-3G database atoms and drives oxygen beyond charge -1, where the
    reference's database ends (KeyError (8, -2) in core/proatomdb.py:282)."""
    import contextlib
    import io

    from horton_part.core.proatomdb import ProAtomDB

    coords, numbers = synthetic.organic_like(natom, seed)
    numbers = np.where(numbers == 7, 8, numbers)
    grid = synthetic_grid(coords, numbers, nrad, nang)
    records, raw = _pack_records((1, 6, 8))
    db = ProAtomDB(records)
    q_gen = np.array([CONFIG2_HI_CHARGES[int(z)] for z in numbers])
    rho = database_promolecule(db, grid.points, coords, numbers, q_gen)
    t0 = time.time()
    with contextlib.redirect_stdout(io.StringIO()):
        part = wpart_schemes("hi")(coords, numbers, numbers.astype(float), grid, rho, proatomdb=db)
        part.do_charges()
    res = _full_out(part, grid)
    res["seconds"] = np.float64(time.time() - t0)
    print(f"  config2 hi: niter={res.get('niter')} q={np.round(part['charges'][:6], 6)} {time.time() - t0:.0f} s", flush=True)
    save("config2_hi.npz", {"hi": res}, coordinates=coords, numbers=numbers, generating_charges=q_gen,
         dens_sample=rho[::997].copy(), aim_weights_sample=grid.aim_weights[::997].copy(),
         grid_spec=np.array(f"BeckeRTransform(1e-4,1.5) o GaussChebyshev({nrad}) x Lebedev{nang}, BeckeWeights(); "
                            f"synthetic.organic_like({natom}, {seed}) with N -> O; density = database promolecule"),
         **{f"record/{k}": v for k, v in raw.items()})  # fmt: skip


def case_config3(natom=24, nrad=150, nang=194, seed=0, maxiter=500):
    """BASELINE.json config 3 reduced in atoms only (24-atom water cluster on the REAL 150 x 194 grid,
    698,400 points, Gaussian promolecule): aLISA `sc` with the gauss and the slater basis, run to
    convergence (the 100-atom runs need hundreds of iterations of minutes each on the CPU)."""
    from horton_part.core.basis import ExpBasisFuncHelper

    coords, numbers = synthetic.water_cluster(natom, seed)
    grid = synthetic_grid(coords, numbers, nrad, nang)
    helper = ExpBasisFuncHelper.from_function_type("gauss")
    rho = synthetic.expbasis_promolecule_host(grid.points, coords, numbers, helper, scale={8: 8.6, 1: 0.7})
    pseudo = numbers.astype(float)
    results = {}
    for tag, kw in (("lisa_sc_gauss", dict(solver="sc", maxiter=maxiter)),
                    ("lisa_sc_slater", dict(solver="sc", basis_func="slater", maxiter=maxiter))):
        t0 = time.time()
        results[tag] = run_reference_light("lisa", coords, numbers, pseudo, grid, rho, **kw)
        results[tag]["seconds"] = np.float64(time.time() - t0)
        print(f"  config3 {tag}: niter={results[tag].get('niter')} q={results[tag]['charges'][:3]} {time.time() - t0:.0f} s",
              flush=True)
    save(f"config3_water{natom}.npz", results, coordinates=coords, numbers=numbers,
         dens_sample=rho[::997].copy(), aim_weights_sample=grid.aim_weights[::997].copy(),
         grid_spec=np.array(f"BeckeRTransform(1e-4,1.5) o GaussChebyshev({nrad}) x Lebedev{nang}, BeckeWeights(); "
                            f"synthetic.water_cluster({natom}, {seed}); maxiter={maxiter}"))  # fmt: skip


#: electrons per element of the config-4 promolecule (not Z: the start, c = initials scaled to Z, is then off the solution)
CONFIG4_SCALE = {1: 0.75, 6: 6.2, 7: 7.3, 8: 8.4}


def case_config4(natom=12, nrad=150, nang=194, seed=0):
    """BASELINE.json config 4 reduced in atoms only (12-atom peptide-like chain, REAL grid, Gaussian
    promolecule of the gauss table's initials): gLISA `newton` (exact Hessian) and `sc`."""
    from horton_part.core.basis import ExpBasisFuncHelper

    coords, numbers = synthetic.peptide_like(natom, seed)
    grid = synthetic_grid(coords, numbers, nrad, nang)
    helper = ExpBasisFuncHelper.from_function_type("gauss")
    rho = synthetic.expbasis_promolecule_host(grid.points, coords, numbers, helper, scale=CONFIG4_SCALE)
    pseudo = numbers.astype(float)
    results = {}
    for tag, kw in (("glisa_newton", dict(solver="newton")), ("glisa_sc", dict(solver="sc", maxiter=60))):
        t0 = time.time()
        try:
            results[tag] = run_reference_light("glisa", coords, numbers, pseudo, grid, rho, **kw)
        except Exception as exc:
            results[tag] = {"raised": np.array(f"{type(exc).__name__}: {exc}")}
            print(f"  config4 {tag}: reference raised {type(exc).__name__}: {exc}", flush=True)
            continue
        results[tag]["seconds"] = np.float64(time.time() - t0)
        print(f"  config4 {tag}: niter={results[tag].get('niter')} q={results[tag]['charges'][:3]} {time.time() - t0:.0f} s",
              flush=True)
    save(f"config4_peptide{natom}.npz", results, coordinates=coords, numbers=numbers,
         dens_sample=rho[::997].copy(), aim_weights_sample=grid.aim_weights[::997].copy(),
         grid_spec=np.array(f"BeckeRTransform(1e-4,1.5) o GaussChebyshev({nrad}) x Lebedev{nang}, BeckeWeights(); "
                            f"synthetic.peptide_like({natom}, {seed})"))  # fmt: skip


CASES = {"h2o": case_h2o, "water6": case_water_cluster, "water6g": case_water_gauss, "hirshfeld": case_hirshfeld,
         "solvers": case_water6_solvers, "convex": case_water6_convex, "convex_radial": case_convex_radial, "numeric": case_numeric, "numeric_glisa": case_numeric_glisa, "proatomdb": case_proatomdb, "algo": case_algo, "postproc": case_postproc,
         "molecules": case_molecules,
         "config2": case_config2, "config2_hi": case_config2_hi, "config3": case_config3, "config4": case_config4}

if __name__ == "__main__":
    for name in sys.argv[1:] or CASES:
        print("case", name)
        CASES[name]()
