"""BASELINE.json configurations 1-4 as measurement cases on one B200: timings plus the roofline block of
the dominant kernel of each (CUDA events on the launching stream, algorithmic flops by SURVEY.md section 8d).
Used by bench.py (extra keys of the JSON line) and tools/bench_configs.py (one JSON line per case).

Flop conventions: exponential pro-atom, unit U1: 16 per evaluated pair (8 for Gaussians: no
square root) + 36 per evaluated shell; spline pro-atom: 57 per pair; gLISA Hessian, unit U2: M (M + 1) Npts.
"""

from __future__ import annotations

import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NRAD, NANG = 150, 194


def _torch():
    import torch

    return torch


def grid_for(coords, numbers, nrad=NRAD, nang=NANG):
    from horton_part_b200 import gridlite

    rgrid = gridlite.BeckeRTransform(1e-4, 1.5).transform_1d_grid(gridlite.GaussChebyshev(nrad))
    return gridlite.MolGrid.from_size(numbers, coords, nang, rgrid, np.ones(len(numbers) * nrad * nang), store=True)


def finish_grid(grid, w):
    grid.aim_weights[:] = w
    grid.weights[:] = grid.atweights * w


def fp64_peak_tflops(dev):
    """DFMA throughput measured live on this GPU (hp_dfma_probe; nominal 148 x 64 x 2 x 1.965 GHz = 37.2)."""
    torch = _torch()
    from horton_part_b200 import _lib

    sink = torch.zeros(1, dtype=torch.float64, device=dev)
    ms, fl = np.zeros(1, np.float32), np.zeros(1, np.float64)
    _lib.call("hp_dfma_probe", 4096, sink, ms, fl, torch.cuda.current_stream(dev).cuda_stream)
    return float(fl[0] / (ms[0] * 1e-3) / 1e12)


_WARM = set()


def warm_up(dev):
    """One tiny partitioning per process and device before anything is timed: CUDA context, module load,
    page-locked staging buffers and the graph machinery are one-time costs of the process, not of a job."""
    if str(dev) in _WARM:
        return
    from horton_part_b200 import MBISWPart, gridlite, synthetic

    coords, numbers = synthetic.water_cluster(3, seed=1)
    rgrid = gridlite.BeckeRTransform(1e-4, 1.5).transform_1d_grid(gridlite.GaussChebyshev(20))
    grid = gridlite.MolGrid.from_size(numbers, coords, 26, rgrid, gridlite.BeckeWeights(), store=True)
    rho = synthetic.slater_promolecule_host(grid.points, coords, numbers)
    MBISWPart(coords, numbers, numbers.astype(float), grid, rho, device=dev, maxiter=3).do_partitioning()
    torch = _torch()
    # cuSOLVER / cuBLAS handles of the gLISA Newton step (Cholesky + refinement on the device)
    a = torch.eye(8, dtype=torch.float64, device=dev) * 2.0
    chol, _ = torch.linalg.cholesky_ex(a)
    (a @ torch.cholesky_solve(a[:, :1], chol)).sum().item()
    torch.cuda.synchronize()
    _WARM.add(str(dev))


def timed_partitioning(part):
    import gc

    torch = _torch()
    gc.collect()  # garbage of the previous case (host arrays of 10^6-10^7 points) is not this job's time
    gc.disable()
    try:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        part.do_partitioning()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    finally:
        gc.enable()
    niter = int(part["niter"])
    return {"niter": niter, "seconds": dt, "ms_per_iteration": 1e3 * dt / niter,
            "gpu_ms_weights_per_iteration": 1e3 * float(np.mean(part.history_time_update_at_weights)),
            "gpu_ms_propars_per_iteration": 1e3 * float(np.mean(part.history_time_update_propars)),
            "device_loop": "device_loop" in part.time_usage}  # fmt: skip


def time_weights_kernel(part, reps=10):
    """Mean CUDA-event time of the fused promolecule / weights / entropy launch on the resident slab, with
    the converged parameters (ms), and the evaluated (pairs, shells) if the kernel counts them."""
    torch = _torch()
    dev = part.slab.device
    part._launch_promol_weights(want_entropy=True)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        part._launch_promol_weights(want_entropy=True)
    e1.record()
    torch.cuda.synchronize(dev)
    t = part._table
    pairs = t.pairs_evaluated() if hasattr(t, "pairs_evaluated") else None
    shells = t.shells_evaluated() if hasattr(t, "shells_evaluated") else None
    return e0.elapsed_time(e1) / reps, pairs, shells


def config1(dev, reps=5):
    """H2O MBIS on the reference's test grid, end to end from host arrays (reference: 0.16 s on one core)."""
    from horton_part_b200 import MBISWPart, gridlite

    warm_up(dev)
    z = np.load(os.path.join(ROOT, "tests", "golden", "h2o_hf_sto3g.npz"))
    coords, numbers, pseudo = z["coordinates"], z["numbers"], z["pseudo_numbers"]
    rgrid = gridlite.ExpRTransform(5e-4, 2e1, 119).transform_1d_grid(gridlite.UniformInteger(120))
    grid = gridlite.MolGrid.from_size(numbers, coords, 110, rgrid, z["aim_weights"], store=True)
    torch = _torch()
    times = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        part = MBISWPart(coords, numbers, pseudo, grid, z["dens"], device=dev)
        part.do_partitioning()
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    ok = bool(np.abs(part["charges"] - z["mbis/charges"]).max() < 1e-8 and part["niter"] == int(z["mbis/niter"]))
    return {"workload": "config 1: H2O MBIS, ExpRTransform(5e-4,2e1,119) x Lebedev110, 39,600 points, HF/STO-3G density",
            "niter": int(part["niter"]), "e2e_seconds_best": min(times), "e2e_seconds_first_call": times[0],
            "gpu_seconds": float(np.sum(part.history_time_update_at_weights) + np.sum(part.history_time_update_propars)),
            "reference_cpu_seconds": 0.16, "reference_source": "SURVEY.md section 6 (unmodified reference, one core)",
            "matches_reference_golden": ok}  # fmt: skip


def config2(dev, peak=None):
    """20-atom organic-like chain, 582,000 points, exact Slater promolecule: ISA (spline pass) and MBIS."""
    from horton_part_b200 import ISAWPart, MBISWPart, synthetic

    warm_up(dev)
    peak = peak or fp64_peak_tflops(dev)
    coords, numbers = synthetic.organic_like(20, seed=0)
    grid = grid_for(coords, numbers)
    rho, w, _, _ = synthetic.slater_promolecule_device(grid, coords, numbers, device=dev)
    finish_grid(grid, w)
    pseudo = numbers.astype(float)
    out = {"workload": "config 2: 20-atom organic-like chain, 150x194 grid/atom, 582,000 points"}
    isa = ISAWPart(coords, numbers, pseudo, grid, rho, device=dev)
    out["isa"] = timed_partitioning(isa)
    ms, _, _ = time_weights_kernel(isa, reps=20)
    pairs = float(len(numbers)) * grid.size
    out["isa"]["roofline"] = {
        "kernel": "spline_build_kernel + promol_weights_spline_kernel<4>", "bound": "fp64", "kernel_ms": ms,
        "pairs_per_launch": pairs, "pairs_per_s": pairs / (ms * 1e-3), "flop_per_pair": 57.0,
        "achieved": pairs * 57.0 / (ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
        "frac": pairs * 57.0 / (ms * 1e-3) / 1e12 / peak,
        "note": "dense (no screening: spline tails are not monotone); 1.2e7 pairs per launch is ~80 us of work on 148 SMs"}  # fmt: skip
    mbis = MBISWPart(coords, numbers, pseudo, grid, rho, device=dev)
    out["mbis"] = timed_partitioning(mbis)
    hi = config2_hirshfeld_i(dev)
    if hi is not None:
        out["hirshfeld_i"] = hi
    return out


def config2_hirshfeld_i(dev):
    """Hirshfeld-I at config-2 size on the inputs of tests/golden/config2_hi.npz (20 atoms, 582,000 points, pro-atom
    database records of the reference's tests/cached/atom_*_pow.npz packed in the fixture; density = promolecule of
    the database pro-atoms at fixed charges).  None when the fixture is not there."""
    from horton_part_b200 import HirshfeldIWPart, gridlite
    from horton_part_b200.core.proatomdb import ProAtomDB, ProAtomRecord

    path = os.path.join(ROOT, "tests", "golden", "config2_hi.npz")
    if not os.path.exists(path):
        return None
    z = np.load(path)
    records = []
    for key in z.files:
        if key.startswith("record/"):
            v = z[key]
            npoint = int(v[5])
            rgrid = gridlite.PowerRTransform(v[3], v[4], npoint - 1).transform_1d_grid(gridlite.UniformInteger(npoint))
            records.append(ProAtomRecord(int(v[0]), int(v[1]), float(v[2]), rgrid, v[6 : 6 + npoint].copy(), v[6 + npoint :].copy()))
    db = ProAtomDB(records)
    coords, numbers = z["coordinates"], z["numbers"]
    rgrid = gridlite.BeckeRTransform(1e-4, 1.5).transform_1d_grid(gridlite.GaussChebyshev(NRAD))
    grid = gridlite.MolGrid.from_size(numbers, coords, NANG, rgrid, gridlite.DeviceBeckeWeights(), store=True)
    rho = np.zeros(grid.size)
    for R, zn, q in zip(coords, numbers, z["generating_charges"]):  # as in tests/test_gpu_configs.py
        ic = int(np.floor(q))
        x = float(q - ic)
        one = (int(zn) - ic) == 1 or x == 0.0
        spline = db.get_spline(int(zn), {ic: 1 - x} if one else {ic: 1 - x, ic + 1: x})
        r = np.linalg.norm(grid.points - R, axis=1)
        rho += np.where(r <= db.get_rgrid(int(zn)).points[-1], np.clip(spline(np.minimum(r, 1e3)), 0.0, None), 0.0)
    part = HirshfeldIWPart(coords, numbers, numbers.astype(float), grid, rho, db, device=dev)
    res = timed_partitioning(part)
    res["niter_reference"] = int(z["hi/niter"])
    res["reference_cpu_seconds"] = float(z["hi/seconds"])  # the unmodified reference on this repo's build container
    res["max_abs_charge_diff_vs_reference_run"] = float(np.abs(part["charges"] - z["hi/charges"]).max())
    return res


def config3(dev, natom=100, peak=None, slater_maxiter=50):
    """aLISA `sc`, gauss and slater basis, 100-atom water cluster, 2.91 M points, Gaussian promolecule."""
    from horton_part_b200 import LinearISAWPart, synthetic
    from horton_part_b200.core.basis import ExpBasisFuncHelper

    warm_up(dev)
    peak = peak or fp64_peak_tflops(dev)
    coords, numbers = synthetic.water_cluster(natom, seed=0)
    grid = grid_for(coords, numbers)
    helper = ExpBasisFuncHelper.from_function_type("gauss")
    rho, w = synthetic.expbasis_promolecule_device(grid, coords, numbers, helper, scale={8: 8.6, 1: 0.7}, device=dev)
    finish_grid(grid, w)
    pseudo = numbers.astype(float)
    out = {"workload": f"config 3: aLISA sc, {natom}-atom water cluster, 150x194 grid/atom, {grid.size} points"}
    for basis, maxiter in (("gauss", 500), ("slater", slater_maxiter)):
        part = LinearISAWPart(coords, numbers, pseudo, grid, rho, solver="sc", basis_func=basis, device=dev, maxiter=maxiter)
        res = timed_partitioning(part)
        res["mean_shells_per_atom"] = float(np.mean([part.bs_helper.get_nshell(int(z)) for z in numbers]))
        ms, pairs, shells = time_weights_kernel(part)
        dist = 8.0 if basis == "gauss" else 16.0
        flop = dist * pairs + 36.0 * shells
        res["roofline_weights"] = {
            "kernel": f"promol_weights_local_kernel<{basis.upper()},dense>", "bound": "fp64", "kernel_ms": ms,
            "pairs_evaluated": pairs, "shells_evaluated": shells, "pairs_evaluated_fraction": pairs / (float(natom) * grid.size),
            "achieved": flop / (ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s", "frac": flop / (ms * 1e-3) / 1e12 / peak,
            "work_counted": f"{dist:.0f} flop per evaluated pair + 36 per evaluated shell"}  # fmt: skip
        res["radial_solver"] = {
            "kernel": "shell_project_kernel + lisa_sc_block_kernel (one 128-thread block per atom)",
            "bound": "latency (sequential fixed point, ~1e4 dependent steps per launch for the slater basis)",
            "ms_per_iteration": res["gpu_ms_propars_per_iteration"],
            "share_of_iteration": res["gpu_ms_propars_per_iteration"]
            / (res["gpu_ms_propars_per_iteration"] + res["gpu_ms_weights_per_iteration"])}  # fmt: skip
        res["charges_head"] = [float(x) for x in part["charges"][:3]]
        out[basis] = res
    return out


def config4(dev, natom=300, peak=None, with_sc=False):
    """gLISA on a peptide-like chain, gauss basis (M = 5 natom), exact-Hessian Newton."""
    torch = _torch()
    from horton_part_b200 import GlobalLinearISAWPart, synthetic
    from horton_part_b200.core.basis import ExpBasisFuncHelper

    warm_up(dev)
    peak = peak or fp64_peak_tflops(dev)
    coords, numbers = synthetic.peptide_like(natom, seed=0)
    grid = grid_for(coords, numbers)
    helper = ExpBasisFuncHelper.from_function_type("gauss")
    rho, w = synthetic.expbasis_promolecule_device(grid, coords, numbers, helper, device=dev,
                                                   scale={1: 0.75, 6: 6.2, 7: 7.3, 8: 8.4})
    finish_grid(grid, w)
    pseudo = numbers.astype(float)
    part = GlobalLinearISAWPart(coords, numbers, pseudo, grid, rho, solver="newton", device=dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    part.do_partitioning()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    M = part._table.nshell
    out = {"workload": f"config 4: gLISA newton, {natom}-atom peptide-like chain, M = {M} basis functions, {grid.size} points",
           "niter": int(part["niter"]), "seconds": dt, "seconds_per_newton_iteration": dt / int(part["niter"]),
           "charges_head": [float(x) for x in part["charges"][:3]]}  # fmt: skip
    part._promol_and_entropy()
    part.hessian()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    part.hessian()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    flop = float(M) * (M + 1) * grid.size
    executed, total, ppt = part.hessian_tiles()
    flop_exec = executed * 2.0 * 64 * 64 * ppt
    out["hessian_screening"] = {
        "quadrants_executed": executed, "quadrants_total": total, "points_per_sub_panel": ppt,
        "frac_executed": executed / max(total, 1),
        "executed_tflops": flop_exec / (ms * 1e-3) / 1e12, "executed_frac_of_peak": flop_exec / (ms * 1e-3) / 1e12 / peak,
        "note": "64 x 64 quadrants of the 128 x 128 tile products whose column blocks are below 2^-64 of the chunk's "
                "largest |Gu| are skipped; executed flop counts full quadrants (padding and both halves of the "
                "diagonal quadrants included)"}  # fmt: skip
    frac_exec = executed / max(total, 1) if total else 1.0
    # basis regeneration: F_U1 = 8 + 36 K per atom x point for the gauss basis, never skipped
    flop_regen = float(grid.size) * (8.0 * natom + 36.0 * M)
    flop_credit = flop * (frac_exec if executed else 1.0) + flop_regen  # screening off: every tile ran
    out["roofline_hessian"] = {
        "kernel": "basis_chunk_kernel + syrk_panel_bulk_kernel (mma.sync m8n8k4 f64, cp.async.bulk + mbarrier operand ring) + hessian_finish_kernel",
        "bound": "fp64 (tensor)", "ms": ms, "flop_algorithmic_dense": flop + flop_regen, "flop_regeneration": flop_regen,
        "flop_credited": flop_credit,
        "achieved": flop_credit / (ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
        "frac": flop_credit / (ms * 1e-3) / 1e12 / peak,
        "dense_equivalent_tflops": (flop + flop_regen) / (ms * 1e-3) / 1e12,
        "work_counted": "M (M + 1) Npts (symmetric half, FMA = 2; SURVEY.md 8d unit U2) x the fraction of quadrant products "
                        "executed -- skipped quadrants earn nothing, as for cut-off pairs -- plus the basis regeneration "
                        "Npts (8 natom + 36 M), which is never skipped; dense_equivalent_tflops = the "
                        "unscreened count over the same time (exceeds the peak because work is skipped); peak = "
                        "measured DFMA rate (B200's FP64 tensor rate is nominally the same)"}  # fmt: skip
    # the same Hessian with the screening off: the tile product's own rate on the full M (M + 1) Npts
    os.environ["HP_B200_HESSIAN_SCREEN"] = "0"
    try:
        part.hessian()
        torch.cuda.synchronize()
        e0.record()
        part.hessian()
        e1.record()
        torch.cuda.synchronize()
    finally:
        os.environ.pop("HP_B200_HESSIAN_SCREEN", None)
    ms_dense = e0.elapsed_time(e1)
    out["roofline_hessian_unscreened"] = {
        "ms": ms_dense, "flop_algorithmic": flop + flop_regen, "achieved": (flop + flop_regen) / (ms_dense * 1e-3) / 1e12,
        "peak": peak, "unit": "TFLOP/s", "frac": (flop + flop_regen) / (ms_dense * 1e-3) / 1e12 / peak,
        "note": "HP_B200_HESSIAN_SCREEN=0: every tile product runs (the lower-left quadrant of the diagonal tiles, never "
                "read, is still skipped)"}  # fmt: skip
    e0.record()
    part._shell_integrals(1)
    e1.record()
    torch.cuda.synchronize()
    out["gradient_pass_ms"] = e0.elapsed_time(e1)
    if with_sc:
        part2 = GlobalLinearISAWPart(coords, numbers, pseudo, grid, rho, solver="sc", device=dev, threshold=1e-4)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        part2.do_partitioning()
        torch.cuda.synchronize()
        out["sc_threshold_1e-4"] = {"niter": int(part2["niter"]), "seconds": time.perf_counter() - t0}
    return out
