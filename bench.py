#!/usr/bin/env python
"""bench.py -- Jacobi GLUP/s of the effective-diffusivity hot path on B200 (BASELINE.json metric).

Workload (config.workload): BASELINE.json configs[1] -- the bundled 00042.jpg (1002 x 2007,
decoded by the reference's own decoder and committed as tests/golden/images.npz) with 4x mesh
amplification = 4008 x 8028 = 32 176 224 cells, shipped input.txt defaults (3-phase, Ds 0,
Df 1, Dg 1 237 500, CL 0, CR 1).  If the fixture is missing a synthetic three-tone image of the
same shape is used and `data` says so.

A "step" is one check interval of the reference loop (Deff2D.cuh:1232-1290): `check_every`
(10 000) damped-Jacobi sweeps followed by the boundary-flux Deff evaluation and the stop-rule
update, on the resident domain.  value = cells x sweeps / device time (CUDA events on the
launching stream, max over ranks), counted per sweep.  The iterate (2 x 257 MB) is larger than
L2 (126 MB), so no L2 flush is needed between steps.

e2e: the same metric through the C ABI with HOST buffers: every step uploads the image
(deff2d_domain_load: H2D, threshold + amplification, FloodFill, tables), runs the reference loop
for check_every + 1 sweeps (deff2d_domain_solve: two checks) and reads Deff back.

The same JSON line carries every BASELINE.json config under "configs", each with a `parity` field:
  c1  bundled 00000.jpg, shipped input.txt defaults, full solve (N = 1): sweeps per stage and Deff against the
      reference's own result (140 007 sweeps, Deff 224 673.610442892); plus the 2-phase run
  c2  the headline domain: Deff after `--cpu-sweeps` sweeps against the CPU oracle on the full 4008 x 8028 domain
  c3  `--batch-images` (512) distinct synthetic 256 x 256 images per GPU (image index rank*count + k), packed batch
      mode, images/s end to end; 4 images of rank 0 re-solved by the CPU oracle (sweep counts equal, Deff <= 1e-4)
  c4  one 16384 x 16384 two-phase domain, MaxIter 10 001 (two checks), row slabs over all N GPUs (strong scaling);
      Deff must be identical for every N (compare the SCALE lines); N > 1 also runs a slab-vs-single-GPU self-check
  c5  2048 x 2048 site percolation, Ds/Df = 1e-4, full solve to the reference stop rule (N = 1); the first
      `--c5-oracle-sweeps` sweeps are checked against the CPU oracle

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--sweeps-per-step S] [--kernel 0|1|2] [--tblock T]
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

AMP = 4
ALG_BYTES_PER_LUP = 16.0       # one FP64 read + one FP64 write of x per lattice update


def load_workload():
    fx = os.path.join(ROOT, "tests", "golden", "images.npz")
    if os.path.exists(fx):
        return np.load(fx)["img00042"], "bundled 00042.jpg (decoded fixture), coefficients/iterate synthetic-free"
    rng = np.random.default_rng(42)
    z = rng.random((2007, 1002))
    for _ in range(6):
        z = (z + np.roll(z, 1, 0) + np.roll(z, -1, 0) + np.roll(z, 1, 1) + np.roll(z, -1, 1)) / 5
    q1, q2 = np.quantile(z, [0.3, 0.7])
    img = np.where(z < q1, 0, np.where(z < q2, 150, 255)).astype(np.uint8)
    return img, "synthetic three-tone 1002x2007 (fixture missing)"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                     nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
            while not self._stop_evt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.1)
        except Exception as e:      # NVML missing: report what we know
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def trace(msg):
    """BENCH_TRACE=1: leg markers on stderr and a stack dump + exit if a leg sits for BENCH_TRACE_TIMEOUT (150) seconds."""
    if not os.environ.get("BENCH_TRACE"):
        return
    import faulthandler
    faulthandler.cancel_dump_traceback_later()
    faulthandler.dump_traceback_later(float(os.environ.get("BENCH_TRACE_TIMEOUT", "150")), exit=True)
    print("[bench rank %s %.1f] %s" % (os.environ.get("RANK", "0"), time.time() % 10000, msg), file=sys.stderr, flush=True)


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


# --------------------------------------------------------------------------------- reference arm

def run_reference(args):
    """The reference's own code on host cores (it ships GPU-only; oracle/_ref is its unmodified
    host logic + kernel body run on OpenMP threads).  Each step = one call of its JacobiGPU on the
    benchmark domain, bounded to `ref_sweeps` sweeps; GLUP/s from its own event timer."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    import _oracle as O
    img, data = load_workload()
    three_phase = True
    ref = O.reference("cpu")
    # all host threads this process may use (torchrun exports OMP_NUM_THREADS=1 for its workers)
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(ncpu))
    except OSError:
        pass
    O.oracle().orc_set_num_threads(int(ncpu))
    nthreads = O.oracle().orc_num_threads()
    if args.ref_crop:
        img = img[:args.ref_crop, :]
    Ds, Df, Dg = 0.0, 1.0, 1237500.0
    D = O.fill_D(img, AMP, AMP, 3, Ds, Df, Dg)
    Ny, Nx = D.shape
    cells = Nx * Ny
    G, _ = O.floodfill(O.grid_mask(img, AMP, AMP, 200))
    x0 = O.init_x(Nx, Ny, 0.0, 1.0)
    sweeps = args.ref_sweeps
    if ref is not None:
        kind = "reference"
        A, b = O.ref_discretize(D, 0.0, 1.0, G if three_phase else None)

        def step():
            r = O.ref_jacobi(A, b, x0, D, 0.0, 1.0, 1e-30, sweeps)
            return r["ms"] / 1000.0, r["iters"]
    else:
        kind = "port"
        A, b = O.discretize(D, 0.0, 1.0, G if three_phase else None)

        def step():
            r = O.jacobi(A, b, x0, D, 0.0, 1.0, 1e-30, sweeps)
            return r["seconds"], r["iters"]
    for _ in range(args.warmup):
        step()
    tot_s, tot_it = 0.0, 0
    for _ in range(args.steps):
        s, it = step()
        tot_s += s
        tot_it += it
    glups = cells * tot_it / tot_s / 1e9
    sample = "%d sweeps per step of %s's JacobiGPU loop on %dx%d cells, %d OpenMP threads" % (
        sweeps, "the reference" if kind == "reference" else "the oracle port", Nx, Ny, nthreads)
    line = {"impl": "reference", "metric": "jacobi_glups", "value": glups, "unit": "GLUP/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_s / args.steps * 1000,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": data,
            "config": {"workload": "configs[1]: 00042.jpg x4 mesh amplification (%dx%d cells), 3-phase shipped defaults; "
                                   "bounded sample: %s" % (Nx, Ny, sample)},
            "cpu_baseline": {"value": glups, "unit": "GLUP/s", "cores": nthreads, "kind": kind, "sample": sample},
            "e2e": {"value": glups, "unit": "GLUP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------- our arm

def cpu_baseline(img, sweeps, three_phase=True):
    import _oracle as O
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    O.oracle().orc_set_num_threads(int(ncpu))
    o = O.make_opts(Ds=0.0, Df=1.0, Dg=1237500.0, ampx=AMP, ampy=AMP, nphase=3 if three_phase else 2)
    d = np.zeros(1)
    img = np.ascontiguousarray(img)
    secs = O.oracle().orc_time_sweeps(O._up(img), img.shape[1], img.shape[0], o, 1 if three_phase else 0, sweeps, O._dp(d))
    cells = img.shape[0] * img.shape[1] * AMP * AMP
    n = O.oracle().orc_num_threads()
    return {"value": cells * sweeps / secs / 1e9, "unit": "GLUP/s", "cores": n, "kind": "port", "deff_raw": float(d[0]),
            "sweeps": int(sweeps),
            "sample": "%d sweeps of the oracle's restatement of updateX_SOR (A[n][5]+b, 80 B/LUP) on the full "
                      "%dx%d domain, %d OpenMP threads, %.1f s" % (sweeps, img.shape[1] * AMP, img.shape[0] * AMP, n, secs)}


def ref_cuda_baseline(img, sweeps):
    """The reference's own CUDA build (nvcc, sm_100a, unmodified JacobiGPU + updateX_SOR) timed on
    this GPU by its own event timer -- a reported baseline (north_star), bounded to `sweeps`."""
    import _oracle as O
    if O.reference("cuda") is None:
        return None
    D = O.fill_D(img, AMP, AMP, 3, 0.0, 1.0, 1237500.0)
    G, _ = O.floodfill(O.grid_mask(img, AMP, AMP, 200))
    A, b = O.discretize(D, 0.0, 1.0, G)
    Ny, Nx = D.shape
    r = O.ref_jacobi(A, b, O.init_x(Nx, Ny, 0.0, 1.0), D, 0.0, 1.0, 1e-30, sweeps, kind="cuda")
    return {"value": Nx * Ny * r["iters"] / (r["ms"] * 1e-3) / 1e9, "unit": "GLUP/s", "kind": "reference CUDA build (sm_100a)",
            "sample": "%d sweeps of the reference's JacobiGPU loop (kernel + sync + D2D copy per sweep, one D2H check) on "
                      "%dx%d cells, its own cudaEvent time %.1f ms" % (r["iters"], Nx, Ny, r["ms"])}


def rel_err(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def leg_c1(ctx):
    """configs[0]: bundled 00000.jpg at native resolution, shipped input.txt defaults (3-phase, 7 continuation
    stages) through the whole-path call, against the reference's own run (BASELINE.md section 2 / tests/golden)."""
    import effectivediffusivityfvm_b200 as E
    img = np.load(os.path.join(ROOT, "tests", "golden", "images.npz"))["img00000"]
    want_iters, want_deff = [80001, 10001, 10001, 10001, 10001, 10001, 10001], 224673.610442892
    ctx.solve_image(img, E.default_params(max_iter=11))                       # warm-up: allocations, graphs
    t0 = time.perf_counter()
    r = ctx.solve_image(img, E.default_params())
    dt = time.perf_counter() - t0
    out = {"workload": "00000.jpg 128x128, shipped input.txt defaults (3-phase, Ds 0, Df 1, Dg 1237500, tol 1e-5, MaxIter 5e5)",
           "iters": r["iters"], "total_sweeps": r["total_iters"], "deff": r["deff"], "reference_deff": want_deff,
           "deff_rel_err": rel_err(r["deff"], want_deff), "seconds": dt, "solve_ms": r["solve_ms"],
           "glups": 128 * 128 * r["total_iters"] / dt / 1e9, "us_per_sweep": dt * 1e6 / max(r["total_iters"], 1)}
    out["parity"] = bool(r["iters"] == want_iters and out["deff_rel_err"] <= 1e-4)
    t0 = time.perf_counter()
    r2 = ctx.solve_image(img, E.default_params(Ds=1e-4, Df=1.0, mode=E.MODE_2PH_BATCH))
    dt2 = time.perf_counter() - t0
    out["two_phase"] = {"workload": "same image, Phases 2, Ds 1e-4, Df 1", "iters": r2["iters"], "deff": r2["deff"],
                        "reference_deff": 0.1816910277372, "deff_rel_err": rel_err(r2["deff"], 0.1816910277372), "seconds": dt2,
                        "parity": bool(r2["iters"] == [100001] and rel_err(r2["deff"], 0.1816910277372) <= 1e-4)}
    out["parity"] = bool(out["parity"] and out["two_phase"]["parity"])
    return out


def leg_c2_parity(ctx, img, p, base):
    """configs[1] parity at full size: Deff after the same number of sweeps from x0 as the CPU oracle ran."""
    n = int(base["sweeps"])
    ctx.domain_load(img, 3, p)
    ctx.sweeps(n)
    d = ctx.flux()[0]
    e = rel_err(d, base["deff_raw"])
    return {"sweeps": n, "deff_raw": d, "oracle_deff_raw": base["deff_raw"], "deff_rel_err": e, "parity": bool(e <= 1e-4),
            "note": "un-normalised Deff of the full 4008x8028 domain after %d sweeps from x0, library vs CPU oracle" % n}


def leg_c3(ctx, rank, world, count, size=256, oracle_checks=4, serial_checks=0):
    """configs[2]: `count` distinct images per GPU (image index rank*count + k) through the packed batch mode,
    host images in, Deff out; per-image sweep histogram; `oracle_checks` images of rank 0 re-solved by the CPU oracle."""
    import effectivediffusivityfvm_b200 as E
    from effectivediffusivityfvm_b200.datasets import c3_image
    first = rank * count
    imgs = np.stack([c3_image(first + k, size) for k in range(count)])
    p = E.default_params(Ds=1e-3, Df=1.0, mode=E.MODE_2PH_BATCH, tol=1e-5, max_iter=500000)
    ctx.solve_batch(imgs[:2], E.default_params(Ds=1e-3, Df=1.0, mode=E.MODE_2PH_BATCH, max_iter=50))
    t0 = time.perf_counter()
    res = ctx.solve_batch(imgs, p)
    dt = time.perf_counter() - t0
    iters = np.array([r["total_iters"] for r in res], dtype=np.int64)
    hist = {}
    for v in iters:
        hist[int(v)] = hist.get(int(v), 0) + 1
    out = {"workload": "%d distinct synthetic two-phase %dx%d images per GPU (c3_image(rank*%d + k)), Ds 1e-3, Df 1, tol 1e-5, "
                       "MaxIter 5e5, packed batch mode" % (count, size, size, count),
           "images": count, "seconds": dt, "images_per_s": count / dt, "glups": float(iters.sum()) * size * size / dt / 1e9,
           "sweeps_total": int(iters.sum()), "sweep_histogram": {str(k): hist[k] for k in sorted(hist)},
           "pathflag_sum": int(sum(r["pathflag"] for r in res))}
    if rank == 0 and oracle_checks > 0:
        import _oracle as O
        ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        O.oracle().orc_set_num_threads(int(ncpu))
        # the cheapest images for the oracle (it runs ~1 GLUP/s): smallest sweep counts, ties by index
        pick = [int(k) for k in np.argsort(iters, kind="stable")[:oracle_checks]]
        checks, ok = [], True
        t1 = time.perf_counter()
        for k in pick:
            ref = O.solve_image(imgs[k], O.make_opts(Ds=1e-3, Df=1.0, nphase=2), O.MODE_2PH_BATCH)
            e = rel_err(res[k]["deff"], ref["deff"])
            good = bool(res[k]["iters"] == ref["iters"] and e <= 1e-4 and res[k]["pathflag"] == ref["pathflag"])
            ok = ok and good
            checks.append({"image": first + k, "iters": res[k]["iters"], "oracle_iters": ref["iters"], "deff": res[k]["deff"],
                           "oracle_deff": ref["deff"], "deff_rel_err": e, "ok": good})
        out["oracle_checks"] = checks
        out["oracle_seconds"] = time.perf_counter() - t1
        out["parity"] = ok
    elif rank == 0 and serial_checks > 0:
        # N > 1: no CPU oracle beside the worker processes (torchrun pins them to one OpenMP thread each); instead the
        # same images once more through the single-image path (a different kernel path, itself oracle-checked at N = 1)
        pick = [int(k) for k in np.argsort(iters, kind="stable")[:serial_checks]]
        ok = True
        for k in pick:
            one = ctx.solve_image(imgs[k], p)
            ok = ok and one["iters"] == res[k]["iters"] and one["deff"] == res[k]["deff"] and one["pathflag"] == res[k]["pathflag"]
        out["serial_checks"] = {"images": [first + k for k in pick], "bitwise_equal": bool(ok)}
        out["parity"] = bool(ok)
    return out


def leg_c4(ctx, rank, local_rank, world, size=16384, sweeps=10001):
    """configs[3]: one size x size two-phase domain, the reference loop with MaxIter `sweeps` (checks at sweeps 1 and
    10 001), row slabs over all ranks with NCCL halo exchange and flux all-reduce.  Strong scaling: fixed total work."""
    import torch
    import effectivediffusivityfvm_b200 as E
    from effectivediffusivityfvm_b200.datasets import c4_image
    img = np.tile(c4_image(4096), (size // 4096, size // 4096))          # periodic generator: a seamless medium
    trace("c4 image generated")
    p = E.default_params(Ds=1e-3, Df=1.0, mode=E.MODE_2PH_BATCH)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))
    if world > 1:
        import torch.distributed as dist
        from effectivediffusivityfvm_b200 import slab as slabmod
        dom = slabmod.SlabDomain(ctx, img, p, rank, world, nphase=2)
        load = dom.reload
        solve = dom.solve
    else:
        def load():
            ctx.domain_load(img, 2, p)
        solve = ctx.solve
        load()
    trace("c4 domain loaded")
    solve(1e-30, 401)                                                     # warm-up incl. CUDA-graph capture
    trace("c4 warm-up solve done")
    load()
    ctx.sync()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    r = solve(1e-30, sweeps)
    e1.record(stream)
    ctx.sync()
    ms = e0.elapsed_time(e1)
    t0 = time.perf_counter()
    load()
    r2 = solve(1e-30, sweeps)
    ctx.sync()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([ms, e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0].item()), float(t[1].item())
    cells = size * size
    # Deff of the same solve on ONE GPU (this script at N = 1, bit-stable across the kernel variants): a slab run must
    # reproduce it to rounding -- the flux sums are all-reduced over the slabs, so the last bit may differ
    n1 = {16384: 0.25900217157938904}.get(size)
    vs_n1 = rel_err(r["deff_raw"], n1) if n1 is not None else None
    return {"workload": "one %dx%d two-phase domain (sigma 8 px blobs, porosity 0.6), Ds 1e-3, Df 1, MaxIter %d: "
                        "%s" % (size, size, sweeps, "single GPU" if world == 1 else "%d row slabs, %s + NCCL flux all-reduce" % (world, "halo rows pushed into peer memory by the sweep kernel" if dom.peer else "NCCL halo exchange")),
            "scaling": "strong", "cells": cells, "sweeps": int(r["iters"]), "ms": ms, "glups": cells * r["iters"] / (ms * 1e-3) / 1e9,
            "e2e_seconds": e2e_s, "e2e_glups": cells * r2["iters"] / e2e_s / 1e9,
            "deff_raw": r["deff_raw"], "deff_raw_hex": float(r["deff_raw"]).hex(), "single_gpu_deff_raw": n1,
            "deff_rel_err_vs_single_gpu": vs_n1,
            "parity": bool(r["iters"] == sweeps and np.isfinite(r["deff_raw"]) and r["deff_raw"] == r2["deff_raw"] and
                           (vs_n1 is None or vs_n1 <= 1e-12)),
            "note": "the same Deff at every N to rounding (the flux sums are all-reduced over the slabs): compare deff_raw across "
                    "the SCALE lines; e2e = image upload from host buffers + device FloodFill + assembly + the same solve"}


def leg_slab_parity(rank, local_rank, world):
    """N > 1: a small 3-phase and a 2-phase domain decomposed into `world` row slabs against the undecomposed run on
    this rank's own GPU: own rows bit for bit after 1 / 4 / 203 sweeps, Deff, and the full reference loop."""
    import torch
    import torch.distributed as dist
    import effectivediffusivityfvm_b200 as E
    from effectivediffusivityfvm_b200.slab import SlabDomain
    rng = np.random.default_rng(11)
    z = rng.random((300, 500))
    for _ in range(3):
        z = (z + np.roll(z, 1, 0) + np.roll(z, -1, 0) + np.roll(z, 1, 1) + np.roll(z, -1, 1)) / 5
    q1, q2 = np.quantile(z, [0.3, 0.7])
    img = np.where(z < q1, 0, np.where(z < q2, 150, 255)).astype(np.uint8)
    ok, worst = True, 0.0
    for nphase, halo, peer in ((3, 16, False), (2, 32, False), (3, 8, True)):
        p = E.default_params(Ds=0.0 if nphase == 3 else 1e-3, Df=1.0, Dg=80.0, CL=0.25, CR=1.5, check_every=400)
        ref, ctx = E.Deff2D(local_rank), E.Deff2D(local_rank)
        ref.domain_load(img, nphase, p)
        dom = SlabDomain(ctx, img, p, rank, world, nphase=nphase, halo=halo, peer=peer)
        L = dom.layout
        for n in (1, 4, 203):
            ref.sweeps(n)
            dom.sweeps(n)
            same = np.array_equal(dom.own_field(), ref.get_field()[L.row0:L.row0 + L.own_rows], equal_nan=True)
            e = rel_err(dom.flux(), ref.flux()[0])
            worst = max(worst, e)
            ok = ok and same and e < 1e-12
        ref.domain_load(img, nphase, p)
        a = ref.solve(1e-4, 6000)
        dom.reload()
        b = dom.solve(1e-4, 6000)
        e = rel_err(b["deff_raw"], a["deff_raw"])
        worst = max(worst, e)
        ok = ok and a["iters"] == b["iters"] and e < 1e-12
        ref.close()
        ctx.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return {"parity": bool(flag.item()), "worst_deff_rel_err_rank0": worst,
            "what": "300x500 domain, 3-phase (halo 16) and 2-phase (halo 32) over NCCL, 3-phase (halo 8) with the peer-memory exchange, %d slabs vs one GPU: own rows bitwise after 1/4/203 sweeps, "
                    "Deff <= 1e-12, same sweep count of the full loop; all ranks agree" % world}


def leg_c5(ctx, oracle_sweeps):
    """configs[4]: 2048 x 2048 site percolation just above threshold, Ds/Df = 1e-4, full solve to the reference stop
    rule; the first `oracle_sweeps` sweeps against the CPU oracle (the full solve is ~35 CPU-minutes)."""
    import effectivediffusivityfvm_b200 as E
    from effectivediffusivityfvm_b200.datasets import c5_image
    img = c5_image()
    p = E.default_params(Ds=1e-4, Df=1.0, mode=E.MODE_2PH_BATCH, tol=1e-5, max_iter=500000)
    ctx.solve_image(img, E.default_params(Ds=1e-4, Df=1.0, mode=E.MODE_2PH_BATCH, max_iter=401))
    t0 = time.perf_counter()
    r = ctx.solve_image(img, p)
    dt = time.perf_counter() - t0
    out = {"workload": "2048x2048 site percolation p = 0.60, Ds 1e-4, Df 1, tol 1e-5, MaxIter 5e5, full solve",
           "iters": r["iters"], "deff": r["deff"], "conv": r["conv"], "pathflag": r["pathflag"], "seconds": dt, "solve_ms": r["solve_ms"],
           "glups": 2048 * 2048 * r["total_iters"] / dt / 1e9}
    ok = bool(np.isfinite(r["deff"]) and r["deff"] > 0 and (r["total_iters"] - 1) % 10000 == 0 or r["total_iters"] == 500000)
    if oracle_sweeps > 0:
        import _oracle as O
        ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        O.oracle().orc_set_num_threads(int(ncpu))
        o = O.make_opts(Ds=1e-4, Df=1.0, Dg=0.0, nphase=2)
        d = np.zeros(1)
        O.oracle().orc_time_sweeps(O._up(np.ascontiguousarray(img)), 2048, 2048, o, 0, oracle_sweeps, O._dp(d))
        ctx.domain_load(img, 2, p)
        ctx.sweeps(oracle_sweeps)
        mine = ctx.flux()[0]
        e = rel_err(mine, float(d[0]))
        out["oracle_prefix"] = {"sweeps": oracle_sweeps, "deff_raw": mine, "oracle_deff_raw": float(d[0]), "deff_rel_err": e}
        ok = ok and e <= 1e-4
    out["parity"] = bool(ok)
    # the opt-in accelerated solver (NON-PARITY mode, csrc/chebyshev.cu) on the same system: sweeps and time to a
    # relative residual of 1e-6, next to the reference iterate above (which stops at MaxIter, unconverged)
    try:
        t0 = time.perf_counter()
        rc = ctx.solve_image(img, E.default_params(Ds=1e-4, Df=1.0, mode=E.MODE_2PH_BATCH, solver=1, residual_tol=1e-6, max_iter=2000000))
        out["chebyshev"] = {"sweeps": rc["total_iters"], "seconds": time.perf_counter() - t0, "deff": rc["deff"],
                            "relative_residual": rc["conv"], "reference_iterate_deff": r["deff"],
                            "note": "non-parity mode: converged solution of the same discretisation; the reference iterate "
                                    "after %d sweeps still changes by %.1e per check" % (r["total_iters"], abs(r["conv"]))}
    except Exception as e:
        out["chebyshev"] = {"error": "%s: %s" % (type(e).__name__, e)}
    return out


def run_ours(args):
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the one JSON line (NCCL prints its version banner there)
    import torch
    import effectivediffusivityfvm_b200 as E
    rank, local_rank, world = dist_env()
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    img, data = load_workload()
    H, W = img.shape
    Nx, Ny = W * AMP, H * AMP
    trace("start")
    ctx = E.Deff2D(local_rank)
    p = E.default_params(amp_x=AMP, amp_y=AMP)          # shipped defaults: 3-phase, Ds 0, Df 1, Dg 1237500
    ctx.set_kernel(args.kernel, args.tblock)
    S = args.sweeps_per_step
    slab_mode = world > 1
    if slab_mode:
        from effectivediffusivityfvm_b200 import slab as slabmod
        dom = slabmod.SlabDomain(ctx, img, p, rank, world, weak=True)
        cells = dom.global_cells
        launch_sweeps, flux = dom.sweeps, dom.flux
    else:
        ctx.domain_load(img, 3, p)
        cells = Nx * Ny
        launch_sweeps, flux = ctx.sweeps, lambda: ctx.flux()[0]

    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        launch_sweeps(S)
        return flux()            # K4 + stop-rule state; blocks for the 8-byte result like the reference's check

    trace("domain loaded")
    for _ in range(max(args.warmup, 3) if not args.allow_short_warmup else args.warmup):
        step()
    barrier()
    trace("warm-up done")
    sampler = ClockSampler(local_rank)
    if not args.no_clock_sampler:
        sampler.start()
    l0 = ctx.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record(stream)
        deff = None
        for _ in range(args.steps):
            deff = step()
        ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.kernel_launches - l0
    clocks = sampler.stop() if not args.no_clock_sampler else {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["not sampled"]}
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = cells * S * args.steps / (ms * 1e-3) / 1e9

    # roofline of the dominant kernel (the sweep): live CUDA-event time of sweep launches only
    sweep_ms = ctx.sweeps_timed(S) if not slab_mode else ms / args.steps
    # slab mode: the timed step also holds the halo exchanges; one sweep launch = `depth` sweeps over the local slab
    depth = args.tblock if args.kernel == 2 and args.tblock > 0 else ctx.default_depth
    sweep_launches_per_step = (launches / args.steps) - 1 if not slab_mode else S / float(depth)   # - 1: the flux launch
    peak, peak_src = measured_peaks()
    local_cells = cells // world
    achieved = ALG_BYTES_PER_LUP * local_cells * S / (sweep_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src, "kernel": "k_sweep_tma (tiled sweep, %d sweeps per HBM pass)" % depth if args.kernel != 1 else "k_sweep_simple",
                "alg_bytes_per_launch": ALG_BYTES_PER_LUP * local_cells * (S / max(sweep_launches_per_step or S, 1)),
                "avg_launch_us": sweep_ms * 1e3 / max(sweep_launches_per_step or S, 1)}
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            roofline["traffic"] = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            pass
    if roofline["traffic"]:
        # frac > 1 is temporal blocking (T sweeps per HBM pass), not magic: the DRAM the kernel really
        # moves (ncu, profiles/traffic.json) over the live launch time, against the same measured peak
        roofline["dram_gbs"] = roofline["traffic"] / (roofline["avg_launch_us"] * 1e-6) / 1e9
        roofline["dram_frac"] = roofline["dram_gbs"] / peak
        roofline["note"] = ("achieved counts 16 algorithmic bytes per lattice update and sweep; one launch runs several sweeps per HBM pass, "
                            "so frac can exceed 1; dram_gbs/dram_frac are the measured DRAM bytes of the same launch")

    # end to end through the C ABI with host buffers
    e2e = None
    if not slab_mode:
        Se = S + 1                                          # sweep 0 + check, S sweeps + check (cuh:1243)
        e_steps = max(1, min(args.steps, 3))
        ctx.domain_load(img, 3, p)
        ctx.solve(1e-30, Se)
        t0 = time.perf_counter()
        for _ in range(e_steps):
            ctx.domain_load(img, 3, p)                      # H2D image + pinned mask, assembly, tables
            r = ctx.solve(1e-30, Se)                        # 2 checks; Deff read back (D2H)
        dt = time.perf_counter() - t0
        e2e = {"value": cells * Se * e_steps / dt / 1e9, "unit": "GLUP/s",
               # image (the FloodFill mask is computed on the device) + weight LUT, dead table and compact table
               "h2d_bytes_per_step": int(img.size + 2048 * 32 + 2048 + 1024 * 32),
               # phase counts (48 B) + the convergence state read at each of the 2 checks
               "d2h_bytes_per_step": int(48 + 2 * 2120), "steps": e_steps, "sweeps_per_step": Se,
               "deff_raw": r["deff_raw"]}
    else:
        # every rank uploads its slab again from host buffers (image rows + pinned mask of its rows),
        # runs the reference loop for S + 1 sweeps (two checks, flux all-reduce) and reads Deff back
        Se = S + 1
        e_steps = max(1, min(args.steps, 3))
        dom.reload()
        dom.solve(1e-30, Se)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            dom.reload()
            r = dom.solve(1e-30, Se)
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": cells * Se * e_steps / float(dt.item()) / 1e9, "unit": "GLUP/s",
               "h2d_bytes_per_step": int(dom.h2d_bytes + 2048 * 32 + 2048 + 1024 * 32), "d2h_bytes_per_step": int(48 + 2 * 2120),
               "steps": e_steps, "sweeps_per_step": Se, "deff_raw": r["deff_raw"],
               "note": "per rank: whole source image uploaded (1 B per pixel), device FloodFill over the global domain, assembly, S+1 sweeps with halo exchange, 2 checks -- the same steps as the N = 1 e2e"}

    line = {"metric": "jacobi_glups", "value": value, "unit": "GLUP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": data,
            "config": {"workload": "configs[1]: 00042.jpg x4 mesh amplification (%dx%d cells%s), 3-phase shipped "
                                   "input.txt defaults; step = %d sweeps + flux/Deff check" %
                                   (Nx, Ny, "" if world == 1 else " per GPU, stacked into %d row slabs, %s" % (world, "halo rows pushed into peer memory by the sweep kernel" if dom.peer else "NCCL halo exchange"), S),
                       "cells": cells, "sweeps_per_step": S, "l2_policy": "working set 2x%.0f MB > 126 MB L2, no flush needed" % (Nx * Ny * 8 / 1e6),
                       "kernel": args.kernel, "tblock": args.tblock},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "deff_raw": deff}
    configs = {}
    trace("main timed region and e2e done")
    if args.batch_images > 0:
        # image batches shard with no communication: every rank solves its own, distinct slice
        bl = leg_c3(ctx, rank, world, args.batch_images, oracle_checks=args.c3_oracle_checks if world == 1 else 0,
                    serial_checks=4 if world > 1 else 0)
        if world > 1:
            t = torch.tensor([bl["seconds"], float(bl["sweeps_total"])], device="cuda", dtype=torch.float64)
            dist.all_reduce(t[:1], op=dist.ReduceOp.MAX)
            dist.all_reduce(t[1:], op=dist.ReduceOp.SUM)
            bl["seconds"] = float(t[0].item())
            bl["images"] = args.batch_images * world
            bl["images_per_s"] = bl["images"] / bl["seconds"]
            bl["sweeps_total"] = int(t[1].item())
            bl["glups"] = bl["sweeps_total"] * 65536 / bl["seconds"] / 1e9
            bl["note"] = "images, seconds (max over ranks) and sweeps are whole-job; histogram and oracle checks are rank 0's slice"
        configs["c3"] = bl
        line["batch"] = {k: bl[k] for k in ("workload", "images", "seconds", "images_per_s", "glups")}
    trace("c3 done")
    if not args.no_c4:
        configs["c4"] = leg_c4(ctx, rank, local_rank, world, size=args.c4_size)
    trace("c4 done")
    if world > 1:
        line["slab_parity"] = leg_slab_parity(rank, local_rank, world)
    trace("slab parity done")
    if world == 1:
        if not args.no_c1:
            configs["c1"] = leg_c1(ctx)
        if not args.no_c5:
            configs["c5"] = leg_c5(ctx, args.c5_oracle_sweeps)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(img, args.cpu_sweeps)
            configs["c2"] = leg_c2_parity(ctx, img, p, line["cpu_baseline"])
            try:
                rc = ref_cuda_baseline(img, args.ref_cuda_sweeps)
            except Exception as e:      # a reported baseline must not take the bench line down
                rc = {"unavailable": "%s: %s" % (type(e).__name__, e)}
            if rc:
                line["ref_cuda_baseline"] = rc
        line["configs"] = configs
        line["parity"] = {k: v.get("parity") for k, v in configs.items()}
        if "slab_parity" in line:
            line["parity"]["slab_vs_single_gpu"] = line["slab_parity"]["parity"]
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sweeps-per-step", type=int, default=10000)
    ap.add_argument("--kernel", type=int, default=0)
    ap.add_argument("--tblock", type=int, default=0)
    ap.add_argument("--cpu-sweeps", type=int, default=40)
    ap.add_argument("--ref-sweeps", type=int, default=20)
    ap.add_argument("--ref-crop", type=int, default=0, help="use only the first N source rows for the reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-cuda-sweeps", type=int, default=1001)
    ap.add_argument("--batch-images", type=int, default=512, help="config 3: images per GPU through the packed batch mode (0: skip)")
    ap.add_argument("--c3-oracle-checks", type=int, default=4, help="config 3: images of rank 0 re-solved by the CPU oracle")
    ap.add_argument("--c4-size", type=int, default=16384)
    ap.add_argument("--c5-oracle-sweeps", type=int, default=60)
    ap.add_argument("--no-c1", action="store_true")
    ap.add_argument("--no-c4", action="store_true")
    ap.add_argument("--no-c5", action="store_true")
    ap.add_argument("--allow-short-warmup", action="store_true")
    ap.add_argument("--no-clock-sampler", action="store_true", help="diagnostic: do not poll NVML during the timed region")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
