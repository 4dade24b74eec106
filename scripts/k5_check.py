#!/usr/bin/env python
"""Cluster-resident sweeps (K5, csrc/resident.cu) against the streaming kernel and the tiled packed batch:
bit-identity on single domains of up to 256 x 256 cells and on a packed batch, then timings.  GPU box only."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import effectivediffusivityfvm_b200 as E  # noqa: E402
from effectivediffusivityfvm_b200.datasets import c3_image  # noqa: E402


def blobs(seed, shape, levels=(0, 150, 255), fracs=(0.3, 0.4), smooth=2):
    rng = np.random.default_rng(seed)
    z = rng.random(shape)
    for _ in range(smooth):
        z = (z + np.roll(z, 1, 0) + np.roll(z, -1, 0) + np.roll(z, 1, 1) + np.roll(z, -1, 1)) / 5
    qs = np.quantile(z, np.cumsum(fracs))
    out = np.full(shape, levels[-1], np.uint8)
    for lv, q in reversed(list(zip(levels[:-1], qs))):
        out[z < q] = lv
    return out


nimg = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ctx = E.Deff2D(0)
ok_all = True
for shape in [(64, 64), (24, 16), (100, 130), (65, 64), (64, 65), (200, 70), (129, 255), (256, 256)]:
    for nphase in (2, 3):
        img = blobs(shape[0] * 7 + nphase, shape)
        p = E.default_params(Ds=0.0 if nphase == 3 else 1e-3, Df=1.0, Dg=80.0, CL=0.25, CR=1.5)
        for n in (1, 2, 3, 29, 1000):
            ctx.set_kernel(1)
            ctx.domain_load(img, nphase, p)
            ctx.sweeps(n)
            ref = ctx.get_field()
            dref = ctx.flux()[0]
            ctx.set_kernel(0)
            ctx.set_resident(0)
            ctx.domain_load(img, nphase, p)
            l0 = ctx.kernel_launches
            ctx.sweeps(n)
            launches = ctx.kernel_launches - l0
            got = ctx.get_field()
            same = np.array_equal(got, ref, equal_nan=True) and (ctx.flux()[0] == dref or np.isnan(dref))
            ok_all = ok_all and same and launches == 1
            if not same or launches != 1:
                print("MISMATCH shape %s nphase %d n %d launches %d maxdiff %g" % (shape, nphase, n, launches, np.nanmax(np.abs(got - ref))), flush=True)
print("single-domain bit identity:", ok_all, flush=True)

# full solves, resident vs tiled
img = blobs(5, (128, 128))
res = {}
for mode in (1, 0):
    ctx.set_resident(mode)
    t0 = time.perf_counter()
    res[mode] = ctx.solve_image(img, E.default_params(Dg=1000.0))
    res[mode]["sec"] = time.perf_counter() - t0
same = res[0]["iters"] == res[1]["iters"] and res[0]["deff"] == res[1]["deff"]
ok_all = ok_all and same
print("3-phase 128x128 solve: same %s iters %s  tiled %.3f s  resident %.3f s" % (same, res[0]["iters"], res[1]["sec"], res[0]["sec"]), flush=True)

# packed batch
imgs = np.stack([c3_image(k) for k in range(nimg)])
p = E.default_params(Ds=1e-3, Df=1.0, mode=E.MODE_2PH_BATCH, tol=1e-5, max_iter=500000)
out = {}
for mode in (1, 2):
    ctx.set_resident(mode)
    ctx.solve_batch(imgs[:2], E.default_params(Ds=1e-3, Df=1.0, mode=E.MODE_2PH_BATCH, max_iter=50))
    t0 = time.perf_counter()
    r = ctx.solve_batch(imgs, p)
    dt = time.perf_counter() - t0
    sweeps = sum(x["total_iters"] for x in r)
    out[mode] = (r, dt, sweeps)
    print("batch of %d, resident_mode %d: %.2f s, %.1f images/s, %.0f GLUP/s" % (nimg, mode, dt, nimg / dt, sweeps * 65536 / dt / 1e9), flush=True)
same = all(a["iters"] == b["iters"] and a["deff"] == b["deff"] and a["conv"] == b["conv"] for a, b in zip(out[2][0], out[1][0]))
ok_all = ok_all and same
print("batch bit identity:", same)
print(json.dumps({"ok": bool(ok_all)}))
sys.exit(0 if ok_all else 1)
