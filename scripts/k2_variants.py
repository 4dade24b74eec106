#!/usr/bin/env python
"""A/B of the tiled sweep's two thread layouts (GPU box only):
   python scripts/k2_variants.py [sweeps] [layout,...] [T,...]      layout 3 = 4 x 4 cells per thread, 4 = 2 x 8
For every layout: bit-identity against the streaming kernel on a small 3-phase domain for T = 1..8, then GLUP/s on
config 2 (00042.jpg x4), a 4096^2 blob medium and a 2048^2 site-percolation medium (every cell its own weights)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import effectivediffusivityfvm_b200 as E  # noqa: E402
from effectivediffusivityfvm_b200.datasets import c4_image, c5_image  # noqa: E402

sweeps = int(sys.argv[1]) if len(sys.argv) > 1 else 240
img2 = np.load(os.path.join(ROOT, "tests", "golden", "images.npz"))["img00042"]
rng = np.random.default_rng(3)
z = rng.random((131, 257))
for _ in range(2):
    z = (z + np.roll(z, 1, 0) + np.roll(z, -1, 0) + np.roll(z, 1, 1) + np.roll(z, -1, 1)) / 5
q1, q2 = np.quantile(z, [0.3, 0.7])
small = np.where(z < q1, 0, np.where(z < q2, 150, 255)).astype(np.uint8)
blob = c4_image(4096)
perc = c5_image()
out = {}
layouts = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [3, 4]
depths = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [4, 5, 6, 7, 8]
for fam in layouts:
    ctx = E.Deff2D(0)
    p3 = E.default_params(Ds=0.0, Df=1.0, Dg=80.0, CL=0.25, CR=1.5)
    ctx.set_kernel(1)
    ctx.domain_load(small, 3, p3)
    ctx.sweeps(29)
    ref = ctx.get_field()
    ok = True
    for T in range(1, 9):
        ctx.set_kernel(fam, T)
        ctx.domain_load(small, 3, p3)
        ctx.sweeps(29)
        ok = ok and np.array_equal(ctx.get_field(), ref, equal_nan=True)
    # long run through the graph path
    ctx.set_kernel(1); ctx.domain_load(small, 3, p3); ctx.sweeps(1000); ref = ctx.get_field()
    ctx.set_kernel(fam, 8); ctx.domain_load(small, 3, p3); ctx.sweeps(1000)
    ok = ok and np.array_equal(ctx.get_field(), ref, equal_nan=True)
    row = {"bit_identical": bool(ok)}
    for name, img, nph, par in (("c2", img2, 3, E.default_params(amp_x=4, amp_y=4)),
                                ("blob4096", blob, 2, E.default_params(Ds=1e-3, Df=1.0, mode=E.MODE_2PH_BATCH)),
                                ("perc2048", perc, 2, E.default_params(Ds=1e-4, Df=1.0, mode=E.MODE_2PH_BATCH))):
        ctx.domain_load(img, nph, par)
        cells = img.size * par.amp_x * par.amp_y
        for T in depths:
            ctx.set_kernel(fam, T)
            n = sweeps // T * T
            ctx.sweeps_timed(4 * T)
            ms = min(ctx.sweeps_timed(n) for _ in range(3))
            row["%s_T%d" % (name, T)] = round(cells * n / ms / 1e6, 1)
    out["layout%d" % fam] = row
    print("layout %d: %s" % (fam, json.dumps(row)), flush=True)
    ctx.close()
print(json.dumps(out))
