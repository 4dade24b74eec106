#!/usr/bin/env python
"""Times deff2d_domain_load (H2D + FloodFill + assembly + tables) with the host and the device
FloodFill on BASELINE config 2 and on synthetic large domains.  GPU box only."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import effectivediffusivityfvm_b200 as E  # noqa: E402
from effectivediffusivityfvm_b200.datasets import c5_image  # noqa: E402

ctx = E.Deff2D(0)
img = np.load(os.path.join(ROOT, "tests", "golden", "images.npz"))["img00042"]
cases = [("config2 00042.jpg x4, 3-phase", img, 3, E.default_params(amp_x=4, amp_y=4)),
         ("config5 2048^2 percolation, 2-phase", c5_image(), 2, E.default_params(Ds=1e-4, Df=1.0)),
         ("8192^2 percolation, 2-phase", c5_image(8192, 6), 2, E.default_params(Ds=1e-4, Df=1.0))]
for name, im, nphase, p in cases:
    for mode, label in ((1, "host"), (2, "device")):
        ctx.set_floodfill(mode)
        ctx.domain_load(im, nphase, p)
        t0 = time.perf_counter()
        ctx.domain_load(im, nphase, p)
        dt = time.perf_counter() - t0
        print("%-40s FloodFill on %-6s: domain_load %.1f ms, PathFlag %d" % (name, label, dt * 1e3, ctx.info()["pathflag"]), flush=True)
