#!/usr/bin/env python
"""Sweep-kernel timing on small domains (launch-bound regime): streaming K3 vs tiled K2 at several depths."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import effectivediffusivityfvm_b200 as E  # noqa: E402
from effectivediffusivityfvm_b200.datasets import c3_image  # noqa: E402

ctx = E.Deff2D(0)
p = E.default_params(Ds=1e-3, Df=1.0)
for size in (64, 128, 256, 512, 1024, 2048, 4096):
    img = c3_image(1, size)
    ctx.domain_load(img, 2, p)
    out = []
    for kernel, T in ((1, 1), (5, 4), (5, 6), (5, 8)):
        ctx.set_kernel(kernel, T)
        n = 4800
        ctx.sweeps_timed(240)
        ms = min(ctx.sweeps_timed(n) for _ in range(3))
        out.append("K%d/T%d %.2f us/sweep" % (kernel, T, ms * 1e3 / n))
    print("%4d^2: %s" % (size, ", ".join(out)), flush=True)
