#!/usr/bin/env python
"""Cycles per sweep of the cluster-resident kernel by cluster shape (GPU box only)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import effectivediffusivityfvm_b200 as E  # noqa: E402
from effectivediffusivityfvm_b200.datasets import c3_image  # noqa: E402

ctx = E.Deff2D(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
for shape in [(64, 64), (64, 128), (128, 64), (128, 128), (256, 128), (256, 256)]:
    img = c3_image(1, 256)[:shape[0], :shape[1]]
    p = E.default_params(Ds=1e-3, Df=1.0, mode=E.MODE_2PH_BATCH)
    for mode in (0, 1):
        ctx.set_resident(mode)
        ctx.domain_load(img, 2, p)
        ctx.sweeps_timed(1000)
        ms = ctx.sweeps_timed(n)
        print("%3dx%3d resident_mode %d: %.3f us/sweep = %.0f cycles at 1.965 GHz" % (shape[1], shape[0], mode, ms * 1e3 / n, ms * 1e-3 / n * 1.965e9), flush=True)
