#!/usr/bin/env python
"""Loads the benchmark domain and runs a few sweeps with one kernel configuration (profiling
target for ncu).  python scripts/run_sweeps.py <kernel> <T> [sweeps] [amp]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import effectivediffusivityfvm_b200 as E  # noqa: E402

kernel, T = int(sys.argv[1]), int(sys.argv[2])
sweeps = int(sys.argv[3]) if len(sys.argv) > 3 else 8 * T
amp = int(sys.argv[4]) if len(sys.argv) > 4 else 4
img = np.load(os.path.join(ROOT, "tests", "golden", "images.npz"))["img00042"]
ctx = E.Deff2D(0)
ctx.domain_load(img, 3, E.default_params(amp_x=amp, amp_y=amp))
ctx.set_kernel(kernel, T)
ms = ctx.sweeps_timed(sweeps)
print("kernel %d T=%d: %d sweeps in %.3f ms -> %.1f GLUP/s, Deff_raw %.12g" %
      (kernel, T, sweeps, ms, img.size * amp * amp * sweeps / ms / 1e6, ctx.flux()[0]))
