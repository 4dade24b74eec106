#!/usr/bin/env python
"""A/B timing of one sweep-kernel configuration: python scripts/ab_sweep.py <kernel> <T> [sweeps] [repeats]
(select the library with DEFF2D_LIB=...).  GPU box only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import effectivediffusivityfvm_b200 as E  # noqa: E402

kernel, T = int(sys.argv[1]), int(sys.argv[2])
sweeps = int(sys.argv[3]) if len(sys.argv) > 3 else 400
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
img = np.load(os.path.join(ROOT, "tests", "golden", "images.npz"))["img00042"]
ctx = E.Deff2D(0)
ctx.domain_load(img, 3, E.default_params(amp_x=4, amp_y=4))
ctx.set_kernel(kernel, T)
ctx.sweeps_timed(40)
ms = [ctx.sweeps_timed(sweeps) for _ in range(reps)]
cells = img.size * 16
print("%s kernel %d T=%d: best %.1f median %.1f GLUP/s" % (os.environ.get("DEFF2D_LIB", "current"), kernel, T,
      cells * sweeps / min(ms) / 1e6, cells * sweeps / float(np.median(ms)) / 1e6), flush=True)
