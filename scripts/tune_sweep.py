#!/usr/bin/env python
"""Times the sweep kernels on the benchmark domain (config 2) for every temporal depth.
GPU box only:  python scripts/tune_sweep.py [sweeps]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import effectivediffusivityfvm_b200 as E  # noqa: E402

sweeps = int(sys.argv[1]) if len(sys.argv) > 1 else 240
img = np.load(os.path.join(ROOT, "tests", "golden", "images.npz"))["img00042"]
ctx = E.Deff2D(0)
p = E.default_params(amp_x=4, amp_y=4)
ctx.domain_load(img, 3, p)
cells = 4008 * 8028
out = {}
for kernel, T in [(1, 1)] + [(2, t) for t in range(1, 9)] + [(3, t) for t in (4, 6)] + [(4, t) for t in range(1, 9)]:
    ctx.set_kernel(kernel, T)
    ctx.sweeps_timed(max(T * 4, 8))
    ms = min(ctx.sweeps_timed(sweeps) for _ in range(3))
    glups = cells * sweeps / ms / 1e6
    out["K%d_T%d" % (kernel, T)] = glups
    print("kernel %d T=%d: %.1f GLUP/s  (%.1f us/sweep, %.0f GB/s algorithmic)" % (kernel, T, glups, ms * 1e3 / sweeps, glups * 16), flush=True)
print(json.dumps(out))
