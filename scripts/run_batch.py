#!/usr/bin/env python
"""Short packed-batch run (profiling target for ncu): python scripts/run_batch.py [count] [size] [max_iter]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import effectivediffusivityfvm_b200 as E  # noqa: E402
from effectivediffusivityfvm_b200.datasets import c3_image  # noqa: E402

count = int(sys.argv[1]) if len(sys.argv) > 1 else 128
size = int(sys.argv[2]) if len(sys.argv) > 2 else 256
max_iter = int(sys.argv[3]) if len(sys.argv) > 3 else 401
imgs = np.stack([c3_image(k, size) for k in range(count)])
ctx = E.Deff2D(0)
p = E.default_params(Ds=1e-3, Df=1.0, mode=E.MODE_2PH_BATCH, tol=1e-5, max_iter=max_iter)
t0 = time.perf_counter()
res = ctx.solve_batch(imgs, p)
dt = time.perf_counter() - t0
print("packed batch: %d images %dx%d, %d sweeps each in %.3f s, Deff[0] %.9g" % (count, size, size, max_iter, dt, res[0]["deff"]))
