#!/usr/bin/env python
"""Deff images/s of the packed batch mode on BASELINE config 3 (synthetic two-phase 256x256
microstructures, SURVEY.md 8(d) generator).  GPU box only.

  python scripts/bench_batch.py [count] [size] [slots] [--serial]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import effectivediffusivityfvm_b200 as E  # noqa: E402
from effectivediffusivityfvm_b200.datasets import c3_image  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
count = int(args[0]) if len(args) > 0 else 64
size = int(args[1]) if len(args) > 1 else 256
slots = int(args[2]) if len(args) > 2 else 0
serial = "--serial" in sys.argv

imgs = np.stack([c3_image(k, size) for k in range(count)])
ctx = E.Deff2D(0)
p = E.default_params(Ds=1e-3, Df=1.0, mode=E.MODE_2PH_BATCH, tol=1e-5, max_iter=500000)
ctx.set_batch_slots(slots)
ctx.solve_batch(imgs[:2], E.default_params(Ds=1e-3, Df=1.0, mode=E.MODE_2PH_BATCH, max_iter=100))      # warm-up
l0 = ctx.kernel_launches
t0 = time.perf_counter()
if serial:
    res = [ctx.solve_image(im, p) for im in imgs]
else:
    res = ctx.solve_batch(imgs, p)
dt = time.perf_counter() - t0
sweeps = np.array([r["total_iters"] for r in res], dtype=np.float64)
lups = float(sweeps.sum()) * size * size
print(json.dumps({"mode": "serial" if serial else "packed", "images": count, "size": size, "seconds": dt,
                  "images_per_s": count / dt, "glups": lups / dt / 1e9, "launches": ctx.kernel_launches - l0,
                  "sweeps_min": int(sweeps.min()), "sweeps_median": int(np.median(sweeps)), "sweeps_max": int(sweeps.max()),
                  "deff_mean": float(np.mean([r["deff"] for r in res]))}))
