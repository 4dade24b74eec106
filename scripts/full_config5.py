#!/usr/bin/env python
"""BASELINE config 5 end to end: 2048 x 2048 site percolation (p = 0.60), Ds/Df = 1e-4, 2-phase batch
semantics, tol 1e-5, MaxIter 5e5 -- the reference's stop rule under slow contraction.  GPU box only."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import effectivediffusivityfvm_b200 as E  # noqa: E402
from effectivediffusivityfvm_b200.datasets import c5_image  # noqa: E402

img = c5_image()
ctx = E.Deff2D(0)
p = E.default_params(Ds=1e-4, Df=1.0, mode=E.MODE_2PH_BATCH, tol=1e-5, max_iter=500000)
ctx.solve_image(img[:256, :256].copy(), E.default_params(Ds=1e-4, Df=1.0, mode=E.MODE_2PH_BATCH, max_iter=100))
t0 = time.perf_counter()
r = ctx.solve_image(img, p)
dt = time.perf_counter() - t0
print(json.dumps({"cells": img.size, "wall_s": dt, "iters": r["iters"], "deff": r["deff"], "conv": r["conv"],
                  "porosity": r["porosity"], "pathflag": r["pathflag"], "solve_ms": r["solve_ms"],
                  "glups": img.size * r["total_iters"] / (r["total_ms"] * 1e-3) / 1e9}))
