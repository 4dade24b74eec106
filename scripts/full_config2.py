#!/usr/bin/env python
"""BASELINE config 2 end to end: bundled 00042.jpg, MeshAmp 4, shipped 3-phase defaults, all 7
continuation stages to the reference's stop rule.  GPU box only (minutes)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import effectivediffusivityfvm_b200 as E  # noqa: E402

amp = int(sys.argv[1]) if len(sys.argv) > 1 else 4
img = np.load(os.path.join(ROOT, "tests", "golden", "images.npz"))["img00042"]
ctx = E.Deff2D(0)
p = E.default_params(amp_x=amp, amp_y=amp)
t0 = time.perf_counter()
r = ctx.solve_image(img, p)
dt = time.perf_counter() - t0
cells = img.size * amp * amp
print(json.dumps({"amp": amp, "cells": cells, "wall_s": dt, "iters": r["iters"], "stage_D": r["stage_D"],
                  "stage_deff_raw": r["stage_deff_raw"], "deff": r["deff"], "conv": r["conv"], "pathflag": r["pathflag"],
                  "SVF": r["SVF"], "LVF": r["LVF"], "solve_ms": r["solve_ms"], "total_ms": r["total_ms"],
                  "glups": cells * r["total_iters"] / (r["total_ms"] * 1e-3) / 1e9}))
