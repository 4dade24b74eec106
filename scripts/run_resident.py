#!/usr/bin/env python
"""Profiling target: n cluster-resident sweeps of one size x size config-3 image.  python scripts/run_resident.py [size] [sweeps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import effectivediffusivityfvm_b200 as E  # noqa: E402
from effectivediffusivityfvm_b200.datasets import c3_image  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
ctx = E.Deff2D(0)
ctx.domain_load(c3_image(1, 256)[:size, :size], 2, E.default_params(Ds=1e-3, Df=1.0, mode=E.MODE_2PH_BATCH))
ms = ctx.sweeps_timed(n)
print("%d x %d: %d sweeps in %.3f ms = %.3f us/sweep, Deff_raw %.12g" % (size, size, n, ms, ms * 1e3 / n, ctx.flux()[0]))
