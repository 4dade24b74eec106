#!/usr/bin/env python
"""Peer-mode slab pass timed on ONE GPU with the rank as its own upper and lower neighbour (diagnostic, GPU box only).
The results are those of a vertically periodic domain, not of the reference; the point is the cost of the fused
exchange protocol (flag waits, pushes, fences, the LIST / PEER kernel variant) without NVLink in the picture.
   python scripts/peer_self.py [sweeps]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import effectivediffusivityfvm_b200 as E  # noqa: E402
from effectivediffusivityfvm_b200 import api  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
sweeps = int(args[0]) if args else 3000
peer_only = "--peer-only" in sys.argv
img = np.load(os.path.join(ROOT, "tests", "golden", "images.npz"))["img00042"]
p = E.default_params(amp_x=4, amp_y=4)
cells = img.size * 16
out = {}

if not peer_only:
    ctx = E.Deff2D(0)
    ctx.domain_load(img, 3, p)
    ctx.sweeps(600); ctx.sync()
    t0 = time.perf_counter(); ctx.sweeps(sweeps); ctx.sync(); dt = time.perf_counter() - t0
    out["single_domain"] = {"us_per_pass": dt / (sweeps / 6) * 1e6, "glups": cells * sweeps / dt / 1e9}
    ctx.close()

for halo in ((8,) if peer_only else (8, 32)):
    ctx = E.Deff2D(0)
    ctx.nccl_init(api.nccl_unique_id(), 0, 1)
    g = np.ascontiguousarray(np.tile(img, (3, 1)))
    own = img.shape[0] * 4
    ctx.domain_load_slab_global(g, 3, p, own, own, halo)
    h = ctx.slab_peer_export()
    ctx.slab_peer_attach(h, h)
    ctx.slab_sweeps(600); ctx.sync()
    t0 = time.perf_counter(); ctx.slab_sweeps(sweeps); ctx.sync(); dt = time.perf_counter() - t0
    import hashlib
    f = ctx.get_field()
    out["self_peer_h%d" % halo] = {"us_per_pass": dt / (sweeps / 6) * 1e6, "glups": cells * sweeps / dt / 1e9,
                                   "own_rows_sha": hashlib.sha1(np.ascontiguousarray(f[halo:halo + own]).tobytes()).hexdigest()[:16]}
    ctx.close()
print(json.dumps(out))
