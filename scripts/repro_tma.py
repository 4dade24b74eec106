import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import effectivediffusivityfvm_b200 as E
rng = np.random.default_rng(0)
img = np.where(rng.random((40, 300)) < 0.6, 0, 255).astype(np.uint8)
ctx = E.Deff2D(0)
p = E.default_params(Ds=1e-3, Df=1.0)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ctx.set_kernel(1)
ctx.domain_load(img, 2, p)
ctx.sweeps(T)
ref = ctx.get_field()
ctx.set_kernel(2, T)
ctx.domain_load(img, 2, p)
ctx.sweeps(T)
got = ctx.get_field()
print("max diff", np.nanmax(np.abs(got - ref)), "equal", np.array_equal(got, ref, equal_nan=True))
