import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
import effectivediffusivityfvm_b200 as E
img = np.load('/root/repo/tests/golden/images.npz')['img00042']
ctx = E.Deff2D(0)
ctx.domain_load(img, 3, E.default_params(amp_x=4, amp_y=4))
ctx.sweeps_timed(240)
r = [img.size * 16 * 1200 / ctx.sweeps_timed(1200) / 1e6 for _ in range(5)]
print(os.environ.get('DEFF2D_LIB', 'new'), ' '.join('%.1f' % v for v in r))
