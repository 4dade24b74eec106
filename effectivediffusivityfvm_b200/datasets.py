"""Synthetic inputs of the BASELINE.json configs; numpy/scipy only, seeded."""
import numpy as np


def _smooth_noise(rng, shape, sigma):
    from scipy.ndimage import gaussian_filter
    return gaussian_filter(rng.standard_normal(shape), sigma=sigma, mode="wrap")


def c3_image(k, size=256):
    """Config 3, image k: Gaussian-filtered (sigma 3 px, periodic) white noise thresholded at its
    eps-quantile, eps ~ U(0.35, 0.75) from the same generator; pore = 0, solid = 255."""
    rng = np.random.default_rng(1234 + k)
    eps = rng.uniform(0.35, 0.75)
    z = _smooth_noise(rng, (size, size), 3.0)
    return np.where(z < np.quantile(z, eps), 0, 255).astype(np.uint8)


def c4_image(size=16384, seed=4, sigma=8.0, eps=0.6):
    """Config 4: one large two-phase porous domain (sigma 8 px, porosity 0.6).  Single precision
    noise and a separable filter row block by row block keep the 268 M-pixel case to seconds."""
    from scipy.ndimage import gaussian_filter1d
    rng = np.random.default_rng(seed)
    z = rng.standard_normal((size, size), dtype=np.float32)
    gaussian_filter1d(z, sigma, axis=1, mode="wrap", output=z)
    gaussian_filter1d(z, sigma, axis=0, mode="wrap", output=z)
    thr = np.partition(z.ravel()[::7], int(eps * (z.size // 7 + 1)))[int(eps * (z.size // 7 + 1))]   # eps-quantile of a 1/7 sample
    return np.where(z < thr, 0, 255).astype(np.uint8)


def c5_image(size=2048, seed=5, p=0.60):
    """Config 5: uncorrelated site percolation just above the threshold (p_c ~ 0.5927)."""
    rng = np.random.default_rng(seed)
    return np.where(rng.random((size, size)) < p, 0, 255).astype(np.uint8)
