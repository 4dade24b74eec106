"""numpy emulation of the device data model (padded iterate + phase codes + weight LUT) used by
the CPU tests to check the host-side tables and the matrix-free formulation against the
oracle's materialised A, b -- the same arithmetic the CUDA kernels perform, minus FMA."""
import numpy as np

XOFF = 16


def phase_codes(img, nphase, ampx=1, ampy=1, grid=None):
    """bits 0-1 phase (0 fluid, 1 solid, 2 gas), bit 2 pinned; amplified (Ny, Nx)."""
    a = np.repeat(np.repeat(np.asarray(img, dtype=np.uint8), ampy, axis=0), ampx, axis=1)
    if nphase == 2:
        c = np.where(a < 150, 0, 1).astype(np.uint8)                    # cuh:1779
    else:
        c = np.where(a > 200, 1, np.where(a < 50, 2, 0)).astype(np.uint8)   # cuh:1565-1576
    if grid is not None:
        c = c | (((grid == 1) | (grid == 2)).astype(np.uint8) << 2)    # cuh:750
    return c


def pad_codes(codes):
    Ny, Nx = codes.shape
    pitch = ((Nx + 2 * XOFF) + 15) // 16 * 16
    p = np.full((Ny + 2, pitch), 3, dtype=np.uint8)
    p[1:Ny + 1, XOFF:XOFF + Nx] = codes
    return p


def pad_field(x):
    Ny, Nx = x.shape
    pitch = ((Nx + 2 * XOFF) + 15) // 16 * 16
    p = np.zeros((Ny + 2, pitch), dtype=np.float64)
    p[1:Ny + 1, XOFF:XOFF + Nx] = x
    p[1:Ny + 1, XOFF - 1] = 1.0
    p[1:Ny + 1, XOFF + Nx] = 1.0
    return p


def lut_index(pc, Nx, Ny):
    """11-bit LUT index of every interior cell from the padded codes."""
    r = slice(1, Ny + 1)
    c = pc[r, XOFF:XOFF + Nx].astype(np.int64)
    w = pc[r, XOFF - 1:XOFF + Nx - 1] & 3
    e = pc[r, XOFF + 1:XOFF + Nx + 1] & 3
    s = pc[2:Ny + 2, XOFF:XOFF + Nx] & 3
    n = pc[0:Ny, XOFF:XOFF + Nx] & 3
    return (c & 3) | (w.astype(np.int64) << 2) | (e.astype(np.int64) << 4) | (s.astype(np.int64) << 6) | \
        (n.astype(np.int64) << 8) | ((c & 4) << 8)


def sweep(px, pc, lut, Nx, Ny, omega=2.0 / 3.0, nsweeps=1):
    """x' = (1-w) x + wW xW + wE xE + wS xS + wN xN on the padded arrays (ghosts untouched)."""
    idx = lut_index(pc, Nx, Ny)
    w = lut[idx]                       # (Ny, Nx, 4)
    om = 1.0 - omega
    cur = px.copy()
    r = slice(1, Ny + 1)
    for _ in range(nsweeps):
        c = cur[r, XOFF:XOFF + Nx]
        acc = om * c
        acc = acc + w[..., 0] * cur[r, XOFF - 1:XOFF + Nx - 1]
        acc = acc + w[..., 1] * cur[r, XOFF + 1:XOFF + Nx + 1]
        acc = acc + w[..., 2] * cur[2:Ny + 2, XOFF:XOFF + Nx]
        acc = acc + w[..., 3] * cur[0:Ny, XOFF:XOFF + Nx]
        nxt = cur.copy()
        nxt[r, XOFF:XOFF + Nx] = acc
        cur = nxt
    return cur
