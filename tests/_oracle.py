"""ctypes bindings onto the CPU oracle (oracle/liboracle.so) and, when present, the
reference itself built on host threads (oracle/_ref/libref_cpu.so) or for the GPU
(oracle/_ref/libref_cuda.so).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by the package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_CPU_SO = os.path.join(ORACLE_DIR, "_ref", "libref_cpu.so")
REF_CUDA_SO = os.path.join(ORACLE_DIR, "_ref", "libref_cuda.so")
REF_CPU_EXE = os.path.join(ORACLE_DIR, "_ref", "deff2d_ref_cpu")
REF_CUDA_EXE = os.path.join(ORACLE_DIR, "_ref", "deff2d_ref_cuda")

c_double_p = C.POINTER(C.c_double)
c_ubyte_p = C.POINTER(C.c_ubyte)
c_uint_p = C.POINTER(C.c_uint)


class OrcOpts(C.Structure):
    _fields_ = [("Ds", C.c_double), ("Df", C.c_double), ("Dg", C.c_double),
                ("ampx", C.c_int), ("ampy", C.c_int),
                ("CL", C.c_double), ("CR", C.c_double),
                ("max_iter", C.c_long), ("tol", C.c_double),
                ("nphase", C.c_int), ("check_every", C.c_int),
                ("omega", C.c_double), ("verbose", C.c_int)]


class OrcResult(C.Structure):
    _fields_ = [("porosity", C.c_double), ("SVF", C.c_double), ("LVF", C.c_double),
                ("deff", C.c_double), ("deff_raw", C.c_double), ("conv", C.c_double),
                ("pathflag", C.c_int), ("nstages", C.c_int),
                ("iters", C.c_long * 16), ("stage_deff_raw", C.c_double * 16),
                ("stage_D", C.c_double * 16), ("total_iters", C.c_long),
                ("loop_seconds", C.c_double), ("nchecks", C.c_long)]


def make_opts(Ds=0.0, Df=1.0, Dg=1237500.0, ampx=1, ampy=1, CL=0.0, CR=1.0, max_iter=500000,
              tol=1e-5, nphase=3, check_every=10000, omega=2.0 / 3.0):
    return OrcOpts(Ds, Df, Dg, ampx, ampy, CL, CR, int(max_iter), tol, nphase, check_every, omega, 0)


def build_oracle():
    """(Re)build liboracle.so if it is missing or older than its source."""
    src = os.path.join(ORACLE_DIR, "deff_oracle.c")
    if (not os.path.exists(ORACLE_SO)) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"], stdout=subprocess.DEVNULL)


_orc = None


def oracle():
    global _orc
    if _orc is None:
        build_oracle()
        L = C.CDLL(ORACLE_SO)
        L.orc_porosity.restype = C.c_double
        L.orc_porosity.argtypes = [c_ubyte_p, C.c_int, C.c_int]
        L.orc_weighted_harmonic_mean.restype = C.c_double
        L.orc_weighted_harmonic_mean.argtypes = [C.c_double] * 4
        L.orc_fracts3.argtypes = [c_double_p, C.c_int, C.c_int, C.c_double, C.c_double, c_double_p, c_double_p]
        L.orc_fill_D.argtypes = [c_ubyte_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_double, C.c_double, C.c_double, c_double_p]
        L.orc_grid_mask.argtypes = [c_ubyte_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_uint_p]
        L.orc_floodfill.argtypes = [c_uint_p, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.orc_discretize.argtypes = [c_double_p, c_double_p, c_double_p, C.c_int, C.c_int,
                                     C.c_double, C.c_double, c_uint_p]
        L.orc_sweep.argtypes = [c_double_p, c_double_p, c_double_p, c_double_p, C.c_long, C.c_int, C.c_double]
        L.orc_flux_deff.restype = C.c_double
        L.orc_flux_deff.argtypes = [c_double_p, c_double_p, C.c_int, C.c_int, C.c_double, C.c_double]
        L.orc_residual.restype = C.c_double
        L.orc_residual.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, c_double_p, c_double_p]
        L.orc_jacobi.restype = C.c_long
        L.orc_jacobi.argtypes = [c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, C.c_int, C.c_int,
                                 C.c_double, C.c_double, C.c_double, C.c_long, C.c_int, C.c_double,
                                 c_double_p, c_double_p, c_double_p, C.c_long, C.POINTER(C.c_long), c_double_p]
        L.orc_init_x.argtypes = [c_double_p, C.c_int, C.c_int, C.c_double, C.c_double]
        L.orc_solve_image.restype = C.c_int
        L.orc_solve_image.argtypes = [c_ubyte_p, C.c_int, C.c_int, C.POINTER(OrcOpts), C.c_int,
                                      C.POINTER(OrcResult), c_double_p, c_double_p, C.c_long]
        L.orc_time_sweeps.restype = C.c_double
        L.orc_time_sweeps.argtypes = [c_ubyte_p, C.c_int, C.c_int, C.POINTER(OrcOpts), C.c_int, C.c_long, c_double_p]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        _orc = L
    return _orc


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _up(a):
    return a.ctypes.data_as(c_ubyte_p)


def _ip(a):
    return a.ctypes.data_as(c_uint_p)


MODE_2PH_SINGLE, MODE_2PH_BATCH, MODE_3PH = 0, 1, 2


def solve_image(img, opts, mode, want_field=False, trace_cap=0):
    """Run one image through the oracle's restatement of the reference drivers."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    H, W = img.shape
    res = OrcResult()
    field = np.empty((H * opts.ampy, W * opts.ampx), dtype=np.float64) if want_field else None
    trace = np.zeros(trace_cap, dtype=np.float64) if trace_cap else None
    rc = oracle().orc_solve_image(_up(img), W, H, C.byref(opts), mode, C.byref(res),
                                  _dp(field) if want_field else None,
                                  _dp(trace) if trace_cap else None, trace_cap)
    assert rc == 0
    out = {"porosity": res.porosity, "SVF": res.SVF, "LVF": res.LVF, "deff": res.deff,
           "deff_raw": res.deff_raw, "conv": res.conv, "pathflag": res.pathflag,
           "nstages": res.nstages, "iters": list(res.iters[:res.nstages]),
           "stage_deff_raw": list(res.stage_deff_raw[:res.nstages]),
           "stage_D": list(res.stage_D[:res.nstages]), "total_iters": res.total_iters,
           "loop_seconds": res.loop_seconds}
    if want_field:
        out["field"] = field
    if trace_cap:
        out["trace"] = trace[:min(trace_cap, res.nchecks)]
    return out


def fill_D(img, ampx, ampy, nphase, Ds, Df, Dg):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    H, W = img.shape
    D = np.empty((H * ampy, W * ampx), dtype=np.float64)
    oracle().orc_fill_D(_up(img), W, H, ampx, ampy, nphase, Ds, Df, Dg, _dp(D))
    return D


def grid_mask(img, ampx, ampy, thr):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    H, W = img.shape
    G = np.empty((H * ampy, W * ampx), dtype=np.uint32)
    oracle().orc_grid_mask(_up(img), W, H, ampx, ampy, thr, _ip(G))
    return G


def floodfill(G):
    G = np.ascontiguousarray(G, dtype=np.uint32).copy()
    Ny, Nx = G.shape
    pf = C.c_int(0)
    oracle().orc_floodfill(_ip(G), Nx, Ny, C.byref(pf))
    return G, pf.value


def discretize(D, CL, CR, Grid=None):
    D = np.ascontiguousarray(D, dtype=np.float64)
    Ny, Nx = D.shape
    A = np.empty((Ny * Nx, 5), dtype=np.float64)
    b = np.empty(Ny * Nx, dtype=np.float64)
    g = None
    if Grid is not None:
        Grid = np.ascontiguousarray(Grid, dtype=np.uint32)
        g = _ip(Grid)
    oracle().orc_discretize(_dp(D), _dp(A), _dp(b), Nx, Ny, CL, CR, g)
    return A, b


def init_x(Nx, Ny, CL, CR):
    x = np.empty((Ny, Nx), dtype=np.float64)
    oracle().orc_init_x(_dp(x), Nx, Ny, CL, CR)
    return x


def sweeps(A, b, x, nsweeps, omega=2.0 / 3.0):
    """nsweeps applications of the reference kernel (cuh:69-92); returns the new field."""
    Ny, Nx = x.shape
    cur = np.ascontiguousarray(x, dtype=np.float64).copy()
    nxt = np.empty_like(cur)
    L = oracle()
    for _ in range(nsweeps):
        L.orc_sweep(_dp(A), _dp(cur), _dp(b), _dp(nxt), Nx * Ny, Nx, omega)
        cur, nxt = nxt, cur
    return cur


def flux_deff(x, D, CL, CR):
    Ny, Nx = x.shape
    return oracle().orc_flux_deff(_dp(np.ascontiguousarray(x)), _dp(np.ascontiguousarray(D)), Nx, Ny, CL, CR)


def jacobi(A, b, x, D, CL, CR, tol, max_iter, check_every=10000, omega=2.0 / 3.0, trace_cap=64):
    Ny, Nx = x.shape
    x = np.ascontiguousarray(x, dtype=np.float64).copy()
    xt = np.empty_like(x)
    deff = C.c_double(0)
    conv = C.c_double(0)
    secs = C.c_double(0)
    nt = C.c_long(0)
    trace = np.zeros(trace_cap)
    it = oracle().orc_jacobi(_dp(A), _dp(b), _dp(x), _dp(xt), _dp(np.ascontiguousarray(D)), Nx, Ny, CL, CR,
                             tol, int(max_iter), check_every, omega, C.byref(deff), C.byref(conv),
                             _dp(trace), trace_cap, C.byref(nt), C.byref(secs))
    return {"iters": it, "deff_raw": deff.value, "conv": conv.value, "field": x,
            "trace": trace[:min(trace_cap, nt.value)], "seconds": secs.value}


# ----------------------------------------------------------------- the reference itself

_ref = {}


def reference(kind="cpu"):
    """The reference's own code (None if oracle/_ref has not been built)."""
    path = REF_CPU_SO if kind == "cpu" else REF_CUDA_SO
    if kind not in _ref:
        if not os.path.exists(path):
            _ref[kind] = None
        else:
            L = C.CDLL(path)
            L.ref_decode.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                     c_ubyte_p, C.c_long]
            L.ref_whm.restype = C.c_double
            L.ref_whm.argtypes = [C.c_double] * 4
            L.ref_porosity.restype = C.c_double
            L.ref_porosity.argtypes = [c_ubyte_p, C.c_int, C.c_int]
            L.ref_fracts3.argtypes = [c_double_p, C.c_int, C.c_int, C.c_double, C.c_double, c_double_p, c_double_p]
            L.ref_floodfill.argtypes = [c_uint_p, C.c_int, C.c_int]
            L.ref_discretize.argtypes = [c_double_p, c_double_p, c_double_p, C.c_int, C.c_int,
                                         C.c_double, C.c_double, c_uint_p]
            L.ref_residual.restype = C.c_double
            L.ref_residual.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, c_double_p, c_double_p]
            L.ref_jacobi.restype = C.c_long
            L.ref_jacobi.argtypes = [c_double_p, c_double_p, c_double_p, c_double_p, C.c_int, C.c_int,
                                     C.c_double, C.c_double, C.c_double, C.c_long, C.c_int,
                                     c_double_p, c_double_p, c_double_p]
            _ref[kind] = L
    return _ref[kind]


def ref_decode(path, kind="cpu"):
    L = reference(kind)
    W, H, nch = C.c_int(0), C.c_int(0), C.c_int(0)
    rc = L.ref_decode(path.encode(), C.byref(W), C.byref(H), C.byref(nch), None, 0)
    if rc != 0:
        raise IOError("reference decoder could not read %s" % path)
    out = np.empty((H.value, W.value), dtype=np.uint8)
    L.ref_decode(path.encode(), C.byref(W), C.byref(H), C.byref(nch), _up(out), out.size)
    return out, nch.value


def ref_floodfill(G, kind="cpu"):
    G = np.ascontiguousarray(G, dtype=np.uint32).copy()
    Ny, Nx = G.shape
    pf = reference(kind).ref_floodfill(_ip(G), Nx, Ny)
    return G, pf


def ref_discretize(D, CL, CR, Grid=None, kind="cpu"):
    D = np.ascontiguousarray(D, dtype=np.float64)
    Ny, Nx = D.shape
    A = np.empty((Ny * Nx, 5), dtype=np.float64)
    b = np.empty(Ny * Nx, dtype=np.float64)
    g = None
    if Grid is not None:
        Grid = np.ascontiguousarray(Grid, dtype=np.uint32)
        g = _ip(Grid)
    reference(kind).ref_discretize(_dp(D), _dp(A), _dp(b), Nx, Ny, CL, CR, g)
    return A, b


def ref_jacobi(A, b, x, D, CL, CR, tol, max_iter, precond=False, kind="cpu"):
    Ny, Nx = x.shape
    x = np.ascontiguousarray(x, dtype=np.float64).copy()
    deff, conv, ms = C.c_double(0), C.c_double(0), C.c_double(0)
    it = reference(kind).ref_jacobi(_dp(np.ascontiguousarray(A)), _dp(np.ascontiguousarray(b)), _dp(x),
                                    _dp(np.ascontiguousarray(D)), Nx, Ny, CL, CR, tol, int(max_iter),
                                    1 if precond else 0, C.byref(deff), C.byref(conv), C.byref(ms))
    return {"iters": it, "deff_raw": deff.value, "conv": conv.value, "field": x, "ms": ms.value}
