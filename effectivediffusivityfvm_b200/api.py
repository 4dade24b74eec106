"""Host-side mirror of the reference's driver interface on top of the C ABI.

Names follow the reference (Deff2D.cuh): `options` keys as in input.txt, the four drivers
`SingleSim`, `BatchSim`, `SingleSim3Phase`, `BatchSim3Phase`, and `JacobiGPU`-level stepping on
a resident domain.  Everything numeric happens in libdeff2d.so (hand-written sm_100a CUDA);
this module only marshals buffers.  No CPU fallback exists.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import MODE_2PH_BATCH, MODE_2PH_SINGLE, MODE_3PH, Input, Params, Result


class Deff2DError(RuntimeError):
    pass


def default_params(**overrides):
    """Shipped input.txt defaults (Deff2DGPU/input.txt:2-18) with keyword overrides.

    Accepts the struct field names (Ds, Df, Dg, amp_x, amp_y, CL, CR, max_iter, tol, mode,
    check_every, omega, solver, verbose, residual_tol, strict_reference)."""
    p = Params()
    _lib.lib().deff2d_default_params(C.byref(p))
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise TypeError("unknown parameter %r" % k)
        setattr(p, k, v)
    return p


def read_input_file(path):
    """readInputFile (Deff2D.cuh:234-324): returns the parsed ctypes `Input`."""
    inp = Input()
    rc = _lib.lib().deff2d_read_input_file(str(path).encode(), C.byref(inp))
    if rc:
        raise Deff2DError("cannot read input file %s (status %d)" % (path, rc))
    return inp


def build_tables(Ds, Df, Dg, Nx, Ny, CL=0.0, CR=1.0, omega=2.0 / 3.0):
    """The per-stage coefficient LUT exactly as uploaded to the device (host code, no GPU)."""
    lut = np.empty((2048, 4), dtype=np.float64)
    dead = np.empty(2048, dtype=np.uint8)
    rc = _lib.lib().deff2d_build_tables(Ds, Df, Dg, Nx, Ny, CL, CR, omega,
                                        lut.ctypes.data_as(_lib.c_double_p), dead.ctypes.data_as(_lib.c_ubyte_p))
    if rc:
        raise Deff2DError("build_tables failed (%d)" % rc)
    return lut, dead


def floodfill(grid):
    """FloodFill (Deff2D.cuh:557-713) on a (Ny, Nx) uint8 solid mask; returns (grid, PathFlag)."""
    g = np.ascontiguousarray(grid, dtype=np.uint8).copy()
    Ny, Nx = g.shape
    pf = _lib.lib().deff2d_floodfill(g.ctypes.data_as(_lib.c_ubyte_p), Nx, Ny)
    if pf < 0:
        raise Deff2DError("floodfill failed (%d)" % pf)
    return g, pf


def tile_geometry(T):
    """(ow, oh, tw, th): output box and input tile of a tiled pass of depth T."""
    v = [C.c_int(0) for _ in range(4)]
    rc = _lib.lib().deff2d_tile_geometry(int(T), *[C.byref(x) for x in v])
    if rc:
        raise Deff2DError("tile_geometry failed (%d)" % rc)
    return tuple(x.value for x in v)


def batch_plan(Nx, Ny, count, limit=0):
    """(GX, GY) slot grid of the packed batch mode."""
    gx, gy = C.c_int(0), C.c_int(0)
    rc = _lib.lib().deff2d_batch_plan(Nx, Ny, count, limit, C.byref(gx), C.byref(gy))
    if rc:
        raise Deff2DError("batch_plan failed (%d)" % rc)
    return gx.value, gy.value


def batch_tile_list(Nx, Ny, GX, GY, active, T):
    """Tile ids whose output box touches an active slot of the packed stack."""
    cap = 1 << 20
    t = np.zeros(cap, dtype=np.uint32)
    a = np.ascontiguousarray(active, dtype=np.int32)
    n = _lib.lib().deff2d_batch_tile_list(Nx, Ny, GX, GY, a.ctypes.data_as(C.POINTER(C.c_int)), len(a), T,
                                          t.ctypes.data_as(C.POINTER(C.c_uint32)), cap)
    if n < 0:
        raise Deff2DError("batch_tile_list failed (%d)" % n)
    return t[:n].copy()


def load_image(path):
    """Decode an image file to (H, W) uint8 gray + the file's channel count (cuh:342)."""
    L = _lib.lib()
    ptr = _lib.c_ubyte_p()
    W, H, ch = C.c_int(0), C.c_int(0), C.c_int(0)
    rc = L.deff2d_load_image(str(path).encode(), C.byref(ptr), C.byref(W), C.byref(H), C.byref(ch))
    if rc:
        raise Deff2DError("cannot decode %s (status %d)" % (path, rc))
    try:
        arr = np.ctypeslib.as_array(ptr, shape=(H.value, W.value)).copy()
    finally:
        L.deff2d_free(ptr)
    return arr, ch.value


class Deff2D:
    """One persistent device context (replaces initializeGPU / unInitializeGPU, cuh:904-1021)."""

    def __init__(self, device=0):
        self._L = _lib.lib()
        self._h = C.c_void_p()
        rc = self._L.deff2d_create(C.byref(self._h), int(device))
        if rc:
            raise Deff2DError("deff2d_create failed (%d): %s" % (rc, self._L.deff2d_last_error(None).decode()))
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.deff2d_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc:
            raise Deff2DError("libdeff2d status %d: %s" % (rc, self._L.deff2d_last_error(self._h).decode()))

    # ---- whole-path calls (host buffers in, Deff out) ------------------------------------
    def solve_image(self, gray, params, want_field=False):
        gray = np.ascontiguousarray(gray, dtype=np.uint8)
        H, W = gray.shape
        res = Result()
        field = None
        fp = None
        if want_field:
            field = np.empty((H * params.amp_y, W * params.amp_x), dtype=np.float64)
            fp = field.ctypes.data_as(_lib.c_double_p)
        self._ck(self._L.deff2d_solve_image(self._h, gray.ctypes.data_as(_lib.c_ubyte_p), W, H,
                                            C.byref(params), C.byref(res), fp))
        out = res.as_dict()
        if want_field:
            out["field"] = field
        return out

    def solve_batch(self, images, params, want_fields=False):
        images = np.ascontiguousarray(images, dtype=np.uint8)
        count, H, W = images.shape
        res = (Result * max(count, 1))()
        fields = None
        fp = None
        if want_fields:
            fields = np.empty((count, H * params.amp_y, W * params.amp_x), dtype=np.float64)
            fp = fields.ctypes.data_as(_lib.c_double_p)
        self._ck(self._L.deff2d_solve_batch(self._h, images.ctypes.data_as(_lib.c_ubyte_p), count, W, H,
                                            C.byref(params), res, fp))
        out = [res[k].as_dict() for k in range(count)]
        if want_fields:
            for k in range(count):
                out[k]["field"] = fields[k]
        return out

    def solve_batch_stream(self, count, shape, params, fetch, done=None, want_fields=False):
        """The packed batch solve as a stream (deff2d_solve_batch_stream).

        fetch(k, wait) -> (H, W) uint8 array of image k, None if it is not ready yet (only when wait is False) or
        StopIteration-like `False` when there is no image k; done(k, result_dict, field_or_None) is called as each
        image finishes (any order).  Returns the number of images solved."""
        H, W = shape
        ncell = H * params.amp_y * W * params.amp_x
        err = []

        def _fetch(_user, k, dst, wait):
            try:
                img = fetch(k, bool(wait))
                if img is None:
                    return 1
                if img is False:
                    return 2
                a = np.ascontiguousarray(img, dtype=np.uint8)
                if a.shape != (H, W):
                    return 2
                C.memmove(dst, a.ctypes.data, a.size)
                return 0
            except Exception as e:      # exceptions must not cross the C frames
                err.append(e)
                return -2

        def _done(_user, k, res, field):
            try:
                if done is not None:
                    f = np.ctypeslib.as_array(field, shape=(H * params.amp_y, W * params.amp_x)).copy() if (want_fields and field) else None
                    done(k, res.contents.as_dict(), f)
                return 0
            except Exception as e:
                err.append(e)
                return 1

        solved = C.c_int(0)
        cb_f, cb_d = _lib.BATCH_FETCH_FN(_fetch), _lib.BATCH_DONE_FN(_done)
        rc = self._L.deff2d_solve_batch_stream(self._h, int(count), W, H, C.byref(params), cb_f, cb_d, None, 1 if want_fields else 0,
                                               C.byref(solved))
        if err:
            raise err[0]
        self._ck(rc)
        del ncell
        return solved.value

    # the reference's four drivers by name (Deff2D.cu:17-50)
    def SingleSim(self, gray, params, want_field=False):
        params.mode = MODE_2PH_SINGLE
        return self.solve_image(gray, params, want_field)

    def BatchSim(self, images, params):
        params.mode = MODE_2PH_BATCH
        return self.solve_batch(images, params)

    def SingleSim3Phase(self, gray, params, want_field=False):
        params.mode = MODE_3PH
        return self.solve_image(gray, params, want_field)

    def BatchSim3Phase(self, images, params, want_fields=False):
        params.mode = MODE_3PH
        return self.solve_batch(images, params, want_fields)

    def run_input_file(self, path="input.txt"):
        """The reference program: input.txt in, CSV / CMAP files out (relative to the CWD)."""
        self._ck(self._L.deff2d_run_input_file(self._h, str(path).encode()))

    # ---- resident-domain stepping --------------------------------------------------------
    def domain_load(self, gray, nphase, params):
        gray = np.ascontiguousarray(gray, dtype=np.uint8)
        H, W = gray.shape
        self._ck(self._L.deff2d_domain_load(self._h, gray.ctypes.data_as(_lib.c_ubyte_p), W, H, int(nphase),
                                            C.byref(params)))

    def domain_load_slab(self, gray, nphase, params, row0, ny_global, halo_rows, own_src_rows, pinned=None):
        gray = np.ascontiguousarray(gray, dtype=np.uint8)
        W = gray.shape[1]
        pp = None
        if pinned is not None:
            pinned = np.ascontiguousarray(pinned, dtype=np.uint8)
            pp = pinned.ctypes.data_as(_lib.c_ubyte_p)
        self._keep = (gray, pinned)
        self._ck(self._L.deff2d_domain_load_slab(self._h, gray.ctypes.data_as(_lib.c_ubyte_p), W, int(own_src_rows),
                                                 int(nphase), C.byref(params), int(row0), int(ny_global),
                                                 int(halo_rows), pp))

    def domain_load_slab_global(self, gray, nphase, params, row0, own_rows, halo_rows):
        """This rank's slab from the WHOLE source image: device FloodFill over the whole domain, no host mask."""
        gray = np.ascontiguousarray(gray, dtype=np.uint8)
        H, W = gray.shape
        self._keep = (gray,)
        self._ck(self._L.deff2d_domain_load_slab_global(self._h, gray.ctypes.data_as(_lib.c_ubyte_p), W, H, int(nphase),
                                                        C.byref(params), int(row0), int(own_rows), int(halo_rows)))

    def set_D(self, Ds, Df, Dg=0.0):
        self._ck(self._L.deff2d_domain_set_D(self._h, Ds, Df, Dg))

    def sweeps(self, n):
        self._ck(self._L.deff2d_domain_sweeps(self._h, int(n)))

    def sweeps_timed(self, n):
        ms = C.c_float(0)
        self._ck(self._L.deff2d_domain_sweeps_timed(self._h, int(n), C.byref(ms)))
        return ms.value

    def flux(self):
        d = C.c_double(0)
        q = (C.c_double * 2)()
        self._ck(self._L.deff2d_domain_flux(self._h, C.byref(d), q))
        return d.value, (q[0], q[1])

    def residual(self):
        r = C.c_double(0)
        self._ck(self._L.deff2d_domain_residual(self._h, C.byref(r)))
        return r.value

    def solve(self, tol, max_iter, trace_cap=256):
        it, nt = C.c_int64(0), C.c_int(0)
        d, cv = C.c_double(0), C.c_double(0)
        tr = np.zeros(trace_cap, dtype=np.float64)
        self._ck(self._L.deff2d_domain_solve(self._h, tol, int(max_iter), C.byref(it), C.byref(d), C.byref(cv),
                                             tr.ctypes.data_as(_lib.c_double_p), trace_cap, C.byref(nt)))
        return {"iters": it.value, "deff_raw": d.value, "conv": cv.value, "trace": tr[:min(nt.value, trace_cap)]}

    def info(self):
        Nx, Ny, pf = C.c_int64(0), C.c_int64(0), C.c_int(0)
        por, svf, lvf = C.c_double(0), C.c_double(0), C.c_double(0)
        self._ck(self._L.deff2d_domain_info(self._h, C.byref(Nx), C.byref(Ny), C.byref(pf), C.byref(por),
                                            C.byref(svf), C.byref(lvf)))
        return {"Nx": Nx.value, "Ny": Ny.value, "pathflag": pf.value, "porosity": por.value,
                "SVF": svf.value, "LVF": lvf.value}

    def get_field(self):
        i = self.info()
        f = np.empty((i["Ny"], i["Nx"]), dtype=np.float64)
        self._ck(self._L.deff2d_domain_get_field(self._h, f.ctypes.data_as(_lib.c_double_p)))
        return f

    def set_field(self, field):
        f = np.ascontiguousarray(field, dtype=np.float64)
        self._ck(self._L.deff2d_domain_set_field(self._h, f.ctypes.data_as(_lib.c_double_p)))

    def get_codes(self):
        i = self.info()
        c = np.empty((i["Ny"], i["Nx"]), dtype=np.uint8)
        self._ck(self._L.deff2d_domain_get_codes(self._h, c.ctypes.data_as(_lib.c_ubyte_p)))
        return c

    def sync(self):
        self._ck(self._L.deff2d_sync(self._h))

    def set_kernel(self, kernel, tblock=0):
        self._ck(self._L.deff2d_set_kernel(self._h, int(kernel), int(tblock)))

    def set_resident(self, mode):
        """0: single domains of up to 256 x 256 cells run cluster-resident (default), 1: never, 2: packed batches too."""
        self._ck(self._L.deff2d_set_resident(self._h, int(mode)))

    def set_floodfill(self, mode):
        """0 automatic, 1 host FloodFill, 2 device FloodFill (same result)."""
        self._ck(self._L.deff2d_set_floodfill(self._h, int(mode)))

    def set_batch_slots(self, max_slots):
        """Cap the images resident at a time in packed batch mode (0 = library default)."""
        self._ck(self._L.deff2d_set_batch_slots(self._h, int(max_slots)))

    @property
    def default_depth(self):
        """Sweeps per HBM pass of the default tiled kernel."""
        return int(self._L.deff2d_default_depth(self._h))

    @property
    def kernel_launches(self):
        return int(self._L.deff2d_kernel_launches(self._h))

    @property
    def stream(self):
        return self._L.deff2d_stream(self._h)

    # ---- multi-GPU slabs ---------------------------------------------------------------
    def nccl_init(self, unique_id, rank, nranks):
        buf = (C.c_ubyte * _lib.NCCL_ID_BYTES).from_buffer_copy(bytes(unique_id))
        self._ck(self._L.deff2d_nccl_init(self._h, buf, int(rank), int(nranks)))

    def slab_peer_export(self):
        buf = (C.c_ubyte * _lib.PEER_HANDLE_BYTES)()
        self._ck(self._L.deff2d_slab_peer_export(self._h, buf))
        return bytes(buf)

    def slab_peer_attach(self, above, below):
        """Handles (bytes) of the ranks above / below, None at the ends.  Collective: ends in a barrier."""
        def arg(h):
            return None if h is None else (C.c_ubyte * _lib.PEER_HANDLE_BYTES).from_buffer_copy(bytes(h))
        self._ck(self._L.deff2d_slab_peer_attach(self._h, arg(above), arg(below)))

    def slab_sweeps(self, n):
        self._ck(self._L.deff2d_slab_sweeps(self._h, int(n)))

    def slab_flux(self):
        d = C.c_double(0)
        self._ck(self._L.deff2d_slab_flux(self._h, C.byref(d)))
        return d.value


def solve_image_slabs(contexts, gray, params, want_field=False):
    """One image over the GPUs of `contexts` (Deff2D objects on different devices) from this one
    process: row slabs, one host thread per device inside the library (csrc/multi.cpp)."""
    gray = np.ascontiguousarray(gray, dtype=np.uint8)
    H, W = gray.shape
    res = Result()
    field = np.empty((H * params.amp_y, W * params.amp_x), dtype=np.float64) if want_field else None
    arr = (C.c_void_p * len(contexts))(*[c._h for c in contexts])
    L = _lib.lib()
    rc = L.deff2d_solve_image_slabs(arr, len(contexts), gray.ctypes.data_as(_lib.c_ubyte_p), W, H, C.byref(params), C.byref(res),
                                    field.ctypes.data_as(_lib.c_double_p) if want_field else None)
    if rc:
        msgs = [L.deff2d_last_error(c._h).decode() for c in contexts]
        raise Deff2DError("deff2d_solve_image_slabs failed (%d): %s" % (rc, "; ".join(m for m in msgs if m)))
    out = res.as_dict()
    if want_field:
        out["field"] = field
    return out


def nccl_unique_id():
    buf = (C.c_ubyte * _lib.NCCL_ID_BYTES)()
    rc = _lib.lib().deff2d_nccl_unique_id(buf)
    if rc:
        raise Deff2DError("deff2d_nccl_unique_id failed (%d)" % rc)
    return bytes(buf)
