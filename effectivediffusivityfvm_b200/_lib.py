"""ctypes binding of libdeff2d.so (include/deff2d.h).  There is no fallback: if the CUDA
library has not been built, importing the solver API raises."""
import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DEFF2D_LIB") or os.path.join(PKG_DIR, "libdeff2d.so")   # DEFF2D_LIB: A/B a development build

MAX_STAGES = 16
NCCL_ID_BYTES = 128
PEER_HANDLE_BYTES = 320

MODE_2PH_SINGLE, MODE_2PH_BATCH, MODE_3PH = 0, 1, 2

c_double_p = C.POINTER(C.c_double)
c_ubyte_p = C.POINTER(C.c_ubyte)


class Params(C.Structure):
    """deff2d_params: the numeric part of the reference's `options` (Deff2D.cuh:18-37)."""
    _fields_ = [("Ds", C.c_double), ("Df", C.c_double), ("Dg", C.c_double),
                ("amp_x", C.c_int), ("amp_y", C.c_int),
                ("CL", C.c_double), ("CR", C.c_double),
                ("max_iter", C.c_int64), ("tol", C.c_double),
                ("mode", C.c_int), ("check_every", C.c_int),
                ("omega", C.c_double), ("solver", C.c_int),
                ("verbose", C.c_int), ("residual_tol", C.c_double), ("strict_reference", C.c_int)]


class Result(C.Structure):
    """deff2d_result: `simulationInfo` (Deff2D.cuh:39-52) plus per-stage bookkeeping."""
    _fields_ = [("porosity", C.c_double), ("SVF", C.c_double), ("LVF", C.c_double),
                ("deff", C.c_double), ("deff_raw", C.c_double), ("conv", C.c_double),
                ("pathflag", C.c_int), ("nstages", C.c_int),
                ("iters", C.c_int64 * MAX_STAGES),
                ("stage_deff_raw", C.c_double * MAX_STAGES),
                ("stage_D", C.c_double * MAX_STAGES),
                ("total_iters", C.c_int64), ("n_cells", C.c_int64),
                ("solve_ms", C.c_double), ("total_ms", C.c_double), ("last_df", C.c_double)]

    def as_dict(self):
        n = min(self.nstages, MAX_STAGES)
        return {"porosity": self.porosity, "SVF": self.SVF, "LVF": self.LVF, "deff": self.deff,
                "deff_raw": self.deff_raw, "conv": self.conv, "pathflag": self.pathflag,
                "nstages": self.nstages, "iters": list(self.iters[:n]),
                "stage_deff_raw": list(self.stage_deff_raw[:n]), "stage_D": list(self.stage_D[:n]),
                "total_iters": self.total_iters, "n_cells": self.n_cells, "solve_ms": self.solve_ms,
                "total_ms": self.total_ms, "last_df": self.last_df}


class Input(C.Structure):
    """deff2d_input: a parsed input.txt (Deff2D.cuh:234-324)."""
    _fields_ = [("p", Params), ("nphase", C.c_int), ("batch", C.c_int), ("num_images", C.c_int),
                ("print_cmap", C.c_int), ("input_name", C.c_char * 1000),
                ("output_name", C.c_char * 1000), ("cmap_name", C.c_char * 1000), ("devices", C.c_int), ("field_npy", C.c_int)]


BATCH_FETCH_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_ubyte), C.c_int)
BATCH_DONE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(Result), C.POINTER(C.c_double))

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libdeff2d.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C effectivediffusivityfvm_b200/csrc`. There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
    sig = {
        "deff2d_version": (i32, []),
        "deff2d_create": (i32, [C.POINTER(vp), i32]),
        "deff2d_destroy": (None, [vp]),
        "deff2d_last_error": (C.c_char_p, [vp]),
        "deff2d_default_params": (None, [C.POINTER(Params)]),
        "deff2d_solve_image": (i32, [vp, c_ubyte_p, i32, i32, C.POINTER(Params), C.POINTER(Result), c_double_p]),
        "deff2d_solve_batch": (i32, [vp, c_ubyte_p, i32, i32, i32, C.POINTER(Params), C.POINTER(Result), c_double_p]),
        "deff2d_solve_batch_stream": (i32, [vp, i32, i32, i32, C.POINTER(Params), BATCH_FETCH_FN, BATCH_DONE_FN, vp, i32, C.POINTER(i32)]),
        "deff2d_batch_supported": (i32, [C.POINTER(Params), i32, i32]),
        "deff2d_solve_image_slabs": (i32, [C.POINTER(vp), i32, c_ubyte_p, i32, i32, C.POINTER(Params), C.POINTER(Result), c_double_p]),
        "deff2d_domain_load": (i32, [vp, c_ubyte_p, i32, i32, i32, C.POINTER(Params)]),
        "deff2d_domain_load_slab": (i32, [vp, c_ubyte_p, i32, i32, i32, C.POINTER(Params), i64, i64, i32, c_ubyte_p]),
        "deff2d_domain_set_D": (i32, [vp, dbl, dbl, dbl]),
        "deff2d_domain_sweeps": (i32, [vp, i64]),
        "deff2d_domain_sweeps_timed": (i32, [vp, i64, C.POINTER(C.c_float)]),
        "deff2d_domain_flux": (i32, [vp, c_double_p, c_double_p]),
        "deff2d_domain_residual": (i32, [vp, c_double_p]),
        "deff2d_domain_solve": (i32, [vp, dbl, i64, C.POINTER(i64), c_double_p, c_double_p, c_double_p, i32, C.POINTER(i32)]),
        "deff2d_domain_get_field": (i32, [vp, c_double_p]),
        "deff2d_domain_set_field": (i32, [vp, c_double_p]),
        "deff2d_domain_get_codes": (i32, [vp, c_ubyte_p]),
        "deff2d_domain_info": (i32, [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i32), c_double_p, c_double_p, c_double_p]),
        "deff2d_sync": (i32, [vp]),
        "deff2d_set_kernel": (i32, [vp, i32, i32]),
        "deff2d_set_resident": (i32, [vp, i32]),
        "deff2d_default_depth": (i32, [vp]),
        "deff2d_set_batch_slots": (i32, [vp, i32]),
        "deff2d_set_floodfill": (i32, [vp, i32]),
        "deff2d_set_graphs": (i32, [vp, i32]),
        "deff2d_get_graphs": (i32, [vp]),
        "deff2d_kernel_launches": (i64, [vp]),
        "deff2d_stream": (vp, [vp]),
        "deff2d_domain_buffers": (i32, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(i64), C.POINTER(i64)]),
        "deff2d_nccl_unique_id": (i32, [c_ubyte_p]),
        "deff2d_nccl_init": (i32, [vp, c_ubyte_p, i32, i32]),
        "deff2d_slab_sweeps": (i32, [vp, i64]),
        "deff2d_slab_flux": (i32, [vp, c_double_p]),
        "deff2d_slab_abort": (i32, [vp]),
        "deff2d_slab_peer_export": (i32, [vp, c_ubyte_p]),
        "deff2d_slab_peer_attach": (i32, [vp, c_ubyte_p, c_ubyte_p]),
        "deff2d_slab_peer_detach": (i32, [vp]),
        "deff2d_domain_load_slab_global": (i32, [vp, c_ubyte_p, i32, i32, i32, C.POINTER(Params), i64, i64, i32]),
        "deff2d_accumulate_fraction": (dbl, [i64, i64]),
        "deff2d_build_tables": (i32, [dbl, dbl, dbl, i64, i64, dbl, dbl, dbl, c_double_p, c_ubyte_p]),
        "deff2d_compact_table": (i32, [c_double_p, i32, c_double_p, C.POINTER(C.c_uint16)]),
        "deff2d_split_table": (i32, [c_double_p, i32, C.POINTER(C.c_uint32)]),
        "deff2d_floodfill": (i32, [c_ubyte_p, i64, i64]),
        "deff2d_tile_geometry": (i32, [i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]),
        "deff2d_batch_plan": (i32, [i64, i64, i32, i32, C.POINTER(i32), C.POINTER(i32)]),
        "deff2d_batch_tile_list": (i32, [i64, i64, i32, i32, C.POINTER(i32), i32, i32, C.POINTER(C.c_uint32), i32]),
        "deff2d_read_input_file": (i32, [C.c_char_p, C.POINTER(Input)]),
        "deff2d_write_csv_single": (i32, [C.POINTER(Input), C.POINTER(Result)]),
        "deff2d_write_csv_batch": (i32, [C.POINTER(Input), C.POINTER(Result), i32]),
        "deff2d_append_csv_batch_row": (i32, [C.POINTER(Input), i32, C.POINTER(Result)]),
        "deff2d_write_cmap": (i32, [C.c_char_p, c_double_p, i64, i64]),
        "deff2d_write_field_npy": (i32, [C.c_char_p, c_double_p, i64, i64]),
        "deff2d_run_input_file": (i32, [vp, C.c_char_p]),
        "deff2d_load_image": (i32, [C.c_char_p, C.POINTER(c_ubyte_p), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]),
        "deff2d_free": (None, [vp]),
    }
    for name, (res, args) in sig.items():
        if os.environ.get("DEFF2D_LIB") and not hasattr(L, name):
            continue                   # development A/B against an older build
        fn = getattr(L, name)          # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    L._declared = sorted(sig)
    _lib = L
    return L
