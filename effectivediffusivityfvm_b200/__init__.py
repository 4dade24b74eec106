"""effectivediffusivityfvm_b200 -- B200-native (sm_100a) effective-diffusivity solve.

The hot path of adama-wzr/EffectiveDiffusivityFVM (image -> phases -> FVM coefficients ->
damped-Jacobi sweeps -> boundary-flux Deff) as hand-written CUDA behind a C ABI
(include/deff2d.h, libdeff2d.so), plus this thin host-side mirror of the reference's driver
interface.  There is no CPU fallback: the API raises if the CUDA library is not built or no
device is present.
"""
from ._lib import (LIB_PATH, MODE_2PH_BATCH, MODE_2PH_SINGLE, MODE_3PH, Input, Params, Result)
from .api import (Deff2D, Deff2DError, batch_plan, batch_tile_list, build_tables, default_params, floodfill, load_image,
                  nccl_unique_id, read_input_file, solve_image_slabs, tile_geometry)

__all__ = ["Deff2D", "Deff2DError", "Params", "Result", "Input", "default_params", "read_input_file",
           "build_tables", "floodfill", "tile_geometry", "batch_plan", "batch_tile_list", "load_image", "nccl_unique_id", "solve_image_slabs", "MODE_2PH_SINGLE", "MODE_2PH_BATCH",
           "MODE_3PH", "LIB_PATH"]
