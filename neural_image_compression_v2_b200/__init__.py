"""B200-native (sm_100a) implementation of the per-texel decode / training-step hot path of
21K1113/Neural_Image_Compression_V2.

Layout (only what the path needs):
  csrc/               CUDA kernels + the C ABI (include/nic.h) -> libnic.so, built by `build.py`
  _lib.py             ctypes binding of the C ABI
  var2.py             the reference's configuration names (Projects/var2.py)
  models.py, utils.py, fp_def.py, image_compression.py
                      host-side mirrors of the reference modules of the same name (hot-path functions only)
  parallel.py         tile/frame sharding of decode across ranks (no collective) and DP training helpers

Importing the package does not need a GPU; calling anything that computes does (no CPU fallback).
"""
from . import _lib, fp_def, models, utils, var2  # noqa: F401
from . import image_compression  # noqa: F401
from . import parallel  # noqa: F401
from ._lib import NicError, launch_count, load_library  # noqa: F401

__all__ = ["_lib", "fp_def", "models", "utils", "var2", "image_compression", "parallel", "NicError", "launch_count",
           "load_library"]
