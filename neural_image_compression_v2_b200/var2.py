"""Configuration — the same UPPER_CASE names, defaults, `KEY=VALUE` override syntax and derived values as the
reference's `Projects/var2.py` (lines cited per item), without its `exec`.  The host mirror in
`image_compression.py` reads these names at call time, exactly as the reference's functions read their globals."""
import os
import sys

import torch

from .utils import bits2dtype_torch

# var2.py:6-36 — the typed override table
over_write_variable_dict = {
    "FP_BITS": "int", "NUM_EPOCHS": "int", "IMAGE_SIZE": "int", "IMAGE_3D_SIZE": "int", "MAX_MIP_LEVEL": "int",
    "FEATURE_PYRAMID_CHANNELS": "int", "PE_CHANNELS": "int", "IMAGE_PATH": "str", "PROJECT_NAME": "str",
    "IMAGE_DTYPE": "str", "COMPRESSION_METHOD": "int", "MLP_NUM_DTYPE": "int", "UNIFORM_DISTRIBUTION_RATE": "float",
    "IMAGE_DIMENSION": "int", "IMAGE_BITS": "int", "OUTPUT_BITS": "int", "HIDDEN_LAYER_CHANNELS": "int",
    "CROP_MIP_LEVEL": "int", "NUM_CROPS": "int", "INTERVAL_PRINT": "int", "INTERVAL_SAVE_MODEL": "int",
    "TF_NO_MIP": "bool", "TF_USE_TRI_PE": "bool", "TF_TRAIN_MODEL": "bool", "TF_SHOW_RESULT": "bool",
    "TF_PRINT_LOG": "bool", "TF_PRINT_PSNR": "bool", "TF_WRITE_TIME": "bool", "TF_WRITE_PSNR": "bool",
    # additions of this implementation (not in the reference)
    "DECODE_PRECISION": "str", "OUTPUT_CHANNELS": "int",
}

_DEFAULTS = dict(                      # var2.py:38-87
    IMAGE_PATH="data/sancho_512.png", PROJECT_NAME="image_compression", IMAGE_DTYPE="image", COMPRESSION_METHOD=1,
    MLP_NUM_DTYPE=32, NUM_EPOCHS=1000, UNIFORM_DISTRIBUTION_RATE=0.05, IMAGE_3D_SIZE=64, IMAGE_SIZE=512,
    IMAGE_DIMENSION=2, MAX_MIP_LEVEL=9, IMAGE_BITS=8, OUTPUT_BITS=8, FEATURE_PYRAMID_CHANNELS=12, PE_CHANNELS=6,
    FP_BITS=8, HIDDEN_LAYER_CHANNELS=64, CROP_MIP_LEVEL=8, NUM_CROPS=8, INTERVAL_PRINT=100, INTERVAL_SAVE_MODEL=100000,
    TF_NO_MIP=True, TF_USE_TRI_PE=True, TF_TRAIN_MODEL=True, TF_SHOW_RESULT=False, TF_PRINT_LOG=True,
    TF_PRINT_PSNR=True, TF_WRITE_TIME=True, TF_WRITE_PSNR=True,
    # "f32" = reference-exact CUDA-core path; "f16"/"bf16" = tcgen05 tensor-core path
    DECODE_PRECISION="f32", OUTPUT_CHANNELS=3,
)


def judge_torf(arg, error_massage=""):
    """utils.py:13-20."""
    value = arg.split("=")[1].lower()
    if value in ["true", "1"]:
        return True
    if value in ["false", "0"]:
        return False
    raise ValueError(f"{error_massage} must be a boolean (True/False or 1/0)")


def judge_value(arg, dtype, error_massage=""):
    """utils.py:23-31 (returns the value itself; the reference returns source text for `exec`)."""
    raw = arg.split("=", 1)[1]
    if dtype == "int":
        return int(raw)
    if dtype == "float":
        return float(raw)
    if dtype == "bool":
        return judge_torf(arg, error_massage)
    return raw


def dtype_from_ext(ext):
    """utils.py:330-336."""
    ext = ext.lower()
    if ext in ("npy", "npz"):
        return "ndarray"
    if ext in ("avi", "mp4"):
        return "movie"
    if ext in ("png", "jpg", "jpeg"):
        return "image"


def _derive():
    """var2.py:100-125."""
    g = globals()
    g["DEVICE"] = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    g["BASENAME"] = os.path.basename(IMAGE_PATH)
    g["IMAGE_EXT"] = os.path.splitext(IMAGE_PATH)[1][1:]
    g["IMAGE_DTYPE"] = dtype_from_ext(g["IMAGE_EXT"]) or IMAGE_DTYPE
    g["FEATURE_PYRAMID_SIZE"] = IMAGE_SIZE // 4
    g["FP_DIMENSION"] = 2 if COMPRESSION_METHOD == 2 else IMAGE_DIMENSION
    if TF_NO_MIP:
        g["MAX_MIP_LEVEL"] = 0
    fpd = g["FP_DIMENSION"]
    corners = pow(2, 2) if COMPRESSION_METHOD == 4 else pow(2, fpd)
    g["DECODER_INPUT_CHANNELS"] = FEATURE_PYRAMID_CHANNELS * (corners + 1) + PE_CHANNELS * fpd + 1
    g["CROP_SIZE"] = pow(2, CROP_MIP_LEVEL)
    g["MLP_DTYPE"] = bits2dtype_torch(MLP_NUM_DTYPE, "float")
    g["SAVE_NAME"] = (f"{PROJECT_NAME}_{g['DEVICE']}_{g['BASENAME']}_{MLP_NUM_DTYPE}_{TF_NO_MIP}_{TF_USE_TRI_PE}_"
                      f"{COMPRESSION_METHOD}_{NUM_EPOCHS}_{FP_BITS}")


def update(*args, **overrides):
    """Apply `KEY=VALUE` strings (var2.py:90-95) and/or keyword overrides, then recompute the derived values.
    Keys start from the reference defaults again, as a fresh `import var2` would."""
    g = globals()
    g.update(_DEFAULTS)
    for arg in args:
        for var, typ in over_write_variable_dict.items():
            if arg.startswith(var + "="):
                g[var] = judge_value(arg, typ, var)
    for k, v in overrides.items():
        if k not in over_write_variable_dict:
            raise KeyError(f"unknown configuration key {k}")
        g[k] = v
    _derive()


update(*[a for a in sys.argv[1:] if "=" in a and a.split("=")[0] in over_write_variable_dict])
