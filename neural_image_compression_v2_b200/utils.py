"""Positional encodings, PSNR and dtype maps — host-side mirror of the hot-path functions of the reference's
`Projects/utils.py` (same names and argument meaning), computed by libnic.so."""
import ctypes as C
import glob
import math
import os
import re

import numpy as np
import torch

from . import _lib as L


# the names `from ... import *` hands to the reference script
# (the host file helpers at the end of this module are NOT exported by `*`: inside the reference script its own stay in place)
__all__ = ["triangular_positional_encoding", "positional_encoding", "tri", "calculate_psnr", "bits2dtype_torch", "bits2dtype_np"]

def _pe(coord, num_channels, device, dtype, kind):
    if isinstance(coord, (tuple, list)):
        coord = torch.stack([c.reshape(-1) for c in coord])
    if dtype not in (torch.float32, None):
        raise TypeError("positional encodings are float32 on this path")
    c = coord.detach().to(device=device, dtype=torch.float32).contiguous()
    if c.dim() != 2 or not 1 <= c.shape[0] <= 3:
        raise ValueError(f"coord must be [D, N] with D in 1..3, got {tuple(c.shape)}")
    d, n = c.shape
    out = torch.empty((num_channels * d, n), dtype=torch.float32, device=c.device)
    div = (C.c_float * 8)(*[float(v) for v in L.sin_div_term(num_channels)]) if num_channels else (C.c_float * 8)()
    h = L.handle(c.device)
    L.check(h, L.load_library().nic_positional_encoding(h, L.ptr(c), d, n, num_channels, kind,
                                                        C.cast(div, C.c_void_p), L.ptr(out), L.stream_ptr(c.device)))
    return out


def triangular_positional_encoding(coord, num_channels, device, dtype=torch.float32):
    """utils.py:211-223 — coord [D, N] -> [num_channels*D, N]."""
    return _pe(coord, num_channels, device, dtype, L.PE_TRIANGULAR)


def positional_encoding(coord, num_channels, device, dtype=torch.float32):
    """utils.py:198-208 — coord = tuple of D vectors -> [num_channels*D, N] (sin/cos interleaved)."""
    return _pe(coord, num_channels, device, dtype, L.PE_SINUSOIDAL)


def tri(x, offset=0.5):
    """utils.py:226-227 (elementwise; one triangular row of a single-octave encoding)."""
    return 2 * torch.abs((x - offset) % 2 - 1) - 1


def calculate_psnr(original, reconstructed, num_bits=8):
    """utils.py:117-130 — 10 log10(2^bits * 2^bits / mse); peak is 2^bits, not 2^bits - 1.
    uint8 CUDA tensors are reduced on the device (integer-exact SSE); float inputs use the same formula."""
    peak = pow(2, num_bits)
    if torch.is_tensor(original) and original.dtype == torch.uint8 and original.is_cuda:
        a, b = original.contiguous(), reconstructed.contiguous()
        sse = torch.zeros(1, dtype=torch.float64, device=a.device)
        h = L.handle(a.device)
        L.check(h, L.load_library().nic_sse_u8(h, L.ptr(a), L.ptr(b), a.numel(), L.ptr(sse), L.stream_ptr(a.device)))
        mse = float(sse.item()) / max(a.numel(), 1)
    elif isinstance(original, np.ndarray):
        mse = float(np.mean((original - reconstructed) ** 2))
    else:
        mse = float(torch.mean((original - reconstructed) ** 2))
    if mse == 0:
        return float("inf")
    return 10 * math.log10(peak * peak / mse)


def bits2dtype_torch(num_bits, dtype="float"):
    """utils.py:301-313."""
    if num_bits <= 8:
        return torch.uint8
    if num_bits == 16:
        return {"int": torch.int16, "uint": torch.uint16, "float": torch.float16}[dtype]
    if num_bits == 32:
        return torch.float32
    if num_bits == 64:
        return torch.float64


def bits2dtype_np(num_bits, dtype="float"):
    """utils.py:316-328."""
    if num_bits <= 8:
        return np.uint8
    if num_bits == 16:
        return {"int": np.int16, "uint": np.uint16, "float": np.float16}[dtype]
    if num_bits == 32:
        return np.float32
    if num_bits == 64:
        return np.float64


# ---------------------------------------------------------------------------------------------------------------------
# Host-side file helpers of the 3-D / movie / LUT workflows.  They are not on the device path:
# `readClip` feeds `flatten_movie_to_atlas` / the 3-D trainers with a uint8 array, `timelaps` and `save_result_to_csv`
# take the decoded uint8 volume.  Same names, arguments and file formats as the reference; OpenCV is imported lazily.
def make_filename_by_seq(dirname, filename, seq_digit=3):
    """utils.py:37-62 — `dirname/<stem>_<n+1><ext>` where n is the highest sequence number already present (-1 if none);
    creates `dirname` if needed."""
    os.makedirs(dirname, exist_ok=True)
    stem, ext = os.path.splitext(filename)
    taken = [-1]
    for f in glob.glob(os.path.join(dirname, f"{stem}_[0-9]*{ext}")):
        m = re.match(f"{re.escape(stem)}_([0-9]*){re.escape(ext)}", os.path.basename(f))
        if m and m.group(1):
            taken.append(int(m.group(1)))
    return f"{dirname}/{stem}_{max(taken) + 1:0{seq_digit}}{ext}"


def readClip(filepass):
    """utils.py:67-81 — every frame of a video file as one uint8 array [frames, height, width, 3] (BGR, as OpenCV decodes)."""
    import cv2
    cap = cv2.VideoCapture(filepass)
    try:
        if not cap.isOpened():
            raise FileNotFoundError(f"cannot open video {filepass!r}")
        frames = []
        while True:
            ok, frame = cap.read()
            if not ok:
                break
            frames.append(frame)
    finally:
        cap.release()
    if not frames:
        raise ValueError(f"{filepass!r} holds no decodable frame")
    return np.stack(frames, axis=0)


def timelaps(movie, saved_name, all_frame=64, width=64, height=64, frame_rate=32):
    """utils.py:86-95 — writes frames movie[0 .. all_frame) (uint8 [height, width, 3]) as an mp4v-coded video."""
    import cv2
    video = cv2.VideoWriter(saved_name, cv2.VideoWriter_fourcc("m", "p", "4", "v"), frame_rate, (width, height))
    try:
        if not video.isOpened():
            raise OSError(f"cannot open a video writer for {saved_name!r}")
        for i in range(all_frame):
            video.write(np.ascontiguousarray(movie[i]))
    finally:
        video.release()


def save_result_to_csv(result, filename):
    """utils.py:98-113 — a [S, S, S, 3] LUT as text: one line per (first, second) index pair holding the 3 S values of
    that row, each followed by a comma."""
    arr = result.detach().cpu().numpy() if torch.is_tensor(result) else np.asarray(result)
    size = arr.shape[0]
    with open(filename, mode="w") as f:
        for a in range(size):
            for b in range(size):
                f.write("".join(f"{v.item()}," for v in arr[a, b, :size, :3].reshape(-1)))
                f.write("\n")
