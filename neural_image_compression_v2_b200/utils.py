"""Positional encodings, PSNR and dtype maps — host-side mirror of the hot-path functions of the reference's
`Projects/utils.py` (same names and argument meaning), computed by libnic.so."""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib as L


# the names `from ... import *` hands to the reference script
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
