"""Quantiser library — host-side mirror of the reference's `Projects/models.py` (same names, arguments and
results), computed by libnic.so kernels.  CUDA tensors only; there is no CPU fallback."""
import numpy as np
import torch

from . import _lib as L


# the names `from ... import *` hands to the reference script
__all__ = ["quantize4fp", "quantize_torch", "quantize", "save4fp", "load4fp", "quantize_clamp", "quantize_to_bit", "output_to_u8"]

def _run_f2f(fn_name, tensor, bits):
    t = tensor.detach()
    if t.dtype != torch.float32:
        raise TypeError(f"{fn_name}: float32 tensor expected, got {t.dtype}")
    t = t.contiguous()
    out = torch.empty_like(t)
    h = L.handle(t.device)
    L.check(h, getattr(L.load_library(), fn_name)(h, L.ptr(t), L.ptr(out), t.numel(), bits, L.stream_ptr(t.device)))
    return out


def quantize4fp(tensor, num_bits):
    """models.py:55-57 — floor(t*(2^b-1)+.5)/(2^b-1)."""
    return _run_f2f("nic_quantize4fp", tensor, num_bits)


def quantize_torch(tensor, num_bits):
    """models.py:17-19 (same arithmetic as quantize4fp)."""
    return _run_f2f("nic_quantize4fp", tensor, num_bits)


def quantize(array, bit):
    """models.py:29-35 — tensor inputs only on this path (numpy inputs belong to the host I/O side)."""
    if isinstance(array, np.ndarray):
        raise TypeError("numpy input: the B200 path handles device tensors; use the u8 output of decode instead")
    return _run_f2f("nic_quantize4fp", array, bit)


def save4fp(tensor, num_bits, dtype=torch.uint8):
    """models.py:61-64 — code = floor(t*(2^b-1)+.5) + 2^(b-1) - 1 as uint8, one code per byte."""
    if dtype != torch.uint8:
        raise TypeError("save4fp stores uint8 codes (bits2dtype_torch(FP_BITS) for FP_BITS <= 8)")
    t = tensor.detach()
    if t.dtype != torch.float32:
        raise TypeError(f"save4fp: float32 tensor expected, got {t.dtype}")
    t = t.contiguous()
    out = torch.empty(t.shape, dtype=torch.uint8, device=t.device)
    h = L.handle(t.device)
    L.check(h, L.load_library().nic_quantize_pack(h, L.ptr(t), L.ptr(out), t.numel(), num_bits, L.stream_ptr(t.device)))
    return out


def load4fp(tensor, num_bits, dtype=torch.float32):
    """models.py:68-71 with the intended float result.  The reference's call site passes a uint8 dtype
    (image_compression.py:396), which makes the subtraction wrap; that bug is not reproduced: a uint8
    `dtype` is treated as float32."""
    if dtype in (torch.uint8, None):
        dtype = torch.float32
    if dtype != torch.float32:
        raise TypeError("load4fp returns float32 grids")
    t = tensor.detach()
    if t.dtype != torch.uint8:
        raise TypeError(f"load4fp: uint8 codes expected, got {t.dtype}")
    t = t.contiguous()
    out = torch.empty(t.shape, dtype=torch.float32, device=t.device)
    h = L.handle(t.device)
    L.check(h, L.load_library().nic_unpack(h, L.ptr(t), L.ptr(out), t.numel(), num_bits, L.stream_ptr(t.device)))
    return out


def quantize_clamp(tensor, num_bits=8):
    """models.py:48-51 — clamp to [-(2^b-1)/2^(b+1), 1/2] (out of place, like torch.clamp)."""
    q_min = -(pow(2, num_bits) - 1) / pow(2, num_bits + 1)
    out = tensor.detach().contiguous().clone()
    h = L.handle(out.device)
    L.check(h, L.load_library().nic_clamp(h, L.ptr(out), out.numel(), q_min, 0.5, L.stream_ptr(out.device)))
    return out


def quantize_to_bit(array, num_bits=8):
    """models.py:38-40 — float image in [0,1] -> float values 0..2^b-1 (round half up).  The script calls it on a HOST numpy
    image (image_compression.py:406): such an input makes the round trip through the device (there is no host arithmetic
    in this package) and comes back as a float32 numpy array, like the reference's."""
    if isinstance(array, np.ndarray):
        dev = torch.device("cuda", torch.cuda.current_device())
        t = torch.as_tensor(np.ascontiguousarray(array, dtype=np.float32), device=dev)
        return output_to_u8(t, num_bits).to(torch.float32).cpu().numpy()
    return output_to_u8(array, num_bits).to(torch.float32)


def output_to_u8(tensor, num_bits=8):
    """quantize_to_bit(...).astype(uint8) of image_compression.py:406-407, on the device."""
    t = tensor.detach()
    if t.dtype != torch.float32:
        raise TypeError(f"output_to_u8: float32 tensor expected, got {t.dtype}")
    t = t.contiguous()
    out = torch.empty(t.shape, dtype=torch.uint8, device=t.device)
    h = L.handle(t.device)
    L.check(h, L.load_library().nic_output_to_u8(h, L.ptr(t), L.ptr(out), t.numel(), num_bits, L.stream_ptr(t.device)))
    return out
