"""Feature-pyramid operations — host-side mirror of the reference's `Projects/fp_def.py` (same names,
argument order and return layouts).  Level tables are host integers; everything that touches grid values
runs in libnic.so."""
from collections import defaultdict

import torch

from . import _lib as L
from .models import load4fp, quantize4fp, save4fp


# the names `from ... import *` hands to the reference script
__all__ = ["return_2_power", "return_pyramid_levels", "create_pyramid_mip_levels", "create_pyramid", "create_pyramid_3d", "fp_quantize_clamp", "fp_quantize", "fp_all_quantize", "fp_savable", "fp_load", "fp_savable_packed", "fp_load_packed", "fp_freeze"]

def return_2_power(base_size):
    """fp_def.py:8-15."""
    count = 0
    x = base_size
    while x != 1:
        x = x // 2
        count += 1
    return count


def return_pyramid_levels(base_size):
    """fp_def.py:18-21."""
    return (return_2_power(base_size) + 1) // 2


def create_pyramid_mip_levels(image_size, base_size):
    """fp_def.py:24-34 — mip level -> pyramid level."""
    count = return_2_power(image_size)
    table = defaultdict(int)
    levels = return_pyramid_levels(base_size)
    for i in range(count + 1):
        table[i] = min(max((i // 2) - 1, 0), levels - 1)
    return table


def _create(base_size, channels, num_bits, device, dtype, no_mip, dim):
    if dtype != torch.float32:
        raise TypeError("grids are float32 (the reference's MLP_NUM_DTYPE=16 path diverges within two steps)")
    levels = 1 if no_mip else return_pyramid_levels(base_size)
    q_min = -(pow(2, num_bits) - 1) / pow(2, num_bits + 1)
    q_max = 1 / 2
    pyramid = []
    for i in range(levels * 2):
        size = base_size // (2 ** i)
        shape = (channels,) + (size + 1,) * dim
        pyramid.append(((q_max - q_min) * torch.rand(shape, device=device, dtype=dtype) + q_min).requires_grad_(True))
    return pyramid, levels


def create_pyramid(base_size, channels, num_bits, device, dtype, no_mip=False):
    """fp_def.py:37-56 — 2*levels learnable grids `[C, s+1, s+1]`, U[q_min, 1/2]."""
    return _create(base_size, channels, num_bits, device, dtype, no_mip, 2)


def create_pyramid_3d(base_size, channels, num_bits, device, dtype, no_mip=False):
    """fp_def.py:59-78."""
    return _create(base_size, channels, num_bits, device, dtype, no_mip, 3)


def fp_quantize_clamp(fp, fl, num_bits):
    """fp_def.py:227-232 — clamp the two active grids in place."""
    q_min = -(pow(2, num_bits) - 1) / pow(2, num_bits + 1)
    lib = L.load_library()
    with torch.no_grad():
        for g in (fp[fl * 2], fp[fl * 2 + 1]):
            if not g.is_contiguous() or g.dtype != torch.float32:
                raise TypeError("grids must be contiguous float32")
            h = L.handle(g.device)
            L.check(h, lib.nic_clamp(h, L.ptr(g), g.numel(), q_min, 0.5, L.stream_ptr(g.device)))


def fp_quantize(fp, fl, num_bits):
    """fp_def.py:235-239."""
    with torch.no_grad():
        fp[fl * 2] = quantize4fp(fp[fl * 2], num_bits)
        fp[fl * 2 + 1] = quantize4fp(fp[fl * 2 + 1], num_bits)


def fp_all_quantize(fp, num_bits):
    """fp_def.py:242-247."""
    return [quantize4fp(g, num_bits) for g in fp]


def fp_savable(fp, num_bits, dtype=torch.uint8):
    """fp_def.py:250-255 — uint8 codes, bit-exact with models.save4fp."""
    return [save4fp(g, num_bits, dtype) for g in fp]


def fp_load(compressed_fp, num_bits, dtype=torch.float32):
    """fp_def.py:258-263 (float result; see models.load4fp about the reference's dtype bug)."""
    return [load4fp(g, num_bits, dtype) for g in compressed_fp]


def fp_savable_packed(fp, num_bits):
    """Extension of fp_savable: the same bit-exact codes, stored 8/num_bits per byte.
    Returns a list of (packed uint8 tensor [ceil(n*bits/8)], shape) pairs; num_bits in {1, 2, 4, 8}."""
    lib = L.load_library()
    out = []
    for g in fp:
        codes = save4fp(g, num_bits).reshape(-1)
        n = codes.numel()
        packed = torch.empty(((n * num_bits + 7) // 8,), dtype=torch.uint8, device=codes.device)
        h = L.handle(codes.device)
        L.check(h, lib.nic_pack_codes(h, L.ptr(codes), L.ptr(packed), n, num_bits, L.stream_ptr(codes.device)))
        out.append((packed, tuple(g.shape)))
    return out


def fp_load_packed(packed_fp, num_bits, dtype=torch.float32):
    """Inverse of fp_savable_packed: float32 grids `(code - 2^(b-1) + 1) / (2^b - 1)` (models.load4fp)."""
    lib = L.load_library()
    out = []
    for packed, shape in packed_fp:
        n = 1
        for d in shape:
            n *= d
        codes = torch.empty((n,), dtype=torch.uint8, device=packed.device)
        h = L.handle(packed.device)
        L.check(h, lib.nic_unpack_codes(h, L.ptr(packed.contiguous()), L.ptr(codes), n, num_bits, L.stream_ptr(packed.device)))
        out.append(load4fp(codes.reshape(shape), num_bits, dtype))
    return out


def fp_freeze(fp):
    """fp_def.py:266-268."""
    for g in fp:
        g.requires_grad = False
