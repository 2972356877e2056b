"""CPU oracle, torch edition: the reference's decode / training-step path restated with the SAME torch CPU ops the
reference itself issues (advanced-index gathers on channel-major grids, elementwise weight chains, `torch.cat`,
a transposed view fed to Linear/GELU/Linear/GELU/Linear/Sigmoid, MSELoss + autograd + torch.optim.Adam).

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  It exists because `/root/reference` cannot travel to the GPU box:
`bench.py`'s `cpu_baseline` and `--impl reference` legs time THIS module on the box's host cores (all intra-op
threads, like the reference on a CPU `DEVICE`), and `tests/test_oracle_golden.py` holds it to the same golden
fixtures (outputs of the unmodified reference) as the numpy oracle.  Parity status: PINNED (see nic_oracle.py).

Citations are `file:line` relative to `/root/reference/Projects/`.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import nic_oracle as O

_CORNERS_2D = ((0, 0), (1, 0), (0, 1), (1, 1))                                   # (dy, dx)  fp_def.py:81-86
_CORNERS_3D = ((0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 0), (0, 0, 1), (1, 0, 1), (0, 1, 1), (1, 1, 1))   # (dz,dy,dx) :96-103
_CORNERS_3D_V2 = ((0, 0, 0), (1, 1, 0), (1, 0, 1), (0, 1, 1))                      # :108-111
_W3D_AS_CODED = ((0, 0, 0), (0, 0, 1), (0, 1, 0), (1, 0, 0), (1, 1, 0), (1, 0, 1), (0, 1, 1), (1, 1, 1))   # :176-183


def tri(x, offset):
    """utils.py:226-227."""
    return 2 * torch.abs((x - offset) % 2 - 1) - 1


def triangular_positional_encoding(coord, num_channels):
    """utils.py:211-223 — coord [D, N] -> [num_channels*D, N]."""
    d, n = coord.shape
    pe = torch.zeros(d, num_channels, n, dtype=coord.dtype)
    for octave in range(num_channels // 2):
        div = pow(2, octave)
        for i, offset in enumerate((0.5, 0.0)):
            if octave + i == 0:
                continue
            pe[:, num_channels - (octave * 2 + i + 1), :] = tri(coord / div, offset)
    return pe.reshape(d * num_channels, n)


def positional_encoding(coords, num_channels):
    """utils.py:198-208 — tuple of D vectors -> [num_channels*D, N]."""
    div_term = torch.exp(torch.arange(0, num_channels, 2, dtype=torch.float32) * -(math.log(10000.0) / num_channels))
    rows = []
    for c in coords:
        pe = torch.zeros(num_channels, c.shape[0], dtype=torch.float32)
        pe[0::2, :] = torch.sin(c.unsqueeze(0) * div_term.unsqueeze(1))
        pe[1::2, :] = torch.cos(c.unsqueeze(0) * div_term.unsqueeze(1))
        rows.append(pe)
    return torch.cat(rows, dim=0)


def _axis(origin, size, step):
    """fp_def.py:116-123 for one axis."""
    u0 = (torch.arange(size) + int(origin)) * step
    i0 = torch.floor(u0).to(torch.int32)
    u1 = u0 / 2
    i1 = torch.floor(u1).to(torch.int32)
    return u0, i0, u1, i1


def decoder_input_one(g0, g1, origin, size, step, mip_level, method, pe_channels=6, use_tri_pe=True):
    """image_compression.py:90-96 + fp_def.py:115-223: one crop / block -> rows [Cin, S^D] (channel-major, as the
    reference builds them before its final `.T`)."""
    dim = 2 if method == 1 else 3
    ax = [_axis(origin[a], size, step) for a in range(dim)]
    mesh = lambda vs: [m.reshape(-1) for m in torch.meshgrid(*vs, indexing="ij")]
    i0, i1, u1 = mesh([a[1] for a in ax]), mesh([a[3] for a in ax]), mesh([a[2] for a in ax])
    interp = O.interp_enabled(step)
    rows = []
    if dim == 2:
        (x0, y0), (x1, y1) = i0, i1
        rows += [g0[:, y0 + dy, x0 + dx] for dy, dx in _CORNERS_2D]
        g1c = [g1[:, y1 + dy, x1 + dx] for dy, dx in _CORNERS_2D]
        if interp:
            kx, ky = u1[0] - x1, u1[1] - y1
            wx, wy = (1 - kx, 1 - kx, kx, kx), (1 - ky, ky, 1 - ky, ky)
            g1c = [g * wx[j] * wy[j] for j, g in enumerate(g1c)]
        rows.append(g1c[0] + g1c[1] + g1c[2] + g1c[3])
        rows.append(triangular_positional_encoding(torch.stack(u1), pe_channels) if use_tri_pe
                    else positional_encoding(u1, pe_channels))
    else:
        (x0, y0, z0), (x1, y1, z1) = i0, i1
        rows += [g0[:, z0 + dz, y0 + dy, x0 + dx] for dz, dy, dx in (_CORNERS_3D if method == 3 else _CORNERS_3D_V2)]
        g1c = [g1[:, z1 + dz, y1 + dy, x1 + dx] for dz, dy, dx in _CORNERS_3D]
        if interp:
            kx, ky, kz = u1[0] - x1, u1[1] - y1, u1[2] - z1
            g1c = [g * (kx if a else 1 - kx) * (ky if b else 1 - ky) * (kz if c else 1 - kz)
                   for g, (a, b, c) in zip(g1c, _W3D_AS_CODED)]
        acc = g1c[0]
        for g in g1c[1:]:
            acc = acc + g
        rows.append(acc)
        rows.append(triangular_positional_encoding(torch.stack(u1), pe_channels) if method == 3
                    else positional_encoding(u1, pe_channels))
    rows.append(torch.ones(1, size ** dim) * mip_level)
    return torch.cat(rows, dim=0)


def create_decoder_input(fp, coord, fl, mip_level, method, size=None, pe_channels=6, use_tri_pe=True, crop_mip_level=8):
    """image_compression.py:71-167 — python loop over crops, cat along samples, transposed VIEW [N, Cin]."""
    size = O.train_sample_number(mip_level, method, crop_mip_level) if size is None else size
    step = O.step_number(mip_level, fl)
    blocks = [decoder_input_one(fp[2 * fl], fp[2 * fl + 1], c, size, step, mip_level, method, pe_channels, use_tri_pe)
              for c in torch.as_tensor(coord).tolist()]
    return torch.cat(blocks, dim=1).T


def finally_decode_input(fp, image_size, mip_level, level_table, method, origin=None, pe_channels=6, use_tri_pe=True):
    """image_compression.py:170-211."""
    fl = level_table[mip_level]
    dim = 2 if method == 1 else 3
    origin = (0,) * dim if origin is None else origin
    return decoder_input_one(fp[2 * fl], fp[2 * fl + 1], origin, image_size, O.step_number(mip_level, fl), mip_level,
                             method, pe_channels, use_tri_pe).T


class ColorDecoder(nn.Module):
    """image_compression.py:54-68."""

    def __init__(self, cin, hidden=64, cout=3):
        super().__init__()
        self.decoder = nn.Sequential(nn.Linear(cin, hidden), nn.GELU(), nn.Linear(hidden, hidden), nn.GELU(),
                                     nn.Linear(hidden, cout), nn.Sigmoid())

    def forward(self, x):
        return self.decoder(x)


def make_decoder(params):
    w1 = torch.as_tensor(params[0])
    dec = ColorDecoder(w1.shape[1], w1.shape[0], torch.as_tensor(params[4]).shape[0])
    with torch.no_grad():
        for p, v in zip(dec.parameters(), params):
            p.copy_(torch.as_tensor(v))
    return dec


def decode_block(fp, decoder, image_size, mip_level, level_table, method, origin=None, pe_channels=6, use_tri_pe=True):
    """image_compression.py:313-327 — single-shot decode of one square / cubic block -> [S,..,S,Cout]."""
    with torch.no_grad():
        x = finally_decode_input(fp, image_size, mip_level, level_table, method, origin, pe_channels, use_tri_pe)
        out = decoder(x)
    dim = 2 if method == 1 else 3
    return out.reshape((image_size,) * dim + (out.shape[1],))


def decode_image_tiled(fp, decoder, image_size, level_table, tile=1024):
    """decode_image's tiled branch (image_compression.py:329-345) for a 2-D frame at mip 0, `tile`^2 texels per
    call, assembled in a CPU tensor; followed by the 8-bit output quantiser (:406-407)."""
    result = torch.zeros(image_size, image_size, 3)
    n = image_size // tile
    for i in range(n * n):
        x, y = i % n, i // n
        result[tile * x:tile * (x + 1), tile * y:tile * (y + 1)] = decode_block(
            fp, decoder, tile, 0, level_table, 1, origin=(tile * x, tile * y))
    return torch.floor(result * 255 + 0.5).to(torch.uint8)


class Trainer:
    """The body of train_models (image_compression.py:215-269) with torch autograd + torch.optim.Adam +
    CosineAnnealingLR on CPU, exactly as the reference configures them (:361-365)."""

    def __init__(self, fp, decoder, num_epochs, bits=8, method=1, level_table=None, crop_mip_level=8):
        self.fp = [torch.as_tensor(g).clone().requires_grad_(True) for g in fp]
        self.decoder = decoder
        self.bits, self.method, self.table, self.crop_mip_level = bits, method, level_table, crop_mip_level
        self.num_epochs = num_epochs
        self.opt = torch.optim.Adam([{"params": self.fp, "lr": 0.01}, {"params": decoder.parameters(), "lr": 0.005}])
        self.sched = torch.optim.lr_scheduler.CosineAnnealingLR(self.opt, T_max=num_epochs)
        self.loss_fn = nn.MSELoss()
        self.epoch = 0

    def step(self, coord, targets, lod, noise=None):
        fl = self.table[lod]
        x = create_decoder_input(self.fp, coord, fl, lod, self.method, crop_mip_level=self.crop_mip_level)
        if self.epoch < self.num_epochs * 0.95:                                    # :248-254
            if noise is None:
                noise = (torch.rand_like(x) - 0.5) / pow(2, self.bits)
            x = x + torch.as_tensor(noise)
        out = self.decoder(x)
        loss = self.loss_fn(out, torch.as_tensor(targets).reshape(out.shape))
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        self.sched.step()
        q_min = -(pow(2, self.bits) - 1) / pow(2, self.bits + 1)
        with torch.no_grad():                                                      # fp_quantize_clamp, fp_def.py:227-232
            self.fp[2 * fl].clamp_(q_min, 0.5)
            self.fp[2 * fl + 1].clamp_(q_min, 0.5)
        self.epoch += 1
        return float(loss.detach())
