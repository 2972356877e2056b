"""CPU oracle: a numpy restatement of the reference's per-texel decode / training-step path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  The product package
(`neural_image_compression_v2_b200`) never does, and has no CPU fallback.

Parity status: PINNED.  The reference has no golden vectors of its own, so every function
below is checked in `tests/test_oracle_golden.py` against fixtures under `tests/golden/*.npz` that were
produced by executing the UNMODIFIED reference in-process (`tests/golden/make_golden.py`, which loads
`/root/reference/Projects/image_compression.py` via runpy).

All citations are `file:line` relative to `/root/reference/Projects/`.

Conventions:
  * grids are `[C, y, x]` (2-D) / `[C, z, y, x]` (3-D) float32, `x` = FIRST image axis;
  * a step's samples are ordered `n = crop*S^D + ix*S^(D-1) + iy*S^(D-2) (+ iz)` (meshgrid 'ij');
  * `method` 1 = 2-D, 3 = 3-D eight-corner, 4 = 3-D "v2" (tetrahedral G0 + sinusoidal PE).
"""
from __future__ import annotations

import math

import numpy as np

try:  # exact erf for the GELU; scipy ships in the image
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf)

F32 = np.float32


# ----------------------------------------------------------------------------- level tables (a1, a2)
def return_2_power(base_size: int) -> int:
    """fp_def.py:8-15 — floor(log2(base_size)) by repeated halving."""
    count, x = 0, base_size
    while x != 1:
        x //= 2
        count += 1
    return count


def return_pyramid_levels(base_size: int) -> int:
    """fp_def.py:18-21."""
    return (return_2_power(base_size) + 1) // 2


def create_pyramid_mip_levels(image_size: int, base_size: int) -> dict:
    """fp_def.py:24-34 — mip level -> pyramid level, clamp(mip//2 - 1, 0, levels-1)."""
    count = return_2_power(image_size)
    levels = return_pyramid_levels(base_size)
    return {i: min(max(i // 2 - 1, 0), levels - 1) for i in range(count + 1)}


def q_range(num_bits: int):
    """fp_def.py:48-49, 228-229 — clamp range of a grid value."""
    return -(2 ** num_bits - 1) / 2 ** (num_bits + 1), 0.5


def create_pyramid(base_size, channels, num_bits, dim, rng, no_mip=False):
    """fp_def.py:37-56 / 59-78 — shapes and init range only (the RNG stream is torch's; not replicated)."""
    levels = 1 if no_mip else return_pyramid_levels(base_size)
    q_min, q_max = q_range(num_bits)
    out = []
    for i in range(levels * 2):
        size = base_size // (2 ** i)
        shape = (channels,) + (size + 1,) * dim
        out.append(((q_max - q_min) * rng.random(shape, dtype=F32) + F32(q_min)).astype(F32))
    return out, levels


def step_number(mip_level: int, fl: int) -> float:
    """image_compression.py:79 — grid nodes advanced per texel: 2^(mip - 2(fl+1))."""
    return float(2.0 ** (mip_level - (fl + 1) * 2))


def cin_for(method: int, channels: int, pe_channels: int) -> int:
    """var2.py:114-118 — decoder input width."""
    dim = 2 if method in (1, 2) else 3
    corners = 4 if method in (1, 2, 4) else 8
    return channels * (corners + 1) + pe_channels * dim + 1


# ----------------------------------------------------------------------------- positional encodings (a7, a8)
def tri(x, offset):
    """utils.py:226-227 — 2*|((x-o) mod 2) - 1| - 1 with floor-mod (np.mod == torch %)."""
    x = np.asarray(x, dtype=F32)
    return (F32(2) * np.abs(np.mod(x - F32(offset), F32(2)) - F32(1)) - F32(1)).astype(F32)


def triangular_positional_encoding(coord, num_channels):
    """utils.py:211-223 — coord [D, N] float32 -> [num_channels*D, N].

    Row r of axis a (row index a*num_channels + r): r = num_channels - (2*octave + i + 1) with
    (i, offset) in ((0, .5), (1, 0.)), skipping (octave 0, i 0); remaining rows stay 0.
    For num_channels=6: [tri(u/4,0), tri(u/4,.5), tri(u/2,0), tri(u/2,.5), tri(u,0), 0].
    """
    coord = np.asarray(coord, dtype=F32)
    dim = coord.shape[0]
    pe = np.zeros((num_channels * dim, coord.shape[1]), dtype=F32)
    for octave in range(num_channels // 2):
        div = F32(2 ** octave)
        for i, offset in enumerate((0.5, 0.0)):
            if octave == 0 and i == 0:
                continue
            r = num_channels - (octave * 2 + i + 1)
            pe[r:dim * num_channels:num_channels, :] = tri(coord / div, offset)
    return pe


def sin_div_term(num_channels):
    """utils.py:202 — exp(arange(0, nc, 2) * -(ln 1e4 / nc)) evaluated in float32."""
    return np.exp(np.arange(0, num_channels, 2, dtype=F32) * F32(-(math.log(10000.0) / num_channels))).astype(F32)


def positional_encoding(coord, num_channels):
    """utils.py:198-208 — sinusoidal; coord = sequence of D vectors [N] -> [num_channels*D, N]."""
    n = coord[0].shape[0]
    pe = np.zeros((n, num_channels * len(coord)), dtype=F32)
    div = sin_div_term(num_channels)
    for i, c in enumerate(coord):
        arg = (np.asarray(c, dtype=F32)[:, None] * div[None, :]).astype(F32)
        pe[:, num_channels * i:num_channels * (i + 1):2] = np.sin(arg)
        pe[:, num_channels * i + 1:num_channels * (i + 1):2] = np.cos(arg)
    return pe.T


# ----------------------------------------------------------------------------- gather (a3-a6, a9)
# corner tables: offsets are (dz, dy, dx); weights say which factor (k or 1-k) each axis contributes
_CORNERS_2D = [(0, 0), (1, 0), (0, 1), (1, 1)]                       # (dy, dx)  fp_def.py:81-86
_CORNERS_3D = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 0),           # (dz, dy, dx) fp_def.py:96-103
               (0, 0, 1), (1, 0, 1), (0, 1, 1), (1, 1, 1)]
_CORNERS_3D_V2 = [(0, 0, 0), (1, 1, 0), (1, 0, 1), (0, 1, 1)]        # fp_def.py:108-111
# AS-CODED 3-D G1 weights (fp_def.py:176-183): per corner, (use_kx, use_ky, use_kz); True -> k, False -> 1-k.
# Corners 3, 4 and 6 do NOT match their own offsets (reference quirk, reproduced on purpose).
_W3D_AS_CODED = [(0, 0, 0), (0, 0, 1), (0, 1, 0), (1, 0, 0), (1, 1, 0), (1, 0, 1), (0, 1, 1), (1, 1, 1)]


def _axis_vectors(origin, size, step):
    """fp_def.py:116-123 for one axis: u0, i0, u1, i1, k (all exact dyadic float32 / int32)."""
    u0 = ((np.arange(size, dtype=np.int64) + int(origin)).astype(F32) * F32(step)).astype(F32)
    i0 = np.floor(u0).astype(np.int32)
    u1 = (u0 / F32(2)).astype(F32)
    i1 = np.floor(u1).astype(np.int32)
    k = (u1 - i1.astype(F32)).astype(F32)
    return u0, i0, u1, i1, k


def interp_enabled(step) -> bool:
    """fp_def.py:136 — `int(1 // (step/2)) != 1`; False only for step == 2."""
    return int(1 // (step / 2)) != 1


def decoder_input_one(g0, g1, origin, size, step, mip_level, method, pe_channels=6, use_tri_pe=True):
    """One crop / one decode block: image_compression.py:90-96 + fp_def.py:115-145 (2-D),
    :148-184 (method 3), :187-223 (method 4).  Returns X [size^D, Cin] float32."""
    dim = 2 if method == 1 else 3
    ax = [_axis_vectors(origin[a], size, step) for a in range(dim)]   # a = 0:x 1:y 2:z
    mesh = lambda vecs: [m.reshape(-1) for m in np.meshgrid(*vecs, indexing="ij")]
    i0 = mesh([a[1] for a in ax])
    i1 = mesh([a[3] for a in ax])
    u1 = mesh([a[2] for a in ax])
    kk = mesh([a[4] for a in ax])
    rows = []
    one = F32(1)
    if dim == 2:
        x0, y0 = i0
        x1, y1 = i1
        for dy, dx in _CORNERS_2D:
            rows.append(g0[:, y0 + dy, x0 + dx])
        g1c = [g1[:, y1 + dy, x1 + dx] for dy, dx in _CORNERS_2D]
        if interp_enabled(step):
            kx, ky = kk
            wx = [one - kx, one - kx, kx, kx]          # fp_def.py:141-144
            wy = [one - ky, ky, one - ky, ky]
            g1c = [(g * wx[j]) * wy[j] for j, g in enumerate(g1c)]
        rows.append(((g1c[0] + g1c[1]) + g1c[2]) + g1c[3])          # image_compression.py:95
        if use_tri_pe:
            rows.append(triangular_positional_encoding(np.stack(u1), pe_channels))
        else:
            rows.append(positional_encoding(u1, pe_channels))
    else:
        x0, y0, z0 = i0
        x1, y1, z1 = i1
        corners0 = _CORNERS_3D if method == 3 else _CORNERS_3D_V2
        for dz, dy, dx in corners0:
            rows.append(g0[:, z0 + dz, y0 + dy, x0 + dx])
        g1c = [g1[:, z1 + dz, y1 + dy, x1 + dx] for dz, dy, dx in _CORNERS_3D]
        if interp_enabled(step):
            kx, ky, kz = kk
            for j, (ux, uy, uz) in enumerate(_W3D_AS_CODED):
                g1c[j] = ((g1c[j] * (kx if ux else one - kx)) * (ky if uy else one - ky)) * (kz if uz else one - kz)
        acc = g1c[0]
        for j in range(1, 8):
            acc = acc + g1c[j]
        rows.append(acc)
        if method == 3:
            rows.append(triangular_positional_encoding(np.stack(u1), pe_channels))   # fp_def.py:169
        else:
            rows.append(positional_encoding(u1, pe_channels))                        # fp_def.py:208
    rows.append(np.full((1, size ** dim), F32(mip_level), dtype=F32))                 # lod_tensor * mip_level
    return np.ascontiguousarray(np.concatenate(rows, axis=0).astype(F32).T)


def train_sample_number(mip_level, method, crop_mip_level=8):
    """image_compression.py:78 (2-D hard-codes 8) / :110, :144 (3-D uses CROP_MIP_LEVEL)."""
    return 2 ** max(0, (8 if method == 1 else crop_mip_level) - mip_level)


def create_decoder_input(fp, coord, fl, mip_level, method, size=None, pe_channels=6, use_tri_pe=True,
                         crop_mip_level=8):
    """image_compression.py:71-100 / 103-134 / 137-167 — all crops of a training step, crop-major rows."""
    size = train_sample_number(mip_level, method, crop_mip_level) if size is None else size
    step = step_number(mip_level, fl)
    blocks = [decoder_input_one(fp[2 * fl], fp[2 * fl + 1], c, size, step, mip_level, method, pe_channels, use_tri_pe)
              for c in np.asarray(coord)]
    return np.concatenate(blocks, axis=0)


def finally_decode_input(fp, image_size, mip_level, level_table, method, origin=None, pe_channels=6,
                         use_tri_pe=True):
    """image_compression.py:170-181 / 184-196 / 199-211."""
    fl = level_table[mip_level]
    dim = 2 if method == 1 else 3
    origin = (0,) * dim if origin is None else origin
    return decoder_input_one(fp[2 * fl], fp[2 * fl + 1], origin, image_size, step_number(mip_level, fl), mip_level,
                             method, pe_channels, use_tri_pe)


# ----------------------------------------------------------------------------- decoder MLP (a11)
def gelu_erf(z):
    z = np.asarray(z)
    return (z * (0.5 * (1.0 + _erf(z / math.sqrt(2.0))))).astype(z.dtype)


def gelu_erf_grad(z):
    z = np.asarray(z)
    cdf = 0.5 * (1.0 + _erf(z / math.sqrt(2.0)))
    pdf = np.exp(-0.5 * z * z) / math.sqrt(2.0 * math.pi)
    return (cdf + z * pdf).astype(z.dtype)


def mlp_forward(x, params, dtype=F32, return_hidden=False):
    """image_compression.py:57-66 — Linear/GELU(erf)/Linear/GELU(erf)/Linear/Sigmoid.
    `params` = (W1 [H,Cin], b1, W2 [H,H], b2, W3 [Co,H], b3), nn.Linear layout."""
    w1, b1, w2, b2, w3, b3 = [np.asarray(p, dtype=dtype) for p in params]
    x = np.asarray(x, dtype=dtype)
    z1 = x @ w1.T + b1
    h1 = gelu_erf(z1)
    z2 = h1 @ w2.T + b2
    h2 = gelu_erf(z2)
    z3 = h2 @ w3.T + b3
    out = (1.0 / (1.0 + np.exp(-z3))).astype(dtype)
    if return_hidden:
        return out, (z1, h1, z2, h2, z3)
    return out


def quantize_noise(shape, num_bits, rng):
    """image_compression.py:250 — (U[0,1) - 0.5) / 2^bits (distribution only; torch's stream is not replicated)."""
    return ((rng.random(shape, dtype=F32) - F32(0.5)) / F32(2 ** num_bits)).astype(F32)


def decode_block(fp, params, image_size, mip_level, level_table, method, origin=None, pe_channels=6,
                 use_tri_pe=True, dtype=F32):
    """image_compression.py:313-327 — one single-shot decode; returns [S,...,S, Cout]."""
    x = finally_decode_input(fp, image_size, mip_level, level_table, method, origin, pe_channels, use_tri_pe)
    out = mlp_forward(x, params, dtype)
    dim = 2 if method == 1 else 3
    return out.reshape((image_size,) * dim + (out.shape[1],))


# ----------------------------------------------------------------------------- loss + backward (a12, a13)
def train_forward_backward(fp, params, coord, targets, fl, mip_level, method, noise=None, size=None,
                           pe_channels=6, use_tri_pe=True, crop_mip_level=8, dtype=np.float64):
    """Hand-derived backward of image_compression.py:239-265 (autograd in the reference).

    Returns loss, out, dict(grads of W1,b1,W2,b2,W3,b3), dG0, dG1 (dense, grid-shaped).
    Evaluated in `dtype` (float64 by default so it can referee fp32 implementations).
    """
    g0, g1 = fp[2 * fl], fp[2 * fl + 1]
    size = train_sample_number(mip_level, method, crop_mip_level) if size is None else size
    step = step_number(mip_level, fl)
    x = create_decoder_input(fp, coord, fl, mip_level, method, size, pe_channels, use_tri_pe, crop_mip_level)
    xt = x.astype(dtype) + (0 if noise is None else np.asarray(noise, dtype=dtype))
    w1, b1, w2, b2, w3, b3 = [np.asarray(p, dtype=dtype) for p in params]
    out, (z1, h1, z2, h2, z3) = mlp_forward(xt, params, dtype, return_hidden=True)
    tgt = np.asarray(targets, dtype=dtype).reshape(out.shape)
    n = out.shape[0]
    diff = out - tgt
    loss = float(np.mean(diff * diff))
    d_out = 2.0 * diff / (n * out.shape[1])
    dz3 = d_out * out * (1.0 - out)
    grads = {"W3": dz3.T @ h2, "b3": dz3.sum(0)}
    dz2 = (dz3 @ w3) * gelu_erf_grad(z2)
    grads["W2"], grads["b2"] = dz2.T @ h1, dz2.sum(0)
    dz1 = (dz2 @ w2) * gelu_erf_grad(z1)
    grads["W1"], grads["b1"] = dz1.T @ xt, dz1.sum(0)
    dx = dz1 @ w1
    c = g0.shape[0]
    dg0 = np.zeros(g0.shape, dtype=dtype)
    dg1 = np.zeros(g1.shape, dtype=dtype)
    dim = 2 if method == 1 else 3
    per = size ** dim
    corners0 = _CORNERS_2D if dim == 2 else (_CORNERS_3D if method == 3 else _CORNERS_3D_V2)
    n0 = len(corners0)
    interp = interp_enabled(step)
    for ci, origin in enumerate(np.asarray(coord)):
        ax = [_axis_vectors(origin[a], size, step) for a in range(dim)]
        mesh = lambda vecs: [m.reshape(-1) for m in np.meshgrid(*vecs, indexing="ij")]
        i0 = mesh([a[1] for a in ax])
        i1 = mesh([a[3] for a in ax])
        kk = [k.astype(dtype) for k in mesh([a[4] for a in ax])]
        d = dx[ci * per:(ci + 1) * per]
        for j, off in enumerate(corners0):
            idx = tuple(i0[dim - 1 - t] + off[t] for t in range(dim))      # off is (dz,)dy,dx ; i0 is x,y(,z)
            np.add.at(dg0, (slice(None),) + idx, d[:, j * c:(j + 1) * c].T)
        dg = d[:, n0 * c:(n0 + 1) * c]
        corners1 = _CORNERS_2D if dim == 2 else _CORNERS_3D
        for j, off in enumerate(corners1):
            if not interp:
                w = np.ones(per, dtype=dtype)
            elif dim == 2:
                dy, dx_ = off
                w = (kk[0] if dx_ else 1 - kk[0]) * (kk[1] if dy else 1 - kk[1])
            else:
                ux, uy, uz = _W3D_AS_CODED[j]
                w = (kk[0] if ux else 1 - kk[0]) * (kk[1] if uy else 1 - kk[1]) * (kk[2] if uz else 1 - kk[2])
            idx = tuple(i1[dim - 1 - t] + off[t] for t in range(dim))
            np.add.at(dg1, (slice(None),) + idx, (dg * w[:, None]).T)
    return loss, out, grads, dg0, dg1


# ----------------------------------------------------------------------------- optimiser + clamp (a14, a15)
def cosine_lr(lr0, step_index, t_max):
    """torch CosineAnnealingLR closed form with eta_min=0 (image_compression.py:365)."""
    return lr0 * (1.0 + math.cos(math.pi * step_index / t_max)) / 2.0


def adam_update(p, g, m, v, t, lr, beta1=0.9, beta2=0.999, eps=1e-8, dtype=F32):
    """torch.optim.Adam (defaults; image_compression.py:361-364), single tensor, step count t >= 1.
    denom = sqrt(v)/sqrt(1-b2^t) + eps ; p -= lr/(1-b1^t) * m/denom.  Returns new (p, m, v)."""
    p, g, m, v = [np.asarray(a, dtype=dtype) for a in (p, g, m, v)]
    m = (m + (g - m) * dtype(1 - beta1)).astype(dtype)                     # lerp_
    v = (v * dtype(beta2) + (g * g) * dtype(1 - beta2)).astype(dtype)      # mul_ + addcmul_
    bc1 = 1.0 - beta1 ** t
    bc2_sqrt = math.sqrt(1.0 - beta2 ** t)
    denom = (np.sqrt(v) / dtype(bc2_sqrt) + dtype(eps)).astype(dtype)
    p = (p - dtype(lr / bc1) * (m / denom)).astype(dtype)
    return p, m, v


def fp_quantize_clamp(fp, fl, num_bits):
    """fp_def.py:227-232 — clamp the two ACTIVE grids in place."""
    q_min, q_max = q_range(num_bits)
    for j in (0, 1):
        np.clip(fp[2 * fl + j], F32(q_min), F32(q_max), out=fp[2 * fl + j])


# ----------------------------------------------------------------------------- quantisers (a16, a17, a19)
def quantize4fp(g, num_bits):
    """models.py:55-57 — floor(g*(2^b-1) + .5)/(2^b-1); multiply and add are SEPARATE fp32 roundings."""
    s = F32(2 ** num_bits - 1)
    return (np.floor(np.asarray(g, dtype=F32) * s + F32(0.5)) / s).astype(F32)


def save4fp(g, num_bits):
    """models.py:61-64 — code = floor(g*(2^b-1)+.5) + 2^(b-1) - 1 as uint8 (one code per byte)."""
    s = F32(2 ** num_bits - 1)
    r = np.floor(np.asarray(g, dtype=F32) * s + F32(0.5)) + F32(2 ** (num_bits - 1)) - F32(1)
    return r.astype(np.uint8)


def load4fp(code, num_bits):
    """models.py:68-71 with the INTENDED float dtype (the reference passes uint8 at
    image_compression.py:396 and wraps; SURVEY a17 flags that as a bug)."""
    s = F32(2 ** num_bits - 1)
    return ((np.asarray(code).astype(F32) - F32(2 ** (num_bits - 1)) + F32(1)) / s).astype(F32)


def fp_all_quantize(fp, num_bits):
    """fp_def.py:242-247."""
    return [quantize4fp(g, num_bits) for g in fp]


def quantize_to_bit(x, num_bits=8):
    """models.py:29-40 — floor(x*(2^b-1)+.5)/(2^b-1)*(2^b-1) (float, 0..2^b-1)."""
    s = F32(2 ** num_bits - 1)
    return ((np.floor(np.asarray(x, dtype=F32) * s + F32(0.5)) / s) * s).astype(F32)


def calculate_psnr(a, b, num_bits=8):
    """utils.py:117-130 — peak is 2^bits (256), not 2^bits - 1."""
    mse = float(np.mean((np.asarray(a, dtype=F32) - np.asarray(b, dtype=F32)) ** 2))
    if mse == 0:
        return float("inf")
    peak = 2.0 ** num_bits
    return 10.0 * math.log10(peak * peak / mse)


# ------------------------------------------------------------------------------------------------ data front end
def pil_bilinear_coeffs(in_size, out_size):
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc for the BILINEAR filter (src/libImaging/Resample.c, Pillow is
    the third-party dependency behind transforms.Resize at image_compression.py:436): per output index the first input
    index, the tap count and the fixed-point (22 fractional bits) weights.  Double arithmetic in Pillow's order."""
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int64)
    kk = np.zeros((out_size, ksize), np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = []
        ww = 0.0
        for x in range(xmax):
            a = (x + xmin - center + 0.5) * ss
            a = -a if a < 0.0 else a
            w = 1.0 - a if a < 1.0 else 0.0
            k.append(w)
            ww += w
        for x in range(xmax):
            if ww != 0.0:
                k[x] /= ww
            kk[xx, x] = int(-0.5 + k[x] * (1 << 22)) if k[x] < 0 else int(0.5 + k[x] * (1 << 22))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def pil_resize_bilinear(img, out_h, out_w):
    """transforms.Resize((out_h, out_w)) on an 8-bit PIL image, img uint8 [H, W, C] -> uint8 [out_h, out_w, C]:
    horizontal pass, rounding to 8 bits, vertical pass (ImagingResample).  image_compression.py:433-442."""
    img = np.asarray(img, np.uint8)
    h, w, _ = img.shape
    cur = img.astype(np.int64)
    if out_w != w:
        b, kk = pil_bilinear_coeffs(w, out_w)
        nxt = np.zeros((h, out_w, img.shape[2]), np.int64)
        for xx in range(out_w):
            lo, n = b[xx]
            acc = (1 << 21) + np.tensordot(cur[:, lo:lo + n, :], kk[xx, :n], axes=([1], [0]))
            nxt[:, xx, :] = np.clip(acc >> 22, 0, 255)
        cur = nxt
    if out_h != h:
        b, kk = pil_bilinear_coeffs(h, out_h)
        nxt = np.zeros((out_h, cur.shape[1], img.shape[2]), np.int64)
        for yy in range(out_h):
            lo, n = b[yy]
            acc = (1 << 21) + np.tensordot(kk[yy, :n], cur[lo:lo + n], axes=([0], [0]))
            nxt[yy] = np.clip(acc >> 22, 0, 255)
        cur = nxt
    return cur.astype(np.uint8)


def to_tensor(img_u8):
    """transforms.ToTensor: uint8 [H, W, C] -> float32 [C, H, W] / 255 (image_compression.py:437)."""
    return (np.transpose(np.asarray(img_u8, np.uint8), (2, 0, 1)).astype(F32) / F32(255.0)).astype(F32)


def philox4x32_10(seed, offset, counter):
    """Philox4x32-10 (Salmon et al. 2011) as the device code evaluates it (csrc/nic_internal.cuh: philox4x32): key = seed,
    counter words (counter lo, counter hi, offset lo, offset hi) -> four 32-bit words."""
    m32 = 0xFFFFFFFF
    k0, k1 = seed & m32, (seed >> 32) & m32
    c0, c1, c2, c3 = counter & m32, (counter >> 32) & m32, offset & m32, (offset >> 32) & m32
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c0, 0xCD9E8D57 * c2
        hi0, lo0, hi1, lo1 = p0 >> 32, p0 & m32, p1 >> 32, p1 & m32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & m32, lo1, (hi0 ^ c3 ^ k1) & m32, lo0
        k0, k1 = (k0 + 0x9E3779B9) & m32, (k1 + 0xBB67AE85) & m32
    return c0, c1, c2, c3


def random_crop_origins(seed, step, num_crops, size, crop):
    """Crop origins of the device-side sampler: origin[a] = floor(U_a * (size[a] - crop[a] + 1)), U_a = word a of
    Philox(seed, offset=step, counter=crop index) / 2^32 — uniform over the integers [0, size - crop], the distribution
    of torch.randint(0, size - crop + 1) at image_compression.py:40-41."""
    out = np.zeros((num_crops, len(size)), np.int64)
    for b in range(num_crops):
        r = philox4x32_10(seed, step, b)
        for a in range(len(size)):
            out[b, a] = (r[a] * (size[a] - crop[a] + 1)) >> 32
    return out


def crop_targets(dataset, coord, crop):
    """Target slices of random_crop_dataset (image_compression.py:42-47): dataset [C, S, S(, S)] -> [NC, crop^D, C]."""
    dim = dataset.ndim - 1
    out = []
    for o in np.asarray(coord):
        sl = (slice(None),) + tuple(slice(int(o[a]), int(o[a]) + crop) for a in range(dim))
        out.append(dataset[sl].reshape(dataset.shape[0], -1).T)
    return np.stack(out)


def atlas_pack(frames, atlas_size):
    """COMPRESSION_METHOD 2 flatten (image_compression.py:453-460): frames uint8 [T, S, S, C] -> [A, A, C]."""
    t, s, _, c = frames.shape
    per = atlas_size // s
    atlas = np.zeros((atlas_size, atlas_size, c), np.uint8)
    for i in range(t):
        row, col = i // per, i % per
        atlas[row * s:(row + 1) * s, col * s:(col + 1) * s, :] = frames[i]
    return atlas


def atlas_unpack(atlas, size, num_frames=None):
    """The inverse on the decoded image (image_compression.py:410-419)."""
    t = size if num_frames is None else num_frames
    per = atlas.shape[0] // size
    out = np.zeros((t, size, size, atlas.shape[2]), np.uint8)
    for x in range(t):
        row, col = x // per, x % per
        out[x] = atlas[row * size:(row + 1) * size, col * size:(col + 1) * size, :]
    return out
