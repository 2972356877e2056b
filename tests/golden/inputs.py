"""Deterministic synthetic inputs shared by `make_golden.py` (which feeds them to the unmodified reference)
and by the tests (which feed them to the oracle and to the CUDA path).  numpy PCG64 streams only, so the
same seed gives the same bytes in the build container and on the GPU box."""
import math

import numpy as np

F32 = np.float32


def q_range(bits):
    return -(2 ** bits - 1) / 2 ** (bits + 1), 0.5


def pyramid_sizes(image_size, no_mip):
    base = image_size // 4
    count = int(math.log2(base))
    levels = 1 if no_mip else (count + 1) // 2
    return [base // (2 ** i) + 1 for i in range(2 * levels)]


def make_grids(image_size, dim, channels=12, bits=8, seed=0, no_mip=False, quantized=False):
    """Seeded U[q_min, 1/2] grids `[C, s+1, ...]` for every pyramid level."""
    rng = np.random.default_rng(seed)
    q_min, q_max = q_range(bits)
    out = []
    for s in pyramid_sizes(image_size, no_mip):
        g = ((q_max - q_min) * rng.random((channels,) + (s,) * dim, dtype=F32) + F32(q_min)).astype(F32)
        if quantized:
            sc = F32(2 ** bits - 1)
            g = (np.floor(g * sc + F32(0.5)) / sc).astype(F32)
        out.append(g)
    return out


def make_mlp(cin, hidden=64, cout=3, seed=1, gain=1.0):
    """nn.Linear-style U(-1/sqrt(fan_in), 1/sqrt(fan_in)) parameters, optionally sharpened by `gain`."""
    rng = np.random.default_rng(seed)

    def lin(o, i):
        b = gain / math.sqrt(i)
        return (rng.uniform(-b, b, (o, i)).astype(F32), rng.uniform(-b, b, (o,)).astype(F32))

    w1, b1 = lin(hidden, cin)
    w2, b2 = lin(hidden, hidden)
    w3, b3 = lin(cout, hidden)
    return [w1, b1, w2, b2, w3, b3]


def make_image(size, dim, channels=3, seed=2):
    """Smooth + noise target in [0,1], `[channels, S, ..]` float32 with values k/255 (8-bit image)."""
    rng = np.random.default_rng(seed)
    axes = np.meshgrid(*[np.arange(size, dtype=np.float64) / size] * dim, indexing="ij")
    img = np.zeros((channels,) + (size,) * dim)
    for c in range(channels):
        for _ in range(6):
            f = rng.uniform(0.5, 6.0, dim)
            ph = rng.uniform(0, 2 * math.pi)
            img[c] += np.sin(2 * math.pi * sum(f[a] * axes[a] for a in range(dim)) + ph) / 6.0
    img = 127.5 + 100.0 * img + rng.uniform(-8, 8, img.shape)
    return (np.clip(np.floor(img + 0.5), 0, 255) / 255.0).astype(F32)


def box_mips(img, max_mip):
    """2x box-filter mip chain of a `[C, S, S]` image (synthetic stand-in for transforms.Resize)."""
    out = [img]
    for _ in range(max_mip):
        a = out[-1]
        if a.shape[1] == 1:
            break
        a = 0.25 * (a[:, 0::2, 0::2] + a[:, 1::2, 0::2] + a[:, 0::2, 1::2] + a[:, 1::2, 1::2])
        out.append(a.astype(F32))
    return out


def make_noise(n, cin, bits, seed):
    rng = np.random.default_rng(seed)
    return ((rng.random((n, cin), dtype=F32) - F32(0.5)) / F32(2 ** bits)).astype(F32)


def subsample_index(shape, count, seed=12345):
    """Flat indices used to store a sparse check of a large tensor in a fixture."""
    n = int(np.prod(shape))
    rng = np.random.default_rng(seed)
    return np.sort(rng.choice(n, size=min(count, n), replace=False))
