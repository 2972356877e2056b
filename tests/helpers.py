"""Shared helpers of the parity tests (tests/ only)."""
import os

import numpy as np
import torch

import inputs as I  # tests/golden/inputs.py (on sys.path via conftest)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def dev():
    return torch.device("cuda:0")


def T(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(dev())


def configure(**kw):
    """Reset the var2 mirror to the reference defaults + overrides (like a fresh `import var2`)."""
    from neural_image_compression_v2_b200 import var2
    var2.update(**kw)
    return var2


def make_decoder(params):
    from neural_image_compression_v2_b200.image_compression import ColorDecoder
    w1 = params[0]
    dec = ColorDecoder(in_channels=w1.shape[1], hidden=w1.shape[0], out_channels=params[4].shape[0]).to(dev())
    sd = dec.state_dict()
    for key, p in zip(["decoder.0.weight", "decoder.0.bias", "decoder.2.weight", "decoder.2.bias",
                       "decoder.4.weight", "decoder.4.bias"], params):
        sd[key] = T(p)
    dec.load_state_dict(sd)
    return dec


def psnr256(a, b):
    """The reference's PSNR (peak 2^8) on 0..255 float images (utils.py:117-130)."""
    mse = float(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2))
    return float("inf") if mse == 0 else 10 * np.log10(256.0 * 256.0 / mse)


def lsb_stats(u8_a, u8_b):
    d = np.abs(u8_a.astype(np.int32) - u8_b.astype(np.int32))
    return float((d <= 1).mean()), float((d == 0).mean()), int(d.max())
