"""Throughput of the GENERAL tensor-core decode kernel: 2-D (fast path disabled), dense 3-D volumes, random-access queries."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import inputs as I  # noqa: E402
from neural_image_compression_v2_b200 import _lib as L  # noqa: E402
from neural_image_compression_v2_b200 import image_compression as ic, var2  # noqa: E402

dev = torch.device("cuda:0")


def decoder(cin, seed):
    dec = ic.ColorDecoder(cin, 64, 3).to(dev)
    with torch.no_grad():
        for p, v in zip(dec.parameters_list(), I.make_mlp(cin, seed=seed, gain=2.0)):
            p.copy_(torch.tensor(v))
    return dec


def timed(fn, n_units, label, reps=5):
    fn()
    torch.cuda.synchronize()
    L.set_option(dev, L.OPT_TIME_KERNELS, 1)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    ms, n = L.kernel_time_ms(dev)
    L.set_option(dev, L.OPT_TIME_KERNELS, 0)
    print(f"{label}: call {s.elapsed_time(e) / reps:.3f} ms, kernel {ms / n:.3f} ms, {n_units / (ms / n * 1e-3) / 1e9:.2f} G/s (kernel), "
          f"{n_units / (s.elapsed_time(e) / reps * 1e-3) / 1e9:.2f} G/s (call)")


# 2-D, general kernel
size = 4096
var2.update(IMAGE_SIZE=size)
fp = [torch.tensor(g, device=dev) for g in I.make_grids(size, 2, seed=0, no_mip=True, quantized=True)]
dec = decoder(73, 1)
out = torch.empty((size, size, 3), dtype=torch.uint8, device=dev)
L.set_option(dev, L.OPT_DISABLE_FAST2D, 1)
timed(lambda: ic.decode(fp, dec, 0, precision="f16", out_dtype=torch.uint8, out=out), size * size, "2-D 4096^2 general kernel")
L.set_option(dev, L.OPT_DISABLE_FAST2D, 0)
del fp, out
# 3-D dense volumes
for method, cin in ((3, 127), (4, 79)):
    size = 256
    var2.update(IMAGE_SIZE=size, IMAGE_DIMENSION=3, COMPRESSION_METHOD=method, CROP_MIP_LEVEL=5)
    fp = [torch.tensor(g, device=dev) for g in I.make_grids(size, 3, seed=2, no_mip=True, quantized=True)]
    dec = decoder(cin, 3)
    out = torch.empty((size, size, size, 3), dtype=torch.uint8, device=dev)
    timed(lambda: ic.decode(fp, dec, 0, precision="f16", out_dtype=torch.uint8, out=out), size ** 3, f"3-D {size}^3 method {method}")
    if method == 3:
        q = torch.randint(0, size, (1 << 24, 3), device=dev)
        timed(lambda: ic.decode_points(fp, dec, q, 0, precision="f16", out_dtype=torch.uint8), q.shape[0], "3-D random access 16.7M queries")
    del fp, out
