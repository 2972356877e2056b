"""One workload of the general / code-resident decode kernels for ncu: python tools/run_profile3d.py dense|query|lut [reps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import inputs as I  # noqa: E402
from neural_image_compression_v2_b200 import _lib as L  # noqa: E402
from neural_image_compression_v2_b200 import fp_def, image_compression as ic, var2  # noqa: E402

mode = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda:0")
var2.update(IMAGE_SIZE=256, IMAGE_DIMENSION=3, COMPRESSION_METHOD=3, CROP_MIP_LEVEL=5)
dec = ic.ColorDecoder(127, 64, 3).to(dev)
with torch.no_grad():
    for p, v in zip(dec.parameters_list(), I.make_mlp(127, seed=3, gain=2.0)):
        p.copy_(torch.tensor(v))
if mode == "lut":
    rng = np.random.default_rng(94)
    lo, hi = I.q_range(8)
    fp = [torch.tensor(rng.uniform(lo, hi, s).astype(np.float32), device=dev) for s in ((12, 18, 18, 18), (12, 10, 10, 10))]
    codes = fp_def.fp_savable(fp, 8)
    q = torch.randint(0, 65, (1 << 24, 3), device=dev)
    fn = lambda: ic.decode_points_codes(codes, dec, q, 8, 0, precision="f16", level_table={0: 0})
    units = q.shape[0]
else:
    fp = [torch.tensor(g, device=dev) for g in I.make_grids(256, 3, seed=2, no_mip=True, quantized=True)]
    if mode == "dense":
        out = torch.empty((256, 256, 256, 3), dtype=torch.uint8, device=dev)
        fn = lambda: ic.decode(fp, dec, 0, precision="f16", out_dtype=torch.uint8, out=out)
        units = 256 ** 3
    else:
        q = torch.randint(0, 256, (1 << 24, 3), device=dev)
        fn = lambda: ic.decode_points(fp, dec, q, 0, precision="f16", out_dtype=torch.uint8)
        units = q.shape[0]
fn()
torch.cuda.synchronize()
L.set_option(dev, L.OPT_TIME_KERNELS, 1)
for _ in range(reps):
    fn()
torch.cuda.synchronize()
ms, n = L.kernel_time_ms(dev)
print(f"{mode}: kernel {ms / n:.3f} ms, {units / (ms / n * 1e-3) / 1e9:.2f} G/s")
