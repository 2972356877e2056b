"""Small driver for profiling the fast-path decode kernel: python tools/run_decode.py [f16|bf16] [reps] [size]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import inputs as I  # noqa: E402
from neural_image_compression_v2_b200 import _lib as L  # noqa: E402
from neural_image_compression_v2_b200 import image_compression as ic, var2  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "f16"
reps = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 3
size = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 4096
dev = torch.device("cuda:0")
var2.update(IMAGE_SIZE=size)
fp = [torch.tensor(g, device=dev) for g in I.make_grids(size, 2, seed=0, no_mip=True, quantized=True)]
dec = ic.ColorDecoder(73, 64, 3).to(dev)
with torch.no_grad():
    for p, v in zip(dec.parameters_list(), I.make_mlp(73, seed=1, gain=2.0)):
        p.copy_(torch.tensor(v))
out = torch.empty((size, size, 3), dtype=torch.uint8, device=dev)
for a in sys.argv:
    if a.startswith("dbg="):
        L.set_option(dev, 100, int(a[4:]))
    if a.startswith("npoly="):
        L.set_option(dev, L.OPT_GELU_POLY, int(a[6:]))
L.set_option(dev, L.OPT_TIME_KERNELS, 1)
for _ in range(reps):
    ic.decode(fp, dec, 0, precision=prec, out_dtype=torch.uint8, out=out)
torch.cuda.synchronize()
ms, n = L.kernel_time_ms(dev)
print(f"{prec}: {ms / n:.4f} ms/launch, {size * size / (ms / n * 1e-3) / 1e9:.2f} Gtexel/s (kernel only)")
