"""Driver for profiling the general tensor-core decode kernel on a dense 3-D volume: python tools/run_3d.py [method] [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import inputs as I  # noqa: E402
from neural_image_compression_v2_b200 import _lib as L  # noqa: E402
from neural_image_compression_v2_b200 import image_compression as ic, var2  # noqa: E402

method = int(sys.argv[1]) if len(sys.argv) > 1 else 3
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
size = 256
cin = 127 if method == 3 else 79
var2.update(IMAGE_SIZE=size, IMAGE_DIMENSION=3, COMPRESSION_METHOD=method, CROP_MIP_LEVEL=5)
fp = [torch.tensor(g, device=dev) for g in I.make_grids(size, 3, seed=2, no_mip=True, quantized=True)]
dec = ic.ColorDecoder(cin, 64, 3).to(dev)
with torch.no_grad():
    for p, v in zip(dec.parameters_list(), I.make_mlp(cin, seed=3, gain=2.0)):
        p.copy_(torch.tensor(v))
out = torch.empty((size, size, size, 3), dtype=torch.uint8, device=dev)
L.set_option(dev, L.OPT_TIME_KERNELS, 1)
for _ in range(reps):
    ic.decode(fp, dec, 0, precision="f16", out_dtype=torch.uint8, out=out)
torch.cuda.synchronize()
ms, n = L.kernel_time_ms(dev)
print(f"method {method}: {ms / n:.3f} ms/launch, {size ** 3 / (ms / n * 1e-3) / 1e9:.2f} Gtexel/s (kernel only)")
