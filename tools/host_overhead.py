"""Host enqueue time vs device time of the fused training step (config 1), with and without device-side sampling:
    python tools/host_overhead.py [steps]
If the host needs longer to ENQUEUE a step than the device needs to run it, the step is host-bound."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import inputs as I  # noqa: E402
from neural_image_compression_v2_b200 import image_compression as ic, var2  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
dev = torch.device("cuda:0")
size, nc, crop = 512, 8, 256
var2.update(IMAGE_SIZE=size)
fp = [torch.tensor(g, device=dev) for g in I.make_grids(size, 2, seed=3, no_mip=True)]
dec = ic.ColorDecoder(73, 64, 3).to(dev)
tr = ic.FusedTrainer(fp, dec, num_epochs=100000, fp_bits=8, seed=1, precision="f16")
img8 = torch.tensor(np.ascontiguousarray(np.floor(np.transpose(I.make_image(size, 2, seed=5), (1, 2, 0)) * 255 + 0.5).astype(np.uint8)), device=dev)
pyr = ic.build_mip_pyramid(img8)
coord = torch.randint(0, size - crop + 1, (nc, 2)).to(dev)
tg = ic.sample_crops(pyr[0], coord, crop)
for name, fn in (("presampled", lambda: tr.step(coord, tg, 0)), ("step_sampled", lambda: tr.step_sampled(pyr))):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    s.record()
    for _ in range(steps):
        fn()
    e.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"{name:14s} host enqueue {1e3 * (t1 - t0) / steps:.4f} ms/step   device {s.elapsed_time(e) / steps:.4f} ms/step")
