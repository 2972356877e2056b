"""Small driver for profiling K1 (gather) alone: python tools/run_gather.py [f16|bf16|f32] [reps]"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import inputs as I  # noqa: E402
from neural_image_compression_v2_b200 import _lib as L  # noqa: E402

dt = sys.argv[1] if len(sys.argv) > 1 else "f16"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
size = 4096
dev = torch.device("cuda:0")
fp = [torch.tensor(g, device=dev) for g in I.make_grids(size, 2, seed=0, no_mip=True, quantized=True)]
td, code = {"f16": (torch.float16, L.DT_F16), "bf16": (torch.bfloat16, L.DT_BF16), "f32": (torch.float32, L.DT_F32)}[dt]
x = torch.empty((size * size, 73), dtype=td, device=dev)
geom = L.make_geom(L.METHOD_2D, fp[0], fp[1], size, 1, -2, 0, 6, L.PE_TRIANGULAR)
h, lib = L.handle(dev), L.load_library()
L.set_option(dev, L.OPT_TIME_KERNELS, 1)
dbg = sum(int(a[4:]) for a in sys.argv if a.startswith("dbg="))     # 16-bit kernel: 1 no row build, 2 no bulk store, 16 / 32: 1 / 4 rows per interval
if dbg:
    L.set_option(dev, L.OPT_DEBUG_KNOCKOUT, dbg)
for _ in range(reps):
    L.check(h, lib.nic_gather(h, C.byref(geom), L.ptr(fp[0]), L.ptr(fp[1]), None, L.ptr(x), code, L.stream_ptr(dev)))
torch.cuda.synchronize()
ms, n = L.kernel_time_ms(dev)
b = 73 * x.element_size() + 12 * 4 * (1 / 16 + 1 / 64)
print(f"{dt} dbg={dbg}: {ms / n:.3f} ms/launch, {b * size * size / (ms / n * 1e-3) / 1e9:.0f} GB/s algorithmic")
