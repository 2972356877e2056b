"""Tiny invocations of every kernel family, for compute-sanitizer: python tools/sanity_small.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import inputs as I  # noqa: E402
from neural_image_compression_v2_b200 import _lib as L, fp_def  # noqa: E402
from neural_image_compression_v2_b200 import image_compression as ic, var2  # noqa: E402

dev = torch.device("cuda:0")
size = 256
var2.update(IMAGE_SIZE=size)
fp = [torch.tensor(g, device=dev) for g in I.make_grids(size, 2, seed=0, no_mip=True, quantized=True)]
dec = ic.ColorDecoder(73, 64, 3).to(dev)
for prec in ("f16", "bf16", "f32"):
    a = ic.decode(fp, dec, 0, precision=prec, out_dtype=torch.uint8)                                   # fast path (16-bit) / fp32
    b = ic.decode(fp, dec, 0, size=(40, 50), origin=(3, 5), precision=prec, out_dtype=torch.uint8)     # general kernel
x = ic.finally_decode_input_2d(fp, 256, 0)                                                            # tile gather
x2 = ic.finally_decode_input_2d(fp, 37, 0, 5, 9)                                                      # flat gather
img = torch.rand(3, size, size, device=dev)
coord = torch.tensor([[0, 0], [128, 128], [77, 13]])
tg = ic.sample_crops(img, coord, 128)
for prec in ("f16", "bf16", "f32"):
    var2.update(IMAGE_SIZE=size, CROP_MIP_LEVEL=7)
    tr = ic.FusedTrainer([g.clone() for g in fp], dec, num_epochs=10, fp_bits=8, precision=prec)
    # crops of 2^(8-lod): lod 1 -> 128
    mip1 = torch.rand(3, 128, 128, device=dev)
    c1 = torch.tensor([[0, 0], [0, 0]])
    t1 = ic.sample_crops(mip1, c1, 128)
    tr.fp = [g for g in tr.fp]
    loss = tr.step(c1, t1, 1)
codes = fp_def.fp_savable_packed(fp, 4)
back = fp_def.fp_load_packed(codes, 4)
torch.cuda.synchronize()
print("sanity ok", float(loss))
