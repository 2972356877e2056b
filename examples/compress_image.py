#!/usr/bin/env python
"""Compress one 2-D image with the fused B200 path, the way `Projects/image_compression.py` does (train -> freeze +
quantise -> save codes -> decode -> PSNR), using the reference's own configuration names.

    python examples/compress_image.py [KEY=VALUE ...]      e.g.  NUM_EPOCHS=2000 FP_BITS=8 IMAGE_SIZE=512

Without an IMAGE_PATH that exists, a synthetic image is used (there are no datasets in this repository)."""
import math
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import neural_image_compression_v2_b200 as nic  # noqa: E402
from neural_image_compression_v2_b200 import fp_def, image_compression as ic, utils, var2  # noqa: E402


def main():
    overrides = dict(a.split("=", 1) for a in sys.argv[1:] if "=" in a)
    typed = {k: (int(v) if v.lstrip("-").isdigit() else (v == "True" if v in ("True", "False") else v)) for k, v in overrides.items()}
    var2.update(**typed)
    dev = torch.device("cuda:0")
    size = var2.IMAGE_SIZE
    if os.path.exists(var2.IMAGE_PATH):
        from PIL import Image
        img = np.asarray(Image.open(var2.IMAGE_PATH).convert("RGB").resize((size, size)), dtype=np.float32).transpose(2, 0, 1) / 255.0
    else:
        import inputs as I
        img = I.make_image(size, 2, seed=0)
    images = [torch.tensor(img, device=dev)]                      # TF_NO_MIP: one level (mips are the caller's business)
    fp, _ = fp_def.create_pyramid(var2.FEATURE_PYRAMID_SIZE, var2.FEATURE_PYRAMID_CHANNELS, var2.FP_BITS, dev, torch.float32,
                                  no_mip=var2.TF_NO_MIP)
    fp = [g.detach() for g in fp]
    decoder = ic.ColorDecoder().to(dev)
    trainer = ic.FusedTrainer(fp, decoder, num_epochs=var2.NUM_EPOCHS, fp_bits=var2.FP_BITS, precision="f16")
    t0 = time.perf_counter()
    for epoch in range(var2.NUM_EPOCHS):
        inputs, coord, lod = ic.random_crop_dataset(images, var2.CROP_SIZE, var2.NUM_CROPS, False)
        loss = trainer.step(coord, inputs, lod)
        if (epoch + 1) % max(1, var2.NUM_EPOCHS // 10) == 0:
            print(f"epoch {epoch + 1:6d}  loss {float(loss):.6f}")
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    codes = fp_def.fp_savable(trainer.fp, var2.FP_BITS)           # what the reference torch.save()s
    frame = ic.decode_codes(codes, decoder, var2.FP_BITS, 0, precision="f16")
    target8 = torch.tensor(np.floor(img.transpose(1, 2, 0) * 255 + 0.5).astype(np.uint8), device=dev)
    psnr = utils.calculate_psnr(target8, frame)
    n = var2.NUM_CROPS * var2.CROP_SIZE ** 2
    bits = sum(c.numel() for c in codes) * var2.FP_BITS + sum(p.numel() for p in decoder.parameters()) * 32
    print(f"{var2.NUM_EPOCHS} steps in {dt:.2f} s ({var2.NUM_EPOCHS * n / dt / 1e6:.0f} Msamples/s), PSNR {psnr:.2f} dB, "
          f"{bits / (size * size):.2f} bit/texel")


if __name__ == "__main__":
    main()
