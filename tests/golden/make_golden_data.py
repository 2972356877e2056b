"""Golden vectors for the data front end (mip-pyramid builder): transforms.Resize + ToTensor exactly as the reference
script applies them to its PIL image (Projects/image_compression.py:433-442), executed here with the torchvision / Pillow
of this container on seeded synthetic images.  Run from the repo root: python tests/golden/make_golden_data.py"""
import os
import sys

import numpy as np
from PIL import Image
from torchvision import transforms

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import inputs as I  # noqa: E402


def resize_ref(img_u8, h, w):
    t = transforms.Compose([transforms.Resize((h, w)), transforms.ToTensor()])       # the reference's transform (:436-439)
    return t(Image.fromarray(img_u8, "RGB")).numpy()


def main():
    out = {}
    # the reference case: a square power-of-two image, every mip down to 1x1 (MAX_MIP_LEVEL = log2(size))
    size = 128
    img = np.ascontiguousarray(np.transpose(np.floor(I.make_image(size, 2, seed=21) * 255 + 0.5), (1, 2, 0)).astype(np.uint8))
    rng = np.random.default_rng(22)
    img = np.clip(img.astype(np.int64) + rng.integers(-40, 41, img.shape), 0, 255).astype(np.uint8)     # exercise rounding
    out["square_src"] = img
    for i in range(8):
        s = size >> i
        f = resize_ref(img, s, s)
        u8 = np.floor(np.transpose(f, (1, 2, 0)) * 255 + 0.5).astype(np.uint8)
        assert np.array_equal((u8.astype(np.float32) / np.float32(255)).transpose(2, 0, 1), f)
        out[f"square_mip{i}"] = u8
    # general geometry: non-square, odd sizes, down- and up-scaling on different axes
    src = rng.integers(0, 256, (61, 45, 3)).astype(np.uint8)
    out["rect_src"] = src
    for (h, w) in ((23, 17), (61, 20), (30, 45), (80, 19), (7, 90)):
        f = resize_ref(src, h, w)
        out[f"rect_{h}x{w}"] = np.floor(np.transpose(f, (1, 2, 0)) * 255 + 0.5).astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "resize.npz"), **out)
    print("wrote resize.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
